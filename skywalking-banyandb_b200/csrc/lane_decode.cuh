// lane_decode.cuh -- the per-lane part of the fast varint decoders of scan_kernels.cu.
//
// Everything here is a plain function of one lane's registers (no shuffles, no shared memory), so the very same source is
// also compiled for the host: tests/native/lane_decode_test.cc runs it against a byte-at-a-time reference decoder
// (full chunks, chunks with bytes outside the page, and the two-chain experiment) without a GPU.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define BYDB_LANE_FN __host__ __device__ __forceinline__
#else
#include <vector_functions.h>  // uint4 / make_uint4 for a plain host compiler
#include <vector_types.h>
#define BYDB_LANE_FN inline
#endif

namespace bydb {

BYDB_LANE_FN uint32_t lane_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return static_cast<uint32_t>(__popc(x));
#else
    return static_cast<uint32_t>(__builtin_popcount(x));
#endif
}
BYDB_LANE_FN int lane_clz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz(static_cast<int>(x));
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
BYDB_LANE_FN int lane_ffs(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs(static_cast<int>(x));
#else
    return __builtin_ffs(static_cast<int>(x));
#endif
}
BYDB_LANE_FN uint32_t lane_byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t src = (static_cast<uint64_t>(b) << 32) | a;  // PRMT without the sign-replicate modes
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= static_cast<uint32_t>((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

BYDB_LANE_FN uint32_t msb4(uint32_t x) {  // gathers the 4 byte-MSBs of x into bits 0..3
    // bit 8j+7 times 2^(21-7j) lands on bit 28+j; no two partial products share a bit, so nothing carries
    return ((x & 0x80808080u) * 0x00204081u) >> 28;
}

// one lane's 32 bytes: local prefix P of its deltas folded over the active rows.
// kFull: all 32 bytes are valid (interior chunk) -> no per-byte validity logic.
// kNeed: bit0 = sum wanted, bit1 = min/max wanted.
enum { kNeedSum = 1, kNeedMinMax = 2 };
constexpr uint32_t kFastLaneBytes = 32;
constexpr uint32_t kFastChunkBytes = 32 * kFastLaneBytes;  // 1 KB per warp iteration

BYDB_LANE_FN uint32_t low_bits(uint32_t n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// 32-bit multiply-add that stays a multiply-add: IMAD runs on the FMA pipe, which this integer kernel otherwise
// leaves idle while LOP3/SHF/SEL/IADD3 saturate the ALU pipe (ncu r01h: alu 82 %, fma 17 %).  Written as inline PTX
// so that neither the front end nor ptxas turns a multiply by a 0/1 flag back into logic ops.
BYDB_LANE_FN uint32_t imad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
#if defined(__CUDA_ARCH__)
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
#else
    d = a * b + c;
#endif
    return d;
}
BYDB_LANE_FN int32_t imad_s32(int32_t a, int32_t b, int32_t c) {
    int32_t d;
#if defined(__CUDA_ARCH__)
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
#else
    d = static_cast<int32_t>(static_cast<uint32_t>(a) * static_cast<uint32_t>(b) + static_cast<uint32_t>(c));
#endif
    return d;
}

// Interior chunk (all 32 bytes valid): the per-byte state machine with its selects, masks and shifts rewritten as
// multiply-adds by 0/1 flags so that the work splits about evenly between the ALU and the FMA pipe.
//   accv += b * mul                (mul = 128^k inside a varint, back to 1 after its terminator)
//   v     = h - s * accv           (zig-zag: accv = 2h + s)
//   P    += v * t ; sumP += P * (t & active)
#ifndef BYDB_UNROLL
#define BYDB_UNROLL 2
#endif
#define BYDB_PRAGMA_(x) _Pragma(#x)
#define BYDB_PRAGMA(x) BYDB_PRAGMA_(x)
#define BYDB_UNROLL_WORDS BYDB_PRAGMA(unroll BYDB_UNROLL)

// kMasked: the chunk holds bytes outside the page (first / last chunk): `reset` = term | ~valid restarts the varint
// state at those bytes too, their payload is zeroed by the caller, and only real terminators (term) count as rows.
// kAllRows (only instantiated by the BYDB_EXP_ALLROWS experiment): every row of the block is active, no window arithmetic.
template <int kNeed, bool kMasked, bool kAllRows = false>
BYDB_LANE_FN void fast_lane_decode_imad(const uint4 &wa, const uint4 &wb, uint32_t term, uint32_t reset, uint32_t aw, uint32_t &accv,
                                                      uint32_t &sh, int32_t &P, int32_t &sumP, int32_t &minP, int32_t &maxP) {
    // 8 words x 4 bytes: the word loop stays rolled so that the body (the hottest code of the whole
    // path) stays resident in the instruction caches of every scheduler
    uint32_t w0 = wa.x, w1 = wa.y, w2 = wa.z, w3 = wa.w, w4 = wb.x, w5 = wb.y, w6 = wb.z, w7 = wb.w;
    uint32_t tm = term, rm = reset;
    uint32_t mul = 1u << sh;
    BYDB_UNROLL_WORDS
    for (int q8 = 0; q8 < 8; ++q8) {
        const uint32_t p = w0 & 0x7f7f7f7fu;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t b = j == 3 ? (p >> 24) : (j == 0 ? (p & 0xffu) : lane_byte_perm(p, 0u, 0x4440u + j));
            const uint32_t t = (tm >> j) & 1u;
            const uint32_t nr = (kMasked ? ((rm >> j) & 1u) : t) ^ 1u;
            accv = imad_u32(b, mul, accv);
            const uint32_t h = accv >> 1, s = accv & 1u;
            const int32_t v = imad_s32(static_cast<int32_t>(s), static_cast<int32_t>(0u - accv), static_cast<int32_t>(h));
            P = imad_s32(v, static_cast<int32_t>(t), P);
            const uint32_t at = kAllRows ? t : (aw & t);  // terminator of an active row
            if (kNeed & kNeedSum) sumP = imad_s32(P, static_cast<int32_t>(at), sumP);
            if (kNeed & kNeedMinMax) {
                // candidate = P at an active terminator, the neutral element otherwise
                const int32_t lo_c = static_cast<int32_t>(imad_u32(at, static_cast<uint32_t>(P) - 0x7fffffffu, 0x7fffffffu));
                const int32_t hi_c = static_cast<int32_t>(imad_u32(at, static_cast<uint32_t>(P) - 0x80000000u, 0x80000000u));
                minP = lo_c < minP ? lo_c : minP;
                maxP = hi_c > maxP ? hi_c : maxP;
            }
            if (!kAllRows) aw >>= t;
            accv = imad_u32(accv, nr, 0u);
            mul = imad_u32(mul, imad_u32(nr, 128u, 0u), nr ^ 1u);
        }
        w0 = w1;
        w1 = w2;
        w2 = w3;
        w3 = w4;
        w4 = w5;
        w5 = w6;
        w6 = w7;
        tm >>= 4;
        if (kMasked) rm >>= 4;
    }
    sh = 31u - static_cast<uint32_t>(lane_clz(mul));
}

// EXPERIMENT (off by default, `make variant EXTRA=-DBYDB_EXP_DUAL`): two independent dependency chains per lane.
// The lane's 32 bytes are decoded as two 16-byte halves with separate state; the second half starts "fresh" and its
// first value is corrected afterwards by what the first half's unfinished tail adds (the same identity that joins
// neighbouring lanes, head_delta).  With P_A the first half's total, a value of the second half has the lane-local
// prefix P_A + (its prefix inside the half), so
//   sumP = sumP_A + cnt_B * P_A + sumP_B,  minP = min(minP_A, P_A + minP_B),  maxP likewise,  P = P_A + P_B,
// and the lane's tail is the second half's.  Equivalence with the single chain was checked on 2e5 random windows
// (tools/sim_dual_chain.py).  Doubles the ILP of the byte loop at the cost of one in-thread head correction.
BYDB_LANE_FN int32_t head_delta(uint32_t w0, uint32_t term, uint32_t prev_acc, uint32_t prev_sh);

template <int kNeed>
BYDB_LANE_FN void imad_byte_step(uint32_t b, uint32_t t, uint32_t &accv, uint32_t &mul, int32_t &P, int32_t &sumP, int32_t &minP,
                                               int32_t &maxP, uint32_t &aw) {
    const uint32_t nr = t ^ 1u;
    accv = imad_u32(b, mul, accv);
    const uint32_t h = accv >> 1, s = accv & 1u;
    const int32_t v = imad_s32(static_cast<int32_t>(s), static_cast<int32_t>(0u - accv), static_cast<int32_t>(h));
    P = imad_s32(v, static_cast<int32_t>(t), P);
    const uint32_t at = aw & t;
    if (kNeed & kNeedSum) sumP = imad_s32(P, static_cast<int32_t>(at), sumP);
    if (kNeed & kNeedMinMax) {
        const int32_t lo_c = static_cast<int32_t>(imad_u32(at, static_cast<uint32_t>(P) - 0x7fffffffu, 0x7fffffffu));
        const int32_t hi_c = static_cast<int32_t>(imad_u32(at, static_cast<uint32_t>(P) - 0x80000000u, 0x80000000u));
        minP = lo_c < minP ? lo_c : minP;
        maxP = hi_c > maxP ? hi_c : maxP;
    }
    aw >>= t;
    accv = imad_u32(accv, nr, 0u);
    mul = imad_u32(mul, imad_u32(nr, 128u, 0u), t);
}

template <int kNeed>
BYDB_LANE_FN void fast_lane_decode_dual(const uint4 &wa, const uint4 &wb, uint32_t term, uint32_t aw, uint32_t &accv, uint32_t &sh,
                                                      int32_t &P, int32_t &sumP, int32_t &minP, int32_t &maxP) {
    uint32_t a0 = wa.x, a1 = wa.y, a2 = wa.z, a3 = wa.w, b0 = wb.x, b1 = wb.y, b2 = wb.z, b3 = wb.w;
    const uint32_t termA = term & 0xffffu, termB = term >> 16;
    const uint32_t nA = lane_popc(termA);
    uint32_t tmA = termA, tmB = termB;
    uint32_t awA = aw, awB = aw >> nA;  // nA <= 16; the first half only ever looks at its own nA low bits
    const uint32_t cntB = lane_popc(awB & low_bits(lane_popc(termB)));
    uint32_t accA = 0, mulA = 1, accB = 0, mulB = 1;
    int32_t PA = 0, sumA = 0, mnA = INT32_MAX, mxA = INT32_MIN, PB = 0, sumB = 0, mnB = INT32_MAX, mxB = INT32_MIN;
    const uint32_t headB = b0;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        const uint32_t pa = a0 & 0x7f7f7f7fu, pb = b0 & 0x7f7f7f7fu;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t ba = j == 3 ? (pa >> 24) : (j == 0 ? (pa & 0xffu) : lane_byte_perm(pa, 0u, 0x4440u + j));
            const uint32_t bb = j == 3 ? (pb >> 24) : (j == 0 ? (pb & 0xffu) : lane_byte_perm(pb, 0u, 0x4440u + j));
            imad_byte_step<kNeed>(ba, (tmA >> j) & 1u, accA, mulA, PA, sumA, mnA, mxA, awA);
            imad_byte_step<kNeed>(bb, (tmB >> j) & 1u, accB, mulB, PB, sumB, mnB, mxB, awB);
        }
        a0 = a1;
        a1 = a2;
        a2 = a3;
        b0 = b1;
        b1 = b2;
        b2 = b3;
        tmA >>= 4;
        tmB >>= 4;
    }
    // the second half's first value continues the first half's unfinished tail
    const uint32_t shA = 31u - static_cast<uint32_t>(lane_clz(mulA));
    if (termB != 0 && shA != 0) {
        const int32_t dlt = head_delta(headB, termB, accA, shA);
        PB += dlt;
        if (kNeed & kNeedSum) sumB += dlt * static_cast<int32_t>(cntB);
        if ((kNeed & kNeedMinMax) && cntB) {
            mnB += dlt;
            mxB += dlt;
        }
    }
    P = PA + PB;
    if (kNeed & kNeedSum) sumP = sumA + static_cast<int32_t>(cntB) * PA + sumB;
    if (kNeed & kNeedMinMax) {
        minP = mnA;
        maxP = mxA;
        if (cntB) {
            const int32_t lo = PA + mnB, hi = PA + mxB;
            minP = lo < minP ? lo : minP;
            maxP = hi > maxP ? hi : maxP;
        }
    }
    accv = accB;
    sh = 31u - static_cast<uint32_t>(lane_clz(mulB));
}


// 4 bits -> 4 byte masks (bit j -> 0xff in byte j): bit j times 2^(7j) lands on bit 8j, nothing else does
BYDB_LANE_FN uint32_t expand4(uint32_t n) { return (((n & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu; }

template <bool kFull, int kNeed>
BYDB_LANE_FN void fast_lane_decode(const uint4 &wa, const uint4 &wb, uint32_t valid, uint32_t term, uint32_t aw, uint32_t &accv,
                                                 uint32_t &sh, int32_t &P, int32_t &sumP, int32_t &minP, int32_t &maxP) {
    if (kFull) {
#ifdef BYDB_EXP_DUAL
        fast_lane_decode_dual<kNeed>(wa, wb, term, aw, accv, sh, P, sumP, minP, maxP);
#else
        fast_lane_decode_imad<kNeed, false>(wa, wb, term, term, aw, accv, sh, P, sumP, minP, maxP);
#endif
    } else {
        const uint4 ma = make_uint4(wa.x & expand4(valid), wa.y & expand4(valid >> 4), wa.z & expand4(valid >> 8), wa.w & expand4(valid >> 12));
        const uint4 mb = make_uint4(wb.x & expand4(valid >> 16), wb.y & expand4(valid >> 20), wb.z & expand4(valid >> 24), wb.w & expand4(valid >> 28));
        fast_lane_decode_imad<kNeed, true>(ma, mb, term, term | ~valid, aw, accv, sh, P, sumP, minP, maxP);
    }
}

// what the previous lane's unfinished tail adds to this lane's first value (narrow mode: the value's own
// bytes are the first <= 3 bytes of the lane, its low bits are prev_acc)
BYDB_LANE_FN int32_t head_delta(uint32_t w0, uint32_t term, uint32_t prev_acc, uint32_t prev_sh) {
    const uint32_t fp = static_cast<uint32_t>(lane_ffs(term) - 1);                 // <= 2
    const uint32_t x = w0 & (0xffffffu >> (8u * (2u - fp)));                     // bytes 0..fp
    const uint32_t hx = (x & 0x7fu) | ((x >> 1) & 0x3f80u) | ((x >> 2) & 0x1fc000u);
    const uint32_t full = prev_acc | (hx << prev_sh);
    const int32_t v_true = static_cast<int32_t>(full >> 1) ^ -static_cast<int32_t>(full & 1u);
    const int32_t v_own = static_cast<int32_t>(hx >> 1) ^ -static_cast<int32_t>(hx & 1u);
    return v_true - v_own;
}


}  // namespace bydb
