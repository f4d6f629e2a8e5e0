// encode_kernels.cu -- the WRITE side of the numeric field pages on the device: int64 and float64 value blocks ->
// the page bytes banyand/measure/column.go:113-234 writes into fv.bin (encodeInt64Column / encodeFloat64Column), i.e.
//     [encode type][decimal exponent int16 BE, float64 only][first value, order-preserving 8 bytes][zig-zag varint body]
// with the encode type chosen like pkg/encoding/int_list.go:27-53 (Const / DeltaConst / DeltaOfDelta / Delta) and float64
// values turned into decimal integers like pkg/encoding/float.go:30-124.  One warp per block; every decision of the reference's
// sequential loops is an order-independent reduction (all-equal, same-sign, reset counts), the varint body is laid out with a warp
// scan of the varint lengths.  A float64 block that needs the reference's general shortest-digits search (strconv 'e', -1), holds
// NaN / Inf or overflows when brought to a common exponent is NOT encoded here: it is flagged and goes to the CPU writer
// (csrc/part_writer.cc), which also owns the EncodeTypePlain fallback page.  This is the building block of a device-side merger:
// decoded blocks in, fv.bin pages out, byte-identical to the reference writer (tests/test_gpu_parity.py::test_device_page_encoder_*).
#include <cuda_runtime.h>

#include <cstdint>

#include "scan_kernels.cuh"

namespace bydb {

namespace {

__device__ __forceinline__ int64_t wsub(int64_t a, int64_t b) { return static_cast<int64_t>(static_cast<uint64_t>(a) - static_cast<uint64_t>(b)); }

__constant__ double c_p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
__constant__ long long c_i10[19] = {1LL,
                                    10LL,
                                    100LL,
                                    1000LL,
                                    10000LL,
                                    100000LL,
                                    1000000LL,
                                    10000000LL,
                                    100000000LL,
                                    1000000000LL,
                                    10000000000LL,
                                    100000000000LL,
                                    1000000000000LL,
                                    10000000000000LL,
                                    100000000000000LL,
                                    1000000000000000LL,
                                    10000000000000000LL,
                                    100000000000000000LL,
                                    1000000000000000000LL};

// float.go:107-124 floatToDecimal, and the short-decimal equivalent of floatToDecimalSlow (:128-190): the smallest k such that an
// integer m of at most 15 digits has fl(m / 10^k) == |f|; m and 10^k are exact doubles, so the correctly rounded quotient is what
// parsing "m e-k" gives, and no decimal with fewer fractional digits round-trips (smaller k failed): that IS the shortest form
// strconv prints.  Returns false when the general search would be needed.
__device__ bool float_to_decimal(double f, int64_t &mant, int &exp) {
    if (isnan(f) || isinf(f)) return false;
    if (f == 0.0) {
        mant = 0;
        exp = 0;
        return true;
    }
    // Go's int64(f) on amd64 (CVTTSD2SQ): out of range -> MinInt64
    const int64_t u0 = (f >= 9223372036854775808.0 || f < -9223372036854775808.0) ? INT64_MIN : static_cast<int64_t>(f);
    if (__ll2double_rn(u0) == f) {
        int64_t u = u0;
        int e = 0;
        while (u != 0 && u % 10 == 0) {
            u /= 10;
            ++e;
        }
        mant = u;
        exp = e;
        return true;
    }
    const double a = fabs(f);
    if (!(a < 9007199254740992.0 && a >= 1e-15)) return false;
    for (int k = 1; k <= 15; ++k) {
        const double t = __dmul_rn(a, c_p10[k]);
        if (t >= 1e15) break;
        const double m0 = rint(t);
        for (int dm = -1; dm <= 1; ++dm) {
            const double m = m0 + static_cast<double>(dm);
            if (m < 1.0 || m >= 9007199254740992.0) continue;
            if (__ddiv_rn(m, c_p10[k]) == a) {
                const int64_t mi = static_cast<int64_t>(m);
                if (mi % 10 == 0) continue;  // would have matched at k-1
                mant = f < 0 ? -mi : mi;
                exp = -k;
                return true;
            }
        }
    }
    return false;
}

// float.go:199-230 mulPow10Fast / mulPow10Large
__device__ bool mul_pow10(int64_t v, int n, int64_t &out) {
    if (n < 0) return false;
    while (n >= 19) {
        if (v > INT64_MAX / c_i10[18] || v < INT64_MIN / c_i10[18]) return false;
        v *= c_i10[18];
        n -= 18;
    }
    if (n > 0) {
        if (v > INT64_MAX / c_i10[n] || v < INT64_MIN / c_i10[n]) return false;
        v *= c_i10[n];
    }
    out = v;
    return true;
}

__device__ __forceinline__ uint32_t varint_len(int64_t v, uint64_t &zz) {
    zz = (static_cast<uint64_t>(v) << 1) ^ static_cast<uint64_t>(v >> 63);  // int.go:81-99
    return zz == 0 ? 1u : (static_cast<uint32_t>(63 - __clzll(static_cast<long long>(zz))) / 7u + 1u);
}
__device__ __forceinline__ void varint_put(uint8_t *dst, uint64_t zz, uint32_t len) {
    for (uint32_t i = 0; i + 1 < len; ++i) {
        dst[i] = static_cast<uint8_t>(zz & 0x7f) | 0x80u;
        zz >>= 7;
    }
    dst[len - 1] = static_cast<uint8_t>(zz);
}

}  // namespace

__global__ void __launch_bounds__(256) encode_pages_kernel(const __grid_constant__ EncodeParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < p.n_blocks; b += n_warps) {
        const uint64_t v0 = p.block_off[b];
        const uint32_t n = static_cast<uint32_t>(p.block_off[b + 1] - v0);
        uint8_t *page = p.slots + p.slot_off[b];
        const int64_t *a = nullptr;
        int min_exp = 0;
        bool ok = n > 0;
        if (p.is_float && ok) {
            // ---- float64 -> decimal integers with a common exponent (float.go:30-66)
            const double *src = static_cast<const double *>(p.values) + v0;
            int64_t *dec = p.scratch + v0;
            int16_t *exps = p.exps + v0;
            int mn = INT32_MAX;
            for (uint32_t i = lane; i < n; i += 32) {
                int64_t m;
                int e;
                if (!float_to_decimal(src[i], m, e)) {
                    ok = false;
                    m = 0;
                    e = 0;
                }
                dec[i] = m;
                exps[i] = static_cast<int16_t>(e);
                mn = e < mn ? e : mn;
            }
            ok = __all_sync(0xffffffffu, ok);
            mn = __reduce_min_sync(0xffffffffu, mn);
            __syncwarp();
            if (ok) {
                for (uint32_t i = lane; i < n; i += 32) {
                    const int diff = static_cast<int16_t>(exps[i] - mn);
                    if (diff != 0) {
                        int64_t s;
                        if (!mul_pow10(dec[i], diff, s)) ok = false;
                        else dec[i] = s;
                    }
                }
                ok = __all_sync(0xffffffffu, ok);
                __syncwarp();
            }
            min_exp = mn;
            a = dec;
        } else {
            a = static_cast<const int64_t *>(p.values) + v0;
        }
        if (!ok) {
            if (lane == 0) {
                p.page_len[b] = 0;
                p.status[b] = 1;
            }
            continue;
        }
        // ---- encode type (int_list.go:27-53, :112-179): every test of the sequential loops as a reduction
        const int64_t a0 = a[0];
        const int64_t d1 = n > 1 ? wsub(a[1], a0) : 0;
        bool all_eq = true, sign_ok = true, all_d = true, bad_reset = false;
        uint32_t resets = 0;
        for (uint32_t i = 1 + lane; i < n; i += 32) {
            const int64_t v = a[i], pv = a[i - 1];
            all_eq = all_eq && v == a0;
            const int64_t d = wsub(v, pv);
            if (i >= 2) {
                sign_ok = sign_ok && ((d >> 63) & 1) == ((d1 >> 63) & 1);
                all_d = all_d && d == d1;
            }
            if (v < pv) {
                if (v < 0 || v > (pv >> 3)) bad_reset = true;
                ++resets;
            }
        }
        all_eq = __all_sync(0xffffffffu, all_eq);
        sign_ok = __all_sync(0xffffffffu, sign_ok);
        all_d = __all_sync(0xffffffffu, all_d);
        bad_reset = __any_sync(0xffffffffu, bad_reset);
        resets = __reduce_add_sync(0xffffffffu, resets);
        int enc;
        if (all_eq) enc = 1;                                   // EncodeTypeConst
        else if (n >= 2 && sign_ok && all_d) enc = 2;          // EncodeTypeDeltaConst
        else if (n >= 2 && sign_ok) enc = 4;                   // isDelta -> EncodeTypeDeltaOfDelta
        else if (n >= 2 && (a0 < 0 || (!bad_reset && (resets <= 2 || resets < (n >> 3))))) enc = 4;  // isIncremental
        else enc = 3;                                          // EncodeTypeDelta
        // ---- header
        const uint32_t hdr = p.is_float ? 11u : 9u;
        if (lane == 0) {
            page[0] = static_cast<uint8_t>(enc);
            if (p.is_float) {
                page[1] = static_cast<uint8_t>(static_cast<uint16_t>(static_cast<int16_t>(min_exp)) >> 8);
                page[2] = static_cast<uint8_t>(static_cast<uint16_t>(static_cast<int16_t>(min_exp)) & 0xff);
            }
            const uint64_t ord = static_cast<uint64_t>(a0) ^ (1ull << 63);  // convert.Int64ToBytes: sign bit flipped, big endian
            for (int k = 0; k < 8; ++k) page[hdr - 8 + k] = static_cast<uint8_t>(ord >> (56 - 8 * k));
        }
        // ---- body
        uint8_t *body = page + hdr;
        uint32_t off = 0;
        if (enc == 2) {
            uint64_t zz;
            const uint32_t len = varint_len(d1, zz);
            if (lane == 0) varint_put(body, zz, len);
            off = len;
        } else if (enc == 3 || enc == 4) {
            // value j of the body (j = 1 .. n-1): Delta: a[j]-a[j-1]; DeltaOfDelta: d1 first, then second differences (delta.go:26-89)
            for (uint32_t base = 1; base < n; base += 32) {
                const uint32_t j = base + lane;
                uint64_t zz = 0;
                uint32_t len = 0;
                if (j < n) {
                    int64_t v = wsub(a[j], a[j - 1]);
                    if (enc == 4 && j >= 2) v = wsub(v, wsub(a[j - 1], a[j - 2]));
                    len = varint_len(v, zz);
                }
                uint32_t incl = len;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, s);
                    if (lane >= s) incl += o;
                }
                if (j < n) varint_put(body + off + incl - len, zz, len);
                off += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (lane == 0) {
            p.page_len[b] = hdr + off;
            p.status[b] = 0;
        }
    }
}

// pages out of their worst-case slots into one compact run
__global__ void __launch_bounds__(256) gather_pages_kernel(const __grid_constant__ EncodeParams p, const uint64_t *out_off, uint8_t *out) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < p.n_blocks; b += n_warps) {
        const uint8_t *src = p.slots + p.slot_off[b];
        uint8_t *dst = out + out_off[b];
        const uint32_t len = p.page_len[b];
        for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
    }
}

void launch_encode_pages(const EncodeParams &p, int grid, cudaStream_t s) {
    if (p.n_blocks) encode_pages_kernel<<<grid, 256, 0, s>>>(p);
}
void launch_gather_pages(const EncodeParams &p, const uint64_t *out_off, uint8_t *out, int grid, cudaStream_t s) {
    if (p.n_blocks) gather_pages_kernel<<<grid, 256, 0, s>>>(p, out_off, out);
}
void preload_encode_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, encode_pages_kernel);
    cudaFuncGetAttributes(&a, gather_pages_kernel);
    cudaGetLastError();
}

}  // namespace bydb
