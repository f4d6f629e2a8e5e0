// zstd_dec.cuh -- a compact Zstandard (RFC 8878) frame decoder usable from device code.
//
// Needed for row a8 of the hot path (pkg/encoding/bytes.go:306-350 decompressBlock): byte-string blocks
// of 128 bytes or more are zstd frames (klauspost/compress v1.18.5, level 1, no checksum), which is how the
// reference stores the fallback pages of numeric columns (null cells, floats that are not short decimals),
// dictionaries with many values and high-cardinality string tags.
//
// One thread decodes one frame sequentially (the slow lane runs it on lane 0 of a warp); all tables live
// in a caller-provided workspace so nothing large sits in local memory.  The same source compiles for the
// host (tests/test_zstd_dec.py drives it through a tiny C wrapper against libzstd-produced frames).
// Supported: single/multi block frames, raw / RLE / compressed blocks, raw / RLE / Huffman (1 or 4 streams,
// direct or FSE-compressed weights, treeless) literals, predefined / RLE / FSE / repeat sequence tables,
// repeat offsets, frame content size, content checksum skipped.  Not supported: dictionaries, skippable frames.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define BYDB_HD __host__ __device__
#else
#define BYDB_HD
#endif

namespace bydb {
namespace zstd {

constexpr int kMaxHufLog = 11;
constexpr int kMaxLLLog = 9, kMaxMLLog = 9, kMaxOFLog = 8;

struct FseEntry {
    uint8_t symbol;
    uint8_t nbits;
    uint16_t base;  // new-state baseline
};

struct Workspace {
    FseEntry ll[1 << kMaxLLLog];
    FseEntry ml[1 << kMaxMLLog];
    FseEntry of[1 << kMaxOFLog];
    FseEntry wt[1 << 6];               // Huffman weight FSE table
    uint16_t huf[1 << kMaxHufLog];     // symbol | nbits << 8
    uint8_t weights[256];
    int16_t norm[64];                  // FSE normalised counts while building a table
    uint32_t rank[kMaxHufLog + 2];
    int ll_log, ml_log, of_log, huf_log;
    int have_huf, have_ll, have_ml, have_of;
    uint32_t rep[3];
};

enum Err : int { kOk = 0, kErrTrunc = -1, kErrCorrupt = -2, kErrUnsupported = -3, kErrDstFull = -4 };

// ------------------------------------------------------------------ bit readers
struct FwdBits {  // forward, LSB first (FSE table descriptions)
    const uint8_t *p, *end;
    uint64_t acc;
    int nb;
    BYDB_HD void init(const uint8_t *s, const uint8_t *e) {
        p = s;
        end = e;
        acc = 0;
        nb = 0;
    }
    BYDB_HD uint32_t peek(int n) {
        while (nb < n) {
            uint64_t b = p < end ? *p : 0;
            ++p;
            acc |= b << nb;
            nb += 8;
        }
        return static_cast<uint32_t>(acc & ((1ull << n) - 1));
    }
    BYDB_HD void skip(int n) {
        acc >>= n;
        nb -= n;
    }
    BYDB_HD uint32_t read(int n) {
        uint32_t v = peek(n);
        skip(n);
        return v;
    }
    BYDB_HD const uint8_t *byte_pos() const { return p - (nb >> 3); }  // first unread whole byte
};

// (BYDB_ZSTD_ALIGNED_IO compiles the device flavour of the two helpers below for the host, so the CPU tests cover it.)
// little-endian 64-bit load from any address.  Device: two aligned loads and a funnel shift (reads up to 15 bytes past
// p+8 inside the same pair of aligned words -- every buffer handed to the decoder carries >= 16 bytes of readable slack);
// host: a plain unaligned load.
BYDB_HD inline uint64_t load64_le(const uint8_t *p) {
#if defined(__CUDA_ARCH__) || defined(BYDB_ZSTD_ALIGNED_IO)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~static_cast<uintptr_t>(7));
    const int sh = static_cast<int>(a & 7) * 8;
    const uint64_t lo = w[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (w[1] << (64 - sh));
#else
    uint64_t v;
    __builtin_memcpy(&v, p, 8);
    return v;
#endif
}

// copies n bytes, regions must not overlap; device version moves 8 bytes at a time through aligned words
BYDB_HD inline void copy_bytes(uint8_t *d, const uint8_t *s, int64_t n) {
#if defined(__CUDA_ARCH__) || defined(BYDB_ZSTD_ALIGNED_IO)
    while (n > 0 && (reinterpret_cast<uintptr_t>(d) & 7)) {
        *d++ = *s++;
        --n;
    }
    if (n >= 8) {
        const int sh = static_cast<int>(reinterpret_cast<uintptr_t>(s) & 7) * 8;
        const uint64_t *sp = reinterpret_cast<const uint64_t *>(s - (sh >> 3));
        uint64_t *dp = reinterpret_cast<uint64_t *>(d);
        const int64_t words = n >> 3;
        if (sh == 0) {
            for (int64_t i = 0; i < words; ++i) dp[i] = sp[i];
        } else {
            uint64_t lo = sp[0];
            for (int64_t i = 0; i < words; ++i) {
                const uint64_t hi = sp[i + 1];
                dp[i] = (lo >> sh) | (hi << (64 - sh));
                lo = hi;
            }
        }
        d += words << 3;
        s += words << 3;
        n &= 7;
    }
    while (n > 0) {
        *d++ = *s++;
        --n;
    }
#else
    __builtin_memcpy(d, s, static_cast<size_t>(n));
#endif
}

// LZ77 match copy: dst[i] = dst[i - offset], forward, may overlap itself
BYDB_HD inline void copy_match(uint8_t *d, int64_t offset, int64_t n) {
    if (offset >= 16) {
        while (n > 0) {  // pieces of at most `offset` bytes never overlap their source
            const int64_t c = n < offset ? n : offset;
            copy_bytes(d, d - offset, c);
            d += c;
            n -= c;
        }
    } else {
        for (int64_t k = 0; k < n; ++k) d[k] = d[k - offset];
    }
}

struct BackBits {  // backward: starts at the last byte, below its highest set bit
    const uint8_t *start;
    int64_t len;
    int64_t bitpos;  // number of unread bits
    uint64_t win;    // stream bits [wbit, wbit + 64)
    int64_t wbit;
    BYDB_HD int init(const uint8_t *s, int64_t n) {
        start = s;
        len = n;
        win = 0;
        wbit = static_cast<int64_t>(1) << 60;  // no window yet
        if (n <= 0) return kErrTrunc;
        uint8_t last = s[n - 1];
        if (last == 0) return kErrCorrupt;
        int hb = 7;
        while (!((last >> hb) & 1)) --hb;
        bitpos = (n - 1) * 8 + hb;
        return kOk;
    }
    // reads n bits (n <= 32); bits below the start of the stream read as zero (allowed at the end)
    BYDB_HD uint32_t read(int n) {
        if (n == 0) return 0;
        const int64_t old = bitpos;
        bitpos -= n;
        const uint64_t mask = (1ull << n) - 1ull;
        if (bitpos >= wbit) return static_cast<uint32_t>((win >> (bitpos - wbit)) & mask);  // window top >= old by construction
        if (bitpos >= 0 && len >= 8) {
            // slide the window so that its top byte is the one holding bit old-1: at least 57 fresh bits
            int64_t wb = ((old + 7) >> 3) - 8;
            if (wb < 0) wb = 0;
            win = load64_le(start + wb);
            wbit = wb << 3;
            return static_cast<uint32_t>((win >> (bitpos - wbit)) & mask);
        }
        return read_slow(n);
    }
    BYDB_HD uint32_t read_slow(int n) {  // bitpos already moved
        uint64_t v = 0;
        int64_t bp = bitpos;
        int got = 0;
        if (bp < 0) {  // partially (or fully) past the beginning: the missing low bits are zero
            got = static_cast<int>(-bp < n ? -bp : n);
            bp = 0;
        }
        while (got < n) {
            const int64_t byte = bp >> 3;
            const int off = static_cast<int>(bp & 7);
            const int take = (8 - off) < (n - got) ? (8 - off) : (n - got);
            v |= static_cast<uint64_t>((start[byte] >> off) & ((1u << take) - 1)) << got;
            got += take;
            bp += take;
        }
        return static_cast<uint32_t>(v);
    }
    BYDB_HD bool overrun() const { return bitpos < -64; }
};

BYDB_HD inline int highbit(uint32_t v) {
    int r = 0;
    while (v >>= 1) ++r;
    return r;
}

// ------------------------------------------------------------------ FSE
// builds a decoding table from normalised counts (RFC 8878 4.1.1); returns 0 or an error
BYDB_HD inline int fse_build(FseEntry *table, int log, const int16_t *norm, int nsym) {
    const int size = 1 << log;
    int high = size - 1;
    for (int s = 0; s < nsym; ++s)
        if (norm[s] == -1) {
            table[high].symbol = static_cast<uint8_t>(s);
            --high;
        }
    const int step = (size >> 1) + (size >> 3) + 3;
    const int mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; ++s) {
        for (int i = 0; i < norm[s]; ++i) {
            table[pos].symbol = static_cast<uint8_t>(s);
            do {
                pos = (pos + step) & mask;
            } while (pos > high);
        }
    }
    if (pos != 0) return kErrCorrupt;
    // per-symbol next-state counters (reuse a small on-stack array: <= 64 symbols)
    uint16_t next[64];
    for (int s = 0; s < nsym; ++s) next[s] = static_cast<uint16_t>(norm[s] == -1 ? 1 : norm[s]);
    for (int i = 0; i < size; ++i) {
        const int s = table[i].symbol;
        const uint32_t ns = next[s]++;
        const int nb = log - highbit(ns);
        table[i].nbits = static_cast<uint8_t>(nb);
        table[i].base = static_cast<uint16_t>((ns << nb) - size);
    }
    return kOk;
}

// reads an FSE table description; returns bytes consumed or an error (<0)
BYDB_HD inline int fse_read(FseEntry *table, int *log_out, int max_log, int max_sym, int16_t *norm, const uint8_t *src, const uint8_t *end) {
    if (end - src < 1) return kErrTrunc;
    FwdBits br;
    br.init(src, end);
    const int log = static_cast<int>(br.read(4)) + 5;
    if (log > max_log) return kErrCorrupt;
    int remaining = (1 << log) + 1;
    int threshold = 1 << log;
    int nbits = log + 1;
    int sym = 0;
    bool prev0 = false;
    while (remaining > 1 && sym <= max_sym) {
        if (prev0) {
            // repeat flags: 2 bits at a time, 3 = keep going
            for (;;) {
                const uint32_t r = br.read(2);
                for (uint32_t i = 0; i < r && sym <= max_sym; ++i) norm[sym++] = 0;
                if (r != 3) break;
            }
            prev0 = false;
            continue;
        }
        const int maxv = 2 * threshold - 1 - remaining;
        int count;
        uint32_t v = br.peek(nbits);
        if (static_cast<int>(v & (threshold - 1)) < maxv) {
            count = static_cast<int>(v & (threshold - 1));
            br.skip(nbits - 1);
        } else {
            count = static_cast<int>(v & (2 * threshold - 1));
            if (count >= threshold) count -= maxv;
            br.skip(nbits);
        }
        --count;  // value 0 means probability -1 ("less than 1")
        remaining -= count < 0 ? -count : count;
        if (sym > max_sym) return kErrCorrupt;
        norm[sym++] = static_cast<int16_t>(count);
        prev0 = count == 0;
        while (remaining < threshold) {
            --nbits;
            threshold >>= 1;
        }
    }
    if (remaining != 1 || sym > max_sym + 1) return kErrCorrupt;
    if (br.byte_pos() > end) return kErrTrunc;
    const int rc = fse_build(table, log, norm, sym);
    if (rc) return rc;
    *log_out = log;
    // consumed bytes: every byte touched, including the partially used last one
    const uint8_t *np = br.p - (br.nb >> 3);
    return static_cast<int>(np - src);
}

BYDB_HD inline void fse_rle(FseEntry *table, int *log_out, uint8_t symbol) {
    table[0].symbol = symbol;
    table[0].nbits = 0;
    table[0].base = 0;
    *log_out = 0;
}

// predefined distributions, RFC 8878 3.1.1.3.2.2
BYDB_HD inline int fse_predefined(Workspace *ws, int which) {
    const int16_t ll[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    const int16_t ml[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                            1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    const int16_t of[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    if (which == 0) {
        ws->ll_log = 6;
        return fse_build(ws->ll, 6, ll, 36);
    }
    if (which == 1) {
        ws->of_log = 5;
        return fse_build(ws->of, 5, of, 29);
    }
    ws->ml_log = 6;
    return fse_build(ws->ml, 6, ml, 53);
}

// ------------------------------------------------------------------ Huffman
BYDB_HD inline int huf_build(Workspace *ws, int nweights) {
    // the last weight is implied: the total of 2^(w-1) must be a power of two
    uint32_t total = 0;
    for (int i = 0; i < nweights; ++i) {
        if (ws->weights[i] > kMaxHufLog) return kErrCorrupt;
        if (ws->weights[i]) total += 1u << (ws->weights[i] - 1);
    }
    if (total == 0) return kErrCorrupt;
    const int log = highbit(total) + 1;
    if (log > kMaxHufLog) return kErrCorrupt;
    const uint32_t rest = (1u << log) - total;
    if (rest == 0 || (rest & (rest - 1))) return kErrCorrupt;
    if (nweights >= 256) return kErrCorrupt;
    ws->weights[nweights] = static_cast<uint8_t>(highbit(rest) + 1);
    const int nsym = nweights + 1;
    for (int w = 0; w <= kMaxHufLog + 1; ++w) ws->rank[w] = 0;
    for (int s = 0; s < nsym; ++s) ws->rank[ws->weights[s]]++;
    // start index of each weight class: lower weights (longer codes) first
    uint32_t next = 0;
    for (int w = 1; w <= log; ++w) {
        const uint32_t cnt = ws->rank[w];
        ws->rank[w] = next;
        next += cnt << (w - 1);
    }
    if (next != (1u << log)) return kErrCorrupt;
    for (int s = 0; s < nsym; ++s) {
        const int w = ws->weights[s];
        if (!w) continue;
        const uint32_t len = 1u << (w - 1);
        const uint16_t e = static_cast<uint16_t>(s | ((log + 1 - w) << 8));
        const uint32_t at = ws->rank[w];
        for (uint32_t i = 0; i < len; ++i) ws->huf[at + i] = e;
        ws->rank[w] += len;
    }
    ws->huf_log = log;
    ws->have_huf = 1;
    return kOk;
}

// Huffman tree description; returns bytes consumed or error
BYDB_HD inline int huf_read_tree(Workspace *ws, const uint8_t *src, const uint8_t *end) {
    if (end - src < 1) return kErrTrunc;
    const int hb = src[0];
    int nweights;
    int used;
    if (hb >= 128) {
        nweights = hb - 127;
        const int bytes = (nweights + 1) / 2;
        if (end - src < 1 + bytes) return kErrTrunc;
        for (int i = 0; i < nweights; ++i) {
            const uint8_t b = src[1 + i / 2];
            ws->weights[i] = (i & 1) ? (b & 0xf) : (b >> 4);
        }
        used = 1 + bytes;
    } else {
        if (end - src < 1 + hb) return kErrTrunc;
        int wlog = 0;
        const int r = fse_read(ws->wt, &wlog, 6, 11, ws->norm, src + 1, src + 1 + hb);
        if (r < 0) return r;
        BackBits bb;
        const int rc = bb.init(src + 1 + r, hb - r);
        if (rc) return rc;
        uint32_t s1 = bb.read(wlog), s2 = bb.read(wlog);
        nweights = 0;
        // two interleaved states; the stream has no end marker: it ends when a state update runs out of bits,
        // and the other state's pending symbol is then the last weight
        for (;;) {
            if (nweights >= 254) return kErrCorrupt;
            ws->weights[nweights++] = ws->wt[s1].symbol;
            s1 = ws->wt[s1].base + bb.read(ws->wt[s1].nbits);
            if (bb.bitpos < 0) {
                ws->weights[nweights++] = ws->wt[s2].symbol;
                break;
            }
            ws->weights[nweights++] = ws->wt[s2].symbol;
            s2 = ws->wt[s2].base + bb.read(ws->wt[s2].nbits);
            if (bb.bitpos < 0) {
                if (nweights >= 255) return kErrCorrupt;
                ws->weights[nweights++] = ws->wt[s1].symbol;
                break;
            }
        }
        used = 1 + hb;
    }
    const int rc = huf_build(ws, nweights);
    return rc ? rc : used;
}

BYDB_HD inline int huf_decode_stream(const Workspace *ws, const uint8_t *src, int64_t len, uint8_t *dst, int64_t n) {
    BackBits bb;
    int rc = bb.init(src, len);
    if (rc) return rc;
    const int log = ws->huf_log;
    const uint32_t smask = (1u << log) - 1;
    uint32_t state = bb.read(log);
    for (int64_t i = 0; i < n; ++i) {
        const uint16_t e = ws->huf[state];
        dst[i] = static_cast<uint8_t>(e & 0xff);
        const int nb = e >> 8;
        state = ((state << nb) & smask) | bb.read(nb);
    }
    // all bits must be consumed: after the last symbol the reader sits `log` bits before the stream start
    if (bb.bitpos != -static_cast<int64_t>(log)) return kErrCorrupt;
    return kOk;
}

// the four streams of a literals section decoded in lock step: four independent dependency chains in one thread
// (table look-up -> bit count -> next state) hide each other's latency -- that is what the format's 4 streams are for
BYDB_HD inline int huf_decode_4streams(const Workspace *ws, const uint8_t *s0, int64_t l0, const uint8_t *s1, int64_t l1, const uint8_t *s2, int64_t l2,
                                       const uint8_t *s3, int64_t l3, uint8_t *dst, int64_t per, int64_t last) {
    BackBits b0, b1, b2, b3;
    int rc = b0.init(s0, l0);
    if (!rc) rc = b1.init(s1, l1);
    if (!rc) rc = b2.init(s2, l2);
    if (!rc) rc = b3.init(s3, l3);
    if (rc) return rc;
    const int log = ws->huf_log;
    const uint32_t smask = (1u << log) - 1;
    const uint16_t *tab = ws->huf;
    uint32_t t0 = b0.read(log), t1 = b1.read(log), t2 = b2.read(log), t3 = b3.read(log);
    uint8_t *d0 = dst, *d1 = dst + per, *d2 = dst + 2 * per, *d3 = dst + 3 * per;
    const int64_t common = last < per ? last : per;
    for (int64_t i = 0; i < common; ++i) {
        const uint16_t e0 = tab[t0], e1 = tab[t1], e2 = tab[t2], e3 = tab[t3];
        d0[i] = static_cast<uint8_t>(e0);
        d1[i] = static_cast<uint8_t>(e1);
        d2[i] = static_cast<uint8_t>(e2);
        d3[i] = static_cast<uint8_t>(e3);
        t0 = ((t0 << (e0 >> 8)) & smask) | b0.read(e0 >> 8);
        t1 = ((t1 << (e1 >> 8)) & smask) | b1.read(e1 >> 8);
        t2 = ((t2 << (e2 >> 8)) & smask) | b2.read(e2 >> 8);
        t3 = ((t3 << (e3 >> 8)) & smask) | b3.read(e3 >> 8);
    }
    for (int64_t i = common; i < per; ++i) {  // the last stream is the short one
        const uint16_t e0 = tab[t0], e1 = tab[t1], e2 = tab[t2];
        d0[i] = static_cast<uint8_t>(e0);
        d1[i] = static_cast<uint8_t>(e1);
        d2[i] = static_cast<uint8_t>(e2);
        t0 = ((t0 << (e0 >> 8)) & smask) | b0.read(e0 >> 8);
        t1 = ((t1 << (e1 >> 8)) & smask) | b1.read(e1 >> 8);
        t2 = ((t2 << (e2 >> 8)) & smask) | b2.read(e2 >> 8);
    }
    for (int64_t i = common; i < last; ++i) {  // (not reachable for conforming frames: last <= per)
        const uint16_t e3 = tab[t3];
        d3[i] = static_cast<uint8_t>(e3);
        t3 = ((t3 << (e3 >> 8)) & smask) | b3.read(e3 >> 8);
    }
    const int64_t endpos = -static_cast<int64_t>(log);
    if (b0.bitpos != endpos || b1.bitpos != endpos || b2.bitpos != endpos || b3.bitpos != endpos) return kErrCorrupt;
    return kOk;
}

// ------------------------------------------------------------------ sequences
BYDB_HD inline void seq_tables(uint32_t code, int which, uint32_t *base, int *bits) {
    const uint32_t ll_base[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
    const uint8_t ll_bits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    const uint32_t ml_base[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                                  35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
    const uint8_t ml_bits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    if (which == 0) {
        *base = ll_base[code];
        *bits = ll_bits[code];
    } else {
        *base = ml_base[code];
        *bits = ml_bits[code];
    }
}

// decodes one compressed block into dst[dpos..]; lit is a scratch area of at least 128 KB for the literals
BYDB_HD inline int64_t decode_block(Workspace *ws, const uint8_t *src, int64_t len, uint8_t *dst, int64_t dpos, int64_t dcap, uint8_t *lit) {
    const uint8_t *p = src, *end = src + len;
    if (len < 1) return kErrTrunc;
    // ---- literals section
    const int ltype = p[0] & 3, sf = (p[0] >> 2) & 3;
    int64_t regen = 0, comp = 0;
    int nstreams = 1;
    const uint8_t *literals = nullptr;
    if (ltype < 2) {
        int hs;
        if (sf == 0 || sf == 2) {
            regen = p[0] >> 3;
            hs = 1;
        } else if (sf == 1) {
            if (len < 2) return kErrTrunc;
            regen = (p[0] >> 4) + (static_cast<int64_t>(p[1]) << 4);
            hs = 2;
        } else {
            if (len < 3) return kErrTrunc;
            regen = (p[0] >> 4) + (static_cast<int64_t>(p[1]) << 4) + (static_cast<int64_t>(p[2]) << 12);
            hs = 3;
        }
        p += hs;
        if (regen > 131072) return kErrCorrupt;
        if (ltype == 0) {
            if (end - p < regen) return kErrTrunc;
            literals = p;
            p += regen;
        } else {
            if (end - p < 1) return kErrTrunc;
            for (int64_t i = 0; i < regen; ++i) lit[i] = p[0];
            literals = lit;
            p += 1;
        }
    } else {
        int hs;
        uint64_t h = 0;
        if (sf < 2) {
            if (len < 3) return kErrTrunc;
            h = p[0] | (static_cast<uint64_t>(p[1]) << 8) | (static_cast<uint64_t>(p[2]) << 16);
            regen = (h >> 4) & 0x3ff;
            comp = (h >> 14) & 0x3ff;
            hs = 3;
            nstreams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            if (len < 4) return kErrTrunc;
            h = p[0] | (static_cast<uint64_t>(p[1]) << 8) | (static_cast<uint64_t>(p[2]) << 16) | (static_cast<uint64_t>(p[3]) << 24);
            regen = (h >> 4) & 0x3fff;
            comp = (h >> 18) & 0x3fff;
            hs = 4;
            nstreams = 4;
        } else {
            if (len < 5) return kErrTrunc;
            h = p[0] | (static_cast<uint64_t>(p[1]) << 8) | (static_cast<uint64_t>(p[2]) << 16) | (static_cast<uint64_t>(p[3]) << 24) |
                (static_cast<uint64_t>(p[4]) << 32);
            regen = (h >> 4) & 0x3ffff;
            comp = (h >> 22) & 0x3ffff;
            hs = 5;
            nstreams = 4;
        }
        p += hs;
        if (end - p < comp || regen > 131072) return kErrTrunc;
        const uint8_t *ls = p, *le = p + comp;
        p += comp;
        if (ltype == 2) {
            const int used = huf_read_tree(ws, ls, le);
            if (used < 0) return used;
            ls += used;
        } else if (!ws->have_huf) {
            return kErrCorrupt;
        }
        if (nstreams == 1) {
            const int rc = huf_decode_stream(ws, ls, le - ls, lit, regen);
            if (rc) return rc;
        } else {
            if (le - ls < 6) return kErrTrunc;
            const int64_t s1 = ls[0] | (ls[1] << 8), s2 = ls[2] | (ls[3] << 8), s3 = ls[4] | (ls[5] << 8);
            ls += 6;
            const int64_t s4 = (le - ls) - s1 - s2 - s3;
            if (s4 < 1) return kErrCorrupt;
            const int64_t per = (regen + 3) / 4;
            const int64_t last = regen - 3 * per;
            if (last < 0) return kErrCorrupt;
            if (s1 < 1 || s2 < 1 || s3 < 1) return kErrCorrupt;
            const int rc = huf_decode_4streams(ws, ls, s1, ls + s1, s2, ls + s1 + s2, s3, ls + s1 + s2 + s3, s4, lit, per, last);
            if (rc) return rc;
        }
        literals = lit;
    }
    // ---- sequences section
    if (end - p < 1) return kErrTrunc;
    int64_t nseq = p[0];
    if (nseq == 0) {
        p += 1;
    } else if (nseq < 128) {
        p += 1;
    } else if (nseq < 255) {
        if (end - p < 2) return kErrTrunc;
        nseq = ((nseq - 128) << 8) + p[1];
        p += 2;
    } else {
        if (end - p < 3) return kErrTrunc;
        nseq = p[1] + (static_cast<int64_t>(p[2]) << 8) + 0x7f00;
        p += 3;
    }
    int64_t lpos = 0;
    if (nseq > 0) {
        if (end - p < 1) return kErrTrunc;
        const int modes = *p++;
        const int mll = (modes >> 6) & 3, mof = (modes >> 4) & 3, mml = (modes >> 2) & 3;
        // tables in the order LL, OF, ML
        for (int t = 0; t < 3; ++t) {
            const int mode = t == 0 ? mll : (t == 1 ? mof : mml);
            FseEntry *tab = t == 0 ? ws->ll : (t == 1 ? ws->of : ws->ml);
            int *log = t == 0 ? &ws->ll_log : (t == 1 ? &ws->of_log : &ws->ml_log);
            int *have = t == 0 ? &ws->have_ll : (t == 1 ? &ws->have_of : &ws->have_ml);
            if (mode == 0) {
                const int rc = fse_predefined(ws, t);
                if (rc) return rc;
                *have = 1;
            } else if (mode == 1) {
                if (end - p < 1) return kErrTrunc;
                fse_rle(tab, log, *p++);
                *have = 1;
            } else if (mode == 2) {
                const int max_log = t == 0 ? kMaxLLLog : (t == 1 ? kMaxOFLog : kMaxMLLog);
                const int max_sym = t == 0 ? 35 : (t == 1 ? 31 : 52);
                const int used = fse_read(tab, log, max_log, max_sym, ws->norm, p, end);
                if (used < 0) return used;
                p += used;
                *have = 1;
            } else if (!*have) {
                return kErrCorrupt;
            }
        }
        BackBits bb;
        const int rc = bb.init(p, end - p);
        if (rc) return rc;
        uint32_t sll = bb.read(ws->ll_log), sof = bb.read(ws->of_log), sml = bb.read(ws->ml_log);
        for (int64_t i = 0; i < nseq; ++i) {
            const uint32_t ofc = ws->of[sof].symbol, mlc = ws->ml[sml].symbol, llc = ws->ll[sll].symbol;
            if (ofc > 31 || mlc > 52 || llc > 35) return kErrCorrupt;
            const uint64_t ofv = (1ull << ofc) + bb.read(static_cast<int>(ofc));
            uint32_t mlb, llb;
            int mlx, llx;
            seq_tables(mlc, 1, &mlb, &mlx);
            seq_tables(llc, 0, &llb, &llx);
            const int64_t mlen = mlb + bb.read(mlx);
            const int64_t llen = llb + bb.read(llx);
            // offset with the repeat-offset rules (RFC 8878 3.1.1.5)
            uint64_t offset;
            if (ofv > 3) {
                offset = ofv - 3;
                ws->rep[2] = ws->rep[1];
                ws->rep[1] = ws->rep[0];
                ws->rep[0] = static_cast<uint32_t>(offset);
            } else {
                uint32_t idx = static_cast<uint32_t>(ofv) - 1;  // 0,1,2
                if (llen == 0) ++idx;
                if (idx == 0) {
                    offset = ws->rep[0];
                } else {
                    offset = idx < 3 ? ws->rep[idx] : ws->rep[0] - 1;
                    if (offset == 0) return kErrCorrupt;
                    if (idx > 1) ws->rep[2] = ws->rep[1];
                    ws->rep[1] = ws->rep[0];
                    ws->rep[0] = static_cast<uint32_t>(offset);
                }
            }
            if (i + 1 < nseq) {
                sll = ws->ll[sll].base + bb.read(ws->ll[sll].nbits);
                sml = ws->ml[sml].base + bb.read(ws->ml[sml].nbits);
                sof = ws->of[sof].base + bb.read(ws->of[sof].nbits);
            }
            if (bb.overrun()) return kErrCorrupt;
            // execute
            if (lpos + llen > regen) return kErrCorrupt;
            if (dpos + llen + mlen > dcap) return kErrDstFull;
            copy_bytes(dst + dpos, literals + lpos, llen);
            dpos += llen;
            lpos += llen;
            if (static_cast<int64_t>(offset) > dpos) return kErrCorrupt;
            copy_match(dst + dpos, static_cast<int64_t>(offset), mlen);
            dpos += mlen;
        }
        if (bb.bitpos != 0) return kErrCorrupt;
    }
    const int64_t rest = regen - lpos;
    if (dpos + rest > dcap) return kErrDstFull;
    copy_bytes(dst + dpos, literals + lpos, rest);
    return dpos + rest;
}

// Frame_Content_Size of the frame at src, or -1 when the header does not carry it / is not a zstd frame.
BYDB_HD inline int64_t frame_content_size(const uint8_t *src, int64_t len) {
    if (len < 6) return -1;
    const uint32_t magic = src[0] | (src[1] << 8) | (src[2] << 16) | (static_cast<uint32_t>(src[3]) << 24);
    if (magic != 0xFD2FB528u) return -1;
    const int fhd = src[4];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    int64_t pos = 5 + (single ? 0 : 1) + (did == 0 ? 0 : (did == 1 ? 1 : (did == 2 ? 2 : 4)));
    const int n = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (n == 0 || len < pos + n) return -1;
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= static_cast<uint64_t>(src[pos + i]) << (8 * i);
    if (n == 2) v += 256;
    return v > (1ull << 40) ? -1 : static_cast<int64_t>(v);
}

// Decodes one frame. Returns the decoded size (>= 0) or an Err (< 0). lit: >= 128 KB scratch.
BYDB_HD inline int64_t decode_frame(Workspace *ws, const uint8_t *src, int64_t len, uint8_t *dst, int64_t dcap, uint8_t *lit) {
    if (len < 6) return kErrTrunc;
    const uint32_t magic = src[0] | (src[1] << 8) | (src[2] << 16) | (static_cast<uint32_t>(src[3]) << 24);
    if (magic != 0xFD2FB528u) return kErrUnsupported;
    const int fhd = src[4];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
    if (fhd & 0x08) return kErrCorrupt;
    int64_t pos = 5;
    if (!single) pos += 1;  // window descriptor (the whole output buffer is our window)
    if (did) return kErrUnsupported;
    const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (len < pos + fcs_bytes) return kErrTrunc;
    pos += fcs_bytes;
    ws->have_huf = ws->have_ll = ws->have_ml = ws->have_of = 0;
    ws->rep[0] = 1;
    ws->rep[1] = 4;
    ws->rep[2] = 8;
    int64_t dpos = 0;
    for (;;) {
        if (len < pos + 3) return kErrTrunc;
        const uint32_t bh = src[pos] | (src[pos + 1] << 8) | (static_cast<uint32_t>(src[pos + 2]) << 16);
        pos += 3;
        const int last = bh & 1, type = (bh >> 1) & 3;
        const int64_t bsize = bh >> 3;
        if (type == 0) {
            if (len < pos + bsize) return kErrTrunc;
            if (dpos + bsize > dcap) return kErrDstFull;
            copy_bytes(dst + dpos, src + pos, bsize);
            dpos += bsize;
            pos += bsize;
        } else if (type == 1) {
            if (len < pos + 1) return kErrTrunc;
            if (dpos + bsize > dcap) return kErrDstFull;
            for (int64_t k = 0; k < bsize; ++k) dst[dpos + k] = src[pos];
            dpos += bsize;
            pos += 1;
        } else if (type == 2) {
            if (len < pos + bsize) return kErrTrunc;
            const int64_t r = decode_block(ws, src + pos, bsize, dst, dpos, dcap, lit);
            if (r < 0) return r;
            dpos = r;
            pos += bsize;
        } else {
            return kErrCorrupt;
        }
        if (last) break;
    }
    (void)checksum;  // 4 trailing bytes, not verified (the reference writes no checksum: zstd.go:70-75)
    return dpos;
}

}  // namespace zstd
}  // namespace bydb
