// part_writer.cc -- host-side measure part writer + synthetic part generator (include/bydb_synth.h).
//
// An independent C++ implementation of the reference's flush path (see the citations in
// bydb_synth.h); tests/test_writer_vs_oracle.py checks it byte-for-byte against the oracle's C
// writer.  Multi-threaded: blocks are encoded in parallel into per-thread buffers and stitched.
#include <dlfcn.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bydb_gpu.h"
#include "../../include/bydb_synth.h"

namespace {

using Bytes = std::vector<uint8_t>;

// banyand/measure/measure.go:41-46
constexpr uint64_t kMaxUncompressedBlock = 2ull * 1024 * 1024;
constexpr uint64_t kMaxUncompressedPrimary = 128ull * 1024;
constexpr uint64_t kMaxBlockLength = 8ull * 1024;

// ------------------------------------------------------------------ zstd (compress side)
using compress_fn = size_t (*)(void *, size_t, const void *, size_t, int);
using bound_fn = size_t (*)(size_t);
using iserr_fn = unsigned (*)(size_t);
struct ZstdC {
    compress_fn compress = nullptr;
    bound_fn bound = nullptr;
    iserr_fn is_error = nullptr;
    bool ok = false;
};
ZstdC &zstdc() {
    static ZstdC z;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        z.compress = reinterpret_cast<compress_fn>(dlsym(h, "ZSTD_compress"));
        z.bound = reinterpret_cast<bound_fn>(dlsym(h, "ZSTD_compressBound"));
        z.is_error = reinterpret_cast<iserr_fn>(dlsym(h, "ZSTD_isError"));
        z.ok = z.compress && z.bound && z.is_error;
    });
    return z;
}
bool zstd_append(Bytes &dst, const uint8_t *src, size_t n) {  // pkg/compress/zstd/zstd.go:54-57, level 1
    ZstdC &z = zstdc();
    if (!z.ok) return false;
    size_t cap = z.bound(n);
    size_t at = dst.size();
    dst.resize(at + cap);
    size_t r = z.compress(dst.data() + at, cap, src, n, 1);
    if (z.is_error(r)) return false;
    dst.resize(at + r);
    return true;
}

// ------------------------------------------------------------------ primitive encoders
inline void put_u64be(Bytes &b, uint64_t u) {
    for (int k = 7; k >= 0; --k) b.push_back(static_cast<uint8_t>(u >> (8 * k)));
}
inline void put_varu(Bytes &b, uint64_t u) {  // pkg/encoding/int.go:152-185
    while (u > 0x7f) {
        b.push_back(static_cast<uint8_t>(0x80 | (u & 0x7f)));
        u >>= 7;
    }
    b.push_back(static_cast<uint8_t>(u));
}
inline void put_varint(Bytes &b, int64_t v) {  // pkg/encoding/int.go:75-99 (zig-zag LEB128)
    put_varu(b, (static_cast<uint64_t>(v) << 1) ^ static_cast<uint64_t>(v >> 63));
}
inline void put_str(Bytes &b, const std::string &s) {  // pkg/encoding/bytes.go:28-32
    put_varu(b, s.size());
    b.insert(b.end(), s.begin(), s.end());
}
inline void put_conv_i64(Bytes &b, int64_t i) {  // pkg/convert/number.go:33-45 (order preserving)
    uint64_t u = i >= 0 ? (static_cast<uint64_t>(i) | (1ull << 63)) : ((1ull << 63) - (0ull - static_cast<uint64_t>(i)));
    put_u64be(b, u);
}
inline int64_t wsub(int64_t a, int64_t b) { return static_cast<int64_t>(static_cast<uint64_t>(a) - static_cast<uint64_t>(b)); }
inline int64_t wadd(int64_t a, int64_t b) { return static_cast<int64_t>(static_cast<uint64_t>(a) + static_cast<uint64_t>(b)); }

// pkg/encoding/int_list.go:27-53 Int64ListToBytes (type selection :112-179, bodies delta.go:26-89)
int encode_int64_list(Bytes &dst, const int64_t *a, size_t n, int64_t &first) {
    first = a[0];
    bool is_const = true;
    for (size_t i = 1; i < n && is_const; ++i) is_const = a[i] == a[0];
    if (is_const) return 1;
    bool is_delta = n >= 2, delta_const = n >= 2;
    if (n >= 2) {
        const int64_t d1 = wsub(a[1], a[0]);
        const int64_t asc = (d1 >> 63) & 1;
        int64_t prev = a[1];
        for (size_t i = 2; i < n; ++i) {
            const int64_t d = wsub(a[i], prev);
            if ((((d >> 63) & 1) ^ asc) == 1) {
                is_delta = delta_const = false;
                break;
            }
            if (d != d1) delta_const = false;
            prev = a[i];
        }
    }
    if (is_delta && delta_const) {
        put_varint(dst, wsub(a[1], a[0]));
        return 2;
    }
    bool dod = is_delta;
    if (!dod && n >= 2) {  // isIncremental
        if (a[0] < 0) {
            dod = true;
        } else {
            size_t resets = 0;
            int64_t vprev = a[0];
            bool ok = true;
            for (size_t i = 1; i < n; ++i) {
                const int64_t v = a[i];
                if (v < vprev) {
                    if (v < 0 || v > (vprev >> 3)) {
                        ok = false;
                        break;
                    }
                    ++resets;
                }
                vprev = v;
            }
            dod = ok && (resets <= 2 || resets < (n >> 3));
        }
    }
    if (dod) {
        int64_t d1 = wsub(a[1], a[0]);
        put_varint(dst, d1);
        int64_t v = a[1];
        for (size_t i = 2; i < n; ++i) {
            const int64_t d2 = wsub(wsub(a[i], v), d1);
            d1 = wadd(d1, d2);
            v = wadd(v, d1);
            put_varint(dst, d2);
        }
        return 4;
    }
    int64_t v = a[0];
    for (size_t i = 1; i < n; ++i) {
        const int64_t d = wsub(a[i], v);
        v = wadd(v, d);
        put_varint(dst, d);
    }
    return 3;
}

// ------------------------------------------------------------------ float64 -> (mantissa, exp10)
// pkg/encoding/float.go:107-190.  The shortest round-trip digits come from std::to_chars, which like
// Go's strconv 'e',-1 emits the shortest decimal that parses back to the same double (closest one).
bool float_to_decimal(double f, int64_t &mant, int &exp) {
    if (std::isnan(f) || std::isinf(f)) return false;
    if (f == 0) {
        mant = 0;
        exp = 0;
        return true;
    }
    int64_t u = (f >= 9223372036854775808.0 || f < -9223372036854775808.0) ? INT64_MIN : static_cast<int64_t>(f);
    if (static_cast<double>(u) == f) {
        int e = 0;
        while (u != 0 && u % 10 == 0) {
            u /= 10;
            ++e;
        }
        mant = u;
        exp = e;
        return true;
    }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, f, std::chars_format::scientific);
    if (res.ec != std::errc()) return false;
    // [-]d[.ddd]e[+-]xx
    const char *p = buf;
    bool neg = false;
    if (*p == '-') {
        neg = true;
        ++p;
    }
    char digits[32];
    int nd = 0, frac = 0;
    bool after_dot = false;
    while (p < res.ptr && *p != 'e') {
        if (*p == '.') {
            after_dot = true;
        } else {
            digits[nd++] = *p;
            if (after_dot) ++frac;
        }
        ++p;
    }
    if (p >= res.ptr) return false;
    int sci = 0;
    {
        ++p;
        bool eneg = false;
        if (*p == '+') ++p;
        else if (*p == '-') {
            eneg = true;
            ++p;
        }
        while (p < res.ptr) sci = sci * 10 + (*p++ - '0');
        if (eneg) sci = -sci;
    }
    while (nd > 1 && digits[nd - 1] == '0') {  // float.go:166-169
        --nd;
        --frac;
    }
    uint64_t m = 0;
    for (int i = 0; i < nd; ++i) {
        const uint64_t dg = static_cast<uint64_t>(digits[i] - '0');
        if (m > (UINT64_MAX - dg) / 10) return false;
        m = m * 10 + dg;
    }
    if (m > static_cast<uint64_t>(INT64_MAX)) return false;
    if (sci > 32767 || sci < -32768) return false;
    exp = static_cast<int16_t>(static_cast<int16_t>(sci) - static_cast<int16_t>(frac));
    mant = neg ? -static_cast<int64_t>(m) : static_cast<int64_t>(m);
    return true;
}

// a decimal given exactly as k / 10^d  (k < 2^53, d <= 15): its float is fl(k/10^d) and the
// shortest digits of that float are k with the trailing zeros stripped (writer notes)
inline void decimal_to_mant_exp(int64_t k, int d, int64_t &mant, int &exp) {
    if (k == 0) {
        mant = 0;
        exp = 0;
        return;
    }
    int e = -d;
    while (k % 10 == 0) {
        k /= 10;
        ++e;
    }
    mant = k;
    exp = e;
}

const int64_t kPow10[19] = {1LL,
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
bool mul_pow10(int64_t v, int n, int64_t &out) {  // float.go:199-230
    if (n < 0) return false;
    while (n >= 19) {
        if (v > INT64_MAX / kPow10[18] || v < INT64_MIN / kPow10[18]) return false;
        v *= kPow10[18];
        n -= 18;
    }
    if (n > 0) {
        if (v > INT64_MAX / kPow10[n] || v < INT64_MIN / kPow10[n]) return false;
        v *= kPow10[n];
    }
    out = v;
    return true;
}

// ------------------------------------------------------------------ bytes block / dictionary / bit packing
void compress_block(Bytes &dst, const uint8_t *src, size_t n) {  // bytes.go:291-304
    if (n < 128) {
        dst.push_back(0);
        dst.push_back(static_cast<uint8_t>(n));
        dst.insert(dst.end(), src, src + n);
        return;
    }
    dst.push_back(1);
    Bytes z;
    zstd_append(z, src, n);
    put_varu(dst, z.size());
    dst.insert(dst.end(), z.begin(), z.end());
}
struct Cell {  // nil when len < 0
    const uint8_t *p;
    int64_t len;
};
void encode_bytes_block(Bytes &dst, const Cell *cells, size_t n) {  // bytes.go:45-72, 209-240
    uint64_t nmax = 0;
    for (size_t i = 0; i < n; ++i) nmax = std::max<uint64_t>(nmax, cells[i].len < 0 ? 0 : static_cast<uint64_t>(cells[i].len) + 1);
    Bytes lens;
    const int w = nmax < (1ull << 8) ? 1 : nmax < (1ull << 16) ? 2 : nmax < (1ull << 32) ? 4 : 8;
    lens.push_back(static_cast<uint8_t>(w == 1 ? 0 : w == 2 ? 1 : w == 4 ? 2 : 3));
    for (size_t i = 0; i < n; ++i) {
        const uint64_t v = cells[i].len < 0 ? 0 : static_cast<uint64_t>(cells[i].len) + 1;
        for (int k = w - 1; k >= 0; --k) lens.push_back(static_cast<uint8_t>(v >> (8 * k)));
    }
    compress_block(dst, lens.data(), lens.size());
    Bytes data;
    for (size_t i = 0; i < n; ++i)
        if (cells[i].len > 0) data.insert(data.end(), cells[i].p, cells[i].p + cells[i].len);
    compress_block(dst, data.data(), data.size());
}
struct BitW {  // pkg/encoding/writer.go:25-96
    Bytes &out;
    uint8_t cache = 0, avail = 8;
    explicit BitW(Bytes &o) : out(o) {}
    void bit(bool b) {
        if (b) cache |= static_cast<uint8_t>(1u << (avail - 1));
        if (--avail == 0) {
            out.push_back(cache);
            cache = 0;
            avail = 8;
        }
    }
    void byte(uint8_t b) {
        out.push_back(static_cast<uint8_t>(cache | (b >> (8 - avail))));
        cache = avail == 8 ? 0 : static_cast<uint8_t>(b << avail);
    }
    void bits(uint64_t u, int n) {
        u <<= (64 - n);
        for (; n >= 8; n -= 8) {
            byte(static_cast<uint8_t>(u >> 56));
            u <<= 8;
        }
        uint8_t rem = static_cast<uint8_t>(u >> 56);
        for (; n > 0; --n) {
            bit((rem & 0x80) != 0);
            rem = static_cast<uint8_t>(rem << 1);
        }
    }
    void flush() {
        if (avail != 8) out.push_back(cache);
        cache = 0;
        avail = 8;
    }
};
// dictionary.go:69-77 Encode with pre-resolved indices (values in first-appearance order)
void encode_dictionary(Bytes &dst, const std::vector<Cell> &values, const std::vector<uint32_t> &idx) {
    put_varu(dst, values.size());
    encode_bytes_block(dst, values.data(), values.size());
    std::vector<uint32_t> rle;
    if (!idx.empty()) {
        uint32_t cur = idx[0], cnt = 1;
        for (size_t i = 1; i < idx.size(); ++i) {
            if (idx[i] == cur) {
                ++cnt;
            } else {
                rle.push_back(cur);
                rle.push_back(cnt);
                cur = idx[i];
                cnt = 1;
            }
        }
        rle.push_back(cur);
        rle.push_back(cnt);
    }
    BitW w(dst);
    if (rle.empty()) {
        w.bits(0, 32);
        w.flush();
        return;
    }
    w.bits(rle.size(), 32);
    uint32_t maxv = 0;
    for (uint32_t v : rle) maxv = std::max(maxv, v);
    const int width = maxv > 0 ? 32 - __builtin_clz(maxv) : 1;
    w.bits(static_cast<uint64_t>(width), 8);
    for (uint32_t v : rle) w.bits(v, width);
    w.flush();
}
// column.go:222-234 encodeDefault over generic cells (dictionary when <= 256 distinct, nil != empty)
void encode_default(Bytes &dst, const std::vector<Cell> &cells) {
    std::vector<Cell> values;
    std::vector<uint32_t> idx(cells.size());
    bool dict_ok = true;
    for (size_t i = 0; i < cells.size() && dict_ok; ++i) {
        size_t k = 0;
        for (; k < values.size(); ++k) {
            const Cell &a = values[k], &b = cells[i];
            if (a.len < 0 && b.len < 0) break;
            if (a.len < 0 || b.len < 0) continue;
            if (a.len == b.len && (a.len == 0 || memcmp(a.p, b.p, static_cast<size_t>(a.len)) == 0)) break;
        }
        if (k == values.size()) {
            if (values.size() == 256) {
                dict_ok = false;
                break;
            }
            values.push_back(cells[i]);
        }
        idx[i] = static_cast<uint32_t>(k);
    }
    if (dict_ok) {
        dst.push_back(10);
        encode_dictionary(dst, values, idx);
    } else {
        dst.push_back(9);
        encode_bytes_block(dst, cells.data(), cells.size());
    }
}

// ------------------------------------------------------------------ column pages (column.go:113-220)
void encode_int64_page(Bytes &dst, const int64_t *v, size_t n) {
    Bytes body;
    int64_t first = 0;
    const int enc = encode_int64_list(body, v, n, first);
    dst.push_back(static_cast<uint8_t>(enc));
    put_conv_i64(dst, first);
    dst.insert(dst.end(), body.begin(), body.end());
}
// mant/exp per value -> page; falls back to the Plain page built from the float cells on overflow
void encode_float64_page(Bytes &dst, std::vector<int64_t> &mant, std::vector<int> &exps, bool ok, const double *values, const int64_t *dec_k, int dec_d,
                         size_t n) {
    if (ok) {
        int min_exp = INT32_MAX;
        for (size_t i = 0; i < n; ++i) min_exp = std::min(min_exp, exps[i]);
        for (size_t i = 0; i < n && ok; ++i) {
            const int diff = static_cast<int16_t>(exps[i] - min_exp);
            if (diff == 0) continue;
            ok = mul_pow10(mant[i], diff, mant[i]);
        }
        if (ok) {
            Bytes body;
            int64_t first = 0;
            const int enc = encode_int64_list(body, mant.data(), n, first);
            dst.push_back(static_cast<uint8_t>(enc));
            const uint16_t e16 = static_cast<uint16_t>(static_cast<int16_t>(min_exp));
            dst.push_back(static_cast<uint8_t>(e16 >> 8));
            dst.push_back(static_cast<uint8_t>(e16 & 0xff));
            put_conv_i64(dst, first);
            dst.insert(dst.end(), body.begin(), body.end());
            return;
        }
    }
    // column.go:203-208: EncodeTypePlain marker + default page over the 8-byte IEEE cells
    Bytes raw(n * 8);
    std::vector<Cell> cells(n);
    static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    for (size_t i = 0; i < n; ++i) {
        const double fv = values ? values[i] : static_cast<double>(dec_k[i]) / p10[dec_d];
        uint64_t u;
        memcpy(&u, &fv, 8);
        for (int k = 0; k < 8; ++k) raw[i * 8 + k] = static_cast<uint8_t>(u >> (56 - 8 * k));
        cells[i] = Cell{raw.data() + i * 8, 8};
    }
    dst.push_back(9);
    encode_default(dst, cells);
}

// ------------------------------------------------------------------ block metadata
struct ColMeta {
    uint64_t off, size;
};
struct BlockMeta {
    uint64_t sid, uncompressed, count;
    uint64_t ts_off, ts_size, ver_off;
    int64_t ts_min, ts_max, ver_first;
    uint8_t ts_enc, ver_enc;
    uint64_t tfm_off, tfm_size;
    std::vector<ColMeta> fields;
};

struct Schema {
    std::vector<std::string> field_names;
    std::vector<int> field_types;
    std::string family;  // empty = none
    std::vector<std::string> tag_names;
    std::vector<int> tag_types;
};

// per-thread output segment
struct Segment {
    Bytes ts, fv, tf, tfm;
    std::vector<BlockMeta> blocks;
};

// view of one block's rows
struct BlockRows {
    uint64_t sid;
    size_t n;
    const int64_t *ts, *ver;
    // fields: one of i64 / f64 / dec per column
    std::vector<const int64_t *> f_i64;
    std::vector<const double *> f_f64;
    std::vector<const int64_t *> f_dec;
    std::vector<int> f_dec_digits;
    // tags
    std::vector<const int64_t *> t_i64;
    std::vector<const uint32_t *> t_idx;
    std::vector<const std::vector<std::string> *> t_values;
};

void encode_block(const Schema &sc, const BlockRows &br, Segment &seg, std::vector<int64_t> &tmp_m, std::vector<int> &tmp_e, std::vector<double> &tmp_f) {
    BlockMeta bm{};
    const size_t n = br.n;
    bm.sid = br.sid;
    bm.count = n;
    // block.go:267-297 uncompressedSizeBytes
    uint64_t unc = static_cast<uint64_t>(n) * 16;
    if (!sc.family.empty()) {
        unc += sc.family.size();
        for (size_t t = 0; t < sc.tag_names.size(); ++t) {
            unc += sc.tag_names[t].size();
            if (sc.tag_types[t] == BYDB_VT_INT64) {
                unc += 8ull * n;
            } else {
                for (size_t i = 0; i < n; ++i) unc += (*br.t_values[t])[br.t_idx[t][i]].size();
            }
        }
    }
    for (size_t f = 0; f < sc.field_names.size(); ++f) unc += (sc.field_names[f].size() + 8) * static_cast<uint64_t>(n);
    bm.uncompressed = unc;
    // block.go:361-379 mustWriteTimestampsTo
    {
        Bytes body;
        int64_t first = 0;
        const int enc = encode_int64_list(body, br.ts, n, first);
        bm.ts_enc = static_cast<uint8_t>(enc + 4);
        bm.ts_min = first;
        bm.ts_max = br.ts[n - 1];
        bm.ts_off = seg.ts.size();
        bm.ver_off = body.size();
        seg.ts.insert(seg.ts.end(), body.begin(), body.end());
        body.clear();
        bm.ver_enc = static_cast<uint8_t>(encode_int64_list(body, br.ver, n, first));
        bm.ver_first = first;
        bm.ts_size = bm.ver_off + body.size();
        seg.ts.insert(seg.ts.end(), body.begin(), body.end());
    }
    // tag family: block.go:184-205 marshalTagFamily
    if (!sc.family.empty()) {
        Bytes cfm;
        put_varu(cfm, sc.tag_names.size());
        for (size_t t = 0; t < sc.tag_names.size(); ++t) {
            const uint64_t off = seg.tf.size();
            if (sc.tag_types[t] == BYDB_VT_INT64) {
                encode_int64_page(seg.tf, br.t_i64[t], n);
            } else {
                // dictionary over the values in first-appearance order (dictionary.go:36-50)
                const auto &vals = *br.t_values[t];
                std::vector<int> remap(vals.size(), -1);
                std::vector<Cell> dict;
                std::vector<uint32_t> idx(n);
                bool dict_ok = true;
                for (size_t i = 0; i < n; ++i) {
                    const uint32_t v = br.t_idx[t][i];
                    if (remap[v] < 0) {
                        if (dict.size() == 256) {
                            dict_ok = false;
                            break;
                        }
                        remap[v] = static_cast<int>(dict.size());
                        dict.push_back(Cell{reinterpret_cast<const uint8_t *>(vals[v].data()), static_cast<int64_t>(vals[v].size())});
                    }
                    idx[i] = static_cast<uint32_t>(remap[v]);
                }
                if (dict_ok) {
                    seg.tf.push_back(10);
                    encode_dictionary(seg.tf, dict, idx);
                } else {
                    std::vector<Cell> cells(n);
                    for (size_t i = 0; i < n; ++i) {
                        const auto &s = vals[br.t_idx[t][i]];
                        cells[i] = Cell{reinterpret_cast<const uint8_t *>(s.data()), static_cast<int64_t>(s.size())};
                    }
                    seg.tf.push_back(9);
                    encode_bytes_block(seg.tf, cells.data(), n);
                }
            }
            put_str(cfm, sc.tag_names[t]);
            cfm.push_back(static_cast<uint8_t>(sc.tag_types[t]));
            put_varu(cfm, off);
            put_varu(cfm, seg.tf.size() - off);
        }
        bm.tfm_off = seg.tfm.size();
        bm.tfm_size = cfm.size();
        seg.tfm.insert(seg.tfm.end(), cfm.begin(), cfm.end());
    }
    // fields
    bm.fields.resize(sc.field_names.size());
    for (size_t f = 0; f < sc.field_names.size(); ++f) {
        const uint64_t off = seg.fv.size();
        if (sc.field_types[f] == BYDB_VT_INT64) {
            encode_int64_page(seg.fv, br.f_i64[f], n);
        } else {
            tmp_m.resize(n);
            tmp_e.resize(n);
            bool ok = true;
            const double *fvals = br.f_f64[f];
            if (br.f_dec[f]) {
                const int d = br.f_dec_digits[f];
                for (size_t i = 0; i < n; ++i) decimal_to_mant_exp(br.f_dec[f][i], d, tmp_m[i], tmp_e[i]);
            } else {
                for (size_t i = 0; i < n && ok; ++i) ok = float_to_decimal(fvals[i], tmp_m[i], tmp_e[i]);
            }
            encode_float64_page(seg.fv, tmp_m, tmp_e, ok, fvals, br.f_dec[f], br.f_dec_digits[f], n);
            (void)tmp_f;
        }
        bm.fields[f] = ColMeta{off, seg.fv.size() - off};
    }
    seg.blocks.push_back(std::move(bm));
}

}  // namespace

struct bydb_part_image {
    std::vector<std::pair<std::string, Bytes>> files;
    uint64_t total_rows = 0, n_blocks = 0;
};

namespace {

// block_metadata.go:113-131 marshal (+ :268-277) with the segment-relative offsets rebased
void marshal_block(Bytes &dst, const Schema &sc, const BlockMeta &bm, uint64_t ts_base, uint64_t fv_base, uint64_t tf_base, uint64_t tfm_base) {
    (void)tf_base;
    put_u64be(dst, bm.sid);
    put_varu(dst, bm.uncompressed);
    put_varu(dst, bm.count);
    put_varu(dst, bm.ts_off + ts_base);
    put_varu(dst, bm.ts_size);
    put_u64be(dst, static_cast<uint64_t>(bm.ts_min));
    put_u64be(dst, static_cast<uint64_t>(bm.ts_max));
    dst.push_back(bm.ts_enc);
    put_varu(dst, bm.ver_off);
    put_u64be(dst, static_cast<uint64_t>(bm.ver_first));
    dst.push_back(bm.ver_enc);
    if (sc.family.empty()) {
        put_varu(dst, 0);
    } else {
        put_varu(dst, 1);
        put_str(dst, sc.family);
        put_varu(dst, bm.tfm_off + tfm_base);
        put_varu(dst, bm.tfm_size);
    }
    put_varu(dst, sc.field_names.size());
    for (size_t f = 0; f < sc.field_names.size(); ++f) {
        put_str(dst, sc.field_names[f]);
        dst.push_back(static_cast<uint8_t>(sc.field_types[f]));
        put_varu(dst, bm.fields[f].off + fv_base);
        put_varu(dst, bm.fields[f].size);
    }
}

// Tag pages reference absolute offsets inside <family>.tf through the .tfm records, which were
// written segment-relative: rebuild each block's columnFamilyMetadata with rebased offsets.
void rebase_tfm(const Schema &sc, Segment &seg, uint64_t tf_base) {
    if (sc.family.empty() || tf_base == 0) return;
    Bytes out;
    for (auto &bm : seg.blocks) {
        const uint8_t *p = seg.tfm.data() + bm.tfm_off;
        const uint8_t *end = p + bm.tfm_size;
        auto rd = [&](uint64_t &v) {
            v = 0;
            for (unsigned s = 0; p < end; s += 7) {
                uint8_t b = *p++;
                v |= static_cast<uint64_t>(b & 0x7f) << s;
                if (b < 0x80) break;
            }
        };
        Bytes rec;
        uint64_t n;
        rd(n);
        put_varu(rec, n);
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t nl, off, size;
            rd(nl);
            put_varu(rec, nl);
            rec.insert(rec.end(), p, p + nl);
            p += nl;
            rec.push_back(*p++);
            rd(off);
            rd(size);
            put_varu(rec, off + tf_base);
            put_varu(rec, size);
        }
        bm.tfm_off = out.size();
        bm.tfm_size = rec.size();
        out.insert(out.end(), rec.begin(), rec.end());
    }
    seg.tfm.swap(out);
}

// stitches the segments (in order) into the part files; block_writer.go:206-285 bookkeeping
bydb_part_image *assemble(const Schema &sc, std::vector<Segment> &segs) {
    auto img = new bydb_part_image();
    Bytes ts, fv, tf, tfm, primary, meta_raw, meta;
    Bytes pblock;
    uint64_t sid_first = 0;
    int64_t pmin = 0, pmax = 0;
    bool has = false;
    auto flush_primary = [&]() {
        if (!pblock.empty()) {
            const uint64_t off = primary.size();
            zstd_append(primary, pblock.data(), pblock.size());
            put_u64be(meta_raw, sid_first);
            put_u64be(meta_raw, static_cast<uint64_t>(pmin));
            put_u64be(meta_raw, static_cast<uint64_t>(pmax));
            put_u64be(meta_raw, off);
            put_u64be(meta_raw, primary.size() - off);
        }
        pblock.clear();
        has = false;
        pmin = pmax = 0;
        sid_first = 0;
    };
    for (auto &seg : segs) {
        rebase_tfm(sc, seg, tf.size());
        const uint64_t ts_base = ts.size(), fv_base = fv.size(), tf_base = tf.size(), tfm_base = tfm.size();
        for (const auto &bm : seg.blocks) {
            if (!has) {
                sid_first = bm.sid;
                has = true;
                pmin = bm.ts_min;
                pmax = bm.ts_max;
            } else {
                pmin = std::min(pmin, bm.ts_min);
                pmax = std::max(pmax, bm.ts_max);
            }
            marshal_block(pblock, sc, bm, ts_base, fv_base, tf_base, tfm_base);
            img->total_rows += bm.count;
            img->n_blocks++;
            if (pblock.size() > kMaxUncompressedPrimary) flush_primary();
        }
        ts.insert(ts.end(), seg.ts.begin(), seg.ts.end());
        fv.insert(fv.end(), seg.fv.begin(), seg.fv.end());
        tf.insert(tf.end(), seg.tf.begin(), seg.tf.end());
        tfm.insert(tfm.end(), seg.tfm.begin(), seg.tfm.end());
        Segment().ts.swap(seg.ts);
        Segment().fv.swap(seg.fv);
        Segment().tf.swap(seg.tf);
        Segment().tfm.swap(seg.tfm);
    }
    flush_primary();
    zstd_append(meta, meta_raw.data(), meta_raw.size());
    img->files.emplace_back("meta.bin", std::move(meta));
    img->files.emplace_back("primary.bin", std::move(primary));
    img->files.emplace_back("timestamps.bin", std::move(ts));
    img->files.emplace_back("fv.bin", std::move(fv));
    if (!sc.family.empty()) {
        img->files.emplace_back(sc.family + ".tf", std::move(tf));
        img->files.emplace_back(sc.family + ".tfm", std::move(tfm));
    }
    return img;
}

unsigned pick_threads(uint32_t want) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    return want ? want : hw;
}

// per-row uncompressed size, part.go:234-249
uint64_t row_uncompressed(const Schema &sc, const BlockRows &all, size_t r) {
    uint64_t n = 16;
    for (size_t f = 0; f < sc.field_names.size(); ++f) n += sc.field_names[f].size() + 8;
    if (!sc.family.empty()) {
        n += sc.family.size();
        for (size_t t = 0; t < sc.tag_names.size(); ++t) {
            n += sc.tag_names[t].size();
            n += sc.tag_types[t] == BYDB_VT_INT64 ? 8 : (*all.t_values[t])[all.t_idx[t][r]].size();
        }
    }
    return n;
}

BlockRows slice(const BlockRows &all, size_t lo, size_t hi, uint64_t sid) {
    BlockRows b = all;
    b.sid = sid;
    b.n = hi - lo;
    b.ts += lo;
    b.ver += lo;
    for (auto &p : b.f_i64)
        if (p) p += lo;
    for (auto &p : b.f_f64)
        if (p) p += lo;
    for (auto &p : b.f_dec)
        if (p) p += lo;
    for (auto &p : b.t_i64)
        if (p) p += lo;
    for (auto &p : b.t_idx)
        if (p) p += lo;
    return b;
}

// cuts one series' rows [lo,hi) into blocks like part.go:192-199 and encodes them
void write_series(const Schema &sc, const BlockRows &all, size_t lo, size_t hi, uint64_t sid, Segment &seg, std::vector<int64_t> &tm, std::vector<int> &te,
                  std::vector<double> &tf) {
    size_t index_prev = lo;
    uint64_t unc = 0;
    for (size_t i = lo; i < hi; ++i) {
        if (unc >= kMaxUncompressedBlock || (i - index_prev) > kMaxBlockLength) {
            encode_block(sc, slice(all, index_prev, i, sid), seg, tm, te, tf);
            index_prev = i;
            unc = 0;
        }
        unc += row_uncompressed(sc, all, i);
    }
    if (hi > index_prev) encode_block(sc, slice(all, index_prev, hi, sid), seg, tm, te, tf);
}

}  // namespace

extern "C" {

int bydb_part_write(const bydb_write_input *in, bydb_part_image **out) {
    if (!in || !out || in->n_rows == 0 || !in->series_ids || !in->timestamps || !in->versions) return BYDB_EINVAL;
    if (!zstdc().ok) return BYDB_EIO;
    Schema sc;
    BlockRows all{};
    all.ts = in->timestamps;
    all.ver = in->versions;
    std::vector<std::vector<std::string>> tag_values(in->n_tags);
    for (uint32_t f = 0; f < in->n_fields; ++f) {
        const bydb_wcolumn &c = in->fields[f];
        if (c.value_type != BYDB_VT_INT64 && c.value_type != BYDB_VT_FLOAT64) return BYDB_EINVAL;
        sc.field_names.push_back(c.name);
        sc.field_types.push_back(c.value_type);
        all.f_i64.push_back(c.value_type == BYDB_VT_INT64 ? c.i64 : nullptr);
        all.f_f64.push_back(c.value_type == BYDB_VT_FLOAT64 ? c.f64 : nullptr);
        all.f_dec.push_back(c.value_type == BYDB_VT_FLOAT64 && c.dec_digits >= 0 ? c.dec_k : nullptr);
        all.f_dec_digits.push_back(c.dec_digits);
        if (c.value_type == BYDB_VT_FLOAT64 && c.dec_digits > 15) return BYDB_EINVAL;
    }
    if (in->tag_family && in->n_tags > 0) {
        sc.family = in->tag_family;
        for (uint32_t t = 0; t < in->n_tags; ++t) {
            const bydb_wcolumn &c = in->tags[t];
            if (c.value_type != BYDB_VT_INT64 && c.value_type != BYDB_VT_STR) return BYDB_EINVAL;
            sc.tag_names.push_back(c.name);
            sc.tag_types.push_back(c.value_type);
            all.t_i64.push_back(c.value_type == BYDB_VT_INT64 ? c.i64 : nullptr);
            all.t_idx.push_back(c.value_type == BYDB_VT_STR ? c.str_idx : nullptr);
            for (uint32_t k = 0; k < c.n_str_values; ++k) tag_values[t].push_back(c.str_values[k]);
            all.t_values.push_back(&tag_values[t]);
        }
    }
    // series boundaries + validation (sorted, unique, no zero timestamp at a series start: part.go:176-190)
    std::vector<size_t> starts;
    for (uint64_t i = 0; i < in->n_rows; ++i) {
        if (i == 0 || in->series_ids[i] != in->series_ids[i - 1]) {
            if (i > 0 && in->series_ids[i] < in->series_ids[i - 1]) return BYDB_EINVAL;
            starts.push_back(i);
        } else if (in->timestamps[i] <= in->timestamps[i - 1]) {
            return BYDB_EINVAL;
        }
    }
    if (in->timestamps[0] == 0 || in->series_ids[0] == 0) return BYDB_EINVAL;
    starts.push_back(in->n_rows);
    const size_t ns = starts.size() - 1;
    const unsigned nt = std::min<unsigned>(pick_threads(in->threads), static_cast<unsigned>(ns));
    std::vector<Segment> segs(nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            std::vector<int64_t> tm;
            std::vector<int> te;
            std::vector<double> tf;
            const size_t a = ns * t / nt, b = ns * (t + 1) / nt;
            for (size_t s = a; s < b; ++s) write_series(sc, all, starts[s], starts[s + 1], in->series_ids[starts[s]], segs[t], tm, te, tf);
        });
    }
    for (auto &x : th) x.join();
    *out = assemble(sc, segs);
    return 0;
}

// ------------------------------------------------------------------ synthetic generator
namespace {
struct Rng {  // splitmix64
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return static_cast<double>(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return next() % n; }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};
}  // namespace

int bydb_synth_part(const bydb_synth_spec *spec, bydb_part_image **out) {
    if (!spec || !out || spec->n_series == 0 || spec->n_points == 0 || spec->sid0 == 0 || spec->t0 == 0 || spec->t_step <= 0) return BYDB_EINVAL;
    if (!zstdc().ok) return BYDB_EIO;
    Schema sc;
    for (uint32_t f = 0; f < spec->n_fields; ++f) {
        sc.field_names.push_back(spec->fields[f].name);
        sc.field_types.push_back(spec->fields[f].kind >= BYDB_SYN_I_DELTA ? BYDB_VT_INT64 : BYDB_VT_FLOAT64);
    }
    std::vector<std::string> region_values, zone_values;
    const bool want_code = (spec->code_tag & 1u) != 0, want_zone = (spec->code_tag & 2u) != 0;
    if (spec->region_values > 0 || spec->code_tag) {
        sc.family = "default";
        if (spec->region_values > 0) {
            sc.tag_names.push_back("region");
            sc.tag_types.push_back(BYDB_VT_STR);
            for (uint32_t k = 0; k < spec->region_values; ++k) region_values.push_back("r" + std::to_string(k));
        }
        if (want_zone) {  // a second dictionary tag (BASELINE config 5: two string predicates + one int64 predicate)
            sc.tag_names.push_back("zone");
            sc.tag_types.push_back(BYDB_VT_STR);
            for (uint32_t k = 0; k < 5; ++k) zone_values.push_back("z" + std::to_string(k));
        }
        if (want_code) {
            sc.tag_names.push_back("code");
            sc.tag_types.push_back(BYDB_VT_INT64);
        }
    }
    const size_t ns = spec->n_series, np = spec->n_points;
    const unsigned nt = std::min<unsigned>(pick_threads(spec->threads), static_cast<unsigned>(ns));
    std::vector<Segment> segs(nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            std::vector<int64_t> tm;
            std::vector<int> te;
            std::vector<double> tf;
            std::vector<int64_t> ts(np), ver(np, 1);
            for (size_t j = 0; j < np; ++j) ts[j] = spec->t0 + static_cast<int64_t>(j) * spec->t_step;
            std::vector<std::vector<int64_t>> ibuf(spec->n_fields, std::vector<int64_t>(np));
            std::vector<std::vector<double>> dbuf(spec->n_fields);
            std::vector<uint32_t> region(np), zone(np);
            std::vector<int64_t> code(np);
            const size_t a = ns * t / nt, b = ns * (t + 1) / nt;
            for (size_t s = a; s < b; ++s) {
                const uint64_t sid = spec->sid0 + s * spec->sid_step;
                Rng rng{spec->seed ^ (sid * 0xD6E8FEB86659FD93ull)};
                BlockRows all{};
                all.ts = ts.data();
                all.ver = ver.data();
                for (uint32_t f = 0; f < spec->n_fields; ++f) {
                    const int kind = spec->fields[f].kind;
                    int64_t *iv = ibuf[f].data();
                    const int64_t *pi = nullptr, *pdec = nullptr;
                    const double *pf = nullptr;
                    int dd = -1;
                    switch (kind) {
                        case BYDB_SYN_F_LATENCY:
                            for (size_t j = 0; j < np; ++j) iv[j] = std::llround((25.0 + 5.0 * rng.normal()) * 100.0);
                            pdec = iv;
                            dd = 2;
                            break;
                        case BYDB_SYN_F_WALK3: {
                            double v = 25.0;
                            for (size_t j = 0; j < np; ++j) {
                                v += (rng.uniform() - 0.5) * 0.2;
                                iv[j] = std::llround(v * 1000.0);
                            }
                            pdec = iv;
                            dd = 3;
                            break;
                        }
                        case BYDB_SYN_F_INT1000:
                            for (size_t j = 0; j < np; ++j) iv[j] = static_cast<int64_t>(rng.below(1000));
                            pdec = iv;
                            dd = 0;
                            break;
                        case BYDB_SYN_F_UNIFORM:
                            dbuf[f].resize(np);
                            for (size_t j = 0; j < np; ++j) dbuf[f][j] = rng.uniform() * 100.0;
                            pf = dbuf[f].data();
                            break;
                        case BYDB_SYN_I_DELTA: {
                            int64_t v = 0;
                            for (size_t j = 0; j < np; ++j) {
                                iv[j] = v;
                                v += static_cast<int64_t>(rng.below(10)) + 1;
                            }
                            pi = iv;
                            break;
                        }
                        case BYDB_SYN_I_FLUCT: {
                            int64_t v = 25;
                            for (size_t j = 0; j < np; ++j) {
                                v += static_cast<int64_t>(rng.below(11)) - 5;
                                iv[j] = v;
                            }
                            pi = iv;
                            break;
                        }
                        case BYDB_SYN_I_RANDOM100:
                            for (size_t j = 0; j < np; ++j) iv[j] = static_cast<int64_t>(rng.below(100));
                            pi = iv;
                            break;
                        case BYDB_SYN_I_COUNTER: {
                            int64_t v = 0;
                            const size_t r1 = np / 3, r2 = 2 * np / 3;
                            for (size_t j = 0; j < np; ++j) {
                                if (j == r1 || j == r2) v = 0;  // a counter reset
                                iv[j] = v;
                                v += static_cast<int64_t>(rng.below(10)) + 1;
                            }
                            pi = iv;
                            break;
                        }
                        default:
                            for (size_t j = 0; j < np; ++j) iv[j] = 0;
                            pi = iv;
                    }
                    all.f_i64.push_back(pi);
                    all.f_f64.push_back(pf);
                    all.f_dec.push_back(pdec);
                    all.f_dec_digits.push_back(dd);
                }
                if (spec->region_values > 0) {
                    if (spec->region_run == 0) {
                        const uint32_t v = static_cast<uint32_t>(sid % spec->region_values);
                        std::fill(region.begin(), region.end(), v);
                    } else {
                        size_t j = 0;
                        while (j < np) {
                            const uint32_t v = static_cast<uint32_t>(rng.below(spec->region_values));
                            size_t run = 1;
                            if (spec->region_run > 1) run = 1 + rng.below(2ull * spec->region_run - 1);
                            for (size_t e = std::min(np, j + run); j < e; ++j) region[j] = v;
                        }
                    }
                    all.t_i64.push_back(nullptr);
                    all.t_idx.push_back(region.data());
                    all.t_values.push_back(&region_values);
                }
                if (want_zone) {
                    size_t j = 0;
                    while (j < np) {
                        const uint32_t v = static_cast<uint32_t>(rng.below(5));
                        const size_t run = 1 + rng.below(127);
                        for (size_t e = std::min(np, j + run); j < e; ++j) zone[j] = v;
                    }
                    all.t_i64.push_back(nullptr);
                    all.t_idx.push_back(zone.data());
                    all.t_values.push_back(&zone_values);
                }
                if (want_code) {
                    for (size_t j = 0; j < np; ++j) code[j] = static_cast<int64_t>(rng.below(6)) * 100;
                    all.t_i64.push_back(code.data());
                    all.t_idx.push_back(nullptr);
                    all.t_values.push_back(nullptr);
                }
                write_series(sc, all, 0, np, sid, segs[t], tm, te, tf);
            }
        });
    }
    for (auto &x : th) x.join();
    *out = assemble(sc, segs);
    return 0;
}

uint32_t bydb_part_image_n_files(const bydb_part_image *p) { return static_cast<uint32_t>(p->files.size()); }
const char *bydb_part_image_file_name(const bydb_part_image *p, uint32_t i) { return p->files[i].first.c_str(); }
const uint8_t *bydb_part_image_file_data(const bydb_part_image *p, uint32_t i, uint64_t *len) {
    if (len) *len = p->files[i].second.size();
    return p->files[i].second.data();
}
void bydb_part_image_counts(const bydb_part_image *p, uint64_t *total_rows, uint64_t *n_blocks) {
    if (total_rows) *total_rows = p->total_rows;
    if (n_blocks) *n_blocks = p->n_blocks;
}
void bydb_part_image_free(bydb_part_image *p) { delete p; }

}  // extern "C"
