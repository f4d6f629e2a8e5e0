// unpack_kernels.cu -- normalisation of the reference's fallback pages when a part enters the HBM cache.
//
// The reference falls back to byte-string pages for everything its int64-list codecs cannot carry:
//   * numeric columns with a null cell or with floats that are not short decimals:
//       [EncodeTypePlain][encodeDefault page]            banyand/measure/column.go:147-153,192-195,203-208
//   * encodeDefault = dictionary (<= 256 distinct values) or a plain bytes block   column.go:222-234
//   * every bytes block is two compressBlocks (lengths, data); a block of >= 128 B is a zstd frame
//                                                          pkg/encoding/bytes.go:45-72,291-350
// None of that is friendly to a 32-lane decoder, and zstd is inherently sequential per frame.  The pages are
// immutable, so the work is done ONCE per part (bydb_part_register, or lazily on the cold host path) instead of
// once per query: classify_pages finds them, unpack_pages rewrites each into a side arena in HBM and repoints the
// page's DevCol at it.  The scan kernels then see only
//   * numeric raw-cell pages   [kEncRawCells][has_nulls][6 pad][n x u64 LE value][n x u8 valid]
//   * dictionary / plain string pages whose compressBlocks are type 0 (short) or kBlockRawLong
//     [2][u32 LE len][bytes] -- the reference's layout with the zstd frames inflated.
// One warp per page: lane 0 runs the zstd decoder (zstd_dec.cuh), all lanes expand / copy.
#include <cstdint>

#include "../../include/bydb_gpu.h"
#include "scan_kernels.cuh"
#include "zstd_dec.cuh"

namespace bydb {

namespace {

constexpr uint32_t kTmpBytes = 131072;  // one zstd block's worth: numeric pages hold <= 8193 cells of 8 B
constexpr uint32_t kWsBytes = (sizeof(zstd::Workspace) + 255u) & ~255u;
constexpr uint32_t kLitBytes = 131072 + 256;
constexpr uint32_t kOffLit = kWsBytes;
constexpr uint32_t kOffTmpA = kOffLit + kLitBytes;
constexpr uint32_t kOffTmpB = kOffTmpA + kTmpBytes;
constexpr uint32_t kOffTable = kOffTmpB + kTmpBytes;  // 256 x u64 values + 256 x u8 valid (numeric dictionaries)
constexpr uint32_t kScratchStride = kOffTable + 256 * 8 + 256;

__device__ __forceinline__ bool is_numeric(uint8_t vt) { return vt == BYDB_VT_INT64 || vt == BYDB_VT_FLOAT64; }

__device__ inline bool read_varuint(const uint8_t *&p, const uint8_t *end, uint64_t &out) {
    uint64_t u = 0;
    for (uint32_t i = 0; i < 10 && p < end; ++i) {
        const uint8_t c = *p++;
        u |= static_cast<uint64_t>(c & 0x7f) << (7 * i);
        if (c < 0x80) {
            out = u;
            return true;
        }
    }
    return false;
}

// One compressBlock header (bytes.go:291-350 + our raw-long form): payload pointer, stored length, decoded length
struct CBlock {
    const uint8_t *payload;
    uint32_t stored;   // bytes of payload in the page
    uint32_t type;     // 0 short raw, 1 zstd, 2 raw-long
    int64_t decoded;   // decoded size (-1 unknown)
};
__device__ inline bool parse_cblock(const uint8_t *&p, const uint8_t *end, CBlock &b) {
    if (end - p < 1) return false;
    b.type = *p++;
    if (b.type == 0) {
        if (end - p < 1) return false;
        b.stored = *p++;
        b.decoded = b.stored;
    } else if (b.type == 1) {
        uint64_t n;
        if (!read_varuint(p, end, n) || n > 0xffffffffull) return false;
        b.stored = static_cast<uint32_t>(n);
    } else if (b.type == kBlockRawLong) {
        if (end - p < 4) return false;
        b.stored = p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24);
        p += 4;
        b.decoded = b.stored;
    } else {
        return false;
    }
    if (static_cast<uint64_t>(end - p) < b.stored) return false;
    b.payload = p;
    if (b.type == 1) b.decoded = zstd::frame_content_size(p, b.stored);
    p += b.stored;
    return true;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
    const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ uint64_t load_be64(const uint8_t *p) {
    uint64_t u = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) u = (u << 8) | p[k];
    return u;
}
// stored cell -> the value the scan aggregates: int64 from the order-preserving form (convert/number.go:93-106),
// float64 as its IEEE bits
__device__ __forceinline__ uint64_t cell_value(uint64_t be, bool is_int) {
    if (!is_int) return be;
    if (be >> 63) return be ^ (1ull << 63);
    return 0ull - ((1ull << 63) - be);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ classify
// thread per block: finds the pages that need rewriting, reserves arena space for each
__global__ void classify_pages_kernel(const __grid_constant__ UnpackParams p) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.n_blocks) return;
    const DevBlock blk = p.blocks[b];
    for (uint32_t c = 0; c < blk.n_cols; ++c) {
        const uint32_t ci = blk.col_begin + c;
        const DevCol col = p.cols[ci];
        if (col.size < 2 || col.file_id == p.arena_file_id) continue;
        const uint8_t *page = p.files[col.file_id] + col.off;
        const uint8_t *end = page + col.size;
        const uint8_t enc = page[0];
        uint64_t cap = 0;
        uint32_t kind = 0;
        if (is_numeric(col.value_type)) {
            if (enc != 9) continue;
            kind = kUnpackNumeric;
            cap = 8ull + 9ull * blk.count;
        } else {
            if (enc != 9 && enc != 10) continue;
            const uint8_t *q = page + 1;
            if (enc == 10) {
                uint64_t nv;
                if (!read_varuint(q, end, nv)) continue;
            }
            CBlock lens, data;
            if (!parse_cblock(q, end, lens) || !parse_cblock(q, end, data)) continue;
            if (lens.type != 1 && data.type != 1) continue;  // nothing to inflate
            if (lens.decoded < 0 || data.decoded < 0) {
                atomicAdd(&p.counters[3], 1ull);  // a frame without Frame_Content_Size: left as it is
                continue;
            }
            kind = kUnpackString;
            cap = static_cast<uint64_t>(q - page) + 16 + static_cast<uint64_t>(lens.decoded) + static_cast<uint64_t>(data.decoded) +
                  static_cast<uint64_t>(end - q);
        }
        cap = (cap + 15ull) & ~15ull;
        if (cap > 0xfffffff0ull) continue;
        const unsigned long long at = atomicAdd(&p.counters[1], static_cast<unsigned long long>(cap));
        const unsigned long long j = atomicAdd(&p.counters[0], 1ull);
        if (j < p.max_jobs) {
            UnpackJob job;
            job.col = ci;
            job.rows = blk.count;
            job.out_off = at;
            job.out_cap = static_cast<uint32_t>(cap);
            job.kind = kind;
            p.jobs[j] = job;
        }
    }
}

// ------------------------------------------------------------------------------------------------ unpack
namespace {

// lane 0: brings one compressBlock into `dst` (cap bytes) or points at it in place; returns decoded length or -1
__device__ inline int64_t inflate_block(const CBlock &b, zstd::Workspace *ws, uint8_t *lit, uint8_t *dst, int64_t cap, const uint8_t *&view) {
    if (b.type != 1) {
        view = b.payload;
        return b.stored;
    }
    const int64_t n = zstd::decode_frame(ws, b.payload, b.stored, dst, cap, lit);
    view = dst;
    return n;
}

struct LensView {
    const uint8_t *p;
    uint32_t width;
};
// encodeUint64List (bytes.go:209-240): [type 0..3][n x 1/2/4/8 bytes big endian]
__device__ __forceinline__ bool lens_view(const uint8_t *raw, int64_t len, uint64_t n, LensView &v) {
    if (len < 1 || raw[0] > 3) return false;
    v.width = 1u << raw[0];
    v.p = raw + 1;
    return static_cast<uint64_t>(len) == 1 + n * v.width;
}
__device__ __forceinline__ uint64_t lens_at(const LensView &v, uint64_t i) {
    uint64_t L = 0;
    for (uint32_t k = 0; k < v.width; ++k) L = (L << 8) | v.p[i * v.width + k];
    return L;
}

// numeric fallback page -> raw cells.  Returns the output size or 0 on failure.
__device__ uint32_t unpack_numeric(const uint8_t *page, uint32_t size, uint32_t rows, bool is_int, uint8_t *out, uint32_t out_cap, uint8_t *scratch,
                                   int lane) {
    zstd::Workspace *ws = reinterpret_cast<zstd::Workspace *>(scratch);
    uint8_t *lit = scratch + kOffLit, *tmpA = scratch + kOffTmpA, *tmpB = scratch + kOffTmpB;
    uint64_t *tab_val = reinterpret_cast<uint64_t *>(scratch + kOffTable);
    uint8_t *tab_ok = scratch + kOffTable + 256 * 8;
    const uint64_t need = 8ull + 9ull * rows;
    if (need > out_cap || size < 3) return 0;
    uint64_t *vals = reinterpret_cast<uint64_t *>(out + 8);
    uint8_t *valid = out + 8 + 8ull * rows;
    const uint8_t *end = page + size;
    const uint8_t inner = page[1];  // encodeDefault's own type byte
    // ---- lane 0: inflate the two blocks
    uint64_t lens_ptr = 0, data_ptr = 0, tail_ptr = 0;
    int64_t lens_len = -1, data_len = -1;
    uint64_t nvals = 0;
    if (lane == 0) {
        const uint8_t *q = page + 2;
        bool ok = inner == 9 || inner == 10;
        if (ok && inner == 10) ok = read_varuint(q, end, nvals) && nvals >= 1 && nvals <= 256;
        CBlock lb, db;
        ok = ok && parse_cblock(q, end, lb) && parse_cblock(q, end, db);
        if (ok) {
            const uint8_t *v;
            lens_len = inflate_block(lb, ws, lit, tmpA, kTmpBytes, v);
            lens_ptr = reinterpret_cast<uint64_t>(v);
            data_len = inflate_block(db, ws, lit, tmpB, kTmpBytes, v);
            data_ptr = reinterpret_cast<uint64_t>(v);
            tail_ptr = reinterpret_cast<uint64_t>(q);
        }
    }
    __syncwarp();
    lens_len = static_cast<int64_t>(shfl64(static_cast<uint64_t>(lens_len), 0));
    data_len = static_cast<int64_t>(shfl64(static_cast<uint64_t>(data_len), 0));
    if (lens_len < 0 || data_len < 0) return 0;
    const uint8_t *lens_raw = reinterpret_cast<const uint8_t *>(shfl64(lens_ptr, 0));
    const uint8_t *data = reinterpret_cast<const uint8_t *>(shfl64(data_ptr, 0));
    const uint8_t *tail = reinterpret_cast<const uint8_t *>(shfl64(tail_ptr, 0));
    nvals = shfl64(nvals, 0);
    bool bad = false, any_null = false;
    if (inner == 9) {
        // plain bytes block: cell i is lens[i]-1 bytes long (0 = nil, bytes.go:50-58); numeric cells are 8 bytes
        LensView lv;
        if (!lens_view(lens_raw, lens_len, rows, lv)) return 0;
        uint32_t carry = 0;
        for (uint32_t base = 0; base < rows; base += 32) {
            const uint32_t r = base + lane;
            uint64_t L = r < rows ? lens_at(lv, r) : 0;
            const bool have = L > 0;
            if (have && L != 9) bad = true;
            const uint32_t bal = __ballot_sync(0xffffffffu, have);
            const uint32_t idx = carry + __popc(bal & ((1u << lane) - 1u));
            carry += __popc(bal);
            if (r < rows) {
                uint64_t v = 0;
                if (have && 8ull * idx + 8 <= static_cast<uint64_t>(data_len)) v = cell_value(load_be64(data + 8ull * idx), is_int);
                else if (have) bad = true;
                vals[r] = v;
                valid[r] = have ? 1 : 0;
                any_null |= !have;
            }
        }
        if (8ull * carry != static_cast<uint64_t>(data_len)) bad = true;
    } else {
        // dictionary (dictionary.go:69-114): value table, then bit-packed (value, run) pairs
        LensView lv;
        if (!lens_view(lens_raw, lens_len, nvals, lv)) return 0;
        if (lane == 0) {
            uint64_t off = 0;
            for (uint32_t k = 0; k < nvals; ++k) {
                const uint64_t L = lens_at(lv, k);
                tab_ok[k] = L > 0;
                tab_val[k] = 0;
                if (L > 0) {
                    if (L != 9 || off + 8 > static_cast<uint64_t>(data_len)) {
                        bad = true;
                        break;
                    }
                    tab_val[k] = cell_value(load_be64(data + off), is_int);
                    off += 8;
                }
            }
        }
        __syncwarp();
        const uint8_t *q = tail;
        if (end - q < 5) return 0;
        const uint32_t nrle = (static_cast<uint32_t>(q[0]) << 24) | (q[1] << 16) | (q[2] << 8) | q[3];
        const uint32_t wbits = q[4];
        q += 5;
        if (nrle == 0 || (nrle & 1u) || wbits == 0 || wbits > 32) return 0;
        if (static_cast<uint64_t>(end - q) * 8 < static_cast<uint64_t>(nrle) * wbits) return 0;
        const uint32_t nruns = nrle >> 1;
        const uint64_t vmask = wbits == 32 ? 0xffffffffull : ((1ull << wbits) - 1ull);
        auto bits_at = [&](uint64_t bo) {
            // up to 5 bytes hold a <= 32-bit field at any bit offset; bytes past the page read as zero
            uint64_t x = 0;
            const uint8_t *s = q + (bo >> 3);
            for (int k = 0; k < 5; ++k) x = (x << 8) | (s + k < end ? s[k] : 0);
            return static_cast<uint32_t>((x >> (40 - (bo & 7) - wbits)) & vmask);
        };
        uint32_t row_carry = 0;
        for (uint32_t base = 0; base < nruns; base += 32) {
            const uint32_t ri = base + lane;
            uint32_t value = 0, cnt = 0;
            if (ri < nruns) {
                value = bits_at(static_cast<uint64_t>(2 * ri) * wbits);
                cnt = bits_at(static_cast<uint64_t>(2 * ri + 1) * wbits);
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += o;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t start = row_carry + incl - cnt;
            if (ri < nruns && value >= nvals) {
                bad = true;
                value = 0;
            }
            if (static_cast<uint64_t>(row_carry) + total > rows) {
                bad = true;  // more cells than rows
            } else if (cnt) {
                const uint64_t v = tab_val[value];
                const uint8_t ok = tab_ok[value];
                any_null |= !ok;
                // long runs are spread over the warp, short ones written by their lane
                if (cnt <= 64) {
                    for (uint32_t r = start; r < start + cnt; ++r) {
                        vals[r] = v;
                        valid[r] = ok;
                    }
                }
            }
            const uint32_t longm = __ballot_sync(0xffffffffu, !bad && cnt > 64 && static_cast<uint64_t>(row_carry) + total <= rows);
            uint32_t lm = longm;
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                const uint32_t s0 = __shfl_sync(0xffffffffu, start, src), n0 = __shfl_sync(0xffffffffu, cnt, src);
                const uint32_t val_id = __shfl_sync(0xffffffffu, value, src);
                const uint64_t v = tab_val[val_id];
                const uint8_t ok = tab_ok[val_id];
                for (uint32_t r = s0 + lane; r < s0 + n0; r += 32) {
                    vals[r] = v;
                    valid[r] = ok;
                }
            }
            row_carry += total;
            if (__any_sync(0xffffffffu, bad)) break;
        }
        if (row_carry != rows) bad = true;
    }
    if (__any_sync(0xffffffffu, bad)) return 0;
    const bool nulls = __any_sync(0xffffffffu, any_null);
    if (lane == 0) {
        out[0] = kEncRawCells;
        out[1] = nulls ? 1 : 0;
        for (int k = 2; k < 8; ++k) out[k] = 0;
    }
    return static_cast<uint32_t>(need);
}

// string page -> same page with the zstd blocks inflated in place of the frames. Returns output size or 0.
__device__ uint32_t unpack_string(const uint8_t *page, uint32_t size, uint8_t *out, uint32_t out_cap, uint8_t *scratch, int lane) {
    zstd::Workspace *ws = reinterpret_cast<zstd::Workspace *>(scratch);
    uint8_t *lit = scratch + kOffLit;
    const uint8_t *end = page + size;
    // lane 0 writes headers and inflates; the raw copies are done by the whole warp afterwards
    struct Span {
        uint64_t src;
        uint32_t dst, len;
    };
    Span sp[3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    uint32_t total = 0;
    if (lane == 0) {
        const uint8_t *q = page + 1;
        bool ok = true;
        if (page[0] == 10) {
            uint64_t nv;
            ok = read_varuint(q, end, nv);
        }
        uint32_t o = static_cast<uint32_t>(q - page);
        if (ok && o <= out_cap) {
            for (uint32_t k = 0; k < o; ++k) out[k] = page[k];  // type byte (+ value count)
            for (int b = 0; b < 2 && ok; ++b) {
                const uint8_t *hdr = q;
                CBlock cb;
                ok = parse_cblock(q, end, cb);
                if (!ok) break;
                if (cb.type == 1) {
                    if (cb.decoded < 0 || static_cast<uint64_t>(o) + 5 + static_cast<uint64_t>(cb.decoded) > out_cap) {
                        ok = false;
                        break;
                    }
                    const int64_t n = zstd::decode_frame(ws, cb.payload, cb.stored, out + o + 5, cb.decoded, lit);
                    if (n != cb.decoded) {
                        ok = false;
                        break;
                    }
                    out[o] = kBlockRawLong;
                    out[o + 1] = static_cast<uint8_t>(n);
                    out[o + 2] = static_cast<uint8_t>(n >> 8);
                    out[o + 3] = static_cast<uint8_t>(n >> 16);
                    out[o + 4] = static_cast<uint8_t>(n >> 24);
                    o += 5 + static_cast<uint32_t>(n);
                } else {
                    const uint32_t whole = static_cast<uint32_t>(q - hdr);
                    if (static_cast<uint64_t>(o) + whole > out_cap) {
                        ok = false;
                        break;
                    }
                    sp[b] = Span{reinterpret_cast<uint64_t>(hdr), o, whole};
                    o += whole;
                }
            }
            if (ok) {
                const uint32_t rest = static_cast<uint32_t>(end - q);
                if (static_cast<uint64_t>(o) + rest > out_cap) ok = false;
                else {
                    sp[2] = Span{reinterpret_cast<uint64_t>(q), o, rest};
                    o += rest;
                }
            }
            total = ok ? o : 0;
        }
    }
    __syncwarp();
    total = __shfl_sync(0xffffffffu, total, 0);
    if (total == 0) return 0;
    for (int b = 0; b < 3; ++b) {
        const uint8_t *src = reinterpret_cast<const uint8_t *>(shfl64(sp[b].src, 0));
        const uint32_t dst = __shfl_sync(0xffffffffu, sp[b].dst, 0), len = __shfl_sync(0xffffffffu, sp[b].len, 0);
        for (uint32_t k = lane; k < len; k += 32) out[dst + k] = src[k];
    }
    return total;
}

}  // namespace

__global__ void __launch_bounds__(128) unpack_pages_kernel(const __grid_constant__ UnpackParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t *scratch = p.scratch + static_cast<size_t>(warp_global) * kScratchStride;
    for (;;) {
        unsigned long long j = 0;
        if (lane == 0) j = atomicAdd(&p.counters[2], 1ull);
        j = shfl64(j, 0);
        if (j >= p.n_jobs) break;
        const UnpackJob job = p.jobs[j];
        const DevCol col = p.cols[job.col];
        const uint8_t *page = p.files[col.file_id] + col.off;
        uint8_t *out = p.arena + job.out_off;
        uint32_t n = 0;
        if (job.kind == kUnpackNumeric) n = unpack_numeric(page, col.size, job.rows, col.value_type == BYDB_VT_INT64, out, job.out_cap, scratch, lane);
        else n = unpack_string(page, col.size, out, job.out_cap, scratch, lane);
        __syncwarp();
        if (lane == 0) {
            if (n) {
                DevCol nc = col;
                nc.off = job.out_off;
                nc.size = n;
                nc.file_id = p.arena_file_id;
                p.cols[job.col] = nc;
                atomicAdd(&p.counters[4], 1ull);
            } else {
                atomicAdd(&p.counters[3], 1ull);
            }
        }
    }
}

size_t unpack_scratch_stride() { return kScratchStride; }

void launch_classify_pages(const UnpackParams &p, cudaStream_t s) {
    if (p.n_blocks == 0) return;
    classify_pages_kernel<<<(p.n_blocks + 127) / 128, 128, 0, s>>>(p);
}

void launch_unpack_pages(const UnpackParams &p, int n_warps, cudaStream_t s) {
    if (p.n_jobs == 0 || n_warps <= 0) return;
    unpack_pages_kernel<<<(n_warps + 3) / 4, 128, 0, s>>>(p);
}

void preload_unpack_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, classify_pages_kernel);
    cudaFuncGetAttributes(&a, unpack_pages_kernel);
    cudaGetLastError();
}

}  // namespace bydb
