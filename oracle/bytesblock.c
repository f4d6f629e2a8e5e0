/*
 * bytesblock.c -- ORACLE (test infrastructure): bytes-block, dictionary, bit packing, zstd wrapper.
 * Restates pkg/encoding/{bytes.go,dictionary.go,writer.go,reader.go} and pkg/compress/zstd/zstd.go.
 */
#include "bydb_oracle.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ zstd via dlopen */
/* third-party: the reference uses github.com/klauspost/compress/zstd v1.18.5 (go.mod:171); any
 * RFC 8878 codec is byte-compatible at the decompressed level. No zstd.h in this image, so the
 * prototypes of the stable libzstd API are declared by hand. */
typedef size_t (*zstd_compress_fn)(void *, size_t, const void *, size_t, int);
typedef size_t (*zstd_decompress_fn)(void *, size_t, const void *, size_t);
typedef size_t (*zstd_bound_fn)(size_t);
typedef unsigned (*zstd_iserr_fn)(size_t);
typedef unsigned long long (*zstd_fcs_fn)(const void *, size_t);
static struct {
    int loaded;
    zstd_compress_fn compress;
    zstd_decompress_fn decompress;
    zstd_bound_fn bound;
    zstd_iserr_fn is_error;
    zstd_fcs_fn frame_content_size;
} Z;
static int zstd_load(void) {
    if (Z.loaded) return Z.loaded > 0 ? 0 : -1;
    void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        Z.loaded = -1;
        return -1;
    }
    Z.compress = (zstd_compress_fn)dlsym(h, "ZSTD_compress");
    Z.decompress = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
    Z.bound = (zstd_bound_fn)dlsym(h, "ZSTD_compressBound");
    Z.is_error = (zstd_iserr_fn)dlsym(h, "ZSTD_isError");
    Z.frame_content_size = (zstd_fcs_fn)dlsym(h, "ZSTD_getFrameContentSize");
    Z.loaded = (Z.compress && Z.decompress && Z.bound && Z.is_error && Z.frame_content_size) ? 1 : -1;
    return Z.loaded > 0 ? 0 : -1;
}
/* zstd.go:54-57 Compress (appends) */
int ob_zstd_compress(ob_buf *dst, const void *src, size_t n, int level) {
    if (zstd_load() != 0) return -1;
    size_t bound = Z.bound(n);
    uint8_t *tmp = (uint8_t *)malloc(bound ? bound : 1);
    size_t r = Z.compress(tmp, bound, src, n, level);
    if (Z.is_error(r)) {
        free(tmp);
        return -1;
    }
    ob_buf_append(dst, tmp, r);
    free(tmp);
    return 0;
}
/* zstd.go:49-52 Decompress (appends) */
int ob_zstd_decompress(ob_buf *dst, const void *src, size_t n) {
    if (zstd_load() != 0) return -1;
    unsigned long long fcs = Z.frame_content_size(src, n);
    size_t cap;
    if (fcs == (unsigned long long)-2) return -1;          /* ZSTD_CONTENTSIZE_ERROR */
    if (fcs == (unsigned long long)-1) cap = n * 64 + 4096; /* unknown: grow below */
    else cap = (size_t)fcs;
    for (int attempt = 0; attempt < 8; attempt++) {
        uint8_t *tmp = (uint8_t *)malloc(cap ? cap : 1);
        size_t r = Z.decompress(tmp, cap, src, n);
        if (!Z.is_error(r)) {
            ob_buf_append(dst, tmp, r);
            free(tmp);
            return 0;
        }
        free(tmp);
        if (fcs != (unsigned long long)-1) return -1;
        cap *= 8;
    }
    return -1;
}

/* ------------------------------------------------------------------ compressBlock */
/* bytes.go:291-304 */
static void compress_block(ob_buf *dst, const uint8_t *src, size_t n) {
    if (n < 128) {
        ob_buf_put(dst, 0);
        ob_buf_put(dst, (uint8_t)n);
        ob_buf_append(dst, src, n);
        return;
    }
    ob_buf_put(dst, 1);
    ob_buf z = {0};
    ob_zstd_compress(&z, src, n, 1);
    ob_varuint64_append(dst, z.len);
    ob_buf_append(dst, z.p, z.len);
    ob_buf_free(&z);
}
/* bytes.go:306-350; returns consumed bytes or -1 */
static int64_t decompress_block(ob_buf *dst, const uint8_t *src, size_t n) {
    if (n < 1) return -1;
    uint8_t t = src[0];
    if (t == 0) {
        if (n < 2) return -1;
        size_t bl = src[1];
        if (n - 2 < bl) return -1;
        ob_buf_append(dst, src + 2, bl);
        return (int64_t)(2 + bl);
    }
    if (t == 1) {
        uint64_t bl;
        size_t used = ob_varuint64_read(src + 1, n - 1, &bl);
        /* reference: BytesToVarUint64 returning (src,0) on failure means blockLen 0 */
        if (used == 0) bl = 0;
        if (n - 1 - used < bl) return -1;
        if (ob_zstd_decompress(dst, src + 1 + used, (size_t)bl) != 0) return -1;
        return (int64_t)(1 + used + bl);
    }
    return -1;
}

/* bytes.go:209-240 encodeUint64List + :173-179 EncodeUint64Block */
static void uint64_block_encode(ob_buf *dst, const uint64_t *a, size_t n) {
    ob_buf bb = {0};
    uint64_t nmax = 0;
    for (size_t i = 0; i < n; i++)
        if (a[i] > nmax) nmax = a[i];
    if (nmax < (1ULL << 8)) {
        ob_buf_put(&bb, 0);
        for (size_t i = 0; i < n; i++) ob_buf_put(&bb, (uint8_t)a[i]);
    } else if (nmax < (1ULL << 16)) {
        ob_buf_put(&bb, 1);
        for (size_t i = 0; i < n; i++) {
            ob_buf_put(&bb, (uint8_t)(a[i] >> 8));
            ob_buf_put(&bb, (uint8_t)a[i]);
        }
    } else if (nmax < (1ULL << 32)) {
        ob_buf_put(&bb, 2);
        for (size_t i = 0; i < n; i++)
            for (int k = 3; k >= 0; k--) ob_buf_put(&bb, (uint8_t)(a[i] >> (8 * k)));
    } else {
        ob_buf_put(&bb, 3);
        for (size_t i = 0; i < n; i++)
            for (int k = 7; k >= 0; k--) ob_buf_put(&bb, (uint8_t)(a[i] >> (8 * k)));
    }
    compress_block(dst, bb.p, bb.len);
    ob_buf_free(&bb);
}
/* bytes.go:182-197 DecodeUint64Block + :242-286 decodeUint64List; returns consumed or -1 */
static int64_t uint64_block_decode(uint64_t *dst, size_t n, const uint8_t *src, size_t srclen) {
    ob_buf bb = {0};
    int64_t used = decompress_block(&bb, src, srclen);
    if (used < 0 || bb.len < 1) {
        ob_buf_free(&bb);
        return -1;
    }
    uint8_t t = bb.p[0];
    const uint8_t *s = bb.p + 1;
    size_t sl = bb.len - 1;
    int w = t == 0 ? 1 : t == 1 ? 2 : t == 2 ? 4 : t == 3 ? 8 : 0;
    if (w == 0 || sl != (size_t)w * n) {
        ob_buf_free(&bb);
        return -1;
    }
    for (size_t i = 0; i < n; i++) {
        uint64_t v = 0;
        for (int k = 0; k < w; k++) v = (v << 8) | s[i * (size_t)w + (size_t)k];
        dst[i] = v;
    }
    ob_buf_free(&bb);
    return used;
}

/* bytes.go:45-72 EncodeBytesBlock */
void ob_bytes_block_encode(ob_buf *dst, const ob_bytes *a, size_t n) {
    uint64_t *lens = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) lens[i] = a[i].len < 0 ? 0 : (uint64_t)a[i].len + 1;
    uint64_block_encode(dst, lens, n);
    free(lens);
    ob_buf bb = {0};
    for (size_t i = 0; i < n; i++)
        if (a[i].len > 0) ob_buf_append(&bb, a[i].p, (size_t)a[i].len);
    compress_block(dst, bb.p, bb.len);
    ob_buf_free(&bb);
}

/* bytes.go:84-127 Decode (allow_tail=0) / :130-170 DecodeWithTail (allow_tail=1).
 * Items are materialised as offsets into arena; out[i].p is fixed up at the end because the arena
 * may be reallocated while appending. */
int64_t ob_bytes_block_decode(ob_bytes *out, size_t n, const uint8_t *src, size_t srclen, ob_buf *arena, int allow_tail) {
    uint64_t *lens = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
    int64_t used = uint64_block_decode(lens, n, src, srclen);
    if (used < 0) {
        free(lens);
        return -1;
    }
    size_t base = arena->len;
    int64_t used2 = decompress_block(arena, src + used, srclen - (size_t)used);
    if (used2 < 0) {
        free(lens);
        return -1;
    }
    if (!allow_tail && (size_t)(used + used2) != srclen) {
        free(lens);
        return -1;
    }
    size_t avail = arena->len - base;
    size_t off = 0;
    for (size_t i = 0; i < n; i++) {
        if (lens[i] == 0) {
            out[i].p = NULL;
            out[i].len = -1;
            continue;
        }
        uint64_t al = lens[i] - 1;
        if (avail - off < al) {
            free(lens);
            return -1;
        }
        out[i].p = (const uint8_t *)(uintptr_t)(base + off); /* offset for now */
        out[i].len = (int64_t)al;
        off += (size_t)al;
    }
    free(lens);
    for (size_t i = 0; i < n; i++)
        if (out[i].len >= 0) out[i].p = arena->p + (size_t)(uintptr_t)out[i].p;
    return used + used2;
}

/* ------------------------------------------------------------------ bit writer (writer.go:25-96) */
void ob_bitw_init(ob_bitw *w, ob_buf *out) {
    w->out = out;
    w->cache = 0;
    w->available = 8;
}
void ob_bitw_bool(ob_bitw *w, int b) {
    if (b) w->cache |= (uint8_t)(1u << (w->available - 1));
    w->available--;
    if (w->available == 0) {
        ob_buf_put(w->out, w->cache);
        w->cache = 0;
        w->available = 8;
    }
}
void ob_bitw_byte(ob_bitw *w, uint8_t b) {
    ob_buf_put(w->out, (uint8_t)(w->cache | (b >> (8 - w->available))));
    w->cache = (uint8_t)(b << w->available); /* Go: byte shift by 8 yields 0 */
    if (w->available == 8) w->cache = 0;
}
void ob_bitw_bits(ob_bitw *w, uint64_t u, int nbits) {
    if (nbits <= 0) return;
    u <<= (64 - (unsigned)nbits);
    for (; nbits >= 8; nbits -= 8) {
        ob_bitw_byte(w, (uint8_t)(u >> 56));
        u <<= 8;
    }
    uint8_t rem = (uint8_t)(u >> 56);
    for (; nbits > 0; nbits--) {
        ob_bitw_bool(w, (rem & 0x80) != 0);
        rem = (uint8_t)(rem << 1);
    }
}
void ob_bitw_flush(ob_bitw *w) {
    if (w->available != 8) ob_buf_put(w->out, w->cache);
    w->cache = 0;
    w->available = 8;
}

/* bit reader (reader.go:25-98) over a byte slice */
typedef struct {
    const uint8_t *p;
    size_t n, pos;
    uint8_t cache, len;
} bitr;
static int bitr_bool(bitr *r, int *out) {
    if (r->len == 0) {
        if (r->pos >= r->n) return -1;
        r->cache = r->p[r->pos++];
        r->len = 8;
    }
    r->len--;
    *out = (r->cache & 0x80) != 0;
    r->cache = (uint8_t)(r->cache << 1);
    return 0;
}
static int bitr_byte(bitr *r, uint8_t *out) {
    if (r->pos >= r->n) return -1;
    uint8_t b = r->p[r->pos++];
    if (r->len == 0) {
        r->cache = b; /* reader.go:81-88: cache assigned but len stays 0 */
        *out = b;
        return 0;
    }
    *out = (uint8_t)(r->cache | (b >> r->len));
    r->cache = (uint8_t)(b << (8 - r->len));
    return 0;
}
static int bitr_bits(bitr *r, int nbits, uint64_t *out) {
    uint64_t res = 0;
    for (; nbits >= 8; nbits -= 8) {
        uint8_t b;
        if (bitr_byte(r, &b)) return -1;
        res = (res << 8) | b;
    }
    for (; nbits > 0; nbits--) {
        int bit;
        if (bitr_bool(r, &bit)) return -1;
        res = (res << 1) | (uint64_t)bit;
    }
    *out = res;
    return 0;
}

/* test hook for pkg/encoding/reader_test.go: runs a script of reads over `data`.
 * ops[i]: 0 = ReadBool, -1 = ReadByte, n > 0 = ReadBits(n); out[i] receives the value. Returns 0 or -1. */
int ob_bit_reader_script(const uint8_t *data, size_t n, const int *ops, size_t n_ops, uint64_t *out) {
    bitr r = {data, n, 0, 0, 0};
    for (size_t i = 0; i < n_ops; i++) {
        if (ops[i] == 0) {
            int bit;
            if (bitr_bool(&r, &bit)) return -1;
            out[i] = (uint64_t)bit;
        } else if (ops[i] < 0) {
            uint8_t b;
            if (bitr_byte(&r, &b)) return -1;
            out[i] = b;
        } else if (bitr_bits(&r, ops[i], &out[i])) {
            return -1;
        }
    }
    return 0;
}

/* dictionary.go:199-219 bitPackingEncoder.encode + :253-261 encodeBitPacking */
void ob_bitpack_encode(ob_buf *dst, const uint32_t *src, size_t n) {
    ob_bitw w;
    ob_bitw_init(&w, dst);
    if (n == 0) {
        ob_bitw_bits(&w, 0, 32);
        ob_bitw_flush(&w);
        return;
    }
    ob_bitw_bits(&w, (uint64_t)n, 32);
    uint32_t maxv = 0;
    for (size_t i = 0; i < n; i++)
        if (src[i] > maxv) maxv = src[i];
    int width = 1;
    if (maxv > 0) width = 32 - __builtin_clz(maxv);
    ob_bitw_bits(&w, (uint64_t)width, 8);
    for (size_t i = 0; i < n; i++) ob_bitw_bits(&w, src[i], width);
    ob_bitw_flush(&w);
}

/* dictionary.go:52-66 valuesEqual */
static int values_equal(const ob_bytes *a, const ob_bytes *b) {
    if (a->len < 0 && b->len < 0) return 1;
    if (a->len < 0 || b->len < 0) return 0;
    return a->len == b->len && (a->len == 0 || memcmp(a->p, b->p, (size_t)a->len) == 0);
}

/* dictionary.go:36-50 Add (x n) + :69-77 Encode + :164-182 encodeRLE */
int ob_dictionary_encode(ob_buf *dst, const ob_bytes *a, size_t n) {
    ob_bytes values[256];
    size_t nv = 0;
    uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        size_t k;
        for (k = 0; k < nv; k++)
            if (values_equal(&values[k], &a[i])) break;
        if (k == nv) {
            if (nv == 256) {
                free(idx);
                return 0;
            }
            values[nv++] = a[i];
        }
        idx[i] = (uint32_t)k;
    }
    ob_varuint64_append(dst, nv);
    ob_bytes_block_encode(dst, values, nv);
    /* RLE: (value,count) pairs */
    uint32_t *rle = (uint32_t *)malloc(sizeof(uint32_t) * (2 * n + 2));
    size_t nr = 0;
    if (n > 0) {
        uint32_t cur = idx[0], cnt = 1;
        for (size_t i = 1; i < n; i++) {
            if (idx[i] == cur) {
                cnt++;
            } else {
                rle[nr++] = cur;
                rle[nr++] = cnt;
                cur = idx[i];
                cnt = 1;
            }
        }
        rle[nr++] = cur;
        rle[nr++] = cnt;
    }
    ob_bitpack_encode(dst, rle, nr);
    free(rle);
    free(idx);
    return 1;
}

/* dictionary.go:90-114 Decode (+ :116-162 decodeBytesBlockWithTail, :184-197 decodeRLE, :230-251 decode) */
int ob_dictionary_decode(ob_bytes *out, size_t n, const uint8_t *src, size_t srclen, ob_buf *arena) {
    uint64_t count;
    size_t used = ob_varuint64_read(src, srclen, &count);
    if (used == 0) count = 0;
    if (count == 0) return n == 0 ? 0 : -1; /* reference returns dst unchanged: zero items */
    if (count > 256) return -1;
    ob_bytes values[256];
    /* the dictionary values must stay valid: decode into arena, but arena may grow only here */
    int64_t u2 = ob_bytes_block_decode(values, (size_t)count, src + used, srclen - used, arena, 1);
    if (u2 < 0) return -1;
    bitr r = {src + used + (size_t)u2, srclen - used - (size_t)u2, 0, 0, 0};
    uint64_t length;
    if (bitr_bits(&r, 32, &length)) return -1;
    size_t produced = 0;
    if (length > 0) {
        uint64_t width;
        if (bitr_bits(&r, 8, &width)) return -1;
        if (length % 2 != 0) return -1; /* reference would index out of range */
        for (uint64_t i = 0; i < length; i += 2) {
            uint64_t value, cnt;
            if (bitr_bits(&r, (int)width, &value)) return -1;
            if (bitr_bits(&r, (int)width, &cnt)) return -1;
            if (value >= count) return -1;
            for (uint64_t j = 0; j < cnt; j++) {
                if (produced >= n) return -1;
                out[produced++] = values[value];
            }
        }
    }
    return produced == n ? 0 : -1;
}
