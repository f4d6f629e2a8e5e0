/*
 * codec.c -- ORACLE (test infrastructure): integer / float list codecs.
 * Restates pkg/encoding/{int.go,int_list.go,delta.go,float.go} and pkg/convert/number.go
 * of the reference.  See bydb_oracle.h for the rules on who may call this.
 */
#include "bydb_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ buffers */
static void ob_buf_grow(ob_buf *b, size_t need) {
    if (b->len + need <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : 256;
    while (nc < b->len + need) nc *= 2;
    b->p = (uint8_t *)realloc(b->p, nc);
    b->cap = nc;
}
void ob_buf_free(ob_buf *b) {
    free(b->p);
    b->p = NULL;
    b->len = b->cap = 0;
}
void ob_buf_reset(ob_buf *b) { b->len = 0; }
void ob_buf_append(ob_buf *b, const void *src, size_t n) {
    if (n == 0) return;
    ob_buf_grow(b, n);
    memcpy(b->p + b->len, src, n);
    b->len += n;
}
void ob_buf_put(ob_buf *b, uint8_t c) {
    ob_buf_grow(b, 1);
    b->p[b->len++] = c;
}

/* ------------------------------------------------------------------ varints */
/* pkg/encoding/int.go:75-99 VarInt64ListToBytes (single value).
 * The 1-byte fast path (|v|<0x40) is the general zig-zag rule restricted to int8. */
void ob_varint64_append(ob_buf *dst, int64_t v) {
    uint64_t u = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);
    while (u > 0x7f) {
        ob_buf_put(dst, (uint8_t)(0x80 | (u & 0x7f)));
        u >>= 7;
    }
    ob_buf_put(dst, (uint8_t)u);
}

/* pkg/encoding/int.go:152-185 VarUint64ToBytes / VarUint64sToBytes */
void ob_varuint64_append(ob_buf *dst, uint64_t u) {
    while (u > 0x7f) {
        ob_buf_put(dst, (uint8_t)(0x80 | (u & 0x7f)));
        u >>= 7;
    }
    ob_buf_put(dst, (uint8_t)u);
}

/* pkg/encoding/int.go:189-211 BytesToVarUint64 (binary.Uvarint semantics on the slow path).
 * Returns bytes consumed; 0 when src is empty/truncated/overflowing (the reference returns
 * (src, 0) unchanged in those cases, i.e. consumes nothing). */
size_t ob_varuint64_read(const uint8_t *src, size_t n, uint64_t *out) {
    uint64_t x = 0;
    unsigned s = 0;
    *out = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t b = src[i];
        if (i == 10) return 0; /* binary.Uvarint: overflow */
        if (b < 0x80) {
            if (i == 9 && b > 1) return 0;
            *out = x | ((uint64_t)b << s);
            return i + 1;
        }
        x |= (uint64_t)(b & 0x7f) << s;
        s += 7;
    }
    return 0;
}

/* pkg/encoding/int.go:111-148 BytesToVarInt64List */
size_t ob_varint64_list_read(const uint8_t *src, size_t n, int64_t *dst, size_t cnt) {
    size_t idx = 0;
    for (size_t i = 0; i < cnt; i++) {
        if (idx >= n) return (size_t)-1;
        uint8_t c = src[idx++];
        if (c < 0x80) {
            int8_t v = (int8_t)((int8_t)(c >> 1) ^ (int8_t)((int8_t)(c << 7) >> 7));
            dst[i] = (int64_t)v;
            continue;
        }
        uint64_t u = (uint64_t)(c & 0x7f);
        size_t start = idx - 1;
        unsigned shift = 0;
        while (c >= 0x80) {
            if (idx >= n) return (size_t)-1;
            if (idx - start > 9) return (size_t)-1;
            c = src[idx++];
            shift += 7;
            u |= (uint64_t)(c & 0x7f) << (shift & 63);
        }
        dst[i] = (int64_t)(u >> 1) ^ ((int64_t)(u << 63) >> 63);
    }
    return idx;
}

/* ------------------------------------------------------------------ int list */
/* wrapping int64 subtraction/addition, as Go's int64 arithmetic */
static inline int64_t wsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
static inline int64_t wadd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }

/* int_list.go:112-123 */
static int is_const(const int64_t *a, size_t n) {
    if (n == 0) return 0;
    for (size_t i = 0; i < n; i++)
        if (a[i] != a[0]) return 0;
    return 1;
}
/* int_list.go:146-148 */
static inline int64_t sign_bit(int64_t n) { return (n >> 63) & 1; }
/* int_list.go:125-144 */
static void is_delta(const int64_t *a, size_t n, int *isd, int *isdc) {
    *isd = *isdc = 0;
    if (n < 2) return;
    int ct = 1;
    int64_t d1 = wsub(a[1], a[0]);
    int64_t asc = sign_bit(d1);
    int64_t prev = a[1];
    for (size_t i = 2; i < n; i++) {
        int64_t d = wsub(a[i], prev);
        if ((sign_bit(d) ^ asc) == 1) return;
        if (ct && d != d1) ct = 0;
        prev = a[i];
    }
    *isd = 1;
    *isdc = ct;
}
/* int_list.go:150-179 */
static int is_incremental(const int64_t *a, size_t n) {
    if (n < 2) return 0;
    size_t resets = 0;
    int64_t vprev = a[0];
    if (vprev < 0) return 1;
    for (size_t i = 1; i < n; i++) {
        int64_t v = a[i];
        if (v < vprev) {
            if (v < 0) return 0;
            if (v > (vprev >> 3)) return 0;
            resets++;
        }
        vprev = v;
    }
    if (resets <= 2) return 1;
    return resets < (n >> 3);
}

/* delta.go:26-43 int64ListDeltaToBytes */
static void delta_encode(ob_buf *dst, const int64_t *src, size_t n) {
    int64_t v = src[0];
    for (size_t i = 1; i < n; i++) {
        int64_t d = wsub(src[i], v);
        v = wadd(v, d);
        ob_varint64_append(dst, d);
    }
}
/* delta.go:72-89 int64sDeltaOfDeltaToBytes */
static void dod_encode(ob_buf *dst, const int64_t *src, size_t n) {
    int64_t d1 = wsub(src[1], src[0]);
    ob_varint64_append(dst, d1);
    int64_t v = src[1];
    for (size_t i = 2; i < n; i++) {
        int64_t d2 = wsub(wsub(src[i], v), d1);
        d1 = wadd(d1, d2);
        v = wadd(v, d1);
        ob_varint64_append(dst, d2);
    }
}

/* int_list.go:27-53 Int64ListToBytes */
int ob_int64_list_encode(ob_buf *dst, const int64_t *a, size_t n, int64_t *first) {
    if (n == 0) return OB_ENC_UNKNOWN; /* reference panics */
    *first = a[0];
    if (is_const(a, n)) return OB_ENC_CONST;
    int isd, isdc;
    is_delta(a, n, &isd, &isdc);
    if (isdc) {
        ob_varint64_append(dst, wsub(a[1], a[0]));
        return OB_ENC_DELTA_CONST;
    }
    if (isd) {
        dod_encode(dst, a, n);
        return OB_ENC_DELTA_OF_DELTA;
    }
    if (is_incremental(a, n)) {
        dod_encode(dst, a, n);
        return OB_ENC_DELTA_OF_DELTA;
    }
    delta_encode(dst, a, n);
    return OB_ENC_DELTA;
}

/* int_list.go:57-101 BytesToInt64List; delta.go:45-70, 91-118 */
int ob_int64_list_decode(int64_t *dst, const uint8_t *src, size_t srclen, int enc, int64_t first, size_t count) {
    switch (enc) {
    case OB_ENC_DELTA: {
        if (count < 1) return -1;
        int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (count ? count : 1));
        size_t used = ob_varint64_list_read(src, srclen, tmp, count - 1);
        if (used == (size_t)-1 || used != srclen) {
            free(tmp);
            return -1;
        }
        int64_t v = first;
        dst[0] = v;
        for (size_t i = 0; i + 1 < count; i++) {
            v = wadd(v, tmp[i]);
            dst[i + 1] = v;
        }
        free(tmp);
        return 0;
    }
    case OB_ENC_DELTA_OF_DELTA: {
        if (count < 2) return -1;
        int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * count);
        size_t used = ob_varint64_list_read(src, srclen, tmp, count - 1);
        if (used == (size_t)-1 || used != srclen) {
            free(tmp);
            return -1;
        }
        int64_t v = first, d1 = tmp[0];
        dst[0] = v;
        v = wadd(v, d1);
        dst[1] = v;
        for (size_t i = 1; i + 1 < count; i++) {
            d1 = wadd(d1, tmp[i]);
            v = wadd(v, d1);
            dst[i + 1] = v;
        }
        free(tmp);
        return 0;
    }
    case OB_ENC_CONST:
        if (srclen > 0) return -1;
        for (size_t i = 0; i < count; i++) dst[i] = first;
        return 0;
    case OB_ENC_DELTA_CONST: {
        int64_t d;
        size_t used = ob_varint64_list_read(src, srclen, &d, 1);
        if (used == (size_t)-1 || used != srclen) return -1;
        int64_t v = first;
        for (size_t i = 0; i < count; i++) {
            dst[i] = v;
            v = wadd(v, d);
        }
        return 0;
    }
    default:
        return -1;
    }
}

/* ------------------------------------------------------------------ pkg/convert/number.go */
/* number.go:33-45 Int64ToBytes: order-preserving, NOT two's complement */
void ob_conv_int64_to_bytes(int64_t i, uint8_t out[8]) {
    uint64_t u;
    if (i >= 0) {
        u = (uint64_t)i | (1ULL << 63);
    } else {
        uint64_t absu = (uint64_t)0 - (uint64_t)i; /* -abs wraps for MinInt64 exactly as Go */
        u = (1ULL << 63) - absu;
    }
    for (int k = 0; k < 8; k++) out[k] = (uint8_t)(u >> (56 - 8 * k));
}
/* number.go:93-106 BytesToInt64 */
int64_t ob_conv_bytes_to_int64(const uint8_t b[8]) {
    uint64_t u = 0;
    for (int k = 0; k < 8; k++) u = (u << 8) | b[k];
    if (b[0] >= 128) {
        u ^= 1ULL << 63;
        return (int64_t)u;
    }
    u = (1ULL << 63) - u;
    return (int64_t)((uint64_t)0 - u);
}

/* ------------------------------------------------------------------ pkg/encoding/float.go */
/* Go math.Pow10 (src/math/pow10.go): pow10tab[n%32] * pow10postab32[n/32] for 0<=n<=308; +Inf above;
 * for -323<=n<0: pow10negtab32[-n/32] / pow10tab[-n%32]; 0 below. */
double ob_pow10(int n) {
    static const double tab[32] = {1e00, 1e01, 1e02, 1e03, 1e04, 1e05, 1e06, 1e07, 1e08, 1e09, 1e10,
                                   1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21,
                                   1e22, 1e23, 1e24, 1e25, 1e26, 1e27, 1e28, 1e29, 1e30, 1e31};
    static const double postab32[10] = {1e00, 1e32, 1e64, 1e96, 1e128, 1e160, 1e192, 1e224, 1e256, 1e288};
    static const double negtab32[11] = {1e-00, 1e-32, 1e-64, 1e-96, 1e-128, 1e-160, 1e-192, 1e-224, 1e-256, 1e-288, 1e-320};
    if (0 <= n && n <= 308) return postab32[n / 32] * tab[n % 32];
    if (-323 <= n && n <= 0) return negtab32[(-n) / 32] / tab[(-n) % 32];
    if (n > 0) return INFINITY;
    return 0;
}

/* float.go:69-93 DecimalIntListToFloat64List (+ computeDivisors :96-103) */
void ob_decimal_list_to_float64(double *dst, const int64_t *vals, size_t n, int16_t exp) {
    if (exp >= 0) {
        volatile double scale = ob_pow10((int)exp);
        for (size_t i = 0; i < n; i++) {
            volatile double r = (double)vals[i] * scale; /* no FMA contraction possible: single op */
            dst[i] = r;
        }
        return;
    }
    double divisors[8];
    int nd = 0;
    int neg = -(int)exp;
    while (neg > 0) {
        int step = neg < 308 ? neg : 308;
        divisors[nd++] = ob_pow10(step);
        neg -= step;
    }
    for (size_t i = 0; i < n; i++) {
        volatile double r = (double)vals[i];
        for (int k = 0; k < nd; k++) r = r / divisors[k];
        dst[i] = r;
    }
}

/* Shortest round-trip decimal digits of a finite non-zero double, as Go's
 * strconv.AppendFloat(f,'e',-1,64): minimal digit count that uniquely identifies f, and among
 * those the decimal closest to f.  digits[] gets the ASCII digits (no dot), *dexp the exponent of
 * the FIRST digit (d.ddd e dexp).  We search precisions upward; at each precision the correctly
 * rounded candidate and its two neighbours are tested for round trip (covers the asymmetric
 * rounding interval at binade boundaries). */
static int roundtrips(const char *digits, int nd, int dexp, int neg, double f) {
    char buf[64];
    int k = 0;
    if (neg) buf[k++] = '-';
    buf[k++] = digits[0];
    buf[k++] = '.';
    for (int i = 1; i < nd; i++) buf[k++] = digits[i];
    if (nd == 1) buf[k++] = '0';
    snprintf(buf + k, sizeof(buf) - (size_t)k, "e%d", dexp);
    double g = strtod(buf, NULL);
    return memcmp(&g, &f, sizeof(double)) == 0;
}
static long double cand_value(const char *digits, int nd, int dexp) {
    char buf[64];
    int k = 0;
    buf[k++] = digits[0];
    buf[k++] = '.';
    for (int i = 1; i < nd; i++) buf[k++] = digits[i];
    if (nd == 1) buf[k++] = '0';
    snprintf(buf + k, sizeof(buf) - (size_t)k, "e%d", dexp);
    return strtold(buf, NULL);
}
/* digit-string +/- 1 in the last place; may change nd/dexp on carry */
static void digits_inc(char *d, int *nd, int *dexp) {
    int i = *nd - 1;
    while (i >= 0 && d[i] == '9') d[i--] = '0';
    if (i >= 0) {
        d[i]++;
        return;
    }
    /* 999 -> 1000: keep nd digits: "100" with exponent+1 */
    d[0] = '1';
    for (int k = 1; k < *nd; k++) d[k] = '0';
    (*dexp)++;
}
static int digits_dec(char *d, int *nd, int *dexp) {
    int i = *nd - 1;
    while (i >= 0 && d[i] == '0') d[i--] = '9';
    if (i < 0) return 0;
    d[i]--;
    if (d[0] == '0') {
        if (*nd == 1) return 0;
        /* 1000 -> 0999: becomes 9990 at exponent-1 (same digit count) */
        memmove(d, d + 1, (size_t)(*nd - 1));
        d[*nd - 1] = '9';
        (*dexp)--;
    }
    return 1;
}
static void shortest_digits(double f, char *digits, int *nd_out, int *dexp_out) {
    double a = fabs(f);
    for (int prec = 0; prec <= 16; prec++) {
        char buf[64];
        snprintf(buf, sizeof buf, "%.*e", prec, a);
        /* parse d.ddddde[+-]xx */
        char cd[24];
        int nd = 0;
        const char *p = buf;
        while (*p && *p != 'e') {
            if (*p >= '0' && *p <= '9') cd[nd++] = *p;
            p++;
        }
        int dexp = atoi(p + 1);
        char best[24];
        int bnd = 0, bexp = 0, found = 0;
        long double bdist = 0;
        for (int c = 0; c < 3; c++) {
            char t[24];
            memcpy(t, cd, (size_t)nd);
            int tnd = nd, texp = dexp;
            if (c == 1) digits_inc(t, &tnd, &texp);
            if (c == 2 && !digits_dec(t, &tnd, &texp)) continue;
            if (!roundtrips(t, tnd, texp, 0, a)) continue;
            long double dist = fabsl(cand_value(t, tnd, texp) - (long double)a);
            if (!found || dist < bdist) {
                memcpy(best, t, (size_t)tnd);
                bnd = tnd;
                bexp = texp;
                bdist = dist;
                found = 1;
            }
        }
        if (found) {
            memcpy(digits, best, (size_t)bnd);
            *nd_out = bnd;
            *dexp_out = bexp;
            return;
        }
    }
    /* 17 significant digits always round-trip */
    char buf[64];
    snprintf(buf, sizeof buf, "%.16e", a);
    int nd = 0;
    const char *p = buf;
    while (*p && *p != 'e') {
        if (*p >= '0' && *p <= '9') digits[nd++] = *p;
        p++;
    }
    *nd_out = nd;
    *dexp_out = atoi(p + 1);
}

int ob_force_slow_float = 0; /* tests: force the general shortest-digits search */

/* float.go:107-124 floatToDecimal + :128-190 floatToDecimalSlow */
int ob_float_to_decimal(double f, int64_t *mant, int16_t *exp) {
    if (isnan(f) || isinf(f)) return -1;
    if (f == 0) {
        *mant = 0;
        *exp = 0;
        return 0;
    }
    /* Go int64(f) on amd64 (CVTTSD2SQ): out-of-range -> MinInt64 */
    int64_t u;
    if (f >= 9223372036854775808.0 || f < -9223372036854775808.0)
        u = INT64_MIN;
    else
        u = (int64_t)f;
    if ((double)u == f) {
        int16_t e = 0;
        while (u != 0 && u % 10 == 0) {
            u /= 10;
            e++;
        }
        *mant = u;
        *exp = e;
        return 0;
    }
    /* Fast equivalent of the shortest-digits search for "short" decimals: the smallest k such that
     * some integer m (|m| < 2^53) has fl(m / 10^k) == f.  m and 10^k (k <= 22) are exact doubles, so
     * the correctly rounded quotient equals strtod("m e-k"): m*10^-k round-trips, and no decimal
     * with fewer fractional digits does (smaller k failed), hence it is the shortest one.  Falls
     * through to the general search when nothing is found (checked against it in the tests). */
    if (!ob_force_slow_float) {
        static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
        double a = fabs(f);
        if (a < 9007199254740992.0 && a >= 1e-15) {
            for (int k = 1; k <= 15; k++) {
                double t = a * p10[k];
                /* only mantissas of <= 15 digits: there the doubles around t are < 1 apart after scaling, so at most ONE
                 * integer m maps back to f.  With 16-17 digits several neighbours round-trip and the shortest-digits rule
                 * (strconv 'e', -1: the candidate closest to the exact value) must decide -- that is the general search. */
                if (t >= 1e15) break;
                double m0 = nearbyint(t);
                for (int dm = -1; dm <= 1; dm++) {
                    double m = m0 + dm;
                    if (m < 1 || m >= 9007199254740992.0) continue;
                    volatile double back = m / p10[k];
                    if (back == a) {
                        int64_t mi = (int64_t)m;
                        if (mi % 10 == 0) continue; /* would have matched at k-1; keep searching */
                        *mant = f < 0 ? -mi : mi;
                        *exp = (int16_t)-k;
                        return 0;
                    }
                }
            }
        }
    }
    /* slow path: shortest 'e' formatting d.ddd e sciExp -> mantissa digits without the dot */
    char digits[24];
    int nd, sci;
    shortest_digits(f, digits, &nd, &sci);
    int frac = nd - 1; /* digits after the dot */
    /* strip trailing zeros (float.go:166-169): keeps at least one digit */
    while (nd > 1 && digits[nd - 1] == '0') {
        nd--;
        frac--;
    }
    /* strconv.ParseInt(...,10,64): overflow -> error */
    uint64_t m = 0;
    for (int i = 0; i < nd; i++) {
        uint64_t dgt = (uint64_t)(digits[i] - '0');
        if (m > (UINT64_MAX - dgt) / 10) return -1;
        m = m * 10 + dgt;
    }
    if (m > (uint64_t)INT64_MAX) return -1;
    if (sci > 32767 || sci < -32768) return -1;
    int e = sci - frac;
    *exp = (int16_t)e; /* int16(sciExp) - fracDigits wraps in int16 like Go */
    *mant = f < 0 ? -(int64_t)m : (int64_t)m;
    return 0;
}

/* float.go:199-230 mulPow10Fast / mulPow10Large */
static int mul_pow10(int64_t v, int n, int64_t *out) {
    static const int64_t tab[19] = {1LL,
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
    if (n < 0) return 0;
    while (n >= 19) {
        if (v > INT64_MAX / tab[18] || v < INT64_MIN / tab[18]) return 0;
        v *= tab[18];
        n -= 18;
    }
    if (n > 0) {
        if (v > INT64_MAX / tab[n] || v < INT64_MIN / tab[n]) return 0;
        v *= tab[n];
    }
    *out = v;
    return 1;
}

/* float.go:30-66 Float64ListToDecimalIntList */
int ob_float64_to_decimal_list(int64_t *dst, const double *src, size_t n, int16_t *exp_out) {
    *exp_out = 0;
    if (n == 0) return 0;
    int16_t *exps = (int16_t *)malloc(sizeof(int16_t) * n);
    int16_t min_exp = INT16_MAX;
    for (size_t i = 0; i < n; i++) {
        int64_t d;
        int16_t e;
        if (ob_float_to_decimal(src[i], &d, &e) != 0) {
            free(exps);
            return -1;
        }
        dst[i] = d;
        exps[i] = e;
        if (e < min_exp) min_exp = e;
    }
    for (size_t i = 0; i < n; i++) {
        int16_t diff = (int16_t)(exps[i] - min_exp);
        if (diff == 0) continue;
        int64_t scaled;
        if (!mul_pow10(dst[i], diff, &scaled)) {
            free(exps);
            return -1;
        }
        dst[i] = scaled;
    }
    free(exps);
    *exp_out = min_exp;
    return 0;
}
