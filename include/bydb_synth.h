/*
 * bydb_synth.h -- C ABI of the host-side measure part WRITER in libbydbgpu.so.
 *
 * Produces byte-exact measure parts (meta.bin, primary.bin, timestamps.bin, fv.bin, <family>.tf,
 * <family>.tfm) the way the reference's flush path does; it is what generates the synthetic parts
 * the benchmark scans and what a GPU-side compaction would reuse.  Reference code it restates:
 *
 *   bydb_part_write   <- banyand/measure/part.go:162-233 (memPart.mustInitFromDataPoints / mustFlush),
 *                        block_writer.go:206-285 (mustWriteBlock / Flush), block.go:139-160,361-379,
 *                        column.go:113-234 (column pages), block_metadata.go:113-131,268-277,
 *                        column_metadata.go:47-106, primary_metadata.go:47-68,
 *                        pkg/encoding/{int.go,int_list.go,delta.go,float.go,bytes.go,dictionary.go,writer.go}
 *   bydb_synth_part   <- the data shapes of banyand/measure/benchmark_encode_test.go:119-201
 *                        (const / incrementing / small_fluctuations / random ...) laid out as
 *                        n_series x n_points regular series
 *
 * Host only (multi-threaded C++); no CUDA involved.  Checked byte-for-byte against the oracle's
 * independent writer in tests/test_writer_vs_oracle.py.
 */
#ifndef BYDB_SYNTH_H
#define BYDB_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bydb_part_image bydb_part_image; /* a finished part held in host memory */

/* one column of already sorted rows */
typedef struct {
    const char *name;
    int32_t value_type;       /* BYDB_VT_INT64 / BYDB_VT_FLOAT64 / BYDB_VT_STR                          */
    int32_t dec_digits;       /* FLOAT64 only: >=0 -> values are given as decimals dec_k[i] / 10^dec_digits */
    const int64_t *i64;       /* INT64 values                                                        */
    const double *f64;        /* FLOAT64 values (dec_digits < 0)                                     */
    const int64_t *dec_k;     /* FLOAT64 decimal numerators (dec_digits >= 0); value = fl(k / 10^d)  */
    const uint32_t *str_idx;  /* STR: index into str_values per row                                  */
    const char *const *str_values;
    uint32_t n_str_values;
    uint32_t reserved;
} bydb_wcolumn;

/* rows sorted by (series id asc, timestamp asc), unique per (series, timestamp), timestamps != 0 */
typedef struct {
    uint64_t n_rows;
    const uint64_t *series_ids;
    const int64_t *timestamps;
    const int64_t *versions;
    uint32_t n_fields;
    const bydb_wcolumn *fields;
    const char *tag_family;   /* NULL = no tag family */
    uint32_t n_tags;
    const bydb_wcolumn *tags;
    uint32_t threads;         /* 0 = hardware concurrency */
    uint32_t reserved;
} bydb_write_input;

int bydb_part_write(const bydb_write_input *in, bydb_part_image **out);

/* synthetic field kinds */
#define BYDB_SYN_F_LATENCY 1      /* float64: round(25 + N(0,5), 2)            -> decimal delta page          */
#define BYDB_SYN_F_WALK3 2        /* float64: random walk +-0.1, 3 decimals    -> decimal delta page          */
#define BYDB_SYN_F_INT1000 3      /* float64: integers 0..999                  -> delta page, exponent >= 0   */
#define BYDB_SYN_F_UNIFORM 4      /* float64: uniform [0,100) full precision   -> EncodeTypePlain fallback    */
#define BYDB_SYN_I_DELTA 10       /* int64: +rand[1,10] per step (monotone)    -> delta-of-delta page         */
#define BYDB_SYN_I_FLUCT 11       /* int64: 25 + walk of rand[-5,5]            -> delta page                  */
#define BYDB_SYN_I_RANDOM100 12   /* int64: rand[0,100)                        -> delta page                  */
#define BYDB_SYN_I_COUNTER 13     /* int64: counter with <= 2 resets           -> delta-of-delta page         */

typedef struct {
    const char *name;
    int32_t kind; /* BYDB_SYN_* */
    int32_t reserved;
} bydb_synth_field;

typedef struct {
    uint64_t n_series;
    uint64_t n_points;       /* per series */
    uint64_t sid0, sid_step; /* series ids sid0 + i*sid_step */
    int64_t t0, t_step;      /* timestamps t0 + j*t_step (regular -> DeltaConst pages) */
    uint32_t n_fields;
    const bydb_synth_field *fields;
    uint32_t region_values;  /* 0 = no tag; else string tag "default"/"region" with values "r0".."r<N-1>" */
    uint32_t region_run;     /* mean run length of equal tag values; 0 = constant per series            */
    uint32_t code_tag;       /* bit 0: also an int64 tag "default"/"code" in {0,100,..,500}; bit 1: also a second
                                string tag "default"/"zone" with values "z0".."z4" in runs of mean length 64    */
    uint32_t threads;        /* 0 = hardware concurrency                                                */
    uint64_t seed;
} bydb_synth_spec;

int bydb_synth_part(const bydb_synth_spec *spec, bydb_part_image **out);

uint32_t bydb_part_image_n_files(const bydb_part_image *p);
const char *bydb_part_image_file_name(const bydb_part_image *p, uint32_t i);
const uint8_t *bydb_part_image_file_data(const bydb_part_image *p, uint32_t i, uint64_t *len);
void bydb_part_image_counts(const bydb_part_image *p, uint64_t *total_rows, uint64_t *n_blocks);
void bydb_part_image_free(bydb_part_image *p);

#ifdef __cplusplus
}
#endif
#endif
