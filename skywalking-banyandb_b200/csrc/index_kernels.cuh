// index_kernels.cuh -- device-side decode of a part's block index (see index_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "part_dir.hpp"

namespace bydb {

constexpr uint32_t kIndexNameMax = 60;     // longest column name interned on the device
constexpr uint32_t kIndexMaxNames = 1024;  // distinct (kind, family, name) triples per part
constexpr uint32_t kIndexCacheSlots = 24;  // positional name cache per walker

enum IndexErr : uint32_t {
    kIdxOk = 0,
    kIdxBadMeta = 1,          // meta.bin is not a zstd frame of 40 B records / records out of order or out of primary.bin
    kIdxBadFrame = 2,         // a primary block does not inflate to its declared size
    kIdxBadBlock = 3,         // corrupt blockMetadata
    kIdxBadEnc = 4,           // unexpected timestamps / version encode type
    kIdxBadColumn = 5,        // corrupt columnMetadata / page outside its file
    kIdxFamily = 6,           // tag family without .tf / .tfm or metadata outside the .tfm
    kIdxOrder = 7,            // blockMetadata out of (series, timestamp) order
    kIdxNames = 8,            // more distinct column names than the device table holds, or a name longer than kIndexNameMax
    kIdxTooManyFamilies = 9,  // more than 16 tag families in a block
};

struct IndexName {
    uint8_t kind;   // 'f' field, 't' tag
    uint8_t fam;    // tag: index into IndexParams::families
    uint8_t len;
    uint8_t bytes[kIndexNameMax + 1];
};
static_assert(sizeof(IndexName) == 64, "IndexName layout");

struct IndexPrimary {
    uint64_t off, size;       // the frame inside primary.bin
    uint64_t raw_off, raw_len;  // its inflated bytes inside the raw arena
    uint64_t block_base, col_base;  // exclusive prefix sums of n_blocks / n_cols (filled by the host between the walks)
    uint32_t n_blocks, n_cols;
};

struct IndexFamily {
    const uint8_t *name;   // device copy of the family name
    const uint8_t *tfm;    // <family>.tfm image on the device
    uint64_t tfm_len, tf_len;
    uint32_t name_len;
    uint8_t file_id;       // slot of <family>.tf in the part's file table
    uint8_t pad[3];
};

struct IndexCtl {
    uint32_t err, err_where;
    uint32_t n_names, name_lock;
    uint32_t n_primary, max_block_rows;
    unsigned long long meta_raw, raw_total, total_rows;
    long long min_ts, max_ts;
};

struct IndexParams {
    const uint8_t *meta, *primary;
    uint64_t meta_len, primary_len, ts_len, fv_len;
    uint8_t *meta_raw;
    uint64_t meta_raw_cap;
    IndexPrimary *pb;
    uint32_t n_primary, n_families;
    uint8_t *raw;             // inflated primary blocks
    uint8_t *scratch;         // zstd workspaces: n_primary * index_scratch_stride() (>= 1 for the meta frame)
    const IndexFamily *families;
    IndexName *names;
    const uint16_t *name_map; // local name id -> the context's interned id (fill pass)
    IndexCtl *ctl;
    DevBlock *blocks;
    DevCol *cols;
    uint64_t n_blocks;
};

size_t index_ws_bytes();
size_t index_scratch_stride();
void launch_index_meta(const IndexParams &p, int phase, cudaStream_t s);
void launch_index_inflate(const IndexParams &p, cudaStream_t s);
void launch_index_walk(const IndexParams &p, bool fill, cudaStream_t s);
void launch_index_order(const IndexParams &p, cudaStream_t s);

}  // namespace bydb
