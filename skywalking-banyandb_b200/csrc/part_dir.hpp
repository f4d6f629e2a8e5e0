// part_dir.hpp -- host-side block directory of one measure part.
//
// Built once per part at bydb_part_register time from the part's index files; it is the GPU path's
// counterpart of the reference's cached blockMetadataArray (banyand/measure/part_iter.go:184-208,
// block_metadata.go:113-168, column_metadata.go:47-122, primary_metadata.go:60-137).  The page
// payloads themselves (timestamps.bin, fv.bin, *.tf) are never parsed on the host: they go to HBM
// verbatim and are decoded by the scan kernel.
#pragma once

#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace bydb {

// Device-visible descriptors (plain structs shared with the kernels).
struct DevBlock {
    uint64_t sid;
    int64_t ts_min, ts_max, ver_first;
    uint64_t ts_off;     // offset of the timestamps page in timestamps.bin
    uint32_t ts_size;    // whole page (timestamps body + versions body)
    uint32_t ver_off;    // timestampsMetadata.versionOffset
    uint32_t count;      // rows
    uint32_t col_begin;  // first DevCol of this block
    uint16_t n_cols;
    uint8_t ts_enc;      // common type 1..4 (WithVersion stripped)
    uint8_t ver_enc;
    uint32_t pad;
};
static_assert(sizeof(DevBlock) == 64, "DevBlock layout");

struct DevCol {
    uint64_t off;        // offset of the page inside its file
    uint32_t size;
    uint16_t name_id;    // interned "f:<field>" or "t:<family>/<tag>" (per context)
    uint8_t value_type;  // pkg/pb/v1/value.go:39-47
    uint8_t file_id;     // index into the part's file table
};
static_assert(sizeof(DevCol) == 16, "DevCol layout");

class NameTable {  // thread-safe: block-index slices of one part are parsed concurrently on the cold path
  public:
    // returns a stable id (>=1); 0 is "unknown"
    uint16_t intern(const std::string &s);
    uint16_t find(const std::string &s) const;
  private:
    mutable std::mutex mu_;
    std::unordered_map<std::string, uint16_t> ids_;
};

struct FileImage {
    std::string name;
    const uint8_t *data;
    uint64_t len;
};

struct PartDir {
    std::vector<DevBlock> blocks;   // ordered (sid, ts_min) as in the part
    std::vector<DevCol> cols;
    std::vector<std::string> files; // file table: [0]=timestamps.bin, [1]=fv.bin, then <family>.tf
    uint64_t total_rows = 0;
    uint32_t max_block_rows = 0;
    int64_t min_ts = 0, max_ts = 0;
};

// Parses meta.bin / primary.bin / *.tfm.  Returns 0 or a negative BYDB_* code; err gets a message.
// batch / n_batches: parse only that slice of the primary blocks (the cold host path pipelines the parsing of one
// slice with the scan of the previous one); the default parses the whole part.
int build_part_dir(const std::vector<FileImage> &files, NameTable &names, PartDir &out, std::string &err, size_t batch = 0, size_t n_batches = 1);

// number of primary blocks (independent zstd frames of primary.bin) of the part
int count_primary_blocks(const std::vector<FileImage> &files, size_t *n, std::string &err);
// concatenates consecutive slices of one part's directory (file ids remapped onto one table)
int merge_part_dirs(std::vector<PartDir> &pieces, PartDir &out, std::string &err);

// zstd frame decompression through the system libzstd (dlopen, no header in the image).
int zstd_decompress(const uint8_t *src, size_t n, std::vector<uint8_t> &dst, std::string &err);

}  // namespace bydb
