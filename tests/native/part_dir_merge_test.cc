// Host-only check of the cold path's block-index slicing (skywalking-banyandb_b200/csrc/part_dir.cc): parsing a part in T pieces
// and merging consecutive pieces must give exactly the directory of a one-shot parse, for every grouping the
// dynamic slicer can form.  Built and run by tests/test_part_dir_native.py with g++ (no CUDA involved).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bydb_synth.h"
#include "part_dir.hpp"

using namespace bydb;

static int fail(const char *what) {
    std::printf("FAIL %s\n", what);
    return 1;
}

static bool same(const PartDir &a, const PartDir &b) {
    return a.blocks.size() == b.blocks.size() && a.cols.size() == b.cols.size() && a.files == b.files && a.total_rows == b.total_rows &&
           a.max_block_rows == b.max_block_rows && a.min_ts == b.min_ts && a.max_ts == b.max_ts &&
           (a.blocks.empty() || std::memcmp(a.blocks.data(), b.blocks.data(), a.blocks.size() * sizeof(DevBlock)) == 0) &&
           (a.cols.empty() || std::memcmp(a.cols.data(), b.cols.data(), a.cols.size() * sizeof(DevCol)) == 0);
}

int main() {
    bydb_synth_field f[3] = {{"latency", BYDB_SYN_F_LATENCY, 0}, {"walk", BYDB_SYN_F_WALK3, 0}, {"calls", BYDB_SYN_I_DELTA, 0}};
    bydb_synth_spec sp{};
    sp.n_series = 700;
    sp.n_points = 9000;  // two blocks per series, several primary blocks in the part
    sp.sid0 = 5;
    sp.sid_step = 3;
    sp.t0 = 1700000000000000000LL;
    sp.t_step = 60000000000LL;
    sp.n_fields = 3;
    sp.fields = f;
    sp.region_values = 6;
    sp.region_run = 9;
    sp.code_tag = 1;
    sp.seed = 77;
    bydb_part_image *img = nullptr;
    if (bydb_synth_part(&sp, &img)) return fail("synth");
    std::vector<FileImage> files;
    for (uint32_t i = 0; i < bydb_part_image_n_files(img); ++i) {
        uint64_t len;
        const uint8_t *d = bydb_part_image_file_data(img, i, &len);
        files.push_back(FileImage{bydb_part_image_file_name(img, i), d, len});
    }
    std::string err;
    size_t n_primary = 0;
    if (count_primary_blocks(files, &n_primary, err)) return fail(err.c_str());
    if (n_primary < 2) return fail("expected several primary blocks");
    NameTable names;
    PartDir whole;
    if (build_part_dir(files, names, whole, err)) return fail(err.c_str());
    if (whole.blocks.size() != 1400 || whole.total_rows != 700ull * 9000ull) return fail("whole directory shape");
    for (size_t T : {size_t{2}, n_primary, size_t{32}}) {
        std::vector<PartDir> pieces(T);
        size_t blocks = 0;
        for (size_t t = 0; t < T; ++t) {
            if (build_part_dir(files, names, pieces[t], err, t, T)) return fail(err.c_str());
            blocks += pieces[t].blocks.size();
        }
        if (blocks != whole.blocks.size()) return fail("pieces do not cover the part");
        // groupings: all at once, first piece alone + rest, pairs
        for (int mode = 0; mode < 3; ++mode) {
            std::vector<PartDir> groups;
            size_t next = 0;
            while (next < T) {
                const size_t take = mode == 0 ? T : (mode == 1 ? (next == 0 ? 1 : T - 1) : 2);
                std::vector<PartDir> g;
                for (size_t k = 0; k < take && next < T; ++k) g.push_back(pieces[next++]);
                PartDir m;
                if (merge_part_dirs(g, m, err)) return fail(err.c_str());
                groups.push_back(m);
            }
            PartDir all;
            if (merge_part_dirs(groups, all, err)) return fail(err.c_str());
            if (!same(all, whole)) return fail("merged directory differs from the one-shot parse");
        }
    }
    // an out-of-order merge must be refused (block_metadata.go:323-336)
    {
        std::vector<PartDir> two(2);
        if (build_part_dir(files, names, two[1], err, 0, 2) || build_part_dir(files, names, two[0], err, 1, 2)) return fail(err.c_str());
        PartDir m;
        if (merge_part_dirs(two, m, err) == 0) return fail("out-of-order pieces were accepted");
    }
    bydb_part_image_free(img);
    std::printf("OK primary_blocks=%zu blocks=%zu cols=%zu\n", n_primary, whole.blocks.size(), whole.cols.size());
    return 0;
}
