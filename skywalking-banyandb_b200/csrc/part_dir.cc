// part_dir.cc -- see part_dir.hpp.
#include "part_dir.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "../../include/bydb_gpu.h"

namespace bydb {

uint16_t NameTable::intern(const std::string &s) {
    auto it = ids_.find(s);
    if (it != ids_.end()) return it->second;
    uint16_t id = static_cast<uint16_t>(ids_.size() + 1);
    ids_.emplace(s, id);
    return id;
}
uint16_t NameTable::find(const std::string &s) const {
    auto it = ids_.find(s);
    return it == ids_.end() ? 0 : it->second;
}

// ------------------------------------------------------------------ zstd (pkg/compress/zstd/zstd.go:49-52)
namespace {
using decompress_fn = size_t (*)(void *, size_t, const void *, size_t);
using iserr_fn = unsigned (*)(size_t);
using fcs_fn = unsigned long long (*)(const void *, size_t);
struct Zstd {
    decompress_fn decompress = nullptr;
    iserr_fn is_error = nullptr;
    fcs_fn content_size = nullptr;
    bool ok = false;
};
Zstd &zstd_lib() {
    static Zstd z;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        z.decompress = reinterpret_cast<decompress_fn>(dlsym(h, "ZSTD_decompress"));
        z.is_error = reinterpret_cast<iserr_fn>(dlsym(h, "ZSTD_isError"));
        z.content_size = reinterpret_cast<fcs_fn>(dlsym(h, "ZSTD_getFrameContentSize"));
        z.ok = z.decompress && z.is_error && z.content_size;
    });
    return z;
}
}  // namespace

int zstd_decompress(const uint8_t *src, size_t n, std::vector<uint8_t> &dst, std::string &err) {
    Zstd &z = zstd_lib();
    if (!z.ok) {
        err = "libzstd.so.1 not available";
        return BYDB_EIO;
    }
    unsigned long long fcs = z.content_size(src, n);
    if (fcs == static_cast<unsigned long long>(-2)) {
        err = "not a zstd frame";
        return BYDB_EINVAL;
    }
    size_t cap = fcs == static_cast<unsigned long long>(-1) ? n * 32 + 4096 : static_cast<size_t>(fcs);
    for (int attempt = 0; attempt < 8; ++attempt) {
        dst.resize(cap ? cap : 1);
        size_t r = z.decompress(dst.data(), cap, src, n);
        if (!z.is_error(r)) {
            dst.resize(r);
            return 0;
        }
        if (fcs != static_cast<unsigned long long>(-1)) break;
        cap *= 8;
    }
    err = "zstd decompress failed";
    return BYDB_EINVAL;
}

// ------------------------------------------------------------------ byte cursor
namespace {
struct Cur {
    const uint8_t *p, *end;
    bool bad = false;
    size_t left() const { return static_cast<size_t>(end - p); }
    uint64_t u64be() {
        if (left() < 8) {
            bad = true;
            return 0;
        }
        uint64_t u = 0;
        for (int k = 0; k < 8; ++k) u = (u << 8) | p[k];
        p += 8;
        return u;
    }
    uint8_t u8() {
        if (left() < 1) {
            bad = true;
            return 0;
        }
        return *p++;
    }
    // LEB128 unsigned (pkg/encoding/int.go:189-211); a truncated value marks the cursor bad
    uint64_t varu() {
        uint64_t x = 0;
        for (unsigned s = 0, i = 0; i < 10; ++i, s += 7) {
            if (p >= end) break;
            uint8_t b = *p++;
            x |= static_cast<uint64_t>(b & 0x7f) << s;
            if (b < 0x80) return x;
        }
        bad = true;
        return 0;
    }
    std::string str() {  // pkg/encoding/bytes.go:35-43 DecodeBytes
        uint64_t n = varu();
        if (bad || left() < n) {
            bad = true;
            return {};
        }
        std::string s(reinterpret_cast<const char *>(p), static_cast<size_t>(n));
        p += n;
        return s;
    }
};

const FileImage *find_file(const std::vector<FileImage> &files, const std::string &name) {
    for (const auto &f : files)
        if (f.name == name) return &f;
    return nullptr;
}
}  // namespace

int build_part_dir(const std::vector<FileImage> &files, NameTable &names, PartDir &out, std::string &err) {
    const FileImage *meta = find_file(files, "meta.bin");
    const FileImage *primary = find_file(files, "primary.bin");
    const FileImage *tsf = find_file(files, "timestamps.bin");
    const FileImage *fvf = find_file(files, "fv.bin");
    if (!meta || !primary || !tsf || !fvf) {
        err = "part needs meta.bin, primary.bin, timestamps.bin and fv.bin";
        return BYDB_ENOENT;
    }
    out.files = {"timestamps.bin", "fv.bin"};
    auto file_id = [&](const std::string &name) -> int {
        for (size_t i = 0; i < out.files.size(); ++i)
            if (out.files[i] == name) return static_cast<int>(i);
        if (!find_file(files, name) || out.files.size() >= 255) return -1;
        out.files.push_back(name);
        return static_cast<int>(out.files.size() - 1);
    };

    // meta.bin = zstd(concat primaryBlockMetadata), 40 B records (primary_metadata.go:60-83)
    std::vector<uint8_t> raw;
    int rc = zstd_decompress(meta->data, meta->len, raw, err);
    if (rc) return rc;
    if (raw.size() % 40 != 0) {
        err = "meta.bin: length is not a multiple of 40";
        return BYDB_EINVAL;
    }
    struct Pbm {
        uint64_t sid;
        int64_t mn, mx;
        uint64_t off, size;
    };
    std::vector<Pbm> pbms(raw.size() / 40);
    {
        Cur c{raw.data(), raw.data() + raw.size()};
        for (auto &p : pbms) {
            p.sid = c.u64be();
            p.mn = static_cast<int64_t>(c.u64be());
            p.mx = static_cast<int64_t>(c.u64be());
            p.off = c.u64be();
            p.size = c.u64be();
        }
        for (size_t i = 1; i < pbms.size(); ++i)
            if (pbms[i].sid < pbms[i - 1].sid) {  // primary_metadata.go:127-134
                err = "primaryBlockMetadata out of order";
                return BYDB_EINVAL;
            }
    }
    out.blocks.clear();
    out.cols.clear();
    out.total_rows = 0;
    out.max_block_rows = 0;
    out.min_ts = INT64_MAX;
    out.max_ts = INT64_MIN;
    std::vector<uint8_t> blk;
    for (const auto &pb : pbms) {
        if (pb.off + pb.size > primary->len) {
            err = "primary block outside primary.bin";
            return BYDB_EINVAL;
        }
        rc = zstd_decompress(primary->data + pb.off, static_cast<size_t>(pb.size), blk, err);
        if (rc) return rc;
        Cur c{blk.data(), blk.data() + blk.size()};
        while (c.p < c.end) {
            // blockMetadata.unmarshal, block_metadata.go:133-168 (+ timestampsMetadata :279-293)
            DevBlock b{};
            b.sid = c.u64be();
            (void)c.varu();  // uncompressedSizeBytes: accounting only
            uint64_t count = c.varu();
            b.ts_off = c.varu();
            uint64_t ts_size = c.varu();
            b.ts_min = static_cast<int64_t>(c.u64be());
            b.ts_max = static_cast<int64_t>(c.u64be());
            uint8_t enc = c.u8();
            uint64_t ver_off = c.varu();
            b.ver_first = static_cast<int64_t>(c.u64be());
            b.ver_enc = c.u8();
            if (c.bad || count == 0 || count > 0x7fffffffu || ts_size > 0xffffffffu || ver_off > ts_size ||
                b.ts_off + ts_size > tsf->len) {
                err = "corrupt blockMetadata (timestamps)";
                return BYDB_EINVAL;
            }
            if (enc < 5 || enc > 8 || b.ver_enc < 1 || b.ver_enc > 4) {  // encoding.go:87-130
                err = "unexpected timestamps encode type";
                return BYDB_EINVAL;
            }
            b.ts_enc = static_cast<uint8_t>(enc - 4);
            b.count = static_cast<uint32_t>(count);
            b.ts_size = static_cast<uint32_t>(ts_size);
            b.ver_off = static_cast<uint32_t>(ver_off);
            b.col_begin = static_cast<uint32_t>(out.cols.size());
            // tag families: name -> dataBlock into <name>.tfm (columnFamilyMetadata of this block)
            uint64_t nfam = c.varu();
            struct Fam {
                std::string name;
                uint64_t off, size;
            };
            std::vector<Fam> fams;
            for (uint64_t i = 0; i < nfam && !c.bad; ++i) {
                Fam f;
                f.name = c.str();
                f.off = c.varu();
                f.size = c.varu();
                fams.push_back(std::move(f));
            }
            // fields: columnFamilyMetadata.unmarshal, column_metadata.go:108-122
            uint64_t nf = c.varu();
            for (uint64_t i = 0; i < nf && !c.bad; ++i) {
                DevCol col{};
                std::string name = c.str();
                col.value_type = c.u8();
                col.off = c.varu();
                uint64_t size = c.varu();
                if (c.bad || size > 0xffffffffu || col.off + size > fvf->len) {
                    err = "corrupt field columnMetadata";
                    return BYDB_EINVAL;
                }
                col.size = static_cast<uint32_t>(size);
                col.name_id = names.intern("f:" + name);
                col.file_id = 1;
                out.cols.push_back(col);
            }
            if (c.bad) {
                err = "corrupt blockMetadata";
                return BYDB_EINVAL;
            }
            for (const auto &f : fams) {
                const FileImage *tfm = find_file(files, f.name + ".tfm");
                int fid = file_id(f.name + ".tf");
                if (!tfm || fid < 0 || f.off + f.size > tfm->len) {
                    err = "tag family '" + f.name + "': missing or truncated .tf/.tfm";
                    return BYDB_EINVAL;
                }
                const FileImage *tf = find_file(files, f.name + ".tf");
                Cur t{tfm->data + f.off, tfm->data + f.off + f.size};
                uint64_t nc = t.varu();
                for (uint64_t i = 0; i < nc && !t.bad; ++i) {
                    DevCol col{};
                    std::string name = t.str();
                    col.value_type = t.u8();
                    col.off = t.varu();
                    uint64_t size = t.varu();
                    if (t.bad || size > 0xffffffffu || col.off + size > tf->len) {
                        err = "corrupt tag columnMetadata";
                        return BYDB_EINVAL;
                    }
                    col.size = static_cast<uint32_t>(size);
                    col.name_id = names.intern("t:" + f.name + "/" + name);
                    col.file_id = static_cast<uint8_t>(fid);
                    out.cols.push_back(col);
                }
                if (t.bad) {
                    err = "corrupt columnFamilyMetadata";
                    return BYDB_EINVAL;
                }
            }
            size_t ncols = out.cols.size() - b.col_begin;
            if (ncols > 0xffff) {
                err = "too many columns in a block";
                return BYDB_EINVAL;
            }
            b.n_cols = static_cast<uint16_t>(ncols);
            // block_metadata.go:323-336 validateBlockMetadataOrder (also across primary blocks)
            if (!out.blocks.empty()) {
                const DevBlock &pre = out.blocks.back();
                if (b.sid < pre.sid || (b.sid == pre.sid && b.ts_min < pre.ts_min)) {
                    err = "blockMetadata out of order";
                    return BYDB_EINVAL;
                }
            }
            out.total_rows += b.count;
            out.max_block_rows = std::max(out.max_block_rows, b.count);
            out.min_ts = std::min(out.min_ts, b.ts_min);
            out.max_ts = std::max(out.max_ts, b.ts_max);
            out.blocks.push_back(b);
        }
    }
    return 0;
}

}  // namespace bydb
