// Host build of the device zstd decoder (skywalking-banyandb_b200/csrc/zstd_dec.cuh) for CPU tests:
// the decoder is __host__ __device__ code, so its algorithm can be checked against libzstd-made frames
// without a GPU.  Built twice by tests/test_zstd_dec.py: as plain host code, and with -DBYDB_ZSTD_ALIGNED_IO, which
// compiles the device flavour of the memory helpers (aligned 64-bit loads + funnel shifts, word-wise copies).
// Test infrastructure only; never linked into libbydbgpu.so.
#include <cstdlib>
#include <cstring>
#include "zstd_dec.cuh"

extern "C" long long zstd_dec_host(const unsigned char *src, long long len, unsigned char *dst, long long cap) {
    // the device contract: every buffer has >= 16 readable bytes of slack on both sides
    constexpr long long kSlack = 32;
    auto *ws = static_cast<bydb::zstd::Workspace *>(std::malloc(sizeof(bydb::zstd::Workspace)));
    auto *lit = static_cast<unsigned char *>(std::calloc(131072 + 2 * kSlack, 1));
    auto *in = static_cast<unsigned char *>(std::calloc(static_cast<size_t>(len + 2 * kSlack), 1));
    auto *out = static_cast<unsigned char *>(std::calloc(static_cast<size_t>(cap + 2 * kSlack), 1));
    std::memcpy(in + kSlack, src, static_cast<size_t>(len));
    const long long r = bydb::zstd::decode_frame(ws, in + kSlack, len, out + kSlack, cap, lit + kSlack);
    bool clean = true;  // nothing may be written outside [0, cap)
    for (long long i = 0; i < kSlack; ++i) clean = clean && out[i] == 0 && out[kSlack + cap + i] == 0;
    if (r > 0) std::memcpy(dst, out + kSlack, static_cast<size_t>(r));
    std::free(ws);
    std::free(lit);
    std::free(in);
    std::free(out);
    return clean ? r : -100;
}
extern "C" int zstd_dec_workspace_bytes() { return static_cast<int>(sizeof(bydb::zstd::Workspace)); }
