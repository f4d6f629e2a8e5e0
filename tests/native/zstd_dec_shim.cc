// Host build of the device zstd decoder (skywalking-banyandb_b200/csrc/zstd_dec.cuh) for CPU tests:
// the decoder is __host__ __device__ code, so its algorithm can be checked against libzstd-made frames
// without a GPU.  Test infrastructure only; never linked into libbydbgpu.so.
#include <cstdlib>
#include "zstd_dec.cuh"

extern "C" long long zstd_dec_host(const unsigned char *src, long long len, unsigned char *dst, long long cap) {
    auto *ws = static_cast<bydb::zstd::Workspace *>(std::malloc(sizeof(bydb::zstd::Workspace)));
    auto *lit = static_cast<unsigned char *>(std::malloc(131072 + 64));
    const long long r = bydb::zstd::decode_frame(ws, src, len, dst, cap, lit);
    std::free(ws);
    std::free(lit);
    return r;
}
extern "C" int zstd_dec_workspace_bytes() { return static_cast<int>(sizeof(bydb::zstd::Workspace)); }
