"""The device zstd decoder (skywalking-banyandb_b200/csrc/zstd_dec.cuh) is __host__ __device__ code: its algorithm is
checked here on the CPU against frames made by the system libzstd (the reference compresses with klauspost/compress
zstd level 1, pkg/compress/zstd/zstd.go; any RFC 8878 encoder produces frames this decoder must read).  The GPU tests
(test_gpu_parity.py::test_fallback_*) then run the same code on the device through the part-admission unpack."""
import ctypes as C
import os
import random
import shutil
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["host-io", "device-io"])
def shim(tmp_path_factory, request):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = tmp_path_factory.mktemp("zstd") / ("zstd_shim_%s.so" % request.param)
    flags = ["-DBYDB_ZSTD_ALIGNED_IO"] if request.param == "device-io" else []
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", *flags, "-I", os.path.join(ROOT, "skywalking-banyandb_b200", "csrc"),
                           "-o", str(out), os.path.join(ROOT, "tests", "native", "zstd_dec_shim.cc")])
    lib = C.CDLL(str(out))
    lib.zstd_dec_host.restype = C.c_longlong
    lib.zstd_dec_host.argtypes = [C.c_char_p, C.c_longlong, C.c_void_p, C.c_longlong]
    return lib


@pytest.fixture(scope="module")
def libzstd():
    try:
        z = C.CDLL("libzstd.so.1")
    except OSError:
        pytest.skip("no libzstd")
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    return z


def _compress(z, data, level):
    cap = z.ZSTD_compressBound(len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), level)
    return buf.raw[:n]


def _decode(shim, frame, cap):
    buf = C.create_string_buffer(max(cap, 1))
    r = shim.zstd_dec_host(frame, len(frame), buf, cap)
    return r, buf.raw[:max(r, 0)]


def _corpus():
    rng = random.Random(1)
    yield b""
    yield b"a"
    yield b"a" * 1000
    yield bytes(range(256)) * 10                                     # raw literals
    yield os.urandom(5000)                                           # incompressible: raw block
    yield b"".join(b"series-%d|region-r%d|" % (rng.randrange(50), rng.randrange(8)) for _ in range(3000))
    yield bytes(rng.choice(b"abcdefgh") for _ in range(20000))        # Huffman, direct weights
    yield bytes(min(255, int(rng.expovariate(0.05))) for _ in range(70000))   # Huffman, FSE-compressed weights
    yield b"".join(int(rng.gauss(0, 1e6)).to_bytes(8, "big", signed=True) for _ in range(8193))   # a numeric cell block
    yield b"".join(struct.pack(">d", rng.random()) for _ in range(8193))
    yield bytes([9]) * 8193                                          # a lengths block without nulls: RLE
    yield os.urandom(100) + b"x" * 300000 + os.urandom(100)          # several blocks, long matches across them
    yield bytes(rng.choice(b"ab") for _ in range(200000))
    for n in (127, 128, 129, 255, 256, 1023, 4096, 131071, 131072, 131073):
        yield bytes(rng.randrange(0, 4) for _ in range(n))


def test_decoder_reads_libzstd_frames(shim, libzstd):
    for data in _corpus():
        for level in (1, 3, 19, -1):
            frame = _compress(libzstd, data, level)
            r, out = _decode(shim, frame, len(data))
            assert r == len(data) and out == data, (len(data), level, r)


def test_decoder_rejects_damage_without_crashing(shim, libzstd):
    rng = random.Random(7)
    data = b"".join(b"series-%d|region-r%d|" % (rng.randrange(50), rng.randrange(8)) for _ in range(800))
    frame = _compress(libzstd, data, 1)
    r, _ = _decode(shim, frame, len(data) - 1)
    assert r < 0                                                     # destination too small
    assert _decode(shim, frame[:-3], len(data))[0] < 0               # truncated
    assert _decode(shim, b"\x00" * 16, 64)[0] < 0                    # not a frame
    for _ in range(2000):                                            # random damage: any answer, no crash, no overrun
        g = bytearray(frame)
        for _k in range(rng.randrange(1, 4)):
            g[rng.randrange(len(g))] ^= 1 << rng.randrange(8)
        if rng.random() < 0.3:
            g = g[:rng.randrange(1, len(g))]
        cap = len(data)
        buf = C.create_string_buffer(cap + 64)
        C.memset(C.addressof(buf) + cap, 0xA5, 64)
        r = shim.zstd_dec_host(bytes(g), len(g), buf, cap)
        assert r <= cap
        assert buf.raw[cap:] == b"\xa5" * 64
