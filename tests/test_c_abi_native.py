"""A plain C99 program that uses the drop-in boundary the way a cgo shim does (tests/native/c_abi_caller.c): every entry point
of include/bydb_gpu.h is referenced through its declared prototype and linked against libbydbgpu.so; the host-only calls run;
bydb_init must refuse to work without a GPU (no CPU fallback)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_caller_links_and_runs(tmp_path, bydb):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.dirname(bydb.library_path())
    exe = tmp_path / "c_abi_caller"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "native", "c_abi_caller.c"), "-L", lib_dir, "-lbydbgpu", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    import torch
    if not torch.cuda.is_available():
        assert "init refused" in out.stdout
