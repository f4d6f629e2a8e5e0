"""A plain C99 program that uses the drop-in boundary the way a cgo shim does (tests/native/c_abi_caller.c): every entry point
of include/bydb_gpu.h is referenced through its declared prototype and linked against libbydbgpu.so; the host-only calls run;
bydb_init must refuse to work without a GPU (no CPU fallback)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_c_caller(tmp_path, bydb):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.dirname(bydb.library_path())
    exe = tmp_path / "c_abi_caller"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "native", "c_abi_caller.c"), "-L", lib_dir, "-lbydbgpu", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    return out.stdout


def test_c_caller_links_and_runs(tmp_path, bydb):
    out = _run_c_caller(tmp_path, bydb)
    import torch
    if not torch.cuda.is_available():
        assert "init refused" in out


@pytest.mark.gpu
def test_c_caller_on_the_device(tmp_path, bydb):
    """The same C99 program on a GPU box: registers a synthetic part and runs the group-by-stored-tag call from plain C."""
    assert "init ok" in _run_c_caller(tmp_path, bydb)


@pytest.mark.gpu
def test_multi_process_reduce_without_torch(tmp_path, bydb):
    """tests/native/comm_ranks.c: one process per rank, mailbox handles over a pipe, bydb_scan_reduce with rotating roots, checked
    against a single context scanning all shards.  One device: the ranks share it (CUDA IPC works within a device); more: round-robin."""
    import torch
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.dirname(bydb.library_path())
    exe = tmp_path / "comm_ranks"
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "native", "comm_ranks.c"), "-L", lib_dir, "-lbydbgpu", "-lm", "-Wl,-rpath," + lib_dir])
    ndev = max(1, torch.cuda.device_count())
    for nranks in sorted({2, min(4, max(2, ndev))}):
        out = subprocess.run([str(exe), str(nranks), str(ndev)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
