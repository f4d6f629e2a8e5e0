"""The C++ host side above the C ABI (include/bydb_operator.hpp: the reference's PullOperator / BatchSchema / AggSpec / Top /
Limit surface, header-only over bydb_gpu.h) driven by tests/native/operator_test.cc.  Without a GPU the program checks the
schema typing and the error contract; with one (pytest -m gpu) it runs grouped aggregations against a synthetic part."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, bydb):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    lib_dir = os.path.dirname(bydb.library_path())
    exe = tmp_path / "operator_test"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "native", "operator_test.cc"), "-L", lib_dir, "-lbydbgpu", "-Wl,-rpath," + lib_dir])
    return exe


def test_operator_contract_without_a_device(tmp_path, bydb):
    out = subprocess.run([str(_build(tmp_path, bydb))], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and ("OK host-only" in out.stdout or "OK full" in out.stdout), out.stdout + out.stderr


@pytest.mark.gpu
def test_operator_full_flow_on_the_device(tmp_path, bydb):
    out = subprocess.run([str(_build(tmp_path, bydb))], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK full" in out.stdout, out.stdout + out.stderr
