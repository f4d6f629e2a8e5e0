"""Host logic of the cold path (no GPU): the block index of a part parsed in pieces and merged (part_dir.cc:
count_primary_blocks / build_part_dir(batch, n) / merge_part_dirs) equals the one-shot parse.  The C++ check lives in
tests/native/part_dir_merge_test.cc and is compiled here with g++ against the product sources (no CUDA)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "skywalking-banyandb_b200", "csrc")


def test_sliced_block_index_equals_one_shot_parse(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = tmp_path / "part_dir_merge_test"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "native", "part_dir_merge_test.cc"),
                           os.path.join(CSRC, "part_dir.cc"), os.path.join(CSRC, "part_writer.cc"), "-ldl", "-lpthread"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
