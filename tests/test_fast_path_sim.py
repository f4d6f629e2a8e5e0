"""tools/sim_fast_path.py simulates the chunk-level algorithm of scan_kernels.cu::delta_page_fast in plain Python (32 lanes x 32
bytes, validity windows of an unaligned page, the narrow-varint check, independent lane decode + head correction, the two
warp scans, carries across chunks, and the early-stop experiment) against a straightforward decode of the same page."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_path_algorithm_simulation():
    spec = importlib.util.spec_from_file_location("sim_fast_path", os.path.join(ROOT, "tools", "sim_fast_path.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
