timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','device_ms_per_step','scan_kernel_ms','blocks_slow_lane','slow_lane_reasons']}, d['roofline']['frac'], d['e2e'])"
