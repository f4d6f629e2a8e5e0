mkdir -p gpurun_out
timeout 300 python -u -m pytest tests -m gpu -x -v --timeout=120 --durations=6 ${QUICK_K:+-k "$QUICK_K"} > gpurun_out/quick_pytest.log 2>&1; tail -4 gpurun_out/quick_pytest.log
if [ -z "$QUICK_NOBENCH" ]; then
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','device_ms_per_step','scan_kernel_ms','blocks_slow_lane','slow_lane_reasons']}, d['roofline']['frac'], d['e2e'], d.get('part_admission'), d.get('fallback_field_query'))"
fi
