timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
WITH_TORCH=1 timeout 100 python tools/e2e_breakdown.py
timeout 100 python tools/e2e_breakdown.py
timeout 600 python bench.py --no-secondary > gpurun_out/bench_r01_e.json 2> gpurun_out/bench_r01_e.err; tail -c 600 gpurun_out/bench_r01_e.err; cat gpurun_out/bench_r01_e.json
