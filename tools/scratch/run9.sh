timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/e2e_breakdown.py
timeout 600 python bench.py > gpurun_out/bench_r01_c.json 2> gpurun_out/bench_r01_c.err; tail -c 600 gpurun_out/bench_r01_c.err; cat gpurun_out/bench_r01_c.json
timeout 100 python tools/kbench.py sphere_t1e4 sphere cylinder ellipsoid free mesh
