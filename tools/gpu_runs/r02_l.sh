#!/bin/bash
# round 2, twelfth GPU pass (1 GPU): ellipsoid with precomputed reciprocals -- parity and rates; then
# the full default bench with the current build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/kbench.py ellipsoid ellipsoid180 sphere_t1e4 cylinder_t1e4 free mesh 2>&1 | grep -v "^  mesh:" | tee gpurun_out/kbench_r02_l.log
DISIMPY_B200_LOWRANK=0 KBENCH_N=1000000 timeout 300 python tools/kbench.py ellipsoid180 2>&1 | tail -1 | tee -a gpurun_out/kbench_r02_l.log
timeout 300 python tools/fuzz_parity.py 60 67 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_r02_l.json 2> gpurun_out/bench_r02_l.err; tail -c 300 gpurun_out/bench_r02_l.err; head -c 400 gpurun_out/bench_r02_l.json
