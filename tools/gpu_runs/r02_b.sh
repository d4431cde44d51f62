#!/bin/bash
# round 2, second GPU pass (1 GPU): new device-list tests, issue-port microbenchmark v2, ncu of the
# config-5 mesh walk at its real step length, full default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
./tools/microbench/fp64_mix 2>&1 | tee gpurun_out/fp64_mix_r02_b.txt
KBENCH_N=250000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_b_config5 python tools/kbench.py config5_shard 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; tail -c 600 gpurun_out/bench_r02_b.err; head -c 1500 gpurun_out/bench_r02_b.json
