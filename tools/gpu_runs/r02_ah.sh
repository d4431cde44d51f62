#!/bin/bash
# final build: ncu launch list of the bench command and ncu --set full of the sphere walk (packed logs)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_ah.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-mesh --no-reference-baselines > gpurun_out/ncu_bench_r02_ah.log 2>&1
export KBENCH_NT=1000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_ah_sphere python tools/kbench.py sphere 2>&1 | tail -2
