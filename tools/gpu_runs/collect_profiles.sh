#!/bin/bash
# What a round's measurements are made of (run from the repo root under gpurun, one B200):
#   gpurun --timeout 3000 -- 'bash tools/gpu_runs/collect_profiles.sh r02_final'
# then, back in the container:  python tools/summarize_profiles.py <tag> gpurun_out/launches_<tag>.csv ...
# Every command is bounded by `timeout`: a hung kernel must not hold the box.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -6
timeout 1500 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 400 gpurun_out/bench_${tag}.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_${tag}.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-mesh --no-reference-baselines > gpurun_out/ncu_bench_${tag}.log 2>&1
export KBENCH_NT=1000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_${tag}_sphere python tools/kbench.py sphere 2>&1 | tail -2
export KBENCH_N=500000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_${tag}_mesh python tools/kbench.py mesh 2>&1 | tail -2
export KBENCH_N=250000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_${tag}_config5 python tools/kbench.py config5_shard 2>&1 | tail -2
export KBENCH_NT=208 KBENCH_N=400000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_${tag}_sphere180 python tools/kbench.py sphere180 2>&1 | tail -2
# the same protocol on the general many-measurement path (tensor-core phase product)
DISIMPY_B200_LOWRANK=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_${tag}_sphere180_general python tools/kbench.py sphere180 2>&1 | tail -2
# the mesh sampler on the config-5 mesh (1.25e7 proposed points per launch)
unset KBENCH_NT KBENCH_N
REPS=1 NTOT=12500000 WALK=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fill_mesh_kernel -c 1 -f \
    -o gpurun_out/prof_${tag}_fill python tools/sampler_bench.py 2>&1 | tail -2
# BASELINE config 5 as one GPU of the 8-GPU job sees it, through simulation()
timeout 400 python tools/config5.py > gpurun_out/config5_${tag}.log 2>&1; tail -3 gpurun_out/config5_${tag}.log
timeout 600 python tools/ab_live_reference.py --out gpurun_out/ab_live_reference_${tag}.json > gpurun_out/ab_live_${tag}.log 2>&1; grep -v "^ \|warn" gpurun_out/ab_live_${tag}.log | tail -9
