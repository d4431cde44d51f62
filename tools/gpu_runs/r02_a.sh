#!/bin/bash
# round 2, first GPU pass: new parity tests, live-reference A/B at full size, baseline kernel rates,
# the two-walkers-per-lane experiment, an ncu capture of the config-5 mesh walk.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/ab_live_reference.py --out gpurun_out/ab_live_reference_r02_a.json > gpurun_out/ab_live_r02_a.log 2>&1; tail -12 gpurun_out/ab_live_r02_a.log
timeout 300 python tools/kbench.py sphere_t1e4 cylinder ellipsoid free mesh mesh_big sphere180 > gpurun_out/kbench_r02_a.log 2>&1
DISIMPY_B200_LOWRANK=0 timeout 300 python tools/kbench.py sphere180 ellipsoid180 mesh180 >> gpurun_out/kbench_r02_a.log 2>&1
for v in two8 two16; do
  DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200_$v.so timeout 300 python tools/kbench.py sphere_t1e4 cylinder ellipsoid >> gpurun_out/kbench_r02_a.log 2>&1
done
cat gpurun_out/kbench_r02_a.log
DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200_two8.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or fresh or zero_normals" 2>&1 | tail -3
KBENCH_N=1000000 KBENCH_NT=200 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_a_config5 python tools/kbench.py config5_shard 2>&1 | tail -2
./tools/microbench/fp64_mix 2>&1 | tee gpurun_out/fp64_mix_r02_a.txt
