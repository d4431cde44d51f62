#!/bin/bash
# round 2, tenth GPU pass (1 GPU): step generator without per-root range tests (sqrt_fast, one
# special-case branch per step) -- self-test, parity, rates at three occupancy settings.
mkdir -p gpurun_out
L=$PWD/disimpy_b200
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
log=gpurun_out/kbench_r02_j.log; : > $log
for v in "" mb6 mb7; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 300 python tools/kbench.py sphere_t1e4 cylinder_t1e4 ellipsoid free mesh sphere180 >> $log 2>&1
done
unset DISIMPY_B200_LIB
grep -v "^  mesh:" $log
timeout 300 python tools/fuzz_parity.py 90 41 2>&1 | tail -1
KBENCH_NT=1000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_j_sphere python tools/kbench.py sphere 2>&1 | tail -1
