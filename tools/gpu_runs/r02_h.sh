#!/bin/bash
# round 2, eighth GPU pass (1 GPU): walker-pool kernel with bounded waiting (age flush, lane serves
# the walker that is behind, 32-byte gradient records) against the one-walker-per-lane kernel.
mkdir -p gpurun_out
L=$PWD/disimpy_b200
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
log=gpurun_out/kbench_r02_h.log; : > $log
for v in "" nopool poolA2 poolA4 poolA16 poolF8A8 poolF8A4; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 300 python tools/kbench.py sphere_t1e4 cylinder_t1e4 ellipsoid sphere180 >> $log 2>&1
done
unset DISIMPY_B200_LIB
cat $log
