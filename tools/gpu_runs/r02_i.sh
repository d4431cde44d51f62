#!/bin/bash
# round 2, ninth GPU pass (1 GPU): walker-pool kernel v3 (walkers in registers, shared memory only on
# collisions) -- parity, then rates against the one-walker-per-lane kernel.
mkdir -p gpurun_out
L=$PWD/disimpy_b200
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
log=gpurun_out/kbench_r02_i.log; : > $log
for v in "" nopool p3b5 p3f8 p3f24 p3f8b5 p3a4; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 300 python tools/kbench.py sphere_t1e4 cylinder_t1e4 ellipsoid sphere180 ellipsoid180 >> $log 2>&1
done
unset DISIMPY_B200_LIB
cat $log
timeout 300 python tools/fuzz_parity.py 60 31 2>&1 | tail -1
