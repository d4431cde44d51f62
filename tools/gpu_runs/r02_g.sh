#!/bin/bash
# round 2, seventh GPU pass (1 GPU): walker-pool kernel -- parity, then rates against the
# one-walker-per-lane kernel and flush-threshold / occupancy variants.
mkdir -p gpurun_out
L=$PWD/disimpy_b200
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
log=gpurun_out/kbench_r02_g.log; : > $log
for v in "" nopool pool8 pool24 pool32 pool16b7 pool16b5; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 300 python tools/kbench.py sphere_t1e4 cylinder ellipsoid >> $log 2>&1
done
unset DISIMPY_B200_LIB
cat $log
timeout 300 python tools/kbench.py sphere180 ellipsoid180 2>&1 | tail -2
timeout 600 python tools/fuzz_parity.py 100 23 2>&1 | tail -2
