#!/bin/bash
# parked-bounce thresholds re-measured with the round-2 step generator
mkdir -p gpurun_out
L=$PWD/disimpy_b200
log=gpurun_out/kbench_r02_u.log; : > $log
for v in "" pe1 pe4 pe8 ps3; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 300 python tools/kbench.py ellipsoid ellipsoid180 sphere_t1e4 cylinder_t1e4 >> $log 2>&1
done
cat $log
