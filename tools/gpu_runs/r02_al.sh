#!/bin/bash
# mesh walk, short-step variant: slots of the 2x2x2 nest in which no lane has a list skipped by the whole warp (sv1) against the default build
mkdir -p gpurun_out
L=$PWD/disimpy_b200
log=gpurun_out/kbench_r02_al.log; : > $log
for v in "" _sv1 "" _sv1; do
  export DISIMPY_B200_LIB=$L/libdisimpy_b200$v.so
  timeout 300 python tools/kbench.py mesh mesh_big config5_shard 2>&1 | grep -v "mesh:" >> $log
done
cat $log
