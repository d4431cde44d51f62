#!/bin/bash
# mesh walk: list range of the start cell requested before the step is drawn, its first entries prefetched (pf1: 128 B, pf2: 256 B)
mkdir -p gpurun_out
L=$PWD/disimpy_b200
log=gpurun_out/kbench_r02_ai.log; : > $log
for v in "" _pf1 _pf2 "" _pf1; do
  export DISIMPY_B200_LIB=$L/libdisimpy_b200$v.so
  timeout 300 python tools/kbench.py mesh mesh_big config5_shard 2>&1 | grep -v "mesh:" >> $log
done
cat $log
