#!/bin/bash
mkdir -p gpurun_out
L=$PWD/disimpy_b200
log=gpurun_out/kbench_r02_v.log; : > $log
for v in pc2 pc3 pc4 pc5; do
  export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so
  timeout 300 python tools/kbench.py cylinder_t1e4 cylinder >> $log 2>&1
done
cat $log
