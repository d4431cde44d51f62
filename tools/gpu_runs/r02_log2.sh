#!/bin/bash
# A/B: logs of the Box-Muller pair as packed f32x2 (DSB_LOG2=1) against scalar; parity first, then kernel times
mkdir -p gpurun_out
L=$PWD/disimpy_b200
log=gpurun_out/kbench_r02_log2.log; : > $log
export DISIMPY_B200_LIB=$L/libdisimpy_b200_log2.so
timeout 900 python -m pytest tests/test_gpu_reference_suite.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 >> $log
for v in log0 log2 log0 log2; do
  export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so
  echo "== $v" >> $log
  timeout 300 python tools/kbench.py sphere_t1e4 cylinder_t1e4 ellipsoid free >> $log 2>&1
done
cat $log
