#!/bin/bash
# chunk length / occupancy of the many-measurement kernels
mkdir -p gpurun_out
L=$PWD/disimpy_b200
log=gpurun_out/kbench_r02_aa.log; : > $log
for v in "" c12b5 c12b4 c8b6 c20; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  KBENCH_N=1000000 DISIMPY_B200_LOWRANK=0 timeout 300 python tools/kbench.py sphere180 ellipsoid180 sphere8 >> $log 2>&1
done
cat $log
DISIMPY_B200_LIB=$L/libdisimpy_b200_c12b5.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fresh_inputs or 180_measurements or low_rank or chunked or partwise" 2>&1 | tail -2
