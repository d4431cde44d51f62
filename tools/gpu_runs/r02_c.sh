#!/bin/bash
# round 2, third GPU pass (1 GPU): cp.async ring / triangle prefetch / occupancy variants of the mesh
# kernel, phase-matrix cache policy variants of the many-measurement kernels, e2e part sizes.
mkdir -p gpurun_out
L=$PWD/disimpy_b200
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
log=gpurun_out/kbench_r02_c.log; : > $log
for v in "" ring0 ring4np ring2 ring6 mesh5; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 300 python tools/kbench.py mesh mesh_big >> $log 2>&1
  KBENCH_N=1000000 timeout 300 python tools/kbench.py config5_shard >> $log 2>&1
done
for v in "" phics mr3 mr3g24; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  KBENCH_N=1000000 DISIMPY_B200_LOWRANK=0 timeout 300 python tools/kbench.py sphere180 ellipsoid180 >> $log 2>&1
  DISIMPY_B200_LOWRANK=0 timeout 300 python tools/kbench.py mesh180 >> $log 2>&1
done
unset DISIMPY_B200_LIB
grep -v "^  mesh:" $log
timeout 300 python tools/e2e_trace.py 2>&1 | tail -8
