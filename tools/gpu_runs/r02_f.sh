#!/bin/bash
# round 2, sixth GPU pass (1 GPU): many-measurement kernel variants (B fragments in registers, A from
# global), ncu of the refined-grid mesh walk (configs 4 and 5), compute-sanitizer.
mkdir -p gpurun_out
L=$PWD/disimpy_b200
log=gpurun_out/kbench_r02_f.log; : > $log
for v in "" bregs aglobal bregs_aglobal bregs_g16; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$L/libdisimpy_b200_$v.so; else unset DISIMPY_B200_LIB; fi
  KBENCH_N=1000000 DISIMPY_B200_LOWRANK=0 timeout 300 python tools/kbench.py sphere180 ellipsoid180 >> $log 2>&1
  DISIMPY_B200_LOWRANK=0 timeout 300 python tools/kbench.py mesh180 sphere8 >> $log 2>&1
done
unset DISIMPY_B200_LIB
grep -v "^  mesh:" $log
DISIMPY_B200_LIB=$L/libdisimpy_b200_bregs.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fresh_inputs or 180_measurements or low_rank or chunked or partwise or randomised" 2>&1 | tail -3
KBENCH_N=500000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_f_mesh python tools/kbench.py mesh 2>&1 | tail -2
KBENCH_N=250000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_f_config5 python tools/kbench.py config5_shard 2>&1 | tail -2
bash tools/gpu_runs/r02_sanitizer.sh
