#!/bin/bash
# Mesh walk, walkers sorted into cell order once per run (DISIMPY_B200_RESORT >= n_t) or every S steps
mkdir -p gpurun_out
log=gpurun_out/kbench_r02_resort2.log; : > $log
for S in 0 100000 500 250 0 100000; do
  echo "== DISIMPY_B200_RESORT=$S" >> $log
  DISIMPY_B200_RESORT=$S timeout 300 python tools/kbench.py mesh mesh_big config5_shard 2>&1 | grep -v "mesh:" >> $log
done
cat $log
