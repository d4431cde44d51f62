#!/bin/bash
# Mesh walk with the walkers re-sorted into cell order every S steps (DISIMPY_B200_RESORT=S, 0 = index order):
# parity with S = 8 forced on every mesh test, then kernel times for config 4 (mesh) and the config-5 shard
mkdir -p gpurun_out
log=gpurun_out/kbench_r02_resort.log; : > $log
DISIMPY_B200_RESORT=8 timeout 900 python -m pytest tests/test_gpu_reference_suite.py tests/test_gpu_parity.py -m gpu -x -q -k "mesh or neuron or fill" 2>&1 | tail -3 >> $log
for S in 0 16 32 64 128; do
  echo "== DISIMPY_B200_RESORT=$S" >> $log
  DISIMPY_B200_RESORT=$S timeout 300 python tools/kbench.py mesh config5_shard 2>&1 | grep -v "mesh:" >> $log
done
cat $log
