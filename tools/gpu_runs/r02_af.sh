#!/bin/bash
# compute-sanitizer over the cell-order sort (cell_order_* kernels, order indirection in the mesh walk, phases_signal_kernel)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --launch-timeout 600 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cell_order" > gpurun_out/sanitizer3_${tool}_full.txt 2>&1
  echo "== $tool rc=$?" | tee -a gpurun_out/sanitizer3_${tool}_full.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer3_${tool}_full.txt | tail -4
done
DISIMPY_B200_RESORT=5 timeout 300 python tools/fuzz_parity.py 150 307 2>&1 | tail -1 | tee gpurun_out/fuzz_r02_af.txt
