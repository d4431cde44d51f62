#!/bin/bash
# compute-sanitizer over the kernels with hand-managed shared memory: the mesh search
# scratch (union range_box / hit, start bits, cp.async ring), the sampler's queue, the mbarrier
# double buffer of the many-measurement kernels.  Summaries -> gpurun_out/sanitizer_*.txt
mkdir -p gpurun_out
SEL='mesh_search_paths or fill_mesh_matches or fill_mesh_column or (fresh_inputs and (7 or 40)) or 180_measurements or chunked_run'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --launch-timeout 600 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_${tool}_full.txt 2>&1
  echo "== $tool rc=$?" | tee -a gpurun_out/sanitizer_${tool}_full.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_${tool}_full.txt | tail -8
done
