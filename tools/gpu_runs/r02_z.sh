#!/bin/bash
# round 2: compute-sanitizer over the round's new host/device paths (device list, dsb_simulate_multi, dsb_fill_mesh_multi,
# uploaded-mesh cache, refined grid with forced factors), and two more fuzz sweeps with other seeds
mkdir -p gpurun_out
SEL='device_list or simulate_multi or uploaded_mesh_cache or default_verbose or real_meshes'
for tool in memcheck racecheck; do
  DISIMPY_B200_REFINE=2,2,3 timeout 1200 compute-sanitizer --tool $tool --launch-timeout 600 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer2_${tool}_full.txt 2>&1
  echo "== $tool rc=$?" | tee -a gpurun_out/sanitizer2_${tool}_full.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer2_${tool}_full.txt | tail -4
done
timeout 300 python tools/fuzz_parity.py 200 211 2>&1 | tail -1 | tee gpurun_out/fuzz_r02_z.txt
DISIMPY_B200_REFINE=3,3,3 timeout 300 python tools/fuzz_parity.py 150 223 2>&1 | tail -1 | tee -a gpurun_out/fuzz_r02_z.txt
