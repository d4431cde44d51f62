#!/bin/bash
# round 2, fifth GPU pass (1 GPU): refined search grid of the mesh walk -- parity with forced odd
# refinement factors, then its effect on the mesh workloads; balanced parts at 1 GPU.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MESHTESTS='mesh or golden or fresh_inputs or randomised or neuron or real_meshes or device_list or 180_measurements or fill'
DISIMPY_B200_REFINE=3,2,4 timeout 900 python -m pytest tests -m gpu -x -q -k "$MESHTESTS" 2>&1 | tail -3
DISIMPY_B200_REFINE=2,5,1 timeout 900 python -m pytest tests -m gpu -x -q -k "$MESHTESTS" 2>&1 | tail -3
log=gpurun_out/kbench_r02_e.log; : > $log
for r in "" 1 "2,2,1" "3,3,1" "2,2,2"; do
  if [ -n "$r" ]; then export DISIMPY_B200_REFINE=$r; else unset DISIMPY_B200_REFINE; fi
  echo "== DISIMPY_B200_REFINE=${r:-auto}" >> $log
  timeout 300 python tools/kbench.py mesh mesh_big >> $log 2>&1
  KBENCH_N=1000000 timeout 300 python tools/kbench.py config5_shard >> $log 2>&1
done
unset DISIMPY_B200_REFINE
grep -v "^  mesh:\|^lib" $log
timeout 600 python tools/fuzz_parity.py 120 11 2>&1 | tail -3
