#!/bin/bash
# round 2, fourteenth GPU pass (1 GPU): mesh range build over non-empty slots only -- parity and rates.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MESHTESTS='mesh or golden or fresh_inputs or randomised or neuron or real_meshes or device_list or 180_measurements'
DISIMPY_B200_REFINE=3,2,4 timeout 900 python -m pytest tests -m gpu -x -q -k "$MESHTESTS" 2>&1 | tail -2
timeout 300 python tools/kbench.py mesh mesh_big mesh180 2>&1 | grep -v "^  mesh:" | tee gpurun_out/kbench_r02_n.log
KBENCH_N=1000000 timeout 300 python tools/kbench.py config5_shard 2>&1 | tail -1 | tee -a gpurun_out/kbench_r02_n.log
timeout 300 python tools/fuzz_parity.py 60 71 2>&1 | tail -1
