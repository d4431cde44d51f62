#!/bin/bash
# round 2: long randomised parity sweeps with the final build (default and forced grid refinements)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python tools/fuzz_parity.py 300 101 2>&1 | tail -1 | tee gpurun_out/fuzz_r02_w.txt
DISIMPY_B200_REFINE=2,3,2 timeout 300 python tools/fuzz_parity.py 150 103 2>&1 | tail -1 | tee -a gpurun_out/fuzz_r02_w.txt
DISIMPY_B200_REFINE=4,1,3 timeout 300 python tools/fuzz_parity.py 150 107 2>&1 | tail -1 | tee -a gpurun_out/fuzz_r02_w.txt
timeout 300 python tools/kbench.py sphere_t1e4 cylinder_t1e4 ellipsoid free mesh 2>&1 | grep -v "^  mesh:" | tee gpurun_out/kbench_r02_w.log
