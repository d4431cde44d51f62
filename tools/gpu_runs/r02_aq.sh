#!/bin/bash
# the walker order kept across dsb_run calls: mesh tests, and the config-5 shard kernel time (unchanged: ~214 ms)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py -m gpu -x -q -k "cell_order or mesh or verbose" 2>&1 | tail -3
timeout 300 python tools/kbench.py config5_shard 2>&1 | grep -v "mesh:" | tee gpurun_out/kbench_r02_aq.log
