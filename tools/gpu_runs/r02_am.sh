#!/bin/bash
# the reference's device-function unit tests on the CUDA device functions, and each of them against the oracle bit for bit
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_functions" 2>&1 | tail -15 | tee gpurun_out/device_functions_r02_am.txt
