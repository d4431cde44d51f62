#!/bin/bash
# mirror of the reference's test__cuda_random_step through one step of free diffusion
timeout 600 python -m pytest tests/test_gpu_reference_suite.py -m gpu -x -q -k "random_step" 2>&1 | tail -15
