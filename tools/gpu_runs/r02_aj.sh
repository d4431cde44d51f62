#!/bin/bash
# final build against the live reference (Numba-CUDA on this GPU) at BASELINE sizes, mesh cases forced into cell order
mkdir -p gpurun_out
DISIMPY_B200_RESORT=100000 timeout 800 python tools/ab_live_reference.py --out gpurun_out/ab_live_reference_r02_aj.json > gpurun_out/ab_live_r02_aj.log 2>&1
grep -v "^ \|warn" gpurun_out/ab_live_r02_aj.log | cut -c1-200 | tail -10
