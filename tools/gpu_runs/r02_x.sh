#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29579"
timeout 600 $TR tools/check_multi_gpu.py 2>&1 | grep "^ok\|Error\|error\|Traceback\|assert" | tee gpurun_out/multi_gpu_check_N${N}_d.txt
