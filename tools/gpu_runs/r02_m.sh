#!/bin/bash
# round 2, thirteenth GPU pass (gpurun --gpus 2): where the mesh e2e time goes, then the 2-GPU bench
# with balanced parts, trace at N = 2.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573"
DISIMPY_B200_DEVICE=0 timeout 300 python tools/e2e_mesh.py 2>&1 | tail -4 | tee gpurun_out/e2e_mesh_r02_m.txt
DISIMPY_B200_DEVICE=0 DISIMPY_B200_TRACE=1 timeout 300 python tools/e2e_mesh.py 2>&1 | grep trace | tail -2 | tee -a gpurun_out/e2e_mesh_r02_m.txt
timeout 300 $TR tools/e2e_trace.py 2>&1 | grep "rank\|trace" | tee gpurun_out/e2e_trace_N${N}_balanced.txt
timeout 600 $TR tools/check_multi_gpu.py 2>&1 | grep "^ok\|Error\|error" | tee gpurun_out/multi_gpu_check_N${N}_b.txt
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_N${N}_b.json 2> gpurun_out/bench_N${N}_b.err; tail -c 300 gpurun_out/bench_N${N}_b.err; head -c 300 gpurun_out/bench_N${N}_b.json
