#!/bin/bash
# round 2, multi-GPU pass (gpurun --gpus 2): sharded simulation() == single GPU, where the e2e time
# goes at N = 2, bench at N = 2 with the mesh configurations on every rank, one process driving 2 GPUs.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
timeout 600 $TR tools/check_multi_gpu.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W1\|^Setting" | tee gpurun_out/multi_gpu_check_N$N.txt
timeout 300 $TR tools/e2e_trace.py 2>&1 | grep "rank\|trace" | tee gpurun_out/e2e_trace_N$N.txt
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_N$N.json 2> gpurun_out/bench_N$N.err; tail -c 300 gpurun_out/bench_N$N.err; head -c 600 gpurun_out/bench_N$N.json
# a plain script: one process, all visible GPUs
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/single_process_N$N.txt
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from disimpy_b200 import gradients, simulations, substrates
n_dev = len(simulations.local_devices())
g, dt = gradients.pgse(10e-3, 30e-3, 10000, [1e9], [[1.0, 0, 0]])
sub = substrates.sphere(10e-6)
n = 1_000_000 * n_dev
for devs in (None, "0"):
    if devs: os.environ["DISIMPY_B200_DEVICES"] = devs
    simulations.simulation(n, 2e-9, g, dt, sub, quiet=True)
    t0 = time.perf_counter(); sig = simulations.simulation(n, 2e-9, g, dt, sub, quiet=True); el = time.perf_counter() - t0
    print("one process, devices %s: %d walkers x 1e4 steps in %.1f ms = %.3e walker-steps/s, signal %.6f"
          % (simulations.local_devices(n), n, 1e3 * el, n * 1e4 / el, sig[0]), flush=True)
PY
