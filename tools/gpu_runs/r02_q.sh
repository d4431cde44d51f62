#!/bin/bash
# round 2, 8-GPU pass (gpurun --gpus 8): sharded == single GPU at 8 ranks, bench at N = 8 (sphere + both mesh
# configurations on every rank: config 5 is then the full 1e8-walker job), one process driving 8 GPUs.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 600 $TR tools/check_multi_gpu.py 2>&1 | grep "^ok\|Error\|error\|Traceback" | tee gpurun_out/multi_gpu_check_N${N}.txt
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_N${N}.json 2> gpurun_out/bench_N${N}.err; tail -c 300 gpurun_out/bench_N${N}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_N${N}.json").read().strip().splitlines()[-1])
print("value %.4e e2e %.4e (%.1f ms) %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["collective"]))
for m in d["mesh"]:
    print(m["config"], "%.3e" % m["value"], "%.1f ms" % m["e2e_ms"])
PY
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/single_process_N8.txt
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from disimpy_b200 import gradients, meshgen, simulations, substrates
n_dev = len(simulations.local_devices())
g, dt = gradients.pgse(10e-3, 30e-3, 10000, [1e9], [[1.0, 0, 0]])
sub = substrates.sphere(10e-6)
n = 1_000_000 * n_dev
simulations.simulation(n, 2e-9, g, dt, sub, quiet=True)
t0 = time.perf_counter(); sig = simulations.simulation(n, 2e-9, g, dt, sub, quiet=True); el = time.perf_counter() - t0
print("one process, devices %s: %d walkers x 1e4 steps in %.1f ms = %.3e walker-steps/s, signal %.6f"
      % (simulations.local_devices(n), n, 1e3 * el, n * 1e4 / el, sig[0]), flush=True)
v, f, pad, _ = meshgen.tube_lattice(8, 8, 5e-6, 12e-6, 40e-6, 64, 12)
mesh = substrates.mesh(v, f, True, padding=pad, init_pos="extra", n_sv=np.array([50, 50, 50]), quiet=True)
g1, dt1 = gradients.pgse(10e-3, 30e-3, 1000, [1e9], [[1.0, 0, 0]])
simulations.simulation(n, 2e-9, g1, dt1, mesh, quiet=True)
t0 = time.perf_counter(); sig = simulations.simulation(n, 2e-9, g1, dt1, mesh, quiet=True); el = time.perf_counter() - t0
print("one process, devices %s, config-4 mesh: %d walkers x 1e3 steps in %.1f ms = %.3e walker-steps/s, signal/n %.6f"
      % (simulations.local_devices(n), n, 1e3 * el, n * 1e3 / el, sig[0] / n), flush=True)
PY
