#!/bin/bash
# single-process device list at 8 GPUs, config-4 mesh: where the time goes (gpurun --gpus 8)
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/single_process_mesh_trace_N8.txt
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from disimpy_b200 import gradients, meshgen, simulations, substrates
v, f, pad, _ = meshgen.tube_lattice(8, 8, 5e-6, 12e-6, 40e-6, 64, 12)
mesh = substrates.mesh(v, f, True, padding=pad, init_pos="extra", n_sv=np.array([50, 50, 50]), quiet=True)
g1, dt1 = gradients.pgse(10e-3, 30e-3, 1000, [1e9], [[1.0, 0, 0]])
g, dt = gradients.pgse(10e-3, 30e-3, 10000, [1e9], [[1.0, 0, 0]])
for devs in ("0,1,2,3,4,5,6,7", "0,1,2,3", "0"):
    os.environ["DISIMPY_B200_DEVICES"] = devs
    n = 1_000_000 * len(devs.split(","))
    os.environ.pop("DISIMPY_B200_TRACE", None)
    simulations.simulation(n, 2e-9, g1, dt1, mesh, quiet=True)
    os.environ["DISIMPY_B200_TRACE"] = "1"
    for rep in range(2):
        t0 = time.perf_counter(); sig = simulations.simulation(n, 2e-9, g1, dt1, mesh, quiet=True); el = time.perf_counter() - t0
        print("mesh, devices %s: %d walkers in %.1f ms = %.3e walker-steps/s" % (devs, n, 1e3 * el, n * 1e3 / el), flush=True)
    os.environ.pop("DISIMPY_B200_TRACE", None)
    simulations.simulation(n, 2e-9, g, dt, substrates.sphere(10e-6), quiet=True)
    t0 = time.perf_counter(); sig = simulations.simulation(n, 2e-9, g, dt, substrates.sphere(10e-6), quiet=True); el = time.perf_counter() - t0
    print("sphere, devices %s: %d walkers x 1e4 steps in %.1f ms = %.3e walker-steps/s" % (devs, n, 1e3 * el, n * 1e4 / el), flush=True)
PY
