"""Where the end-to-end time of simulation() goes (development tool, run under gpurun)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get('WITH_TORCH'):
    import torch
    torch.cuda.init(); _x = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
from disimpy_b200 import gradients, simulations, substrates

n, n_t = 1_000_000, int(os.environ.get("NT", 10000))
g, dt = gradients.pgse(10e-3, 30e-3, n_t, [1e9], [[1.0, 0.0, 0.0]])
sub = substrates.sphere(10e-6)
step_l = np.sqrt(6 * 2e-9 * dt)
simulations.simulation(n, 2e-9, g, dt, sub, quiet=True)
for rep in range(2):
    T = {}
    t0 = time.perf_counter(); pos = simulations._fill_sphere(n, 10e-6, 123); T["fill_sphere"] = time.perf_counter() - t0
    t0 = time.perf_counter(); p, keep = simulations.make_params(sub, n, 0, g, dt, step_l, 123, 1000, 1e-13); T["make_params"] = time.perf_counter() - t0
    t0 = time.perf_counter(); w = simulations.Walk(p, g); T["create"] = time.perf_counter() - t0
    t0 = time.perf_counter(); w.set_positions(pos); T["set_positions"] = time.perf_counter() - t0
    t0 = time.perf_counter(); w.run(); T["launch"] = time.perf_counter() - t0
    t0 = time.perf_counter(); sig, nv = w.signal(); T["signal(sync)"] = time.perf_counter() - t0
    t0 = time.perf_counter(); w.close(); T["close"] = time.perf_counter() - t0
    t0 = time.perf_counter(); simulations.simulation(n, 2e-9, g, dt, sub, quiet=True); T["simulation() total"] = time.perf_counter() - t0
    print({k: round(v * 1e3, 2) for k, v in T.items()})
