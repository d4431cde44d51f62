"""End-to-end time of simulation() on the config-4 mesh and where it goes (development tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disimpy_b200 import gradients, meshgen, simulations, substrates

n = 1_000_000
v, f, pad, _ = meshgen.tube_lattice(8, 8, 5e-6, 12e-6, 40e-6, 64, 12)
t0 = time.perf_counter()
sub = substrates.mesh(v, f, True, padding=pad, init_pos="extra", n_sv=np.array([50, 50, 50]), quiet=True)
print("substrates.mesh: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
g, dt = gradients.pgse(10e-3, 30e-3, 1000, [1e9], [[1.0, 0, 0]])
simulations.simulation(1000, 2e-9, g, dt, sub, quiet=True)
for rep in range(2):
    t0 = time.perf_counter(); pos = simulations._fill_mesh(n, sub, False, 123); t_fill = time.perf_counter() - t0
    step_l = np.sqrt(6 * 2e-9 * dt)
    t0 = time.perf_counter(); p, keep = simulations.make_params(sub, n, 0, g, dt, step_l, 123, 1000, 1e-13); w = simulations.Walk(p, g); t_create = time.perf_counter() - t0
    t0 = time.perf_counter(); w.set_positions(pos); w.run(); sig, nv = w.signal(); t_walk = time.perf_counter() - t0
    w.close()
    t0 = time.perf_counter(); simulations.simulation(n, 2e-9, g, dt, sub, quiet=True); t_all = time.perf_counter() - t0
    print("fill_mesh %.1f ms, create (mesh upload) %.1f ms, walk %.1f ms, simulation() total %.1f ms"
          % (1e3 * t_fill, 1e3 * t_create, 1e3 * t_walk, 1e3 * t_all))
