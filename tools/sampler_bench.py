"""Where the time of a mesh simulation's set-up goes on the config-5 mesh (development tool):
handle creation (mesh upload), the initial-position sampler (dsb_fill_mesh_sim), the walk.

    python tools/sampler_bench.py            # REPS=2, NTOT=12500000,100000000

NTOT: how many points the sampler draws in total (a rank of a multi-GPU run that does not split
the sampler draws all of them and keeps its 1.25e7).  Under ncu (-k regex:fill_mesh_kernel) this
is the command behind profiles/*_fill_mesh_ncu.md.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disimpy_b200 import gradients, meshgen, simulations, substrates  # noqa: E402

v, f, pad, _ = meshgen.tube_lattice(16, 16, 5e-6, 12e-6, 40e-6, 128, 16)
sub = substrates.mesh(v, f, True, padding=pad, init_pos="extra", n_sv=np.array([100, 100, 50]), quiet=True)
dirs = meshgen.fibonacci_sphere(60)
g, dt = gradients.pgse(10e-3, 30e-3, 1000, [1e9] * 60 + [2e9] * 60 + [3e9] * 60, np.vstack([dirs, dirs, dirs]))
step_l = np.sqrt(6 * 2e-9 * dt)
n_local = 12_500_000
walk_too = os.environ.get("WALK", "1") != "0"
for rep in range(int(os.environ.get("REPS", 2))):
    for n_total in [int(x) for x in os.environ.get("NTOT", "12500000,100000000").split(",")]:
        t0 = time.perf_counter()
        p, keep = simulations.make_params(sub, n_local, 0, g, dt, step_l, 123, 1000, 1e-13)
        w = simulations.Walk(p, g)
        w.sync()
        t1 = time.perf_counter()
        w.fill_mesh(sub.voxel_size, False, 123, n_total, 0, 128)
        w.sync()
        t2 = time.perf_counter()
        if walk_too:
            w.run()
            w.signal()
        t3 = time.perf_counter()
        w.close()
        print("sampler draws %d points: create %.0f ms, fill %.0f ms, walk + signal %.0f ms"
              % (n_total, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
