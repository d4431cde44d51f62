"""Kernel-only timing of the walk for a few workloads (development tool, run under gpurun).

    python tools/kbench.py [case ...]      cases: sphere cylinder ellipsoid free mesh sphere180

Prints kernel ms (CUDA events inside the library) and walker-steps/s per case; set
DISIMPY_B200_LIB to compare alternative builds of the library.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disimpy_b200 import gradients, meshgen, simulations, substrates, utils  # noqa: E402


def make(case):
    n, n_t, n_meas = 1_000_000, 2000, 1
    if case == "sphere":
        sub = substrates.sphere(10e-6)
    elif case == "sphere_t1e4":
        sub, n_t = substrates.sphere(10e-6), 10000
    elif case == "sphere180":
        sub, n_meas, n_t, n = substrates.sphere(10e-6), 180, 1000, 200_000
    elif case == "ellipsoid180":
        sub = substrates.ellipsoid(np.array([10e-6, 5e-6, 2.5e-6]),
                                   utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([1.0, 1.0, 1.0])))
        n_meas, n_t, n = 180, 1000, 200_000
    elif case == "sphere8":
        sub, n_meas, n_t = substrates.sphere(10e-6), 8, 1000
    elif case == "cylinder":
        sub = substrates.cylinder(5e-6, np.array([0.0, 0.0, 1.0]))
    elif case == "cylinder_t1e4":
        sub, n_t = substrates.cylinder(5e-6, np.array([0.0, 0.0, 1.0])), 10000
    elif case == "ellipsoid":
        sub = substrates.ellipsoid(np.array([10e-6, 5e-6, 2.5e-6]),
                                   utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([1.0, 1.0, 1.0])))
    elif case == "free":
        sub = substrates.free()
    elif case in ("mesh", "mesh_small", "mesh180", "mesh_big", "mesh_big50", "mesh_coarse", "mesh_verycoarse", "config5_shard"):
        k = 2 if case == "mesh_small" else 8
        if case in ("mesh_big", "mesh_big50", "config5_shard"):  # BASELINE config 5's mesh: ~1e6 triangles
            v, f, pad, _ = meshgen.tube_lattice(16, 16, 5e-6, 12e-6, 40e-6, 128, 16)
            n_sv = np.array([50, 50, 50]) if case == "mesh_big50" else np.array([100, 100, 50])
        elif case in ("mesh_coarse", "mesh_verycoarse"):  # the config-4 mesh on coarse grids: long lists per cell
            v, f, pad, _ = meshgen.tube_lattice(k, k, 5e-6, 12e-6, 40e-6, 64, 12)
            n_sv = np.array([16, 16, 8]) if case == "mesh_coarse" else np.array([4, 4, 2])
        else:
            v, f, pad, _ = meshgen.tube_lattice(k, k, 5e-6, 12e-6, 40e-6, 64, 12)
            n_sv = np.array([50, 50, 50])
        t0 = time.time()
        sub = substrates.mesh(v, f, True, padding=pad, init_pos="uniform", n_sv=n_sv, quiet=True)
        print("  mesh: %d triangles, %d cell entries, built in %.2f s" % (len(f), len(sub.triangle_indices),
                                                                        time.time() - t0))
        n_t, n = 1000, 1_000_000
        if case == "mesh_verycoarse":
            n_t, n = 100, 100_000
        if case == "config5_shard":  # a sixth of one GPU's shard of config 5: 2e6 of 1.25e7 walkers, 180 waveforms
            n_meas, n, n_t = 180, 2_000_000, 1000
        if case == "mesh180":
            n_meas, n = 180, 200_000
    else:
        raise SystemExit("unknown case " + case)
    n_t = int(os.environ.get("KBENCH_NT", n_t))
    n = int(os.environ.get("KBENCH_N", n))
    bvecs = meshgen.fibonacci_sphere(n_meas) if n_meas > 1 else [[1.0, 0, 0]]
    g, dt = gradients.pgse(10e-3, 30e-3, n_t, [1e9] * n_meas, bvecs)
    return sub, g, float(dt), n


def main():
    cases = sys.argv[1:] or ["sphere", "cylinder", "ellipsoid", "free"]
    print("lib:", os.environ.get("DISIMPY_B200_LIB", "default"))
    for case in cases:
        sub, g, dt, n = make(case)
        step_l = np.sqrt(6 * 2e-9 * dt)
        np.random.seed(123)
        if sub.type == "sphere":
            pos = simulations._fill_sphere(n, sub.radius, 123)
        elif sub.type == "cylinder":
            pos = simulations._initial_positions_cylinder(n, sub.radius, np.eye(3), 123)
        elif sub.type == "ellipsoid":
            pos = simulations._initial_positions_ellipsoid(n, sub.semiaxes, sub.R, 123)
        elif sub.type == "mesh":
            pos = np.random.random((n, 3)) * sub.voxel_size
        else:
            pos = np.zeros((n, 3))
        p, keep = simulations.make_params(sub, n, 0, g, dt, step_l, 123, 1000, 1e-13)
        walk = simulations.Walk(p, g)
        best = None
        for rep in range(3):
            walk.set_positions(pos)
            walk.run()
            sig, n_valid = walk.signal()
            ms, nl = walk.run_stats()
            best = ms if best is None else min(best, ms)
        print("%-12s n=%d T=%d M=%d  kernel %.2f ms  %.3e walker-steps/s  signal[0]=%.10f valid=%d"
              % (case, n, g.shape[1], g.shape[0], best, n * g.shape[1] / (best * 1e-3), sig[0], n_valid),
              flush=True)
        walk.close()


if __name__ == "__main__":
    main()
