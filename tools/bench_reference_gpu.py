"""Time the UNMODIFIED reference (its Numba-CUDA kernels) on the GPU box, for context next to
bench.py's numbers (north_star: "reported next to the reference's Numba-CUDA path on the same
B200"; a reported baseline, not an optimisation target).

Test/measurement infrastructure only: imports the reference from oracle/_ref (git-ignored pip
install of /root/reference).  Run:  gpurun -- python tools/bench_reference_gpu.py
Writes gpurun_out/reference_numba_gpu.json.  With --json (what bench.py's `baselines` leg runs) the
progress lines go to stderr and ONE JSON line to stdout.

Two figures per workload:
  e2e          wall time of disimpy.simulations.simulation(..., quiet=True) after a JIT warm-up
  kernel_only  the reference's own step kernel driven for all time steps between CUDA events,
               without the per-step stream.synchronize() and without the mesh loop's
               time.sleep(1e-2) (simulations.py:1394)
"""
import json
import math
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from numba import cuda  # noqa: E402
from numba.cuda.random import create_xoroshiro128p_states  # noqa: E402
from disimpy import gradients, simulations as S, substrates  # noqa: E402
from disimpy_b200 import meshgen  # noqa: E402

D = 2e-9
out = {}
JSON_ONLY = "--json" in sys.argv


def say(*a):
    print(*a, file=sys.stderr if JSON_ONLY else sys.stdout, flush=True)


def pgse(n_t, n_meas=1):
    bvecs = meshgen.fibonacci_sphere(n_meas) if n_meas > 1 else [[1.0, 0, 0]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g, dt = gradients.pgse(10e-3, 30e-3, n_t, np.array([1e9] * n_meas), np.array(bvecs, dtype=float))
    return g, float(dt)


def e2e(name, sub, n, n_t, n_meas=1):
    g, dt = pgse(n_t, n_meas)
    S.simulation(min(n, 2000), D, g[:, :5], dt, sub, quiet=True)
    cuda.synchronize()
    t0 = time.time()
    S.simulation(n, D, g, dt, sub, quiet=True)
    cuda.synchronize()
    el = time.time() - t0
    out[name + "_e2e"] = dict(n_walkers=n, n_t=n_t, n_meas=n_meas, seconds=el,
                              walker_steps_per_s=n * n_t / el)
    say(name, "e2e", out[name + "_e2e"])


def kernel_only(name, sub, n, n_t, positions, n_meas=1):
    g, dt = pgse(n_t, n_meas)
    bs = 128
    gs = int(math.ceil(n / bs))
    stream = cuda.stream()
    rng = create_xoroshiro128p_states(gs * bs, seed=123, stream=stream)
    d_gx = cuda.to_device(np.ascontiguousarray(g[:, :, 0]), stream=stream)
    d_gy = cuda.to_device(np.ascontiguousarray(g[:, :, 1]), stream=stream)
    d_gz = cuda.to_device(np.ascontiguousarray(g[:, :, 2]), stream=stream)
    d_ph = cuda.to_device(np.zeros((n_meas, n)), stream=stream)
    d_exc = cuda.to_device(np.zeros(n).astype(bool))
    d_pos = cuda.to_device(positions, stream=stream)
    step_l = np.sqrt(6 * D * dt)
    if sub.type == "sphere":
        def launch(t):
            S._cuda_step_sphere[gs, bs, stream](d_pos, d_gx, d_gy, d_gz, d_ph, rng, t, step_l, dt,
                                                sub.radius, d_exc, 1000, 1e-13)
    else:
        dv = cuda.to_device(sub.vertices, stream=stream)
        df = cuda.to_device(sub.faces, stream=stream)
        dxs, dys, dzs = (cuda.to_device(a, stream=stream) for a in (sub.xs, sub.ys, sub.zs))
        dti = cuda.to_device(sub.triangle_indices, stream=stream)
        dsi = cuda.to_device(sub.subvoxel_indices, stream=stream)
        dn = cuda.to_device(sub.n_sv, stream=stream)

        def launch(t):
            S._cuda_step_mesh[gs, bs, stream](d_pos, d_gx, d_gy, d_gz, d_ph, rng, t, step_l, dt, dv,
                                              df, dxs, dys, dzs, dsi, dti, d_exc, 1000, dn, 1e-13,
                                              sub.perm_prob)
    launch(0)
    stream.synchronize()
    e0, e1 = cuda.event(), cuda.event()
    e0.record(stream)
    for t in range(1, n_t):
        launch(t)
    e1.record(stream)
    e1.synchronize()
    ms = cuda.event_elapsed_time(e0, e1)
    out[name + "_kernel_only"] = dict(n_walkers=n, n_t=n_t - 1, n_meas=n_meas, ms=ms,
                                      walker_steps_per_s=n * (n_t - 1) / (ms * 1e-3))
    say(name, "kernel-only", out[name + "_kernel_only"])


def main():
    import numba
    out["versions"] = dict(numba=numba.__version__, device=str(cuda.get_current_device().name))
    out["what"] = ("the unmodified reference (disimpy 0.3.0, its Numba-CUDA kernels) on this GPU: *_e2e = its "
                   "simulation(..., quiet=True) call; *_kernel_only = its step kernels launched for all time steps "
                   "between CUDA events, without its per-step stream.synchronize() / the mesh loop's time.sleep(1e-2)")
    sph = substrates.sphere(10e-6)
    e2e("sphere_1e6x1e4", sph, 1_000_000, 10_000)
    kernel_only("sphere_1e6x1e4", sph, 1_000_000, 10_000, S._fill_sphere(1_000_000, 10e-6))
    v, f, pad, _ = meshgen.tube_lattice(8, 8, 5e-6, 12e-6, 40e-6, 64, 12)
    t0 = time.time()
    mesh = substrates.mesh(v, f, True, padding=pad, init_pos="uniform", n_sv=np.array([50, 50, 50]),
                           quiet=True)
    out["mesh_subdivision_seconds"] = time.time() - t0
    say("reference mesh subdivision: %.1f s for %d triangles" % (time.time() - t0, len(f)))
    np.random.seed(123)
    pos = np.random.random((1_000_000, 3)) * mesh.voxel_size
    kernel_only("mesh_98k_1e6x1e3", mesh, 1_000_000, 1000, pos)
    e2e("mesh_98k_1e5x200", mesh, 100_000, 200)
    if JSON_ONLY:
        print(json.dumps(out), flush=True)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "reference_numba_gpu.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
