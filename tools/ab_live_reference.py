"""Same-seed A/B of this package against the LIVE, unmodified reference (its Numba-CUDA kernels)
on the GPU box, at the sizes of the BASELINE configurations (north_star: "Correctness is checked
against the reference's own Numba kernels on the same seed and inputs": final positions within
1e-9 relative -- asserted bit for bit here --, signals within 1e-6 relative, iter_exc equal).

Test infrastructure only: the reference is imported from oracle/_ref (git-ignored pip install of
/root/reference, see oracle/README.md) with the matplotlib stub of oracle/stubs.

    gpurun -- python tools/ab_live_reference.py [--quick] [--out FILE]

--quick: a tenth of the walkers (what tests/test_gpu_live_reference.py runs); default: config 2a/2b
at 1e6 walkers x 1e3 steps, config 3 (ellipsoid, 60 directions x 3 shells) at 1e5 x 1e3, config 4
(periodic tube lattice, 98 304 triangles, init_pos='extra') at 1e5 x 200, plus the reference's
neuron model ('intra', float32 vertices).  Prints one line per case and a JSON summary; exit code
1 on any mismatch.
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

D = 2e-9
SEED = 123


def run(sim, n, g, dt, sub, **kw):
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        t0 = time.time()
        sig, pos = sim(n, D, g, dt, sub, seed=SEED, final_pos=True, quiet=True, **kw)
        el = time.time() - t0
    warned = [str(x.message) for x in w if "Maximum number of iterations" in str(x.message)]
    return np.asarray(sig), np.asarray(pos), warned, el


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_live_reference.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    scale = 10 if args.quick else 1

    import numba
    from numba import cuda
    import disimpy.gradients as rg
    import disimpy.simulations as rs
    import disimpy.substrates as rsub
    import disimpy.utils as rutils
    from disimpy_b200 import gradients, meshgen, simulations, substrates, utils

    report = {"numba": numba.__version__, "device": str(cuda.get_current_device().name), "quick": args.quick,
              "cases": {}}
    dirs = meshgen.fibonacci_sphere(60)
    bvals180 = np.array([1e9] * 60 + [2e9] * 60 + [3e9] * 60)
    bvecs180 = np.vstack([dirs, dirs, dirs])

    def protocol(n_t, many):
        # both packages build their own gradient array; the reference's is the input of both runs
        # (the product's pgse/set_b are compared with it separately below)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if many:
                g, dt = rg.pgse(10e-3, 30e-3, n_t, bvals180, bvecs180)
                g2, dt2 = gradients.pgse(10e-3, 30e-3, n_t, bvals180, bvecs180)
            else:
                g, dt = rg.pgse(10e-3, 30e-3, n_t, np.array([1e9]), np.array([[1.0, 0, 0]]))
                g2, dt2 = gradients.pgse(10e-3, 30e-3, n_t, [1e9], [[1.0, 0.0, 0.0]])
        return np.ascontiguousarray(g), float(dt), bool(np.array_equal(g, g2) and dt == dt2)

    v4, f4, pad4, _ = meshgen.tube_lattice(8, 8, 5e-6, 12e-6, 40e-6, 64, 12)
    real = np.load(os.path.join(ROOT, "tests", "golden", "ref_real_meshes.npz"))
    Rell = rutils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([1.0, 1.0, 1.0]))
    assert np.array_equal(Rell, utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([1.0, 1.0, 1.0])))

    def mesh_pair(v, f, **kw):
        t0 = time.time()
        a = rsub.mesh(v, f, **kw)
        t_ref = time.time() - t0
        t0 = time.time()
        b = substrates.mesh(v, f, **kw)
        return a, b, {"reference_substrates_mesh_s": t_ref, "product_substrates_mesh_s": time.time() - t0}

    cases = {
        "config2a_sphere": lambda: (rsub.sphere(10e-6), substrates.sphere(10e-6), 1_000_000 // scale, 1000, False, {}),
        "config2b_cylinder": lambda: (rsub.cylinder(5e-6, np.array([0.0, 0.0, 1.0])),
                                      substrates.cylinder(5e-6, np.array([0.0, 0.0, 1.0])), 1_000_000 // scale, 1000, False, {}),
        "config3_ellipsoid_180": lambda: (rsub.ellipsoid(np.array([10e-6, 5e-6, 2.5e-6]), Rell),
                                          substrates.ellipsoid(np.array([10e-6, 5e-6, 2.5e-6]), Rell),
                                          100_000 // scale, 1000, True, {}),
        "config4_mesh_extra": lambda: mesh_pair(v4, f4, periodic=True, padding=pad4, init_pos="extra",
                                                n_sv=np.array([50, 50, 50]), quiet=True)[:2] + (100_000 // scale, 200, False, {}),
        "config4_mesh_extra_180": lambda: mesh_pair(v4, f4, periodic=True, padding=pad4, init_pos="extra",
                                                    n_sv=np.array([50, 50, 50]), quiet=True)[:2] + (20_000 // scale, 100, True, {}),
        "neuron_model_intra": lambda: mesh_pair(real["neuron_model_vertices"], real["neuron_model_faces"].astype(np.int64),
                                                periodic=True, init_pos="intra", quiet=True)[:2] + (20_000 // scale, 100, False, {}),
        "sphere_small_iterexc": lambda: (rsub.sphere(0.4e-6), substrates.sphere(0.4e-6), 20_000 // scale, 300, False,
                                         {"max_iter": 3}),
    }
    ok_all = True
    for name, make in cases.items():
        if args.only and args.only not in name:
            continue
        ref_sub, sub, n, n_t, many, kw = make()
        g, dt, same_g = protocol(n_t, many)
        # JIT warm-up of the reference kernel (not timed)
        rs.simulation(256, D, g[:, :3], dt, ref_sub, seed=1, quiet=True, **kw) if ref_sub.type != "mesh" else None
        r_sig, r_pos, r_warn, r_s = run(rs.simulation, n, g, dt, ref_sub, **kw)
        p_sig, p_pos, p_warn, p_s = run(simulations.simulation, n, g, dt, sub, **kw)
        pos_equal = bool(np.array_equal(r_pos, p_pos))
        denom = np.maximum(np.abs(r_pos), 1e-300)
        pos_rel = float(np.max(np.abs(r_pos - p_pos) / denom)) if not pos_equal else 0.0
        sig_rel = float(np.max(np.abs(r_sig - p_sig) / np.maximum(np.abs(r_sig), 1e-300)))
        ok = pos_equal and sig_rel <= 1e-6 and r_warn == p_warn
        ok_all &= ok
        report["cases"][name] = {
            "n_walkers": n, "n_t": n_t, "n_meas": int(g.shape[0]), "ok": ok, "positions_array_equal": pos_equal,
            "positions_max_rel_diff": pos_rel, "signals_max_rel_diff": sig_rel,
            "iter_exc_warning_equal": r_warn == p_warn, "n_flagged_warning": len(r_warn),
            "gradient_arrays_identical": same_g,
            "reference_seconds": r_s, "product_seconds": p_s,
            "reference_walker_steps_per_s": n * n_t / r_s, "product_walker_steps_per_s": n * n_t / p_s,
        }
        print("%-26s n=%-8d T=%-5d M=%-4d positions %s  signals rel %.2e  iter_exc %s  ref %.2fs  here %.3fs  %s"
              % (name, n, n_t, g.shape[0], "EQUAL" if pos_equal else "DIFFER (max rel %.2e)" % pos_rel, sig_rel,
                 "equal" if r_warn == p_warn else "DIFFER", r_s, p_s, "ok" if ok else "MISMATCH"), flush=True)
    report["ok"] = bool(ok_all)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(report, fh, indent=1)
    print(json.dumps({"ok": report["ok"], "cases": {k: v["ok"] for k, v in report["cases"].items()}}))
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
