"""Generate golden input/output vectors from the UNMODIFIED reference running its
own Numba-CUDA kernels on a real GPU (the only valid 1e-9 oracle).

Test infrastructure only.  Run on the GPU box:

    gpurun -- python tools/gen_golden_gpu.py

The reference is imported from ``oracle/_ref`` (git-ignored ``pip install
--target`` of /root/reference; see oracle/README.md) with the matplotlib stub in
``oracle/stubs``.  Output goes to ``gpurun_out/golden/*.npz`` + ``info.json``;
the files are then copied to ``tests/golden/`` and committed.
"""
import json
import os
import sys
import time
import traceback
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)

info = {"cases": {}, "errors": {}, "timing": {}}


def save_info():
    with open(os.path.join(OUT, "info.json"), "w") as f:
        json.dump(info, f, indent=1, default=str)


def run_case(name, fn):
    t0 = time.time()
    try:
        fn()
        info["cases"][name] = round(time.time() - t0, 3)
    except Exception:
        info["errors"][name] = traceback.format_exc()
        print("FAILED", name, info["errors"][name], flush=True)
    save_info()


def main():
    import numba
    import llvmlite
    from numba import cuda

    info["numba"] = numba.__version__
    info["llvmlite"] = llvmlite.__version__
    info["numpy"] = np.__version__
    dev = cuda.get_current_device()
    info["device"] = dev.name.decode() if isinstance(dev.name, bytes) else str(dev.name)
    info["cc"] = list(dev.compute_capability)
    try:
        info["nvvm"] = str(cuda.cudadrv.nvvm.NVVM().get_version())
        info["runtime"] = str(cuda.runtime.get_version())
        info["driver"] = str(cuda.cudadrv.driver.driver.get_version())
    except Exception as e:  # noqa
        info["version_error"] = repr(e)
    save_info()

    import disimpy
    from disimpy import gradients, simulations, substrates, utils
    from numba.cuda.random import (create_xoroshiro128p_states,
                                   xoroshiro128p_normal_float64,
                                   xoroshiro128p_uniform_float64)
    from disimpy_b200 import meshgen

    D = 2e-9

    # ---------------------------------------------------------------- RNG
    def case_rng():
        @cuda.jit
        def draw(states, normals, uniforms):
            i = cuda.grid(1)
            if i < normals.shape[0]:
                for k in range(normals.shape[1]):
                    normals[i, k] = xoroshiro128p_normal_float64(states, i)
                for k in range(uniforms.shape[1]):
                    uniforms[i, k] = xoroshiro128p_uniform_float64(states, i)

        out = {}
        for seed in (0, 123, 2**31 + 12345):
            st = create_xoroshiro128p_states(128, seed=seed)
            h = st.copy_to_host()
            out["states_seed%d_s0" % seed] = h["s0"].copy()
            out["states_seed%d_s1" % seed] = h["s1"].copy()
            normals = np.zeros((128, 24))
            uniforms = np.zeros((128, 4))
            d_n = cuda.to_device(normals)
            d_u = cuda.to_device(uniforms)
            draw[1, 128](st, d_n, d_u)
            out["normals_seed%d" % seed] = d_n.copy_to_host()
            out["uniforms_seed%d" % seed] = d_u.copy_to_host()
            h2 = st.copy_to_host()
            out["after_seed%d_s0" % seed] = h2["s0"].copy()
            out["after_seed%d_s1" % seed] = h2["s1"].copy()
        # a far-away subsequence start (what a shard of a multi-GPU run uses)
        st = create_xoroshiro128p_states(8, seed=123, subsequence_start=1000003)
        h = st.copy_to_host()
        out["states_seed123_off1000003_s0"] = h["s0"].copy()
        out["states_seed123_off1000003_s1"] = h["s1"].copy()
        np.savez_compressed(os.path.join(OUT, "rng.npz"), **out)

    run_case("rng", case_rng)

    # ------------------------------------------------------- simulations
    def pgse(n_t, bvals, bvecs, delta=10e-3, DELTA=30e-3):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g, dt = gradients.pgse(delta, DELTA, n_t, np.asarray(bvals, float),
                                   np.asarray(bvecs, float))
        return g, float(dt)

    def sim_case(name, substrate, n_walkers, n_t, bvals, bvecs, seed=123, traj=False,
                 extra=None, **kw):
        def fn():
            g, dt = pgse(n_t, bvals, bvecs)
            tp = os.path.join(OUT, name + ".traj.txt") if traj else None
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                sig, pos = simulations.simulation(
                    n_walkers, D, g, dt, substrate, seed=seed, traj=tp, final_pos=True,
                    quiet=True, **kw)
                warned = [str(x.message) for x in w
                          if "Maximum number of iterations" in str(x.message)]
            allsig = simulations.simulation(
                n_walkers, D, g, dt, substrate, seed=seed, all_signals=True,
                quiet=True, **kw)
            out = dict(gradient=g, dt=dt, diffusivity=D, seed=seed, n_walkers=n_walkers,
                       signals=sig, positions=pos, all_signals=allsig,
                       iter_exc_warning=np.array(warned))
            for k, v in kw.items():
                out["kw_" + k] = v
            if extra:
                out.update(extra)
            if traj:
                out["traj"] = np.loadtxt(tp).reshape(n_t + 1, n_walkers, 3)
                os.remove(tp)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        run_case(name, fn)

    b3 = [1e9, 2e9, 0.5e9]
    v3 = [[1.0, 0, 0], [0, 1.0, 0], [1.0, 1.0, 1.0]]
    v3 = [list(np.array(v) / np.linalg.norm(v)) for v in v3]

    sim_case("free_traj", substrates.free(), 16, 64, b3, v3, traj=True)
    sim_case("free", substrates.free(), 1000, 200, b3, v3)
    sim_case("sphere_traj", substrates.sphere(2e-6), 16, 100, b3, v3, traj=True)
    sim_case("sphere", substrates.sphere(10e-6), 1024, 1000, b3, v3)
    sim_case("sphere_small", substrates.sphere(1e-6), 512, 300, b3, v3)
    sim_case("sphere_long", substrates.sphere(5e-6), 128, 10000, [1e9], [[1.0, 0, 0]])
    sim_case("sphere_iterexc", substrates.sphere(0.4e-6), 256, 50, b3, v3, max_iter=3)
    ori = np.array([1.0, 2.0, 3.0])
    sim_case("cylinder_traj", substrates.cylinder(2e-6, ori), 16, 100, b3, v3, traj=True,
             extra=dict(radius=2e-6, orientation=ori))
    sim_case("cylinder", substrates.cylinder(5e-6, ori), 1024, 1000, b3, v3,
             extra=dict(radius=5e-6, orientation=ori))
    oz = np.array([0.0, 0.0, 1.0])
    sim_case("cylinder_z", substrates.cylinder(5e-6, oz), 512, 500, b3, v3,
             extra=dict(radius=5e-6, orientation=oz))
    ox = np.array([1.0, 0.0, 0.0])
    sim_case("cylinder_x", substrates.cylinder(3e-6, ox), 256, 300, b3, v3,
             extra=dict(radius=3e-6, orientation=ox))
    sim_case("cylinder_long", substrates.cylinder(2e-6, ori), 128, 10000, [1e9],
             [[1.0, 0, 0]], extra=dict(radius=2e-6, orientation=ori))
    semi = np.array([10e-6, 5e-6, 2.5e-6])
    Rell = utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([1.0, 1.0, 1.0]))
    sim_case("ellipsoid_traj", substrates.ellipsoid(semi / 4, Rell), 16, 100, b3, v3,
             traj=True, extra=dict(semiaxes=semi / 4, R=Rell))
    sim_case("ellipsoid", substrates.ellipsoid(semi, Rell), 1024, 1000, b3, v3,
             extra=dict(semiaxes=semi, R=Rell))
    sim_case("ellipsoid_eye", substrates.ellipsoid(semi / 2), 512, 500, b3, v3,
             extra=dict(semiaxes=semi / 2, R=np.eye(3)))
    sim_case("ellipsoid_long", substrates.ellipsoid(semi / 2, Rell), 128, 10000, [1e9],
             [[1.0, 0, 0]], extra=dict(semiaxes=semi / 2, R=Rell))

    # meshes: synthetic (my generator) so the fixture carries its own inputs
    def mesh_case(name, vertices, faces, periodic, padding, init_pos, n_sv, n_walkers,
                  n_t, perm_prob=0, traj=False, seed=123, **kw):
        def build():
            return substrates.mesh(vertices, faces, periodic, padding=padding,
                                   init_pos=init_pos, n_sv=np.asarray(n_sv), quiet=True,
                                   perm_prob=perm_prob)
        try:
            sub = build()
        except Exception:
            info["errors"][name] = traceback.format_exc()
            save_info()
            return
        extra = dict(mesh_vertices_in=vertices, mesh_faces_in=faces, periodic=periodic,
                     padding=padding, n_sv=np.asarray(n_sv), perm_prob=perm_prob,
                     init_pos=(init_pos if isinstance(init_pos, np.ndarray)
                               else np.array(init_pos)),
                     sub_vertices=sub.vertices, sub_faces=sub.faces,
                     sub_voxel_size=sub.voxel_size, sub_xs=sub.xs, sub_ys=sub.ys,
                     sub_zs=sub.zs, sub_triangle_indices=sub.triangle_indices,
                     sub_subvoxel_indices=sub.subvoxel_indices)
        sim_case(name, sub, n_walkers, n_t, b3, v3, traj=traj, extra=extra, seed=seed, **kw)

    tv, tf, tpad, tcen = meshgen.tube_lattice(2, 2, 2e-6, 5e-6, 6e-6, 16, 3)
    mesh_case("mesh_tubes_uniform", tv, tf, True, tpad, "uniform", [10, 10, 6], 1024, 300)
    mesh_case("mesh_tubes_traj", tv, tf, True, tpad, "uniform", [10, 10, 6], 16, 100,
              traj=True)
    mesh_case("mesh_tubes_perm", tv, tf, True, tpad, "uniform", [10, 10, 6], 1024, 300,
              perm_prob=0.3)
    mesh_case("mesh_tubes_extra", tv, tf, True, tpad, "extra", [10, 10, 6], 1000, 300)
    mesh_case("mesh_tubes_intra", tv, tf, True, tpad, "intra", [10, 10, 6], 1000, 300)
    sv, sf = meshgen.icosphere(3e-6, 2)
    spad = np.array([0.5e-6, 0.25e-6, 1e-6])
    mesh_case("mesh_sphere_np_uniform", sv, sf, False, spad, "uniform", [8, 9, 10], 1024,
              300)
    mesh_case("mesh_sphere_np_intra", sv, sf, False, spad, "intra", [8, 9, 10], 1000, 300)
    mesh_case("mesh_sphere_np_extra", sv, sf, False, spad, "extra", [8, 9, 10], 1000, 300)
    mesh_case("mesh_sphere_p_intra", sv, sf, True, spad, "intra", [8, 9, 10], 1000, 300)
    rng = np.random.RandomState(7)
    ip = (rng.random_sample((512, 3)) * 0.5 + 0.25) * (2 * 3e-6 + 2 * spad)
    mesh_case("mesh_sphere_np_given", sv, sf, False, spad, ip, [8, 9, 10], 512, 300)
    mesh_case("mesh_sphere_long", sv, sf, False, spad, "intra", [8, 9, 10], 128, 3000)

    # ------------------------------------------------- reference timings
    def timing(name, substrate, n_walkers, n_t, n_meas=1):
        def fn():
            bv = [1e9] * n_meas
            bvec = meshgen.fibonacci_sphere(n_meas) if n_meas > 1 else [[1.0, 0, 0]]
            g, dt = pgse(n_t, bv, bvec)
            simulations.simulation(min(n_walkers, 1000), D, g[:, :10], dt, substrate,
                                   quiet=True)  # JIT warm-up
            cuda.synchronize()
            t0 = time.time()
            simulations.simulation(n_walkers, D, g, dt, substrate, quiet=True)
            cuda.synchronize()
            el = time.time() - t0
            info["timing"][name] = dict(n_walkers=n_walkers, n_t=n_t, n_meas=n_meas,
                                        seconds=el, walker_steps_per_s=n_walkers * n_t / el)
            print(name, info["timing"][name], flush=True)
        run_case("timing_" + name, fn)

    timing("free_1e5x1e3", substrates.free(), 100000, 1000)
    timing("sphere_1e5x1e3", substrates.sphere(10e-6), 100000, 1000)
    timing("sphere_1e6x1e3", substrates.sphere(10e-6), 1000000, 1000)
    timing("cylinder_1e6x1e3", substrates.cylinder(5e-6, oz), 1000000, 1000)
    timing("sphere_1e5x1e3_m180", substrates.sphere(10e-6), 100000, 1000, 180)

    # live PTX of the sphere kernel (version check against the offline dump)
    def dump_asm():
        for nm in ("_cuda_step_sphere", "_cuda_step_free"):
            k = getattr(simulations, nm)
            for sig, asm in k.inspect_asm().items():
                with open(os.path.join(OUT, nm + ".live.ptx"), "w") as f:
                    f.write(asm)
                break
    run_case("dump_asm", dump_asm)
    save_info()
    print(json.dumps({k: v for k, v in info.items() if k != "errors"}, indent=1, default=str))
    print("errors:", list(info["errors"]))


if __name__ == "__main__":
    main()
