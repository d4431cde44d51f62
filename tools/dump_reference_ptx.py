"""Dump the reference's GPU arithmetic (PTX, and SASS via ptxas) WITHOUT a GPU.

Test infrastructure only.  Imports the reference from oracle/_ref (a git-ignored
``pip install --target`` of /root/reference) and asks Numba/NVVM to compile each
``_cuda_step_*`` kernel for compute_90 (the highest CC Numba 0.65's NVVM wrapper
knows; on a B200 this PTX is JIT-ed by the driver).  The FMA contraction seen in
the output is what disimpy_b200/csrc/*.cuh and oracle/disimpy_oracle.c restate
with explicit fma/mul/add.

Usage: python tools/dump_reference_ptx.py [outdir]   (default gpurun_out/refptx)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "refptx")
    os.makedirs(out, exist_ok=True)
    import numba.cuda.dispatcher as D

    D.get_current_device = lambda: type("Dev", (), {"compute_capability": (9, 0)})()
    from numba import cuda, float64 as f8, int64 as i8, boolean as b1, types
    from numba.cuda.random import xoroshiro128p_type
    import disimpy.simulations as S

    f2 = f8[:, ::1]
    f1 = f8[::1]
    i2 = i8[:, ::1]
    i1 = i8[::1]
    rng = types.Array(xoroshiro128p_type, 1, "C")
    bb = b1[::1]
    common = (f2, f2, f2, f2, f2, rng, i8, f8, f8)
    sigs = {
        "free": common,
        "sphere": common + (f8, bb, i8, f8),
        "cylinder": common + (f8, f2, f2, bb, i8, f8),
        "ellipsoid": common + (f1, f2, f2, bb, i8, f8),
        "mesh": common + (f2, i2, f1, f1, f1, i2, i1, bb, i8, i1, f8, i8),
        "mesh_pp": common + (f2, i2, f1, f1, f1, i2, i1, bb, i8, i1, f8, f8),
    }
    kernels = {
        "free": S._cuda_step_free,
        "sphere": S._cuda_step_sphere,
        "cylinder": S._cuda_step_cylinder,
        "ellipsoid": S._cuda_step_ellipsoid,
        "mesh": S._cuda_step_mesh,
        "mesh_pp": S._cuda_step_mesh,
    }
    for name, sig in sigs.items():
        ptx, _ = cuda.compile_ptx(kernels[name].py_func, sig, cc=(9, 0))
        p = os.path.join(out, name + ".ptx")
        with open(p, "w") as f:
            f.write(ptx)
        cubin = os.path.join(out, name + ".cubin")
        r = subprocess.run(["ptxas", "-arch=sm_100a", "-v", p, "-o", cubin],
                           capture_output=True, text=True)
        sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True)
        with open(os.path.join(out, name + ".sass"), "w") as f:
            f.write(sass.stdout)
        print(name, len(ptx), r.stderr.strip().splitlines()[-1] if r.stderr else "")


if __name__ == "__main__":
    main()
