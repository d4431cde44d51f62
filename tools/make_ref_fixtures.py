"""Turn the reference's own test fixtures for the hot path into small committed golden files.

Run HERE (the only place /root/reference exists):  python tools/make_ref_fixtures.py
Writes tests/golden/ref_*.npz.  Sources (all under /root/reference/disimpy/tests/):
  test_traj.txt                     golden GPU trajectory, free diffusion, seed 123, 10 walkers
  misst_*_signal_*.txt              MISST reference signals (tests/test_simulations.py:503-654)
  sphere_mesh.pkl + desired_*.npy   mesh subdivision golden (tests/test_substrates.py:366-400)
  cylinder_mesh_closed/open.pkl     meshes used by the reference's mesh physics tests
"""
import os
import pickle

import numpy as np

REF = "/root/reference/disimpy/tests"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    tr = np.loadtxt(os.path.join(REF, "test_traj.txt")).reshape(1000, 10, 3)
    np.savez_compressed(os.path.join(OUT, "ref_test_traj.npz"), traj=tr,
                        step_l=np.linalg.norm(tr[1, 0] - tr[0, 0]))
    misst = {}
    for shape in ("cylinder", "sphere"):
        for sd in (1, 30):
            name = "misst_%s_signal_smalldelta_%dms_bigdelta_40ms_radius_5um" % (shape, sd)
            misst["%s_%dms" % (shape, sd)] = np.loadtxt(os.path.join(REF, name + ".txt"))
    np.savez_compressed(os.path.join(OUT, "ref_misst_signals.npz"), **misst)
    meshes = {}
    for name in ("sphere_mesh", "cylinder_mesh_closed", "cylinder_mesh_open"):
        with open(os.path.join(REF, name + ".pkl"), "rb") as f:
            d = pickle.load(f)
        meshes[name + "_vertices"] = d["vertices"]
        meshes[name + "_faces"] = d["faces"]
    meshes["desired_triangle_indices"] = np.load(os.path.join(REF, "desired_triangle_indices.npy"))
    meshes["desired_subvoxel_indices"] = np.load(os.path.join(REF, "desired_subvoxel_indices.npy"))
    np.savez_compressed(os.path.join(OUT, "ref_meshes.npz"), **meshes)
    for f in sorted(os.listdir(OUT)):
        if f.startswith("ref_"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
