"""Turn the reference's own test fixtures for the hot path into small committed golden files.

Run HERE (the only place /root/reference exists):  python tools/make_ref_fixtures.py
Writes tests/golden/ref_*.npz.  Sources (all under /root/reference/disimpy/tests/):
  test_traj.txt                     golden GPU trajectory, free diffusion, seed 123, 10 walkers
  misst_*_signal_*.txt              MISST reference signals (tests/test_simulations.py:503-654)
  sphere_mesh.pkl + desired_*.npy   mesh subdivision golden (tests/test_substrates.py:366-400)
  cylinder_mesh_closed/open.pkl     meshes used by the reference's mesh physics tests
  neuron-model.pkl, example_mesh.pkl, fibre_mesh.pkl
                                    irregular meshes of the reference's tests (tests/test_simulations.py:814-832,
                                    tests/test_substrates.py); faces stored as int32 (values < 2^31), vertices in
                                    their own dtype (the neuron model's are float32, which is part of the test)
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
    real = {}
    for name in ("neuron-model", "example_mesh", "fibre_mesh"):
        with open(os.path.join(REF, name + ".pkl"), "rb") as f:
            d = pickle.load(f)
        key = name.replace("-", "_")
        real[key + "_vertices"] = d["vertices"]
        assert d["faces"].max() < 2 ** 31 and d["faces"].min() >= 0
        real[key + "_faces"] = d["faces"].astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "ref_real_meshes.npz"), **real)
    # gradient helpers of the reference (disimpy/gradients.py) on fixed inputs: the product's pgse / set_b /
    # calc_b / interpolate_gradient must return the same arrays bit for bit
    import sys
    import warnings
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(here, "oracle", "stubs"), "/root/reference"]
    import disimpy.gradients as rg
    rs = np.random.RandomState(0)
    w = rs.normal(size=(5, 40, 3))
    bvecs = rs.normal(size=(7, 3))
    bvecs /= np.linalg.norm(bvecs, axis=1)[:, None]
    bvals = np.linspace(5e8, 3e9, 7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pg, pdt = rg.pgse(5e-3, 20e-3, 37, bvals, bvecs)
        ig, idt = rg.interpolate_gradient(w, 1e-3, 93)
    np.savez_compressed(os.path.join(OUT, "ref_gradients.npz"), w=w, bvecs=bvecs, bvals=bvals, pgse=pg, pgse_dt=pdt,
                        calc_b=rg.calc_b(w, 1e-3), calc_q=rg.calc_q(w, 1e-3), set_b=rg.set_b(w, 1e-3, np.arange(1, 6) * 1e9),
                        interp=ig, interp_dt=idt)
    for f in sorted(os.listdir(OUT)):
        if f.startswith("ref_"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
