import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """The shared library and the C checker must exist before any test runs."""
    import __graft_entry__ as entry
    entry.build()


# radii of the sphere goldens (the .npz files carry every other input themselves)
SPHERE_RADIUS = {"sphere_traj": 2e-6, "sphere": 10e-6, "sphere_small": 1e-6, "sphere_long": 5e-6,
                 "sphere_iterexc": 0.4e-6}

SIM_CASES = [
    "free_traj", "free", "sphere_traj", "sphere", "sphere_small", "sphere_long", "sphere_iterexc",
    "cylinder_traj", "cylinder", "cylinder_z", "cylinder_x", "cylinder_long",
    "ellipsoid_traj", "ellipsoid", "ellipsoid_eye", "ellipsoid_long",
    "mesh_tubes_uniform", "mesh_tubes_traj", "mesh_tubes_perm", "mesh_tubes_extra",
    "mesh_tubes_intra", "mesh_sphere_np_uniform", "mesh_sphere_np_intra", "mesh_sphere_np_extra",
    "mesh_sphere_p_intra", "mesh_sphere_np_given", "mesh_sphere_long",
]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_kwargs(g):
    kw = {}
    if "kw_max_iter" in g:
        kw["max_iter"] = int(g["kw_max_iter"])
    return kw


def oracle_substrate(name, g):
    """Attribute bag with the reference's substrate attributes, straight from a golden file
    (no product code involved)."""
    NS = types.SimpleNamespace
    if name.startswith("free"):
        return NS(type="free")
    if name.startswith("sphere"):
        return NS(type="sphere", radius=SPHERE_RADIUS[name])
    if name.startswith("cylinder"):
        o = g["orientation"]
        return NS(type="cylinder", radius=float(g["radius"]), orientation=o / np.linalg.norm(o))
    if name.startswith("ellipsoid"):
        return NS(type="ellipsoid", semiaxes=g["semiaxes"], R=g["R"])
    ip = g["init_pos"]
    return NS(type="mesh", vertices=g["sub_vertices"], faces=g["sub_faces"],
              voxel_size=g["sub_voxel_size"], xs=g["sub_xs"], ys=g["sub_ys"], zs=g["sub_zs"],
              triangle_indices=g["sub_triangle_indices"],
              subvoxel_indices=g["sub_subvoxel_indices"], n_sv=g["n_sv"],
              perm_prob=float(g["perm_prob"]), periodic=bool(g["periodic"]),
              init_pos=ip if ip.ndim == 2 else str(ip))


def product_substrate(name, g):
    """The same substrate built through the product's public constructors."""
    from disimpy_b200 import substrates
    if name.startswith("free"):
        return substrates.free()
    if name.startswith("sphere"):
        return substrates.sphere(SPHERE_RADIUS[name])
    if name.startswith("cylinder"):
        return substrates.cylinder(float(g["radius"]), g["orientation"])
    if name.startswith("ellipsoid"):
        return substrates.ellipsoid(g["semiaxes"], g["R"])
    ip = g["init_pos"]
    pp = float(g["perm_prob"])
    return substrates.mesh(g["mesh_vertices_in"], g["mesh_faces_in"], bool(g["periodic"]),
                           padding=g["padding"], init_pos=ip if ip.ndim == 2 else str(ip),
                           n_sv=g["n_sv"], quiet=True, perm_prob=0 if pp == 0 else pp)
