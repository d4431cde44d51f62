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


def check_device_function_known_answers(unit):
    """The reference's unit tests of its device functions (disimpy/tests/test_simulations.py:23-109,
    142-360) with `unit(name, rows_of_arguments) -> rows_of_results` standing in for the test kernels:
    dot / cross / normalize / triangle normal / mat_mul against NumPy on 100 random cases, and the
    known answers for the intersection checks, the reflection and the crossing (7 decimals, like
    npt.assert_almost_equal there)."""
    import numpy.testing as npt
    rs = np.random.RandomState(123)
    a, b = rs.random_sample((100, 3)) - 0.5, rs.random_sample((100, 3)) - 0.5
    npt.assert_almost_equal(unit("dot_product", np.hstack([a, b]))[:, 0], np.einsum("ij,ij->i", a, b))
    npt.assert_almost_equal(unit("cross_product", np.hstack([a, b])), np.cross(a, b))
    npt.assert_almost_equal(unit("normalize_vector", a), a / np.linalg.norm(a, axis=1)[:, None])
    tri = rs.random_sample((100, 3, 3)) - 0.5
    n = np.cross(tri[:, 0] - tri[:, 1], tri[:, 0] - tri[:, 2])
    npt.assert_almost_equal(unit("triangle_normal", tri.reshape(100, 9)), n / np.linalg.norm(n, axis=1)[:, None])
    R = rs.random_sample((100, 3, 3)) - 0.5
    npt.assert_almost_equal(unit("mat_mul", np.hstack([R.reshape(100, 9), a])), np.einsum("nij,nj->ni", R, a))
    s = np.array([1.0, 1.0, 0.0]) / np.linalg.norm([1.0, 1.0, 0.0])
    npt.assert_almost_equal(unit("line_circle_intersection", [[-0.1, -0.1, s[0], s[1], 1.0]])[0, 0], 1.1414213562373097)
    npt.assert_almost_equal(unit("line_sphere_intersection", [[-0.1, -0.1, 0.0, *s, 1.0]])[0, 0], 1.1414213562373097)
    npt.assert_almost_equal(unit("line_ellipsoid_intersection", [[-0.1, -0.1, 0.0, *s, 1.0, 1.0, 1.0]])[0, 0],
                            1.1414213562373097)
    triangle = [2.0, 0, 0, 0, 2.0, 0, 0.0, 0, 0]
    r0s = [[0.1, 0.1, 1.0]] * 4 + [[10.0, 10.0, 0.0]]
    steps = [[0, 0, -1.0], [0, 0, 1.0], [0, 0, -0.1], [1.0, 1.0, 0], [0, 0, 1.0]]
    ds = unit("ray_triangle_intersection_check", [triangle + r + st for r, st in zip(r0s, steps)])
    npt.assert_almost_equal(ds, np.array([[1, -1, 10, np.nan, np.nan]]).T)
    normal = np.array([0.0, 1.0, 1.0]) / np.linalg.norm([0.0, 1.0, 1.0])
    # (the reference's kernel flips its `normal` array in place so that it points against the step, and the
    # second expectation there is written with the array the first call left behind: -normal)
    for eps in (0.0, 0.5):
        for nrm in (normal, -normal):
            out = unit("reflection", [[0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.5, *nrm, eps]])[0]
            npt.assert_almost_equal(out[3:], [0.0, -1.0, 0.0])
            npt.assert_almost_equal(out[:3], np.array([0.0, 0.0, 0.5]) - normal * eps)
    # a bounce off the triangle (0,0,0) (1,0,0) (0,1,0) from above: hit at d = 0.5, back up, eps above the plane
    tri = [0.0, 0, 0, 1.0, 0, 0, 0, 1.0, 0]
    r0, step, eps = [0.0, 0.0, 0.5], [0.0, 0.0, -1.0], 1e-10
    d = unit("ray_triangle_intersection_check", [tri + r0 + step])[0, 0]
    assert 0 < d < 1.0
    nrm = unit("triangle_normal", [tri])[0]
    out = unit("reflection", [r0 + step + [d] + list(nrm) + [eps]])[0]
    npt.assert_almost_equal(out[3:], [0.0, 0.0, 1.0])
    npt.assert_almost_equal(out[:3], [0.0, 0.0, eps])
    assert out[2] == eps
    # through the triangle (0,0,1) (1,0,1) (0,1,1) from below: eps beyond the membrane
    tri = [0.0, 0, 1.0, 1.0, 0, 1.0, 0, 1.0, 1.0]
    r0, step = [0.0, 0.0, 0.0], [0.0, 0.0, 1.0]
    d = unit("ray_triangle_intersection_check", [tri + r0 + step])[0, 0]
    assert 0 < d < 2
    nrm = unit("triangle_normal", [tri])[0]
    out = unit("crossing", [r0 + step + [d] + list(nrm) + [eps]])[0]
    npt.assert_almost_equal(out, [0.0, 0.0, 1 + eps], decimal=12)


def device_function_random_rows(name, n, seed=5):
    """Random argument rows for `name` (collisions that really happen where the function expects
    one), for bit-for-bit comparisons between two implementations."""
    rs = np.random.RandomState(seed)
    u = lambda *shape: rs.random_sample(shape) - 0.5
    unit_vec = lambda: (lambda v: v / np.linalg.norm(v, axis=1)[:, None])(rs.normal(size=(n, 3)))
    if name in ("dot_product", "cross_product"):
        return np.hstack([u(n, 3), u(n, 3)])
    if name == "normalize_vector":    # + zero, overflowing and underflowing squared lengths
        rows = u(n, 3) * 10.0 ** rs.randint(-8, 3, size=(n, 1))
        rows[:6] = [[0, 0, 0], [1e200, 1e200, 0], [1e-200, 0, 1e-200], [0, -0.0, 3e-160], [1e154, 1e154, 1e154], [5e-324, 0, 0]]
        return rows
    if name == "triangle_normal":     # + triangles without area (NaN normals)
        rows = u(n, 9) * 1e-5
        rows[0, 3:6] = rows[0, 0:3]
        rows[1, 6:9] = rows[1, 3:6]
        rows[2] = 0.0
        return rows
    if name == "mat_mul":
        return np.hstack([u(n, 9), u(n, 3) * 1e-5])
    if name == "line_circle_intersection":
        sv = unit_vec()
        return np.hstack([u(n, 2) * 1e-5, sv[:, 1:], np.full((n, 1), 1e-5)])
    if name == "line_sphere_intersection":   # + starting points outside the sphere (negative discriminants: NaN)
        rows = np.hstack([u(n, 3) * 1e-5, unit_vec(), np.full((n, 1), 1e-5)])
        rows[:50, :3] *= 4.0
        return rows
    if name == "line_ellipsoid_intersection":
        return np.hstack([u(n, 3) * 1e-6, unit_vec(), np.tile([1e-5, 5e-6, 2.5e-6], (n, 1))])
    if name == "ray_triangle_intersection_check":   # + rays in the triangle's plane (zero determinant)
        rows = np.hstack([u(n, 9) * 1e-5, u(n, 3) * 1e-5, unit_vec()])
        rows[:20, 2:9:3] = 0.0
        rows[:20, 14] = 0.0
        return rows
    if name in ("reflection", "crossing"):
        return np.hstack([u(n, 3) * 1e-5, unit_vec(), rs.random_sample((n, 1)) * 1e-6, unit_vec(), np.full((n, 1), 1e-13)])
    if name.endswith("subvoxel_overlap") or name.endswith("subvoxel_overlap_periodic"):
        # np.linspace boundaries like substrates.mesh makes them (2 to 16 of them), a few stretched ones;
        # coordinates up to three voxels away on either side, on boundaries and images of boundaries
        rows = np.zeros((n, 19))
        for i in range(n):
            ln = rs.randint(2, 17)
            V = 10.0 ** rs.uniform(-6, -3)
            xs = np.linspace(0, V, ln)
            if i % 7 == 0 and ln > 2:
                xs[1:-1] += rs.uniform(-0.3, 0.3, size=ln - 2) * V / (ln - 1)
            x = rs.uniform(-3, 3, size=2) * V
            if i % 3 == 0:
                x[0] = xs[rs.randint(ln)] + rs.randint(-3, 4) * V
            if i % 5 == 0:
                x[1] = rs.randint(-3, 4) * V
            rows[i, :2], rows[i, 2], rows[i, 3:3 + ln] = x, ln, xs
        return rows
    raise KeyError(name)


def subvoxel_overlap_restated(name, row):
    """disimpy/simulations.py:616-679 in plain Python (the loops as written there), for one argument row."""
    import math
    ln = int(row[2])
    xs, x1, x2 = row[3:3 + ln], row[0], row[1]

    def ll(xmin):
        if xmin <= xs[0]:
            return 0
        if xmin >= xs[-1]:
            return len(xs) - 1
        for i, x in enumerate(xs):
            if x > xmin:
                return i - 1
        return 0

    def ul(xmax):
        if xmax >= xs[-1]:
            return len(xs) - 1
        if xmax <= xs[0]:
            return 0
        for i, x in enumerate(xs):
            if not x < xmax:
                return i
        return len(xs) - 1

    if name == "ll_subvoxel_overlap":
        return ll(min(x1, x2))
    if name == "ul_subvoxel_overlap":
        return ul(max(x1, x2))
    voxel_size = abs(xs[-1] - xs[0])
    x = min(x1, x2) if name.startswith("ll") else max(x1, x2)
    nq = math.floor(x / voxel_size)
    shifted = x - nq * voxel_size
    return (ll(shifted) if name.startswith("ll") else ul(shifted)) + nq * (len(xs) - 1)
