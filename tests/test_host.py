"""CPU-only tests: the C ABI library loads and exports what include/disimpy_b200.h declares,
host-side logic (validation, substrates, mesh binning, initial positions, sharding) and the
reference's CPU-level known answers (tests/test_substrates.py, tests/test_gradients.py,
tests/test_utils.py of the reference).  No compute call touches a GPU here."""

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden, oracle_substrate, product_substrate


def test_library_exports_every_declared_symbol():
    from disimpy_b200 import _lib
    header = open(os.path.join(ROOT, "include", "disimpy_b200.h")).read()
    declared = set(re.findall(r"\b(dsb_[a-z_0-9]+)\s*\(", header))
    declared -= {"dsb_reset"}  # mentioned in the mapping comment only
    L = _lib.lib()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), "missing export %s" % name
    assert set(_lib.EXPORTS) == declared
    assert b"sm_100a" in L.dsb_version()


def test_struct_layout_matches_header():
    """ctypes mirror of dsb_params / dsb_mesh has the C layout (checked via a tiny C program)."""
    from disimpy_b200 import _lib
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "disimpy_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(dsb_params), sizeof(dsb_mesh),
               offsetof(dsb_params, seed), offsetof(dsb_params, R), offsetof(dsb_params, mesh),
               offsetof(dsb_mesh, n_sv), offsetof(dsb_mesh, perm_prob));
        return 0;
    }'''
    exe = os.path.join(ROOT, "oracle", "_build", "layout_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src.encode(), check=True)
    got = [int(v) for v in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    P, M = _lib.DsbParams, _lib.DsbMesh
    assert got == [ctypes.sizeof(P), ctypes.sizeof(M), P.seed.offset, P.R.offset, P.mesh.offset,
                   M.n_sv.offset, M.perm_prob.offset]


def test_no_gpu_fails_loudly():
    """Without a CUDA device simulation() raises instead of computing anything on the CPU."""
    from disimpy_b200 import _lib, gradients, simulations, substrates
    count = ctypes.c_int32(0)
    rc = _lib.lib().dsb_device_count(ctypes.byref(count))
    if rc == 0 and count.value > 0:
        pytest.skip("a GPU is present")
    g, dt = gradients.pgse(5e-3, 20e-3, 10, [1e9], [[1.0, 0, 0]])
    with pytest.raises(Exception, match="unable to detect a CUDA GPU"):
        simulations.simulation(10, 2e-9, g, dt, substrates.free(), quiet=True)
    with pytest.raises(_lib.DsbError):
        simulations.rng_states(1, 4)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "disimpy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "meshgen.py", f


# ------------------------------------------------------------------ substrates

def test_substrate_validation():
    from disimpy_b200 import substrates
    assert substrates.free().type == "free"
    s = substrates.sphere(1e-6)
    assert (s.type, s.radius) == ("sphere", 1e-6)
    for bad in (1, -1.0, "r", 0.0):
        with pytest.raises(ValueError):
            substrates.sphere(bad)
    c = substrates.cylinder(2e-6, np.array([0.0, 0.0, 2.0]))
    assert np.array_equal(c.orientation, [0, 0, 1.0])
    for bad_o in ([0, 0, 1.0], np.array([0, 0, 1]), np.zeros(2)):
        with pytest.raises(ValueError):
            substrates.cylinder(2e-6, bad_o)
    e = substrates.ellipsoid(np.array([1e-6, 2e-6, 3e-6]))
    assert np.array_equal(e.R, np.eye(3))
    with pytest.raises(ValueError):
        substrates.ellipsoid(np.array([1e-6, 2e-6, 3e-6]), 2 * np.eye(3))
    with pytest.raises(ValueError):
        substrates.ellipsoid(np.array([1, 2, 3]))


def test_mesh_validation_and_layout():
    from disimpy_b200 import meshgen, substrates
    v, f = meshgen.icosphere(1e-6, 1)
    with pytest.raises(ValueError):
        substrates.mesh(v.astype(int), f, True, quiet=True)
    with pytest.raises(ValueError):
        substrates.mesh(v, f.astype(float), True, quiet=True)
    with pytest.raises(ValueError):
        substrates.mesh(v, f, 1, quiet=True)
    with pytest.raises(ValueError):
        substrates.mesh(v, f, True, init_pos="inside", quiet=True)
    with pytest.raises(ValueError):
        substrates.mesh(v, f, True, perm_prob=2.0, quiet=True)
    with pytest.raises(ValueError):
        substrates.mesh(v, f, True, n_sv=np.array([5.0, 5, 5]), quiet=True)
    pad = np.array([1e-7, 2e-7, 3e-7])
    s = substrates.mesh(v, f, False, padding=pad, n_sv=np.array([4, 5, 6]), quiet=True)
    assert np.allclose(s.vertices[:-8].min(axis=0), pad)
    assert np.allclose(s.voxel_size, s.vertices[:-8].max(axis=0) + pad)
    assert len(s.faces) == len(f) + 12 and len(s.vertices) == len(v) + 8
    assert np.array_equal(s.vertices[-8], [0, 0, 0]) and np.array_equal(s.vertices[-5], s.voxel_size)
    assert s.subvoxel_indices.shape == (120, 2) and s.subvoxel_indices[-1, 1] == len(s.triangle_indices)
    assert len(s.xs) == 5 and len(s.ys) == 6 and len(s.zs) == 7
    p = substrates.mesh(v, f, True, quiet=True)
    assert len(p.faces) == len(f)


def test_reference_unit_known_answers():
    """tests/test_substrates.py:273-363 and :450-480 of the reference."""
    from disimpy_b200 import substrates
    tri = np.array([[0.5, 0.7, 0.3], [0.9, 0.5, 0.2], [0.6, 0.9, 0.8]])
    assert not substrates._triangle_box_overlap(tri, np.array([[0.1, 0.3, 0.1], [0.4, 0.7, 0.5]]))
    tri = np.array([[0.4, 0.7, 0.2], [0.9, 0.5, 0.2], [0.6, 0.9, 0.2]])
    assert not substrates._triangle_box_overlap(tri, np.array([[0.4, 0.4, 0.3], [0.5, 0.8, 0.6]]))
    tri = np.array([[0.63149023, 0.44235872, 0.77212144], [0.25125724, 0.00087658, 0.66026559],
                    [0.8319006, 0.52731735, 0.22859846]])
    box = np.array([[0.33109806, 0.16637023, 0.91545459], [0.79806038, 0.83915475, 0.38118002]])
    assert substrates._triangle_box_overlap(tri, box)
    xs = np.arange(11)
    assert substrates._interval_sv_overlap(xs, 0, 0) == (0, 1)
    assert substrates._interval_sv_overlap(xs, 0, 1.5) == (0, 2)
    assert substrates._interval_sv_overlap(xs, 9.5, 1.5) == (1, 10)
    assert substrates._interval_sv_overlap(xs, -1.1, 0.5) == (0, 1)
    assert substrates._interval_sv_overlap(xs, 9.5, 11.5) == (9, 10)
    # :273-290 (cross / dot against NumPy), :347-365 (bounding box, subvoxel range), :450-480 (box -> mesh)
    rs = np.random.RandomState(123)
    for _ in range(100):
        a, b = rs.random_sample(3) - 0.5, rs.random_sample(3) - 0.5
        assert np.allclose(substrates._cross_product(a, b), np.cross(a, b), atol=1e-7)
        assert np.isclose(substrates._dot_product(a, b), np.dot(a, b), atol=1e-7)
    tri = np.array([[0.5, 0.7, 0.3], [0.9, 0.5, 0.2], [0.6, 0.9, 0.8]])
    assert np.array_equal(substrates._triangle_aabb(tri), np.vstack((tri.min(axis=0), tri.max(axis=0))))
    box = np.array([[2.5, 5.0, 2.2], [9.2, 9.5, 20]])
    assert np.array_equal(substrates._box_subvoxel_overlap(box, np.arange(6), np.arange(11), np.arange(21)),
                          np.array([[2, 5], [5, 10], [2, 20]]))
    vertices = np.array([[2.5, 5.0, 2.2], [9.2, 5.0, 2.2], [9.2, 9.5, 2.2], [9.2, 9.5, 20.0], [2.5, 9.5, 20.0],
                         [2.5, 5.0, 20.0], [2.5, 9.5, 2.2], [9.2, 5.0, 20.0]])
    faces = np.array([[0, 1, 2], [0, 6, 2], [5, 7, 3], [5, 4, 3], [1, 2, 3], [1, 7, 3], [0, 6, 4], [0, 5, 4],
                      [0, 1, 7], [0, 5, 7], [6, 2, 3], [6, 4, 3]])
    v, f = substrates._aabb_to_mesh(box[0], box[1])
    assert np.array_equal(v, vertices) and np.array_equal(f, faces)


def test_mesh_subdivision_reference_golden():
    """tests/test_substrates.py:366-400: sphere_mesh.pkl, n_sv = [2, 5, 10]."""
    from disimpy_b200 import substrates
    rm = load_golden("ref_meshes")
    s = substrates.mesh(rm["sphere_mesh_vertices"], rm["sphere_mesh_faces"], True,
                        n_sv=np.array([2, 5, 10]), quiet=True)
    assert np.array_equal(s.triangle_indices, rm["desired_triangle_indices"])
    assert np.array_equal(s.subvoxel_indices, rm["desired_subvoxel_indices"])


@pytest.mark.parametrize("name", ["mesh_tubes_uniform", "mesh_sphere_np_uniform",
                                  "mesh_sphere_p_intra", "mesh_tubes_perm"])
def test_mesh_substrate_equals_reference_arrays(name):
    """Substrate arrays the unmodified reference built for the same input mesh."""
    g = load_golden(name)
    s = product_substrate(name, g)
    for k in ("vertices", "faces", "voxel_size", "xs", "ys", "zs", "triangle_indices",
              "subvoxel_indices"):
        assert np.array_equal(getattr(s, k), g["sub_" + k]), k


def test_mesh_subdivision_threaded_equals_serial():
    """> 2048 faces takes the multi-threaded path; the cell lists must stay face-ascending."""
    from disimpy_b200 import meshgen, substrates
    v, f, pad, _ = meshgen.tube_lattice(3, 3, 2e-6, 5e-6, 8e-6, 32, 6)
    s = substrates.mesh(v, f, True, padding=pad, n_sv=np.array([9, 9, 5]), quiet=True)
    assert len(f) > 2048
    for a, b in s.subvoxel_indices:
        cell = s.triangle_indices[a:b]
        assert np.all(np.diff(cell) > 0)
    # every listed pair really overlaps, and the brute-force count agrees on a sample of cells
    rs = np.random.RandomState(0)
    for c in rs.choice(len(s.subvoxel_indices), 12, replace=False):
        x, rem = divmod(c, 9 * 5)
        y, z = divmod(rem, 5)
        box = np.array([[s.xs[x], s.ys[y], s.zs[z]], [s.xs[x + 1], s.ys[y + 1], s.zs[z + 1]]])
        hits = [i for i in range(len(f)) if substrates._triangle_box_overlap(s.vertices[s.faces[i]], box)]
        a, b = s.subvoxel_indices[c]
        assert list(s.triangle_indices[a:b]) == hits


# ------------------------------------------------------------------ host math

def test_vec2vec_rotmat():
    from disimpy_b200 import utils
    rs = np.random.RandomState(123)
    for _ in range(200):
        v, k = rs.random_sample(3) - 0.5, rs.random_sample(3) - 0.5
        R = utils.vec2vec_rotmat(v, k)
        assert np.allclose(R @ (v / np.linalg.norm(v)), k / np.linalg.norm(k))
        assert np.isclose(np.linalg.det(R), 1)
    assert np.array_equal(utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([2.0, 0, 0])), np.eye(3))
    assert np.array_equal(utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([-1.0, 0, 0])), -np.eye(3))


def test_gradients_known_answers():
    """tests/test_gradients.py:20-111 of the reference."""
    from disimpy_b200 import gradients
    T = 80e-3
    g = np.zeros((1, 1000, 3))
    g[0, 1:201, 0] = 0.1
    g[0, -201:-1, 0] = -0.1
    dt = T / (g.shape[1] - 1)
    assert np.allclose(gradients.calc_b(g, dt), 1.07507347e10, rtol=1e-7)
    g2, dt2 = gradients.interpolate_gradient(g, dt, 5000)
    assert g2.shape == (1, 5000, 3) and np.isclose(dt2, T / 4999)
    assert np.isclose(gradients.calc_b(g2, dt2) / gradients.calc_b(g, dt), 1, atol=1e-3)
    bs = np.array([1e9, 3e9])
    gg = gradients.set_b(np.concatenate([g2, g2]), dt2, bs)
    assert np.allclose(gradients.calc_b(gg, dt2), bs)
    bvecs = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]])
    bvals = np.array([1e9, 2e9, 3e9])
    p, pdt = gradients.pgse(10e-3, 40e-3, 500, bvals, bvecs)
    assert p.shape == (3, 500, 3)
    assert np.allclose(p.sum(axis=1), 0, atol=1e-9)
    assert np.allclose(gradients.calc_b(p, pdt), bvals)
    for i in range(3):
        off = [j for j in range(3) if j != i]
        assert np.allclose(p[i][:, off], 0)


def test_initial_positions_match_reference_stream():
    """Host samplers: the first accepted points of the MT19937 stream (any failure here would
    also break every golden comparison of final positions)."""
    from disimpy_b200 import simulations
    from oracle import oracle as O
    import types
    pts = simulations._fill_sphere(5000, 3e-6, 123)
    assert np.all(np.linalg.norm(pts, axis=1) < 3e-6)
    seq = np.random.RandomState(123)
    manual = []
    while len(manual) < 50:
        p = (seq.random_sample(3) - 0.5) * 2 * 3e-6
        if np.linalg.norm(p) < 3e-6:
            manual.append(p)
    assert np.array_equal(pts[:50], np.array(manual))
    sub = types.SimpleNamespace(type="sphere", radius=3e-6)
    assert np.array_equal(pts, O.initial_positions(sub, 5000, 123))
    ax = np.array([3e-6, 2e-6, 1e-6])
    e = simulations._fill_ellipsoid(2000, ax, 5)
    assert np.all(((e / ax) ** 2).sum(axis=1) < 1)
    c = simulations._fill_circle(2000, 1e-6, 5)
    assert c.shape == (2000, 2) and np.all(np.linalg.norm(c, axis=1) < 1e-6)
    # ellipsoid / disc streams against the oracle's NumPy restatement
    esub = types.SimpleNamespace(type="ellipsoid", semiaxes=ax, R=np.eye(3))
    assert np.array_equal(e, O.initial_positions(esub, 2000, 5))
    csub = types.SimpleNamespace(type="cylinder", radius=1e-6, orientation=np.array([1.0, 0, 0]))
    assert np.array_equal(c, O.initial_positions(csub, 2000, 5)[:, 1:3])
    with pytest.raises(ValueError):
        simulations._fill_sphere(10, 1e-6, 2 ** 32)


def test_host_samplers_like_reference():
    """The reference's tests of its host samplers (disimpy/tests/test_simulations.py:363-425), same
    sizes, same assertions."""
    import numpy.testing as npt
    from scipy.stats import kstest
    from disimpy_b200 import simulations, utils
    radius, N = 5e-6, int(1e5)
    for fill in (simulations._fill_circle, simulations._fill_sphere):
        points = fill(N, radius)
        assert np.max(np.linalg.norm(points, axis=1)) < radius
        npt.assert_almost_equal(np.mean(points, axis=0), 0)
        _, p = kstest((points.ravel() + radius) / radius, "uniform")
        npt.assert_almost_equal(p, 0)
    a, b, c = 10e-6, 2e-6, 5e-6
    points = simulations._fill_ellipsoid(N, np.array([a, b, c]))
    assert np.all(np.max(points, axis=0) < [a, b, c]) and np.all(np.min(points, axis=0) > [-a, -b, -c])
    npt.assert_almost_equal(np.mean(points, axis=0), 0)
    for i, r in enumerate([a, b, c]):
        _, p = kstest((points[:, i].ravel() + r) / r, "uniform")
        npt.assert_almost_equal(p, 0)
    N, r = int(1e3), 5e-6
    R = utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([0, 1.0, 0]))
    R_inv = np.linalg.inv(R)
    pos = simulations._initial_positions_cylinder(N, r, R)
    npt.assert_almost_equal(pos[:, 1], np.zeros(N))
    npt.assert_almost_equal(np.matmul(R_inv, pos.T)[0], np.zeros(N))
    pos = simulations._initial_positions_ellipsoid(N, np.array([r, r, 1e-22]), R)
    npt.assert_almost_equal(pos[:, 2], np.zeros(N))
    npt.assert_almost_equal(np.matmul(R_inv, pos.T)[2], np.zeros(N))
    # _set_seed + a call without a seed == the explicit seed == RandomState(seed)'s rejection stream
    simulations._set_seed(77)
    pts = simulations._fill_sphere(1000, r)
    assert np.array_equal(pts, simulations._fill_sphere(1000, r, 77))
    seq, manual = np.random.RandomState(77), []
    while len(manual) < 20:
        q = (seq.random_sample(3) - 0.5) * 2 * r
        if np.linalg.norm(q) < r:
            manual.append(q)
    assert np.array_equal(pts[:20], np.array(manual))
    simulations._SEED = None


def test_traj_line_equals_python_str():
    """simulations.py:1043-1048 writes str(value) + " " per value; the native formatter must produce the
    same characters for every double (shortest round-trip digits, Python's layout), on its serial
    and its threaded path."""
    from disimpy_b200 import simulations
    rs = np.random.RandomState(0)
    v = np.concatenate([rs.normal(size=200000) * 10.0 ** rs.randint(-320, 300, size=200000),
                        rs.normal(size=100000) * 1e-5, rs.randint(-1000, 1000, size=1000).astype(float),
                        np.arange(-20, 20) * 0.1, 10.0 ** np.arange(-25, 25), -(10.0 ** np.arange(-25, 25)) * 1.5,
                        [0.0, -0.0, np.nan, np.inf, -np.inf, 1e16, 1e15, 9999999999999998.0, 1e-4, 9.9e-5, 5e-324,
                         -5e-324, 1.7976931348623157e308, -1.7976931348623157e308, 123456789012345680.0,
                         2.2250738585072014e-308, 0.1, 1 / 3, 2 / 3, 1e22, 1e23]])
    for part in (v, v[:1000], v[-7:], v[:1], v[:0]):
        want = ("".join(str(x) + " " for x in part) + "\n").encode()
        assert bytes(simulations._traj_line(part)) == want
    back = np.array(bytes(simulations._traj_line(v[:200000])).split(), dtype=float)
    assert np.array_equal(back, v[:200000])       # and every value reads back to the same double
    # the file: one line per call, whatever the chunking
    import tempfile
    path = os.path.join(tempfile.mkdtemp(), "traj.txt")
    pos = v[:3000].reshape(1000, 3)
    old_chunk, simulations._TRAJ_CHUNK = simulations._TRAJ_CHUNK, 700
    try:
        simulations._write_traj(path, "w", pos)
        simulations._write_traj(path, "a", pos[:5])
    finally:
        simulations._TRAJ_CHUNK = old_chunk
    want = "".join(str(x) + " " for x in pos.ravel()) + "\n" + "".join(str(x) + " " for x in pos[:5].ravel()) + "\n"
    assert open(path).read() == want


def test_simulation_argument_validation():
    """simulations.py:1127-1153: same ValueErrors (checked before any GPU work would start)."""
    from disimpy_b200 import _lib, gradients, simulations, substrates
    count = ctypes.c_int32(0)
    if _lib.lib().dsb_device_count(ctypes.byref(count)) != 0 or count.value < 1:
        pytest.skip("validation runs after GPU detection, like the reference")
    g, dt = gradients.pgse(5e-3, 20e-3, 10, [1e9], [[1.0, 0, 0]])
    sub = substrates.free()
    for kwargs in (dict(n_walkers=1.5), dict(diffusivity=2), dict(gradient=g[0]), dict(dt=1),
                   dict(substrate="free"), dict(seed=-1), dict(traj=5), dict(quiet=1),
                   dict(cuda_bs=0), dict(max_iter=0)):
        args = dict(n_walkers=10, diffusivity=2e-9, gradient=g, dt=dt, substrate=sub, quiet=True)
        args.update(kwargs)
        with pytest.raises(ValueError):
            simulations.simulation(**args)


def test_shard_ranges_cover_all_walkers():
    from disimpy_b200 import simulations
    for n in (1, 7, 1000, 10 ** 8 + 3):
        for w in (1, 2, 4, 8):
            r = [simulations.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


_GLOO_WORKER = r'''
import os, sys, types
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
from disimpy_b200 import simulations
from oracle import oracle as O
rank, world, d = simulations._dist()
assert world == 2 and d is not None
n = 301
lo, hi = simulations.shard_range(n, rank, world)
sub = types.SimpleNamespace(type="sphere", radius=1e-6)
rs = np.random.RandomState(3)
grad = rs.normal(size=(3, 25, 3)) * 0.05
pos0 = O.initial_positions(sub, n, 9)
# each rank walks its shard with the ORACLE standing in for the GPU kernels (host logic test)
part = O.run_walk(sub, grad, 1e-4, 2e-9, pos0[lo:hi], seed=9, walker_offset=lo)
sig = O.signals_from_phases(part["phases"], part["iter_exc"])
total = simulations._allreduce_sum(sig, d)
allpos = simulations._assemble_rows([(simulations.owned_ranges(n, rank, world), part["positions"])], n, d)
full = O.run_walk(sub, grad, 1e-4, 2e-9, pos0, seed=9)
assert np.array_equal(allpos, full["positions"])
assert np.allclose(total, O.signals_from_phases(full["phases"], full["iter_exc"]), rtol=1e-13)
# the round-robin deal of parts (what a pipelined multi-GPU simulation() uses): every part with its
# own RNG offset, rows gathered back into global order
owned = simulations.owned_ranges(n, rank, world, interleaved=True, part=64)
assert sum(b - a for a, b, _ in owned) > 0 and [la for _, _, la in owned][0] == 0
rows = np.zeros((sum(b - a for a, b, _ in owned), 3))
sig2 = np.zeros(3)
for a, b, la in owned:
    piece = O.run_walk(sub, grad, 1e-4, 2e-9, pos0[a:b], seed=9, walker_offset=a)
    rows[la:la + b - a] = piece["positions"]
    sig2 += O.signals_from_phases(piece["phases"], piece["iter_exc"])
assert np.array_equal(simulations._assemble_rows([(owned, rows)], n, d), full["positions"])
assert np.allclose(simulations._allreduce_sum(sig2, d), O.signals_from_phases(full["phases"], full["iter_exc"]), rtol=1e-13)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_sharding(tmp_path):
    """world_size 2 on CPU: shard ranges, RNG offsets, the signal all-reduce and the row gather
    used by simulation() reproduce the single-process result."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_position_parts_equal_one_shot_sampling():
    """simulation() draws the initial positions of analytic substrates part by part (to overlap
    the sequential host stream with the GPU); the parts must be the one-shot arrays, bit for bit,
    also for shards that start in the middle of the stream."""
    from disimpy_b200 import simulations, substrates, utils
    n = 300_000
    R = utils.vec2vec_rotmat(np.array([1.0, 0, 0]), np.array([0.3, 1.0, -0.4]))
    subs = [substrates.sphere(3e-6), substrates.cylinder(2e-6, np.array([0.2, -1.0, 0.5])),
            substrates.ellipsoid(np.array([3e-6, 2e-6, 1e-6]), R)]
    for sub in subs:
        if sub.type == "sphere":
            full = simulations._fill_sphere(n, sub.radius, 9)
        elif sub.type == "cylinder":
            Rc = utils.vec2vec_rotmat(sub.orientation, np.array([1.0, 0, 0]))
            full = simulations._initial_positions_cylinder(n, sub.radius, np.linalg.inv(Rc), 9)
        else:
            full = simulations._initial_positions_ellipsoid(n, sub.semiaxes, sub.R, 9)
        for lo, hi in ((0, n), (123_457, n)):
            edges = list(range(lo, hi, 65_536)) + [hi]
            got = np.vstack(list(simulations._stream_stretches(sub, 9, list(zip(edges[:-1], edges[1:])))))
            assert np.array_equal(got, full[lo:hi]), sub.type
        # stretches with gaps (a rank's round-robin parts): the walkers in between are drawn and dropped
        picked = [(0, 1000), (5000, 70_000), (200_000, 200_001)]
        for (a, b), pts in zip(picked, simulations._stream_stretches(sub, 9, picked)):
            assert np.array_equal(pts, full[a:b]), sub.type


def test_round_robin_parts_cover_all_walkers():
    from disimpy_b200 import simulations
    for n, world, part in ((1000, 3, 64), (131072 * 5 + 17, 2, None), (64, 4, 64), (10, 1, 4)):
        seen = np.zeros(n, dtype=int)
        for rank in range(world):
            owned = simulations.owned_ranges(n, rank, world, interleaved=True, part=part)
            local = 0
            for a, b, la in owned:
                assert la == local and 0 <= a < b <= n
                seen[a:b] += 1
                local += b - a
        assert np.all(seen == 1)


def test_protocol_factorisation():
    """The rank-revealing factorisation behind the virtual-measurement path (host code): PGSE-type
    protocols have rank <= 3 whatever the number of measurements, the factors reproduce the
    gradient to 1e-13, near-low-rank and full-rank protocols are refused."""
    import ctypes
    from disimpy_b200 import _lib, gradients
    L = _lib.lib()

    def factor(g, max_rank=16):
        g = _lib.f64(g)
        m, t = g.shape[:2]
        rank = ctypes.c_int32(-1)
        u, v = np.zeros((m, max_rank)), np.zeros((max_rank, t, 3))
        _lib.check(L.dsb_protocol_factor(_lib.ptr(g), m, t, max_rank, ctypes.byref(rank), _lib.ptr(u), _lib.ptr(v)),
                   "dsb_protocol_factor")
        r = rank.value
        return r, u.ravel()[:m * r].reshape(m, r), v.ravel()[:r * t * 3].reshape(r, t, 3)

    rs = np.random.RandomState(2)
    bvecs = rs.normal(size=(180, 3))
    g, dt = gradients.pgse(10e-3, 30e-3, 400, np.repeat([1e9, 2e9, 3e9], 60), bvecs)
    r, u, v = factor(g)
    assert r == 3
    scale = np.abs(g).max()
    assert np.abs(np.einsum("mr,rtc->mtc", u, v) - g).max() < 1e-12 * scale
    g2, _ = gradients.pgse(5e-3, 35e-3, 400, np.repeat([1e9, 2e9, 3e9], 60), bvecs)
    r, u, v = factor(np.concatenate([g[:90], g2[90:]]))           # two timings
    assert r == 6 and np.abs(np.einsum("mr,rtc->mtc", u, v) - np.concatenate([g[:90], g2[90:]])).max() < 1e-12 * scale
    assert factor(g[:1] * np.linspace(0.1, 1, 50)[:, None, None])[0] == 1   # one waveform, 50 amplitudes
    assert factor(g + 1e-9 * scale * rs.normal(size=g.shape))[0] == 0       # noise at 1e-9: not low rank
    assert factor(rs.normal(size=(40, 30, 3)), max_rank=16)[0] == 0         # arbitrary waveforms
    assert factor(np.zeros((8, 10, 3)))[0] == 0                             # nothing to factor
    bad = g.copy()
    bad[3, 7, 1] = np.nan
    assert factor(bad)[0] == 0                                              # non-finite samples: general path


def test_sharded_sampler_row_assembly():
    """simulations._round_rows: the accepted points of every rank and round, concatenated round
    after round in rank order and cut to n points, land in the right local rows of every rank."""
    from disimpy_b200 import simulations
    rs = np.random.RandomState(3)
    for world in (1, 2, 3, 8):
        n = int(rs.randint(50, 400))
        rounds = []          # per round, per rank: the global "point ids" that rank accepted
        next_id, total = 0, 0
        while total < n:
            counts = [int(c) for c in rs.randint(0, 40, size=world)]
            rounds.append([list(range(next_id + sum(counts[:r]), next_id + sum(counts[:r + 1]))) for r in range(world)])
            next_id += sum(counts)
            total += sum(counts)
        for rank in range(world):
            lo, hi = simulations.shard_range(n, rank, world)
            mine = np.full(hi - lo, -1)
            have = 0
            for per_rank in rounds:
                counts = [len(x) for x in per_rank]
                for r, src, dst, k in simulations._round_rows(counts, have, lo, hi - lo):
                    mine[dst:dst + k] = per_rank[r][src:src + k]
                have += sum(counts)
            assert np.array_equal(mine, np.arange(lo, hi))


def test_tools_and_entry_points_compile():
    """Every script a maintainer or the driver runs on the GPU box is at least valid Python here
    (they cannot be executed without a GPU)."""
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "tools", "*.py"))) + [os.path.join(root, "bench.py"),
                                                                      os.path.join(root, "__graft_entry__.py")]
    assert len(files) > 5
    for f in files:
        py_compile.compile(f, doraise=True)


def test_gradient_helpers_bit_equal_to_reference(tmp_path):
    """pgse / set_b / calc_b / calc_q / interpolate_gradient return the reference's arrays bit for bit
    (tests/golden/ref_gradients.npz: outputs of disimpy/gradients.py on fixed inputs), so a script
    ported to this package feeds the walk the same gradient samples."""
    from disimpy_b200 import gradients
    r = np.load(os.path.join(GOLDEN, "ref_gradients.npz"))
    g, dt = gradients.pgse(5e-3, 20e-3, 37, r["bvals"], r["bvecs"])
    assert np.array_equal(g, r["pgse"]) and dt == float(r["pgse_dt"])
    assert np.array_equal(gradients.calc_b(r["w"], 1e-3), r["calc_b"])
    assert np.array_equal(gradients.calc_q(r["w"], 1e-3), r["calc_q"])
    assert np.array_equal(gradients.set_b(r["w"], 1e-3, np.arange(1, 6) * 1e9), r["set_b"])
    g, dt = gradients.interpolate_gradient(r["w"], 1e-3, 93)
    assert np.array_equal(g, r["interp"]) and dt == float(r["interp_dt"])
    # Camino scheme files (disimpy/gradients.py:182-212)
    path = tmp_path / "scheme.txt"
    rows = np.hstack([np.full((3, 1), 4.0), np.full((3, 1), 2e-3), np.arange(36.0).reshape(3, 12)])
    path.write_text("VERSION: GRADIENT_WAVEFORM\n" + "\n".join(" ".join(repr(float(x)) for x in row) for row in rows) + "\n")
    g, dt = gradients.load_camino_scheme_file(str(path))
    assert g.shape == (3, 4, 3) and dt == 2e-3 and np.array_equal(g.ravel(), np.arange(36.0))
    path.write_text("VERSION: 1\n")
    with pytest.raises(Exception):
        gradients.load_camino_scheme_file(str(path))


def test_part_edges_and_device_list(monkeypatch):
    """Host logic of the pipelined / device-list run: parts hold up to 131072 walkers (smaller first parts with several GPUs), are
    multiples of the kernel's 128-walker block, and tile the walkers; the device list follows the
    environment."""
    from disimpy_b200 import simulations as S
    for n, slots in ((1, 1), (16384, 1), (1_000_000, 1), (8_000_000, 8), (300_000, 3)):
        e = S.part_edges(n, slots)
        assert e[0] == 0 and e[-1] == n and all(a < b for a, b in zip(e[:-1], e[1:]))
        assert all(x % 128 == 0 for x in e[:-1])
        sizes = np.diff(e)
        assert sizes.max() <= 131072 and (slots == 1 or len(sizes) <= slots or sizes[slots - 1] <= 32768)
        seen = np.zeros(n, dtype=int)
        for k in range(slots):
            local = 0
            for a, b, la in S.owned_ranges(n, k, slots, interleaved=True):
                assert la == local and la % 128 == 0
                seen[a:b] += 1
                local += b - a
        assert np.all(seen == 1)
    monkeypatch.setattr(S, "_device_count", lambda: 4)
    for k in ("DISIMPY_B200_DEVICES", "DISIMPY_B200_DEVICE", "LOCAL_RANK", "DISIMPY_B200_MIN_WALKERS_PER_DEVICE"):
        monkeypatch.delenv(k, raising=False)
    assert S.local_devices() == [0, 1, 2, 3]                       # a plain script: every visible GPU
    assert S.local_devices(300_000) == [0, 1]                      # ... that gets at least 131072 walkers
    assert S.local_devices(10) == [0]
    monkeypatch.setenv("LOCAL_RANK", "6")
    assert S.local_devices() == [2] and S._device() == 2           # a launcher's rank, modulo the visible devices
    monkeypatch.setenv("DISIMPY_B200_DEVICE", "3")
    assert S.local_devices() == [3]
    monkeypatch.setenv("DISIMPY_B200_DEVICES", "1,3")
    assert S.local_devices() == [1, 3]
    monkeypatch.setenv("DISIMPY_B200_DEVICES", "all")
    assert S.local_devices() == [0, 1, 2, 3]
    # rows of several local handles and interleaved ranges go back into global order
    rows = S._assemble_rows([([(0, 2, 0), (4, 5, 2)], np.array([[0.0], [1.0], [4.0]])),
                             ([(2, 4, 0)], np.array([[2.0], [3.0]]))], 5, None)
    assert np.array_equal(rows[:, 0], np.arange(5.0))


def test_library_binds_nccl_at_run_time():
    """The path's one collective lives in the library (dsb_allreduce_signal); NCCL is bound by dlopen,
    so drawing a unique id needs no GPU -- and a missing NCCL is an error code with a text, not a crash."""
    import ctypes
    from disimpy_b200 import _lib, simulations
    L = _lib.lib()
    ident = np.zeros(128, dtype=np.uint8)
    path = simulations._nccl_library_path()
    rc = L.dsb_nccl_unique_id(path.encode() if path else None, _lib.ptr(ident))
    if rc == 0:
        assert ident.any()
    else:
        assert rc == 4 and b"NCCL" in L.dsb_last_error()
    assert L.dsb_nccl_init(None, 0, 3, 2, _lib.ptr(ident), ctypes.byref(ctypes.c_void_p())) == 1   # rank >= world: EINVAL
