"""ctypes front-end of the CPU parity checker (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product
(``disimpy_b200``) never does.  The arithmetic lives in ``disimpy_oracle.c``
(each function cites the reference file:line it restates); this file restates
the host-side orchestration of ``disimpy/simulations.py:1163-1429`` with plain
NumPy so that the golden vectors under ``tests/golden`` can be replayed from
the same inputs the reference was given.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdisimpy_oracle.so")
_LIB = None

SUBSTRATE_CODE = {"free": 0, "sphere": 1, "cylinder": 2, "ellipsoid": 3, "mesh": 4}


class Params(ctypes.Structure):
    _fields_ = [
        ("substrate", ctypes.c_int32),
        ("n_threads", ctypes.c_int32),
        ("n_walkers", ctypes.c_int64),
        ("n_meas", ctypes.c_int64),
        ("n_t", ctypes.c_int64),
        ("walker_offset", ctypes.c_int64),
        ("seed", ctypes.c_uint64),
        ("max_iter", ctypes.c_int64),
        ("step_l", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("epsilon", ctypes.c_double),
        ("radius", ctypes.c_double),
        ("R", ctypes.c_double * 9),
        ("R_inv", ctypes.c_double * 9),
        ("semiaxes", ctypes.c_double * 3),
        ("vertices", ctypes.c_void_p),
        ("faces", ctypes.c_void_p),
        ("xs", ctypes.c_void_p),
        ("ys", ctypes.c_void_p),
        ("zs", ctypes.c_void_p),
        ("len_xs", ctypes.c_int64),
        ("len_ys", ctypes.c_int64),
        ("len_zs", ctypes.c_int64),
        ("subvoxel_indices", ctypes.c_void_p),
        ("triangle_indices", ctypes.c_void_p),
        ("n_sv", ctypes.c_int64 * 3),
        ("perm_prob", ctypes.c_double),
    ]


def build():
    """Compile the C restatement (oracle/Makefile); building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
                os.path.join(_HERE, "disimpy_oracle.c")):
            build()
        _LIB = ctypes.CDLL(_SO)
        _LIB.oracle_simulate.restype = ctypes.c_int
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


UNIT_OPS = {"dot_product": (0, 6, 1), "cross_product": (1, 6, 3), "normalize_vector": (2, 3, 3),
            "triangle_normal": (3, 9, 3), "mat_mul": (4, 12, 3), "line_circle_intersection": (5, 5, 1),
            "line_sphere_intersection": (6, 7, 1), "line_ellipsoid_intersection": (7, 9, 1),
            "ray_triangle_intersection_check": (8, 15, 1), "reflection": (9, 11, 6), "crossing": (10, 11, 3),
            "ll_subvoxel_overlap": (11, 19, 1), "ul_subvoxel_overlap": (12, 19, 1),
            "ll_subvoxel_overlap_periodic": (13, 19, 1), "ul_subvoxel_overlap_periodic": (14, 19, 1)}


def unit(name, args):
    """One of the reference's device functions (disimpy/simulations.py:23-343, `_cuda_<name>`) on
    the rows of `args` (layout: oracle_unit in disimpy_oracle.c)."""
    op, n_in, n_out = UNIT_OPS[name]
    a = np.ascontiguousarray(np.atleast_2d(np.asarray(args, dtype=np.float64)))
    assert a.shape[1] == n_in, (name, a.shape)
    out = np.zeros((a.shape[0], n_out))
    L = lib()
    L.oracle_unit.restype = ctypes.c_int
    rc = L.oracle_unit(ctypes.c_int(op), ctypes.c_int64(a.shape[0]), _ptr(a), _ptr(out))
    assert rc == 0
    return out


def rng_states(seed, n, subsequence_start=0):
    """(n, 2) uint64 xoroshiro128+ states; numba/cuda/random.py:225-241."""
    out = np.zeros((n, 2), dtype=np.uint64)
    lib().oracle_rng_states(ctypes.c_uint64(seed), ctypes.c_uint64(subsequence_start),
                            ctypes.c_int64(n), _ptr(out))
    return out


def draw(state, n_normals, n_uniforms):
    """Advance one state: n_normals normal_float64 draws then n_uniforms uniform_float64."""
    st = np.array(state, dtype=np.uint64).copy()
    normals = np.zeros(max(n_normals, 1))
    uniforms = np.zeros(max(n_uniforms, 1))
    lib().oracle_draw(_ptr(st), ctypes.c_int64(n_normals), _ptr(normals),
                      ctypes.c_int64(n_uniforms), _ptr(uniforms))
    return normals[:n_normals], uniforms[:n_uniforms], st


def _fill_ball(rs, n, scale, dim, accept):
    """Sequential rejection sampling from the MT19937 stream, vectorised in blocks
    (same accept order as disimpy/simulations.py:353-399)."""
    out = np.zeros((0, dim))
    while len(out) < n:
        k = max(1024, int((n - len(out)) * 2.2))
        p = (rs.random_sample((k, dim)) - 0.5) * 2 * scale
        out = np.concatenate([out, p[accept(p)]])
    return out[:n]


def initial_positions(substrate, n_walkers, seed):
    """disimpy/simulations.py:1192, 1221-1227, 1262, 1297-1302 (host samplers use Numba's
    CPU MT19937 seeded by _set_seed(seed) == np.random.RandomState(seed))."""
    rs = np.random.RandomState(seed)
    if substrate.type == "free":
        return np.zeros((n_walkers, 3))
    if substrate.type == "sphere":
        r = substrate.radius
        return _fill_ball(rs, n_walkers, r, 3, lambda p: np.sqrt((p * p).sum(1)) < r)
    if substrate.type == "cylinder":
        r = substrate.radius
        R, R_inv = cylinder_rotations(substrate.orientation)
        pos = np.zeros((n_walkers, 3))
        pos[:, 1:3] = _fill_ball(rs, n_walkers, r, 2, lambda p: np.sqrt((p * p).sum(1)) < r)
        return np.matmul(R_inv, pos.T).T
    if substrate.type == "ellipsoid":
        ax = substrate.semiaxes
        q = _fill_ball(rs, n_walkers, ax, 3,
                       lambda p: ((p / ax) ** 2)[:, 0] + ((p / ax) ** 2)[:, 1] + ((p / ax) ** 2)[:, 2] < 1)
        return np.matmul(substrate.R, q.T).T
    raise ValueError(substrate.type)


def vec2vec_rotmat(v, k):
    """Rodrigues rotation taking v to k (disimpy/utils.py:11-42)."""
    v = v / np.linalg.norm(v)
    k = k / np.linalg.norm(k)
    axis = np.cross(v, k)
    if np.linalg.norm(axis) < np.finfo(float).eps:
        return -np.eye(3) if np.linalg.norm(v - k) > np.linalg.norm(v) else np.eye(3)
    axis /= np.linalg.norm(axis)
    angle = np.arccos(np.dot(v, k))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * np.matmul(K, K)


def cylinder_rotations(orientation):
    """disimpy/simulations.py:1221-1222."""
    R = vec2vec_rotmat(orientation, np.array([1.0, 0, 0]))
    return R, np.linalg.inv(R)


def make_params(substrate, n_walkers, n_meas, n_t, step_l, dt, seed, max_iter, epsilon,
                walker_offset=0, n_threads=1):
    p = Params()
    p.substrate = SUBSTRATE_CODE[substrate.type]
    p.n_threads = n_threads
    p.n_walkers, p.n_meas, p.n_t = n_walkers, n_meas, n_t
    p.walker_offset = walker_offset
    p.seed = seed
    p.max_iter = max_iter
    p.step_l, p.dt, p.epsilon = step_l, dt, epsilon
    keep = []
    if substrate.type == "sphere":
        p.radius = substrate.radius
    elif substrate.type == "cylinder":
        p.radius = substrate.radius
        R, R_inv = cylinder_rotations(substrate.orientation)
        p.R[:] = list(np.ascontiguousarray(R).ravel())
        p.R_inv[:] = list(np.ascontiguousarray(R_inv).ravel())
    elif substrate.type == "ellipsoid":
        p.semiaxes[:] = list(substrate.semiaxes)
        R_inv = substrate.R
        p.R_inv[:] = list(np.ascontiguousarray(R_inv).ravel())
        p.R[:] = list(np.ascontiguousarray(np.linalg.inv(R_inv)).ravel())
    elif substrate.type == "mesh":
        arrs = dict(
            vertices=np.ascontiguousarray(substrate.vertices, dtype=np.float64),
            faces=np.ascontiguousarray(substrate.faces, dtype=np.int64),
            xs=np.ascontiguousarray(substrate.xs, dtype=np.float64),
            ys=np.ascontiguousarray(substrate.ys, dtype=np.float64),
            zs=np.ascontiguousarray(substrate.zs, dtype=np.float64),
            subvoxel_indices=np.ascontiguousarray(substrate.subvoxel_indices, dtype=np.int64),
            triangle_indices=np.ascontiguousarray(substrate.triangle_indices, dtype=np.int64))
        for k, a in arrs.items():
            setattr(p, k, a.ctypes.data)
            keep.append(a)
        p.len_xs, p.len_ys, p.len_zs = len(arrs["xs"]), len(arrs["ys"]), len(arrs["zs"])
        p.n_sv[:] = [int(v) for v in substrate.n_sv]
        p.perm_prob = float(substrate.perm_prob)
    return p, keep


def run_walk(substrate, gradient, dt, diffusivity, positions, seed=123, max_iter=1000,
             epsilon=1e-13, walker_offset=0, traj=False, n_threads=1, rng=None):
    """Run the walk for all time steps.  Returns dict(positions, phases, iter_exc[, traj])."""
    gradient = np.ascontiguousarray(gradient, dtype=np.float64)
    n_meas, n_t = gradient.shape[:2]
    n = positions.shape[0]
    step_l = np.sqrt(6 * diffusivity * dt)
    p, keep = make_params(substrate, n, n_meas, n_t, step_l, dt, seed, max_iter, epsilon,
                          walker_offset, n_threads)
    pos = np.ascontiguousarray(positions, dtype=np.float64).copy()
    phases = np.zeros((n_meas, n))
    iter_exc = np.zeros(n, dtype=np.uint8)
    tr = np.zeros((n_t + 1, n, 3)) if traj else None
    st = None if rng is None else np.ascontiguousarray(rng, dtype=np.uint64).copy()
    rc = lib().oracle_simulate(ctypes.byref(p), _ptr(gradient), _ptr(pos), _ptr(phases),
                               _ptr(iter_exc), None if st is None else _ptr(st),
                               None if tr is None else _ptr(tr))
    if rc != 0:
        raise RuntimeError("oracle_simulate failed: %d" % rc)
    out = dict(positions=pos, phases=phases, iter_exc=iter_exc.astype(bool), rng=st)
    if traj:
        out["traj"] = tr
    return out


COUNTER_NAMES = ("checks", "collisions", "tri_tests", "cells", "searches", "steps")


def work_counters(reset=True):
    """What the reference's algorithm did since the last reset, summed over walkers and threads:
    distance checks (analytic substrates), collisions, ray-triangle tests, grid cells visited,
    collision searches (mesh) and walker-steps.  bench.py turns them into the algorithmic work
    per walker-step of its rooflines."""
    out = np.zeros(len(COUNTER_NAMES), dtype=np.int64)
    lib().oracle_counters(_ptr(out), ctypes.c_int(1 if reset else 0))
    return dict(zip(COUNTER_NAMES, (int(v) for v in out)))


def signals_from_phases(phases, iter_exc, all_signals=False):
    """disimpy/simulations.py:1413-1421."""
    ph = phases.copy()
    ph[:, np.where(iter_exc)[0]] = np.nan
    if all_signals:
        return np.real(np.exp(1j * ph))
    return np.real(np.nansum(np.exp(1j * ph), axis=1))


def fill_mesh(n_points, substrate, intra, seed, cuda_bs=128):
    """disimpy/simulations.py:505-579 (host driver) around oracle_fill_mesh_round."""
    import types
    gs = int(np.ceil(float(n_points) / cuda_bs))
    states = rng_states(seed, gs * cuda_bs)
    sub = substrate
    if not substrate.periodic:  # strip the 12 wall triangles, disimpy/simulations.py:531-546
        vertices = np.copy(substrate.vertices[0:-8])
        faces = np.copy(substrate.faces[0:-12])
        tri = np.copy(substrate.triangle_indices)
        svi = np.copy(substrate.subvoxel_indices)
        # The reference's loop shifts BOTH ends of every cell range whose end lies past a
        # deleted entry (so a cell that held a wall triangle also picks up its predecessor's
        # last entries), then clamps at 0; closed form of that loop:
        keep_mask = tri < len(faces)
        removed_before = np.concatenate([[0], np.cumsum(~keep_mask)])
        svi = svi - removed_before[svi[:, 1]][:, None]
        svi[svi < 0] = 0
        tri = tri[keep_mask]
        sub = types.SimpleNamespace(type="mesh", vertices=vertices, faces=faces, xs=substrate.xs,
                                    ys=substrate.ys, zs=substrate.zs, subvoxel_indices=svi,
                                    triangle_indices=tri, n_sv=substrate.n_sv, perm_prob=0.0)
    p, keep = make_params(sub, n_points, 1, 1, 0.0, 0.0, seed, 1, 0.0)
    voxel = np.ascontiguousarray(substrate.voxel_size, dtype=np.float64)
    points = np.zeros((0, 3))
    while points.shape[0] < n_points:
        new = np.full((n_points, 3), np.inf)
        lib().oracle_fill_mesh_round(ctypes.byref(p), ctypes.c_int(1 if intra else 0), _ptr(voxel),
                                     _ptr(new), ctypes.c_int64(n_points), _ptr(states))
        points = np.vstack((points, new[~np.isinf(new)[:, 0]]))
    return points[0:n_points]


def simulation(n_walkers, diffusivity, gradient, dt, substrate, seed=123, final_pos=False,
               all_signals=False, max_iter=1000, epsilon=1e-13, traj=False, n_threads=1):
    """Oracle equivalent of disimpy.simulations.simulation (disimpy/simulations.py:1051-1429)
    for parity tests: same seeds, same initial positions, same signal definition."""
    np.random.seed(seed)
    if substrate.type == "mesh":
        if isinstance(substrate.init_pos, np.ndarray):
            positions = substrate.init_pos
        elif substrate.init_pos == "uniform":
            positions = np.random.random((n_walkers, 3)) * substrate.voxel_size
        else:
            positions = fill_mesh(n_walkers, substrate, substrate.init_pos == "intra", seed)
    else:
        positions = initial_positions(substrate, n_walkers, seed)
    res = run_walk(substrate, gradient, dt, diffusivity, positions, seed, max_iter, epsilon,
                   traj=traj, n_threads=n_threads)
    res["initial_positions"] = positions
    res["signals"] = signals_from_phases(res["phases"], res["iter_exc"], all_signals)
    return res
