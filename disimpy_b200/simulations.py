"""``simulation()``: drop-in for disimpy.simulations.simulation (simulations.py:1051-1429).

Same signature, validation, prints, warnings, seeding side effects and return values as
the reference.  What changes is everything between "inputs are validated" and "signal is
returned": instead of one Numba launch + stream sync per time step, the walk, the phase
accumulation and the sum of cos(phase) run inside libdisimpy_b200.so (hand-written sm_100a
CUDA behind the C ABI in include/disimpy_b200.h).  There is no CPU fallback.

Multi-GPU: when ``torch.distributed`` is initialised, every rank simulates the contiguous
walker range ``[rank*N/W, (rank+1)*N/W)`` (or, when the initial positions come from the
sequential host stream, every W-th part of 131072 walkers) with the walkers' global RNG
subsequences, and the signal (+ valid-walker count) is summed with one all-reduce; the mesh
sampler's threads are dealt to the ranks and its accepted points all-gathered.  Results do not
depend on the number of ranks (up to floating-point summation order of the signal).
"""

import ctypes
import math
import os
import sys
import warnings

import numpy as np

from . import _lib, substrates, utils
from .gradients import GAMMA  # noqa: F401  (same module-level name as the reference)


# ----------------------------------------------------------------------------------------
# host-side initial positions (simulations.py:346-418): sequential rejection sampling from
# the MT19937 stream seeded with ``seed`` -- Numba's CPU generator after _set_seed(seed) is
# np.random.RandomState(seed) -- done natively in csrc/dsb_hostfill.cpp, same acceptance order.

def _host_fill(shape, n, seed, scale, dim):
    out = np.zeros((n, dim))
    sc = _lib.f64(np.atleast_1d(scale))
    rc = _lib.lib().dsb_host_fill(shape, n, seed, _lib.ptr(sc), _lib.ptr(out))
    if rc != 0:
        raise ValueError("Seed must be between 0 and 2**32 - 1")
    return out


class _HostSampler:
    """The same MT19937 rejection stream as _host_fill, handed out in consecutive stretches, so
    that simulation() can start the walk of the first walkers while the positions of the next
    ones are still being drawn (the stream is sequential; the GPU would otherwise idle)."""

    def __init__(self, shape, seed, scale, dim):
        self._h = ctypes.c_void_p()
        self._dim = dim
        sc = _lib.f64(np.atleast_1d(scale))
        if _lib.lib().dsb_host_sampler_create(shape, seed, _lib.ptr(sc), ctypes.byref(self._h)) != 0:
            raise ValueError("Seed must be between 0 and 2**32 - 1")

    def next(self, n):
        out = np.zeros((n, self._dim))
        _lib.check(_lib.lib().dsb_host_sampler_next(self._h, n, _lib.ptr(out)), "dsb_host_sampler_next")
        return out

    def skip(self, n, block=1 << 20):
        while n > 0:
            self.next(min(n, block))
            n -= block

    def close(self):
        if self._h:
            _lib.lib().dsb_host_sampler_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _fill_circle(n, radius, seed):
    """n points uniform in a disc (simulations.py:353-366) from the MT19937 stream of ``seed``."""
    return _host_fill(0, n, seed, radius, 2)


def _fill_sphere(n, radius, seed):
    """n points uniform in a ball (simulations.py:369-382)."""
    return _host_fill(1, n, seed, radius, 3)


def _fill_ellipsoid(n, semiaxes, seed):
    """n points uniform in an axis-aligned ellipsoid (simulations.py:385-399)."""
    return _host_fill(2, n, seed, semiaxes, 3)


def _initial_positions_cylinder(n_walkers, radius, R, seed):
    """Points in the cross-section of a cylinder, rotated to the lab frame by R
    (simulations.py:402-409)."""
    positions = np.zeros((n_walkers, 3))
    positions[:, 1:3] = _fill_circle(n_walkers, radius, seed)
    return np.matmul(R, positions.T).T


def _initial_positions_ellipsoid(n_walkers, semiaxes, R, seed):
    """Points in an ellipsoid, rotated to the lab frame by R (simulations.py:412-418)."""
    return np.matmul(R, _fill_ellipsoid(n_walkers, semiaxes, seed).T).T


def _fill_mesh(n_points, substrate, intra, seed, cuda_bs=128):
    """Uniform points inside / outside the closed surface of a mesh substrate
    (simulations.py:505-579), sampled on the GPU with the reference's RNG streams and accept
    order.  Non-periodic substrates are sampled against the mesh without its 12 wall
    triangles, with the reference's index bookkeeping (simulations.py:531-546)."""
    if substrate.periodic:
        m, keep = _lib.mesh_struct(substrate)
    else:
        n_faces = len(substrate.faces) - 12
        tri = np.asarray(substrate.triangle_indices)
        svi = np.array(substrate.subvoxel_indices, dtype=np.int64)
        is_wall = tri >= n_faces
        # the reference shifts both ends of a cell range by the number of removed entries
        # that precede the range's END, then clamps at 0
        removed_before = np.concatenate([[0], np.cumsum(is_wall)])
        svi = svi - removed_before[svi[:, 1]][:, None]
        svi[svi < 0] = 0
        m, keep = _lib.mesh_struct(substrate, vertices=substrate.vertices[0:-8],
                                   faces=substrate.faces[0:-12], triangle_indices=tri[~is_wall],
                                   subvoxel_indices=svi)
    voxel = _lib.f64(substrate.voxel_size)
    points = np.zeros((n_points, 3))
    _lib.check(_lib.lib().dsb_fill_mesh(_device(), ctypes.byref(m), _lib.ptr(voxel),
                                        1 if intra else 0, seed, n_points, cuda_bs,
                                        _lib.ptr(points)), "dsb_fill_mesh")
    return points


# ----------------------------------------------------------------------------------------

def add_noise_to_data(data, sigma, seed=None):
    """Add Rician noise to data (simulations.py:1016-1040)."""
    if seed:
        np.random.seed(seed)
    real = np.random.normal(size=data.shape, scale=sigma, loc=0)
    imag = np.random.normal(size=data.shape, scale=sigma, loc=0)
    return np.abs(data + real + 1j * imag)


def _write_traj(traj, mode, positions):
    """One line per time point: x y z of walker 1, walker 2, ... (simulations.py:1043-1048)."""
    with open(traj, mode) as f:
        f.write("".join(str(v) + " " for v in positions.ravel()))
        f.write("\n")


def _device():
    """CUDA device ordinal of this process: DISIMPY_B200_DEVICE, else LOCAL_RANK, else 0."""
    for name in ("DISIMPY_B200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(name, "") != "":
            return int(os.environ[name])
    return 0


def _require_gpu():
    count = ctypes.c_int32(0)
    try:
        rc = _lib.lib().dsb_device_count(ctypes.byref(count))
    except ImportError:
        raise
    if rc != 0 or count.value < 1:
        raise Exception(
            "disimpy_b200 was unable to detect a CUDA GPU. To run the simulation, a B200 "
            "(sm_100a) and a working CUDA driver are required; there is no CPU fallback.")


def _dist():
    """(rank, world_size, torch.distributed or None)"""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover - torch is part of the image
        return 0, 1, None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size(), dist
    return 0, 1, None


def shard_range(n_walkers, rank, world_size):
    """Contiguous global walker range of a rank."""
    return n_walkers * rank // world_size, n_walkers * (rank + 1) // world_size


def owned_ranges(n_walkers, rank, world_size, interleaved=False, part=None):
    """Which global walkers a rank holds and where: [(global_lo, global_hi, local_lo), ...].

    Contiguous (default): the one range of shard_range.  Interleaved: the parts of ``part``
    walkers are dealt round-robin to the ranks -- used when the initial positions come from the
    sequential host stream, so that every rank can start on its first part after 1/world_size of
    the wait a contiguous shard at the end of the stream would have."""
    if not interleaved:
        lo, hi = shard_range(n_walkers, rank, world_size)
        return [(lo, hi, 0)]
    part = _PART if part is None else part
    out, local = [], 0
    for k, a in enumerate(range(0, n_walkers, part)):
        if k % world_size == rank:
            b = min(a + part, n_walkers)
            out.append((a, b, local))
            local += b - a
    return out


def make_params(substrate, n_walkers, walker_offset, gradient, dt, step_l, seed, max_iter,
                epsilon, device=None):
    """Fill the C ABI's dsb_params for one shard; returns (params, keep_alive)."""
    p = _lib.DsbParams()
    p.substrate = _lib.SUBSTRATE_CODE[substrate.type]
    p.device = _device() if device is None else device
    p.n_walkers = n_walkers
    p.walker_offset = walker_offset
    p.n_meas, p.n_t = gradient.shape[0], gradient.shape[1]
    p.seed = seed
    p.max_iter = max_iter
    p.step_l, p.dt, p.epsilon = float(step_l), float(dt), float(epsilon)
    keep = []
    if substrate.type in ("sphere", "cylinder"):
        p.radius = substrate.radius
    if substrate.type == "cylinder":
        # lab -> cylinder frame and back (simulations.py:1221-1222)
        R = utils.vec2vec_rotmat(substrate.orientation, np.array([1.0, 0, 0]))
        R_inv = np.linalg.inv(R)
        p.R[:] = list(_lib.f64(R).ravel())
        p.R_inv[:] = list(_lib.f64(R_inv).ravel())
    if substrate.type == "ellipsoid":
        # substrate.R is ellipsoid -> lab (simulations.py:1297-1299)
        p.semiaxes[:] = list(substrate.semiaxes)
        p.R_inv[:] = list(_lib.f64(substrate.R).ravel())
        p.R[:] = list(_lib.f64(np.linalg.inv(substrate.R)).ravel())
    if substrate.type == "mesh":
        p.mesh, keep = _lib.mesh_struct(substrate)
    return p, keep


class Walk:
    """Thin RAII wrapper of one ``dsb_sim`` handle (one shard on one GPU)."""

    def __init__(self, params, gradient):
        self.params = params
        self.n_walkers, self.n_meas, self.n_t = params.n_walkers, params.n_meas, params.n_t
        self._h = ctypes.c_void_p()
        self._L = _lib.lib()
        g = _lib.f64(gradient)
        _lib.check(self._L.dsb_create(ctypes.byref(params), _lib.ptr(g), ctypes.byref(self._h)),
                   "dsb_create")

    def set_positions(self, positions):
        pos = _lib.f64(positions)
        _lib.check(self._L.dsb_set_positions(self._h, _lib.ptr(pos)), "dsb_set_positions")

    def set_positions_dev(self, dev_ptr):
        _lib.check(self._L.dsb_set_positions_dev(self._h, ctypes.c_void_p(dev_ptr)),
                   "dsb_set_positions_dev")

    def run(self, t0=0, t1=None):
        _lib.check(self._L.dsb_run(self._h, t0, self.n_t if t1 is None else t1), "dsb_run")

    def rewind(self):
        _lib.check(self._L.dsb_rewind(self._h), "dsb_rewind")

    def set_positions_part(self, w0, w1, positions):
        pos = _lib.f64(positions)
        _lib.check(self._L.dsb_set_positions_part(self._h, w0, w1, _lib.ptr(pos)), "dsb_set_positions_part")

    def run_part(self, w0, w1):
        _lib.check(self._L.dsb_run_part(self._h, w0, w1), "dsb_run_part")

    def finish(self):
        _lib.check(self._L.dsb_finish(self._h), "dsb_finish")

    def set_rng_part(self, w0, w1, global_offset):
        """Local walkers [w0, w1) are the global walkers global_offset ... (their RNG subsequences)."""
        _lib.check(self._L.dsb_set_rng_part(self._h, w0, w1, global_offset), "dsb_set_rng_part")

    def sync(self):
        _lib.check(self._L.dsb_sync(self._h), "dsb_sync")

    def signal(self):
        sig = np.zeros(self.n_meas)
        n_valid = ctypes.c_int64(0)
        _lib.check(self._L.dsb_get_signal(self._h, _lib.ptr(sig), ctypes.byref(n_valid)),
                   "dsb_get_signal")
        return sig, n_valid.value

    def positions(self):
        out = np.zeros((self.n_walkers, 3))
        _lib.check(self._L.dsb_get_positions(self._h, _lib.ptr(out)), "dsb_get_positions")
        return out

    def phases(self):
        out = np.zeros((self.n_meas, self.n_walkers))
        _lib.check(self._L.dsb_get_phases(self._h, _lib.ptr(out)), "dsb_get_phases")
        return out

    def iter_exc(self):
        out = np.zeros(self.n_walkers, dtype=np.uint8)
        _lib.check(self._L.dsb_get_iter_exc(self._h, _lib.ptr(out)), "dsb_get_iter_exc")
        return out.astype(bool)

    def rng_states(self):
        out = np.zeros((self.n_walkers, 2), dtype=np.uint64)
        _lib.check(self._L.dsb_get_rng_states(self._h, _lib.ptr(out)), "dsb_get_rng_states")
        return out

    def fill_mesh(self, voxel_size, intra, seed, n_points, first, cuda_bs):
        """Initial positions = points [first, first + n_walkers) of the reference's mesh sampler
        (simulations.py:505-579), drawn on the GPU against the mesh this handle holds and left
        there."""
        voxel = _lib.f64(voxel_size)
        _lib.check(self._L.dsb_fill_mesh_sim(self._h, _lib.ptr(voxel), 1 if intra else 0, seed, n_points, first,
                                             cuda_bs), "dsb_fill_mesh_sim")

    def fill_shard_begin(self, seed, thread_begin, thread_end):
        _lib.check(self._L.dsb_fill_shard_begin(self._h, seed, thread_begin, thread_end), "dsb_fill_shard_begin")

    def fill_shard_round(self, voxel_size, intra, accepted_dev_ptr):
        """One round of the sampler for this handle's threads; returns how many points were
        accepted (written to device memory at ``accepted_dev_ptr``, thread order)."""
        voxel = _lib.f64(voxel_size)
        n = ctypes.c_int64(0)
        _lib.check(self._L.dsb_fill_shard_round(self._h, _lib.ptr(voxel), 1 if intra else 0,
                                                ctypes.c_void_p(accepted_dev_ptr), ctypes.byref(n)),
                   "dsb_fill_shard_round")
        return n.value

    def fill_shard_end(self):
        _lib.check(self._L.dsb_fill_shard_end(self._h), "dsb_fill_shard_end")

    def protocol_rank(self):
        """r > 0: the protocol's gradient matrix has rank r <= 4 and the walk carries r virtual
        measurements (see include/disimpy_b200.h); 0: general path."""
        return int(self._L.dsb_protocol_rank(self._h))

    def set_rng_states(self, states):
        st = np.ascontiguousarray(states, dtype=np.uint64)
        _lib.check(self._L.dsb_set_rng_states(self._h, _lib.ptr(st)), "dsb_set_rng_states")

    def run_stats(self):
        ms = ctypes.c_double(0)
        n = ctypes.c_int64(0)
        _lib.check(self._L.dsb_get_run_stats(self._h, ctypes.byref(ms), ctypes.byref(n)),
                   "dsb_get_run_stats")
        return ms.value, n.value

    def timer_start(self):
        _lib.check(self._L.dsb_timer_start(self._h), "dsb_timer_start")

    def timer_stop(self):
        ms = ctypes.c_double(0)
        _lib.check(self._L.dsb_timer_stop(self._h, ctypes.byref(ms)), "dsb_timer_stop")
        return ms.value

    def close(self):
        if self._h:
            self._L.dsb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rng_states(seed, n, subsequence_start=0, device=None):
    """(n, 2) uint64 xoroshiro128+ states, what numba's create_xoroshiro128p_states(n, seed,
    subsequence_start) holds (numba/cuda/random.py:225-292), derived on the GPU."""
    out = np.zeros((n, 2), dtype=np.uint64)
    _lib.check(_lib.lib().dsb_rng_states(_device() if device is None else device, seed,
                                         subsequence_start, n, _lib.ptr(out)), "dsb_rng_states")
    return out


def simulation(
    n_walkers,
    diffusivity,
    gradient,
    dt,
    substrate,
    seed=123,
    traj=None,
    final_pos=False,
    all_signals=False,
    quiet=False,
    cuda_bs=128,
    max_iter=int(1e3),
    epsilon=1e-13,
):
    """Simulate a diffusion-weighted MR experiment and generate signal.

    Parameters and return values are those of the reference (simulations.py:1066-1113):
    ``gradient`` has shape (n_measurements, n_time_points, 3) in T/m, ``dt`` is the time step
    in s, ``substrate`` comes from :mod:`disimpy_b200.substrates`.  Returns ``signals``
    (n_measurements,) -- or the per-walker signals (n_measurements, n_walkers) when
    ``all_signals`` -- and additionally the final positions (n_walkers, 3) when ``final_pos``.
    ``cuda_bs`` is accepted for compatibility; it only fixes the number of RNG streams of the
    'intra'/'extra' mesh sampler like it does in the reference.
    """
    _require_gpu()

    if not isinstance(n_walkers, int) or n_walkers <= 0:
        raise ValueError("Incorrect value (%s) for n_walkers" % n_walkers)
    if not isinstance(diffusivity, float) or diffusivity <= 0:
        raise ValueError("Incorrect value (%s) for diffusivity" % diffusivity)
    if (not isinstance(gradient, np.ndarray) or gradient.ndim != 3 or gradient.shape[2] != 3
            or not np.issubdtype(gradient.dtype, np.floating)):
        raise ValueError("Incorrect value (%s) for gradient" % gradient)
    if not isinstance(dt, float) or dt <= 0:
        raise ValueError("Incorrect value (%s) for dt" % dt)
    if not isinstance(substrate, substrates._Substrate):
        raise ValueError("Incorrect value (%s) for substrate" % substrate)
    if not isinstance(seed, int) or seed < 0:
        raise ValueError("Incorrect value (%s) for seed" % seed)
    if traj:
        if not isinstance(traj, str):
            raise ValueError("Incorrect value (%s) for traj" % traj)
    if not isinstance(quiet, bool):
        raise ValueError("Incorrect value (%s) for quiet" % quiet)
    if not isinstance(cuda_bs, int) or cuda_bs <= 0:
        raise ValueError("Incorrect value (%s) for cuda_bs" % cuda_bs)
    if not isinstance(max_iter, int) or max_iter < 1:
        raise ValueError("Incorrect value (%s) for max_iter" % max_iter)
    if substrate.type not in _lib.SUBSTRATE_CODE:
        raise ValueError("Incorrect value (%s) for substrate" % substrate)

    n_t = gradient.shape[1]
    if not quiet:
        print("Starting simulation")
        if traj:
            print("The trajectories file will be up to %s GB" % (n_t * n_walkers * 3 * 25 / 1e9))

    # Same seeding side effect as the reference (simulations.py:1169-1170): NumPy's global
    # generator is reseeded; the host samplers use an MT19937 stream with the same seed.
    np.random.seed(seed)
    step_l = np.sqrt(6 * diffusivity * dt)

    if not quiet:
        print("Number of random walkers = %s" % n_walkers)
        print("Number of steps = %s" % n_t)
        print("Step length = %s m" % step_l)
        print("Step duration = %s s" % dt)

    # Analytic substrates draw their initial positions from one sequential host stream.  When
    # nothing needs all of them up front (no trajectory file, no progress display) they are
    # drawn part by part while the GPU already walks the earlier parts.
    pipelined = substrate.type in ("sphere", "cylinder", "ellipsoid") and not traj and quiet
    positions = None
    device_fill = False
    if pipelined:
        pass
    elif substrate.type == "free":
        positions = np.zeros((n_walkers, 3))
    elif substrate.type == "cylinder":
        R = utils.vec2vec_rotmat(substrate.orientation, np.array([1.0, 0, 0]))
        positions = _initial_positions_cylinder(n_walkers, substrate.radius, np.linalg.inv(R), seed)
    elif substrate.type == "sphere":
        positions = _fill_sphere(n_walkers, substrate.radius, seed)
    elif substrate.type == "ellipsoid":
        positions = _initial_positions_ellipsoid(n_walkers, substrate.semiaxes, substrate.R, seed)
    else:
        if isinstance(substrate.init_pos, np.ndarray):
            if n_walkers != substrate.init_pos.shape[0]:
                raise ValueError("n_walkers must be equal to the number of initial positions")
            positions = substrate.init_pos
        else:
            if not quiet:
                print("Calculating initial positions")
            if substrate.init_pos == "uniform":
                positions = np.random.random((n_walkers, 3)) * substrate.voxel_size
            elif substrate.periodic:
                # drawn on the GPU against the mesh the walk's handle holds, and left there
                device_fill = True
            elif substrate.init_pos == "intra":
                positions = _fill_mesh(n_walkers, substrate, True, seed, cuda_bs)
            else:
                positions = _fill_mesh(n_walkers, substrate, False, seed, cuda_bs)
            if not quiet:
                print("Finished calculating initial positions")

    rank, world, dist = _dist()
    # the parts of a pipelined run are dealt round-robin to the ranks (if there are enough of them)
    interleaved = pipelined and world > 1 and (n_walkers + _PART - 1) // _PART >= world
    owned = owned_ranges(n_walkers, rank, world, interleaved)
    lo, hi = owned[0][0], owned[0][1]          # (the contiguous shard when not interleaved)
    n_local = sum(b - a for a, b, _ in owned)
    if traj and rank == 0 and not device_fill:
        _write_traj(traj, "w", positions)

    params, keep = make_params(substrate, n_local, lo, gradient, dt, step_l, seed, max_iter,
                               epsilon)
    walk = Walk(params, gradient)
    try:
        if interleaved:
            for a, b, la in owned:
                walk.set_rng_part(la, la + b - a, a)
        if pipelined:
            _walk_pipelined(walk, substrate, owned, seed)
        elif device_fill:
            if world > 1 and dist.get_backend() == "nccl" and n_walkers >= _SHARDED_FILL_MIN:
                _fill_mesh_sharded(walk, substrate, n_walkers, lo, n_local, seed, rank, world, dist)
            else:
                walk.fill_mesh(substrate.voxel_size, substrate.init_pos == "intra", seed, n_walkers, lo, cuda_bs)
            if traj:
                start = _gather_rows(walk.positions(), n_walkers, owned, dist)
                if rank == 0:
                    _write_traj(traj, "w", start)
        else:
            walk.set_positions(positions[lo:hi])
        if pipelined:
            pass
        elif traj:
            for t in range(n_t):
                walk.run(t, t + 1)
                step_pos = _gather_rows(walk.positions(), n_walkers, owned, dist)
                if rank == 0:
                    _write_traj(traj, "a", step_pos)
                if not quiet:
                    sys.stdout.write(f"\r{np.round((t / n_t) * 100, 1)}%")
                    sys.stdout.flush()
        elif quiet:
            walk.run(0, n_t)
        else:  # a handful of launches so that progress can be shown (cut on multiples of 8 steps:
            # the many-measurement kernels work in 8-step chunks)
            edges = np.linspace(0, n_t, min(n_t, 20) + 1).astype(int)
            edges[1:-1] = (edges[1:-1] + 4) // 8 * 8
            edges = np.unique(np.clip(edges, 0, n_t))
            for t0, t1 in zip(edges[:-1], edges[1:]):
                sys.stdout.write(f"\r{np.round((t0 / n_t) * 100, 1)}%")
                sys.stdout.flush()
                walk.run(int(t0), int(t1))
                walk.sync()

        # The signal kernel also counts the walkers whose iter_exc flag is clear, so the
        # per-walker flags only travel to the host when something was flagged (or when the
        # caller asked for per-walker output).
        if all_signals:
            iter_exc = walk.iter_exc()
            n_flagged = int(iter_exc.sum())
        else:
            signals, n_valid = walk.signal()
            n_flagged = n_local - n_valid
            iter_exc = None
        if dist is not None:
            n_flagged = int(round(_allreduce_sum(np.array([float(n_flagged)]), dist)[0]))
        if n_flagged > 0:
            if iter_exc is None:
                iter_exc = walk.iter_exc()
            iter_exc_all = _gather_rows(iter_exc, n_walkers, owned, dist)
            warnings.warn(
                "Maximum number of iterations was exceeded in the intersection "
                + "check algorithm for walkers %s" % np.where(iter_exc_all)[0])

        if all_signals:
            phases = walk.phases()
            phases[:, np.where(iter_exc)[0]] = np.nan
            signals = np.real(np.exp(1j * phases))
            signals = _gather_rows(signals.T, n_walkers, owned, dist).T
        elif dist is not None:
            signals = _allreduce_sum(signals, dist)
        if not quiet:
            sys.stdout.write("\rSimulation finished\n")
            sys.stdout.flush()
        if final_pos:
            final = _gather_rows(walk.positions(), n_walkers, owned, dist)
            return signals, final
        return signals
    finally:
        walk.close()


_PART = 131072  # walkers per part: about one full wave of 128-walker blocks on a B200


def _position_parts(substrate, lo, hi, seed, part=_PART, wanted=None):
    """Initial positions of global walkers [lo, hi) of an analytic substrate, part by part:
    yields (a, b, positions of local walkers [a, b)).  Concatenated, the parts are what
    _fill_sphere / _initial_positions_cylinder / _initial_positions_ellipsoid return in one go
    (simulations.py:346-418).  Parts for which ``wanted(a, b)`` is false are drawn (the stream is
    sequential) but not yielded."""
    if substrate.type == "sphere":
        sampler, to_lab = _HostSampler(1, seed, substrate.radius, 3), None
    elif substrate.type == "cylinder":
        R = utils.vec2vec_rotmat(substrate.orientation, np.array([1.0, 0, 0]))
        sampler, to_lab = _HostSampler(0, seed, substrate.radius, 2), np.linalg.inv(R)
    else:
        sampler, to_lab = _HostSampler(2, seed, substrate.semiaxes, 3), substrate.R
    try:
        sampler.skip(lo)
        n = hi - lo
        for a in range(0, n, part):
            b = min(a + part, n)
            pts = sampler.next(b - a)
            if wanted is not None and not wanted(a, b):
                continue
            if substrate.type == "cylinder":
                body = np.zeros((b - a, 3))
                body[:, 1:3] = pts
                pts = body
            if to_lab is not None:
                pts = np.matmul(to_lab, pts.T).T
            yield a, b, pts
    finally:
        sampler.close()


def _walk_pipelined(walk, substrate, owned, seed):
    """Walks every part over all time steps as soon as its positions are there: the host draws
    the next part while the GPU works.  Same positions, same walk, same signal as drawing
    everything first.  ``owned``: the rank's walkers (owned_ranges)."""
    walk.rewind()
    if len(owned) == 1:   # one contiguous shard: skip to it, then part by part
        lo, hi, _ = owned[0]
        for a, b, pts in _position_parts(substrate, lo, hi, seed):
            walk.set_positions_part(a, b, pts)
            walk.run_part(a, b)
    else:                 # parts dealt round-robin: draw the whole stream, keep this rank's parts
        local_of = {a: la for a, _, la in owned}
        for a, b, pts in _position_parts(substrate, 0, owned[-1][1], seed, wanted=lambda a, b: a in local_of):
            la = local_of[a]
            walk.set_positions_part(la, la + b - a, pts)
            walk.run_part(la, la + b - a)
    walk.finish()


# from this many walkers on, the ranks of a multi-GPU run share the work of the mesh sampler
_SHARDED_FILL_MIN = 1 << 20


def _round_rows(counts, have, lo, n_local):
    """Where a round's accepted points go: rank r accepted counts[r] points, in thread order; they
    become the global points have, have + 1, ... in rank order.  Yields (rank, first row in that
    rank's block, first local row, how many) for the pieces that fall into [lo, lo + n_local)."""
    start = have
    for r, c in enumerate(counts):
        a, b = max(start, lo), min(start + c, lo + n_local)
        if a < b:
            yield r, a - start, a - lo, b - a
        start += c


def _fill_mesh_sharded(walk, substrate, n_points, lo, n_local, seed, rank, world, dist):
    """The reference's mesh sampler (disimpy/simulations.py:505-579) with its threads dealt to the
    ranks: in every round rank r evaluates threads shard_range(n_points, r, world), the accepted
    points of all ranks are all-gathered (NCCL) and concatenated in rank = thread order, and each
    rank keeps the rows [lo, lo + n_local) of the first n_points.  Same points, same order as one
    GPU drawing everything; 1/world of the ray tests per rank."""
    import torch
    dev = torch.device("cuda", _device())
    intra = substrate.init_pos == "intra"
    t0, t1 = shard_range(n_points, rank, world)
    cap = max(b - a for a, b in (shard_range(n_points, r, world) for r in range(world)))
    mine = torch.empty((n_local, 3), dtype=torch.float64, device=dev)
    accepted = torch.empty((cap, 3), dtype=torch.float64, device=dev)
    everyone = torch.empty((world, cap, 3), dtype=torch.float64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    walk.fill_shard_begin(seed, t0, t1)
    try:
        have = 0
        for _ in range(100000):
            n_acc = walk.fill_shard_round(substrate.voxel_size, intra, accepted.data_ptr())
            dist.all_gather_into_tensor(counts, torch.tensor([n_acc], dtype=torch.int64, device=dev))
            dist.all_gather_into_tensor(everyone, accepted)
            round_counts = counts.tolist()
            for r, src, dst, k in _round_rows(round_counts, have, lo, n_local):
                mine[dst:dst + k] = everyone[r, src:src + k]
            have += sum(round_counts)
            # the library writes `accepted` on its own stream in the next round: the all-gather
            # and the copies above must be through with it
            torch.cuda.synchronize(dev)
            if have >= n_points:
                break
        else:
            raise RuntimeError("fill_mesh: no acceptable points (is the surface closed?)")
    finally:
        walk.fill_shard_end()
    walk.set_positions_dev(mine.data_ptr())
    walk.sync()


def _allreduce_sum(values, dist):
    """Sum a small float64 vector over ranks (NCCL on the GPU when that backend is active)."""
    import torch
    dev = "cuda:%d" % _device() if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(np.ascontiguousarray(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def _gather_rows(local, n_total, owned, dist):
    """Assemble per-walker rows from all ranks (only used for final_pos / all_signals / traj
    / the iter_exc warning -- per-shard host gathers, no device collective).  ``owned``: this
    rank's walkers as owned_ranges returns them."""
    if dist is None:
        return local
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, (owned, local))
    full = np.zeros((n_total,) + local.shape[1:], dtype=local.dtype)
    for ranges, rows in out:
        for a, b, la in ranges:
            full[a:b] = rows[la:la + b - a]
    return full
