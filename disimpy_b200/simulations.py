"""``simulation()``: drop-in for disimpy.simulations.simulation (simulations.py:1051-1429).

Same signature, validation, prints, warnings, seeding side effects and return values as
the reference.  What changes is everything between "inputs are validated" and "signal is
returned": instead of one Numba launch + stream sync per time step, the walk, the phase
accumulation and the sum of cos(phase) run inside libdisimpy_b200.so (hand-written sm_100a
CUDA behind the C ABI in include/disimpy_b200.h).  There is no CPU fallback.

Multi-GPU, two ways (results do not depend on either, up to the floating-point summation
order of the signal):
* one process, several GPUs (a plain script; the reference user never launches with torchrun):
  the walkers are split over the devices of ``local_devices()`` -- every visible GPU unless
  DISIMPY_B200_DEVICES / DISIMPY_B200_DEVICE say otherwise --, one handle per device driven
  from this thread (all launches are asynchronous), signals summed on the host;
* one process per GPU under ``torch.distributed``: every rank simulates the contiguous walker
  range ``[rank*N/W, (rank+1)*N/W)`` (or, when the initial positions come from the sequential
  host stream, every W-th part of the stream) with the walkers' global RNG subsequences, and
  the signal (+ valid-walker count) is summed with ONE all-reduce on the device buffer the
  library fills; the mesh sampler's threads are dealt to the ranks and its accepted points
  all-gathered.
"""

import ctypes
import math
import os
import sys
import warnings

import numpy as np

from . import _lib, substrates, utils
from .gradients import GAMMA  # noqa: F401  (same module-level name as the reference)
from .substrates import _aabb_to_mesh  # noqa: F401  (simulations.py:582-613 is a second copy of it there)


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


_SEED = None


def _set_seed(seed):
    """simulations.py:346-350: the seed of the host samplers below when they are called without
    one, as the reference's are (there it seeds Numba's CPU generator)."""
    global _SEED
    _SEED = int(seed)


def _stream_seed(seed):
    """The MT19937 stream a sampler call draws from: the given seed (what simulation() passes), else
    the one of _set_seed, else a fresh one from NumPy's global generator (the reference's samplers
    then continue whatever state Numba's generator is in; here every call starts its stream)."""
    if seed is not None:
        return seed
    return _SEED if _SEED is not None else int(np.random.randint(0, 2 ** 32, dtype=np.uint64))


def _fill_circle(n, radius, seed=None):
    """n points uniform in a disc (simulations.py:353-366) from the MT19937 stream of ``seed``."""
    return _host_fill(0, n, _stream_seed(seed), radius, 2)


def _fill_sphere(n, radius, seed=None):
    """n points uniform in a ball (simulations.py:369-382)."""
    return _host_fill(1, n, _stream_seed(seed), radius, 3)


def _fill_ellipsoid(n, semiaxes, seed=None):
    """n points uniform in an axis-aligned ellipsoid (simulations.py:385-399)."""
    return _host_fill(2, n, _stream_seed(seed), semiaxes, 3)


def _initial_positions_cylinder(n_walkers, radius, R, seed=None):
    """Points in the cross-section of a cylinder, rotated to the lab frame by R
    (simulations.py:402-409)."""
    positions = np.zeros((n_walkers, 3))
    positions[:, 1:3] = _fill_circle(n_walkers, radius, seed)
    return np.matmul(R, positions.T).T


def _initial_positions_ellipsoid(n_walkers, semiaxes, R, seed=None):
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


def _traj_line(positions):
    """The text of one time point as the reference writes it (simulations.py:1043-1048): str(v) + " "
    for every value of positions.ravel(), then a newline -- formatted natively (dsb_format_traj_line:
    the same characters, ~40 times faster than a str() per value)."""
    v = _lib.f64(positions).ravel()
    buf = np.empty(26 * v.size + 1, dtype=np.uint8)
    n = ctypes.c_int64(0)
    _lib.check(_lib.lib().dsb_format_traj_line(_lib.ptr(v), v.size, _lib.ptr(buf), buf.size, ctypes.byref(n)), "dsb_format_traj_line")
    return memoryview(buf[:n.value])


_TRAJ_CHUNK = 1 << 22   # values formatted per call (26 bytes of buffer each)


def _write_traj(traj, mode, positions):
    """One line per time point: x y z of walker 1, walker 2, ... (simulations.py:1043-1048)."""
    v = _lib.f64(positions).ravel()
    with open(traj, mode + "b") as f:
        for a in range(0, max(v.size, 1), _TRAJ_CHUNK):
            line = _traj_line(v[a:a + _TRAJ_CHUNK])
            f.write(line if a + _TRAJ_CHUNK >= v.size else line[:-1])   # one newline, after the last value


def _device_count():
    count = ctypes.c_int32(0)
    if _lib.lib().dsb_device_count(ctypes.byref(count)) != 0:
        return 0
    return count.value


def _device():
    """CUDA device ordinal of this process: DISIMPY_B200_DEVICE, else LOCAL_RANK (modulo the number
    of visible devices: launchers that give every task one visible GPU still count LOCAL_RANK up),
    else 0."""
    if os.environ.get("DISIMPY_B200_DEVICE", "") != "":
        return int(os.environ["DISIMPY_B200_DEVICE"])
    if os.environ.get("LOCAL_RANK", "") != "":
        return int(os.environ["LOCAL_RANK"]) % max(_device_count(), 1)
    return 0


# a device takes part in a single-process run only if it gets at least this many walkers
# (DISIMPY_B200_MIN_WALKERS_PER_DEVICE overrides: tests run several handles on small problems)
_MIN_WALKERS_PER_DEVICE = 131072


def local_devices(n_walkers=None):
    """The device list of a single-process run.  DISIMPY_B200_DEVICES = "all" or a
    comma-separated list of ordinals picks it; otherwise a process that was given one device
    (DISIMPY_B200_DEVICE, or LOCAL_RANK from a launcher) uses that one, and a plain script uses
    every visible GPU.  With ``n_walkers``: no more devices than keep every shard at
    _MIN_WALKERS_PER_DEVICE walkers or more."""
    env = os.environ.get("DISIMPY_B200_DEVICES", "").strip()
    if env and env != "all":
        devs = [int(x) for x in env.split(",") if x.strip() != ""]
    elif env == "all" or (os.environ.get("DISIMPY_B200_DEVICE", "") == "" and os.environ.get("LOCAL_RANK", "") == ""):
        devs = list(range(max(_device_count(), 1)))
    else:
        devs = [_device()]
    if n_walkers is not None:
        floor = int(os.environ.get("DISIMPY_B200_MIN_WALKERS_PER_DEVICE", _MIN_WALKERS_PER_DEVICE))
        devs = devs[:max(1, min(len(devs), n_walkers // max(floor, 1)))]
    return devs


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


_COMMS = {}   # (world, rank, device) -> library communicator (ctypes.c_void_p), or None when unavailable


def _nccl_library_path():
    """The NCCL the ranks should all load: the one bundled with torch when there is one (it is the
    one torch.distributed itself runs on), else whatever the loader finds (libnccl.so.2)."""
    try:
        import nvidia.nccl
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


def library_comm(dist, rank, world):
    """The library's own NCCL communicator of this process for the path's one collective
    (dsb_allreduce_signal), created on first use: rank 0 draws the NCCL unique id, torch.distributed
    only carries its 128 bytes to the other ranks.  None (on every rank alike) when
    DISIMPY_B200_NCCL=torch or when any rank failed to set it up: the caller then reduces through
    torch.distributed."""
    key = (world, rank, _device())
    if key in _COMMS:
        return _COMMS[key]
    import torch
    comm, ok = ctypes.c_void_p(), os.environ.get("DISIMPY_B200_NCCL", "") != "torch"
    path = _nccl_library_path()
    cpath = path.encode() if path else None
    ident = np.zeros(128, dtype=np.uint8)
    if ok and rank == 0:
        ok = _lib.lib().dsb_nccl_unique_id(cpath, _lib.ptr(ident)) == 0
    box = [ident.tobytes() if ok else None]
    dist.broadcast_object_list(box, src=0)
    ok = ok and box[0] is not None
    if ok:
        ident = np.frombuffer(box[0], dtype=np.uint8).copy()
        ok = _lib.lib().dsb_nccl_init(cpath, _device(), rank, world, _lib.ptr(ident), ctypes.byref(comm)) == 0
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda:%d" % _device())
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)     # all ranks or none
    if int(flag.item()) != 1:
        if ok:
            _lib.lib().dsb_nccl_destroy(comm)
        comm = None
    _COMMS[key] = comm
    return comm


def _destroy_comms():
    for comm in _COMMS.values():
        if comm:
            try:
                _lib.lib().dsb_nccl_destroy(comm)
            except Exception:
                pass
    _COMMS.clear()


import atexit  # noqa: E402
atexit.register(_destroy_comms)


def shard_range(n_walkers, rank, world_size):
    """Contiguous global walker range of a rank."""
    return n_walkers * rank // world_size, n_walkers * (rank + 1) // world_size


_PART = 131072        # walkers per part: about one full wave of 128-walker blocks on a B200
_PART_FIRST = 32768   # several GPUs: the first parts are smaller, so that every GPU has work within a few ms


def part_edges(n_walkers, n_slots=1, part=None):
    """Boundaries of the parts a pipelined run is dealt in: [0, e1, e2, ..., n_walkers].  A fixed
    ``part`` gives equal parts.  By default one GPU gets parts of _PART walkers (measured on a B200:
    smaller first parts leave the SMs underfilled for the first milliseconds, 198.4 against 194.2 ms
    for 1e6 walkers x 1e4 steps); with several GPUs the first ``n_slots`` parts hold _PART_FIRST
    walkers and every further round of ``n_slots`` parts twice as many, up to _PART, so that the last
    GPU does not wait for n_slots full parts of the sequential stream.  All multiples of 128, the walk
    kernel's block."""
    if part is not None:
        return list(range(0, n_walkers, part)) + [n_walkers]
    edges, size = [0], int(os.environ.get("DISIMPY_B200_PART_FIRST", _PART if n_slots == 1 else _PART_FIRST))
    while edges[-1] < n_walkers:
        left = n_walkers - edges[-1]
        if left < n_slots * size:
            # the last round is split evenly: a round-robin deal of equal parts with a ragged end gives one
            # GPU up to a whole part more than another (measured at 2 GPUs x 1e6 walkers: 1 048 576 against
            # 951 424 walkers, 204 ms against 192 ms)
            size = max(128, (left + 128 * n_slots - 1) // (128 * n_slots) * 128)
        for _ in range(n_slots):
            if edges[-1] < n_walkers:
                edges.append(min(edges[-1] + size, n_walkers))
        size = min(2 * size, _PART)
    return edges


def owned_ranges(n_walkers, rank, world_size, interleaved=False, part=None):
    """Which global walkers a rank (or a device of a single-process run) holds and where:
    [(global_lo, global_hi, local_lo), ...].

    Contiguous (default): the one range of shard_range.  Interleaved: the parts of part_edges are
    dealt round-robin -- used when the initial positions come from the sequential host stream, so
    that everybody can start on a first part at once instead of after everything that precedes a
    contiguous shard in the stream."""
    if not interleaved:
        lo, hi = shard_range(n_walkers, rank, world_size)
        return [(lo, hi, 0)]
    edges = part_edges(n_walkers, world_size, part)
    out, local = [], 0
    for k, (a, b) in enumerate(zip(edges[:-1], edges[1:])):
        if k % world_size == rank:
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

    def allreduce_signal(self, comm):
        """Global (signal, valid count): NCCL all-reduce of the handle's result buffer inside the library."""
        sig, n_valid = np.zeros(self.n_meas), ctypes.c_int64(0)
        _lib.check(self._L.dsb_allreduce_signal(self._h, comm, _lib.ptr(sig), ctypes.byref(n_valid)), "dsb_allreduce_signal")
        return sig, n_valid.value

    def copy_signal_to(self, dev_ptr):
        """The n_meas + 1 result doubles (sum cos, valid count) into a device buffer of the caller."""
        _lib.check(self._L.dsb_copy_signal_dev(self._h, ctypes.c_void_p(dev_ptr)), "dsb_copy_signal_dev")

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
    # nothing needs all of them up front (no trajectory file) they are drawn part by part while
    # the GPU already walks the earlier parts (the progress display then counts walkers, not
    # time steps).
    pipelined = substrate.type in ("sphere", "cylinder", "ellipsoid") and not traj
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

    trace = _Trace()
    shards = _Shards(substrate, n_walkers, gradient, dt, step_l, seed, max_iter, epsilon, pipelined)
    dist, rank = shards.dist, shards.rank
    trace("handles")
    try:
        if traj and rank == 0 and not device_fill:
            _write_traj(traj, "w", positions)
        if pipelined:
            _walk_pipelined(shards, substrate, seed, trace, progress=not quiet and rank == 0)
        elif device_fill:
            shards.fill_mesh(substrate, seed, cuda_bs)
            if traj:
                start = shards.rows(lambda w: w.positions())
                if rank == 0:
                    _write_traj(traj, "w", start)
        else:
            shards.set_positions(positions)
        trace("positions")
        if pipelined:
            pass
        elif traj:
            for t in range(n_t):
                shards.run(t, t + 1)
                step_pos = shards.rows(lambda w: w.positions())
                if rank == 0:
                    _write_traj(traj, "a", step_pos)
                if not quiet:
                    sys.stdout.write(f"\r{np.round((t / n_t) * 100, 1)}%")
                    sys.stdout.flush()
        elif quiet:
            shards.run(0, n_t)
        else:  # a handful of launches so that progress can be shown, cut on multiples of 16 steps (the
            # many-measurement kernels work in chunks of 16 steps, 8 for a mesh)
            edges = np.linspace(0, n_t, min(n_t, 20) + 1).astype(int)
            edges[1:-1] = (edges[1:-1] + 8) // 16 * 16
            edges = np.unique(np.clip(edges, 0, n_t))
            for t0, t1 in zip(edges[:-1], edges[1:]):
                sys.stdout.write(f"\r{np.round((t0 / n_t) * 100, 1)}%")
                sys.stdout.flush()
                shards.run(int(t0), int(t1))
                shards.sync()
        trace("submitted")

        # The signal kernel also counts the walkers whose iter_exc flag is clear, so the per-walker
        # flags only travel to the host when something was flagged (or when the caller asked for
        # per-walker output).  One collective for signal and count together.
        signals, n_valid = shards.signal(trace)
        n_flagged = n_walkers - n_valid
        iter_exc_all = None
        if n_flagged > 0 or all_signals:
            iter_exc_all = shards.rows(lambda w: w.iter_exc())
        if n_flagged > 0:
            warnings.warn(
                "Maximum number of iterations was exceeded in the intersection "
                + "check algorithm for walkers %s" % np.where(iter_exc_all)[0])

        if all_signals:
            phases = shards.rows(lambda w: w.phases().T).T
            phases[:, np.where(iter_exc_all)[0]] = np.nan
            signals = np.real(np.exp(1j * phases))
        if not quiet:
            sys.stdout.write("\rSimulation finished\n")
            sys.stdout.flush()
        if final_pos:
            final = shards.rows(lambda w: w.positions())
            trace("done")
            return signals, final
        trace("done")
        return signals
    finally:
        shards.close()
        trace.report()


class _Trace:
    """DISIMPY_B200_TRACE=1: host time stamps of the phases of a simulation() call, printed to stderr
    as one JSON line per call (where does the end-to-end time go next to the kernel time)."""

    def __init__(self):
        self.on = os.environ.get("DISIMPY_B200_TRACE", "") not in ("", "0")
        if self.on:
            import time
            self._clock = time.perf_counter
            self.t0 = self._clock()
            self.marks = []

    def __call__(self, label):
        if self.on:
            self.marks.append((label, round(1e3 * (self._clock() - self.t0), 3)))

    def report(self):
        if self.on:
            import json
            sys.stderr.write("disimpy_b200 trace (ms since entry): " + json.dumps(self.marks) + "\n")
            sys.stderr.flush()


class _Shards:
    """The handles this process drives: one per device of local_devices() in a single-process run, or
    the one of a rank under torch.distributed.  Slot k of n_slots (= ranks x local devices) holds the
    global walkers owned_ranges(n_walkers, k, n_slots, interleaved)."""

    def __init__(self, substrate, n_walkers, gradient, dt, step_l, seed, max_iter, epsilon, pipelined):
        self.rank, self.world, self.dist = _dist()
        devices = [_device()] if self.dist is not None else local_devices(n_walkers)
        self.n_walkers, self.n_meas = n_walkers, gradient.shape[0]
        self.n_slots = self.world * len(devices)
        # the parts of a pipelined run are dealt round-robin (if there are enough of them)
        self.interleaved = (pipelined and self.n_slots > 1
                            and len(part_edges(n_walkers, self.n_slots)) - 1 >= self.n_slots)
        self.walks, self.owned = [], []
        try:
            for j, dev in enumerate(devices):
                owned = owned_ranges(n_walkers, self.rank * len(devices) + j, self.n_slots, self.interleaved)
                n_local = sum(b - a for a, b, _ in owned)
                walk = None
                if n_local > 0:   # (fewer walkers than slots: the empty ones contribute nothing)
                    params, keep = make_params(substrate, n_local, owned[0][0], gradient, dt, step_l, seed, max_iter,
                                               epsilon, device=dev)
                    walk = Walk(params, gradient)
                    if self.interleaved:
                        for a, b, la in owned:
                            walk.set_rng_part(la, la + b - a, a)
                self.walks.append(walk)
                self.owned.append(owned)
        except Exception:
            self.close()
            raise

    def live(self):
        return [(w, o) for w, o in zip(self.walks, self.owned) if w is not None]

    def set_positions(self, positions):
        for w, owned in self.live():
            (lo, hi, _), = owned
            w.set_positions(positions[lo:hi])

    def run(self, t0, t1):
        for w, _ in self.live():
            w.run(t0, t1)

    def sync(self):
        for w, _ in self.live():
            w.sync()

    def fill_mesh(self, substrate, seed, cuda_bs):
        """Initial positions of a periodic mesh, drawn on the GPU(s) and left there."""
        intra = substrate.init_pos == "intra"
        live = self.live()
        if self.dist is not None:
            if self.dist.get_backend() == "nccl" and self.n_walkers >= _SHARDED_FILL_MIN and all(w is not None for w in self.walks):
                (lo, hi, _), = self.owned[0]
                _fill_mesh_sharded(self.walks[0], substrate, self.n_walkers, lo, hi - lo, seed, self.rank, self.world,
                                   self.dist)
            else:
                for w, ((lo, _, _),) in live:
                    w.fill_mesh(substrate.voxel_size, intra, seed, self.n_walkers, lo, cuda_bs)
        elif len(live) == 1:
            live[0][0].fill_mesh(substrate.voxel_size, intra, seed, self.n_walkers, live[0][1][0][0], cuda_bs)
        else:   # the device list of one process: the sampler's threads are dealt to the devices (in the library)
            handles = (ctypes.c_void_p * len(live))(*[w._h for w, _ in live])
            voxel = _lib.f64(substrate.voxel_size)
            _lib.check(_lib.lib().dsb_fill_mesh_multi(handles, len(live), _lib.ptr(voxel), 1 if intra else 0, seed,
                                                      self.n_walkers), "dsb_fill_mesh_multi")

    def signal(self, trace=None):
        """(sum over ALL walkers of cos(phase) per measurement, number of unflagged walkers): local
        handles summed in slot order, then one all-reduce over the ranks."""
        total = np.zeros(self.n_meas + 1)
        live = self.live()
        comm = library_comm(self.dist, self.rank, self.world) if (
            self.dist is not None and self.dist.get_backend() == "nccl") else None
        if comm:
            # the library reduces its own result buffer over the ranks (NCCL, in place, on its stream)
            sig, n_valid = np.zeros(self.n_meas), ctypes.c_int64(0)
            if live:
                _lib.check(_lib.lib().dsb_allreduce_signal(live[0][0]._h, comm, _lib.ptr(sig), ctypes.byref(n_valid)),
                           "dsb_allreduce_signal")
            else:
                _lib.check(_lib.lib().dsb_allreduce_zeros(_device(), comm, self.n_meas, _lib.ptr(sig), ctypes.byref(n_valid)),
                           "dsb_allreduce_zeros")
            total[:-1], total[-1] = sig, n_valid.value
            if trace:
                trace("walk finished + all-reduce")
        elif self.dist is not None and self.dist.get_backend() == "nccl":
            # (fallback: torch.distributed reduces a tensor the library copies its result doubles into)
            import torch
            t = torch.zeros(self.n_meas + 1, dtype=torch.float64, device="cuda:%d" % _device())
            if live:
                torch.cuda.current_stream(t.device).synchronize()   # (the zero fill precedes the library's copy)
                live[0][0].copy_signal_to(t.data_ptr())
            if trace:
                trace("walk finished")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            total = t.cpu().numpy()
        else:
            for w, _ in live:
                sig, n_valid = w.signal()
                total[:-1] += sig
                total[-1] += n_valid
            if trace:
                trace("walk finished")
            if self.dist is not None:
                total = _allreduce_sum(total, self.dist)
        if trace:
            trace("signal reduced")
        return total[:-1].copy(), int(round(total[-1]))

    def rows(self, fetch):
        """Per-walker rows (fetch(walk) -> array with one row per local walker) of all walkers, in
        global order, on every rank (only used for final_pos / all_signals / traj / the iter_exc
        warning: per-shard host gathers, no device collective)."""
        pieces = [(owned, fetch(w)) for w, owned in self.live()]
        if self.dist is not None and self.dist.get_backend() == "nccl" and len(self.walks) == 1:
            return self._rows_nccl(pieces)
        return _assemble_rows(pieces, self.n_walkers, self.dist)

    def _rows_nccl(self, pieces):
        """The same gather as _assemble_rows over NCCL as ONE all-gather of equally padded blocks instead
        of pickled objects; every rank's walker ranges follow from owned_ranges, so only rows travel."""
        import torch
        ranges = [owned_ranges(self.n_walkers, r, self.world, self.interleaved) for r in range(self.world)]
        cap = max(sum(b - a for a, b, _ in rr) for rr in ranges)
        # the row layout (trailing shape, dtype) is the same on every rank; a rank without walkers learns it from the others
        meta = [(pieces[0][1].shape[1:], pieces[0][1].dtype.str) if pieces else None]
        allmeta = [None] * self.world
        self.dist.all_gather_object(allmeta, meta[0])
        shape, dtype = next(m for m in allmeta if m is not None)
        dtype = np.dtype(dtype)
        width = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        dev = torch.device("cuda", _device())
        block = torch.zeros((cap, max(width, 1)), dtype=torch.uint8, device=dev)
        if pieces:
            raw = np.ascontiguousarray(pieces[0][1]).view(np.uint8).reshape(len(pieces[0][1]), -1)
            block[:len(raw)] = torch.from_numpy(raw).to(dev)
        everyone = torch.empty((self.world, cap, max(width, 1)), dtype=torch.uint8, device=dev)
        self.dist.all_gather_into_tensor(everyone, block)
        host = everyone.cpu().numpy()
        full = np.zeros((self.n_walkers,) + tuple(shape), dtype=dtype)
        flat = full.view(np.uint8).reshape(self.n_walkers, -1) if width else None
        for r, rr in enumerate(ranges):
            for a, b, la in rr:
                if width:
                    flat[a:b] = host[r, la:la + b - a, :width]
        return full

    def close(self):
        for w in self.walks:
            if w is not None:
                w.close()
        self.walks = []


def _stream_sampler(substrate, seed):
    """(sampler over the sequential host stream of an analytic substrate, function that turns a
    stretch of the stream into lab-frame positions): simulations.py:346-418."""
    if substrate.type == "sphere":
        sampler, to_lab = _HostSampler(1, seed, substrate.radius, 3), None
    elif substrate.type == "cylinder":
        R = utils.vec2vec_rotmat(substrate.orientation, np.array([1.0, 0, 0]))
        sampler, to_lab = _HostSampler(0, seed, substrate.radius, 2), np.linalg.inv(R)
    else:
        sampler, to_lab = _HostSampler(2, seed, substrate.semiaxes, 3), substrate.R

    def finish(pts, alone=False):
        if substrate.type == "cylinder":
            body = np.zeros((len(pts), 3))
            body[:, 1:3] = pts
            pts = body
        if to_lab is not None:
            if len(pts) == 1 and not alone:
                # BLAS multiplies a single column through another routine (gemv) than a matrix (gemm), with a
                # last-bit difference: a one-walker stretch of a longer run must round like the reference's product
                # over all walkers, so it is multiplied as a two-column matrix
                pts = np.matmul(to_lab, np.vstack([pts, pts]).T).T[:1]
            else:
                pts = np.matmul(to_lab, pts.T).T
        return pts
    return sampler, finish


def _stream_stretches(substrate, seed, stretches, n_total=None):
    """Initial positions of the global walkers [a, b) of every (a, b) in ``stretches`` (ascending, disjoint),
    from ONE pass over the sequential host stream: yields one (b - a, 3) array per stretch.  Concatenated
    over [0, n) they are what _fill_sphere / _initial_positions_cylinder / _initial_positions_ellipsoid
    return in one go (simulations.py:346-418); walkers between the stretches are drawn and dropped.  ``n_total``:
    the number of walkers of the whole run (a run of ONE walker rotates its position like the reference does)."""
    sampler, finish = _stream_sampler(substrate, seed)
    try:
        at = 0
        for a, b in stretches:
            sampler.skip(a - at)
            yield finish(sampler.next(b - a), alone=n_total == 1)
            at = b
    finally:
        sampler.close()


def _walk_pipelined(shards, substrate, seed, trace=None, progress=False):
    """Walks every part over all time steps as soon as its positions are there: the host draws
    the next part while the GPUs work.  Same positions, same walk, same signal as drawing
    everything first.  One pass over the sequential stream serves every local handle; stretches
    that belong to other ranks are drawn and dropped."""
    jobs = []   # (global lo, global hi, handle, local lo), in stream order
    for w, owned in shards.live():
        w.rewind()
        if shards.interleaved:
            jobs += [(a, b, w, la) for a, b, la in owned]
        else:      # one contiguous shard per handle, walked in parts of growing size
            (lo, hi, _), = owned
            edges = part_edges(hi - lo)
            jobs += [(lo + a, lo + b, w, a) for a, b in zip(edges[:-1], edges[1:])]
    jobs.sort(key=lambda j: j[0])
    todo, done = max(sum(b - a for a, b, _, _ in jobs), 1), 0
    stretches = _stream_stretches(substrate, seed, [(a, b) for a, b, _, _ in jobs], shards.n_walkers)
    for k, (a, b, w, la) in enumerate(jobs):
        if progress:   # the reference's progress line (simulations.py:1209), by walkers handed to the GPU
            sys.stdout.write(f"\r{np.round((done / todo) * 100, 1)}%")
            sys.stdout.flush()
        w.set_positions_part(la, la + b - a, next(stretches))
        w.run_part(la, la + b - a)
        done += b - a
        if trace and k == 0:
            trace("first part submitted")
    stretches.close()
    for w, _ in shards.live():
        w.finish()


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


def _assemble_rows(pieces, n_total, dist):
    """pieces: [(owned ranges, rows of those walkers), ...] held by this process -> the rows of all
    n_total walkers in global order, on every rank."""
    if dist is not None:
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, pieces)
        pieces = [p for rank_pieces in out for p in rank_pieces]
    shape, dtype = next(((rows.shape[1:], rows.dtype) for _, rows in pieces), ((), np.float64))
    full = np.zeros((n_total,) + shape, dtype=dtype)
    for ranges, rows in pieces:
        for a, b, la in ranges:
            full[a:b] = rows[la:la + b - a]
    return full

