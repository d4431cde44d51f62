"""ctypes binding of libdisimpy_b200.so (the C ABI in include/disimpy_b200.h).

There is no fallback: if the shared library is missing the import fails loudly, and
if it is present but no CUDA device is, every GPU entry point returns an error code
that is raised as an exception.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DISIMPY_B200_LIB lets kernel experiments (tools/kbench.py) load an alternative build
LIB_PATH = os.environ.get("DISIMPY_B200_LIB") or os.path.join(_HERE, "libdisimpy_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class DsbMesh(ctypes.Structure):
    _fields_ = [
        ("vertices", ctypes.c_void_p),
        ("n_vertices", ctypes.c_int64),
        ("faces", ctypes.c_void_p),
        ("n_faces", ctypes.c_int64),
        ("xs", ctypes.c_void_p),
        ("ys", ctypes.c_void_p),
        ("zs", ctypes.c_void_p),
        ("subvoxel_indices", ctypes.c_void_p),
        ("triangle_indices", ctypes.c_void_p),
        ("n_triangle_indices", ctypes.c_int64),
        ("n_sv", ctypes.c_int64 * 3),
        ("perm_prob", ctypes.c_double),
    ]


class DsbParams(ctypes.Structure):
    _fields_ = [
        ("substrate", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("n_walkers", ctypes.c_int64),
        ("walker_offset", ctypes.c_int64),
        ("n_meas", ctypes.c_int64),
        ("n_t", ctypes.c_int64),
        ("seed", ctypes.c_uint64),
        ("max_iter", ctypes.c_int64),
        ("step_l", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("epsilon", ctypes.c_double),
        ("radius", ctypes.c_double),
        ("R", ctypes.c_double * 9),
        ("R_inv", ctypes.c_double * 9),
        ("semiaxes", ctypes.c_double * 3),
        ("mesh", DsbMesh),
    ]


SUBSTRATE_CODE = {"free": 0, "sphere": 1, "cylinder": 2, "ellipsoid": 3, "mesh": 4}

# every symbol include/disimpy_b200.h declares
EXPORTS = [
    "dsb_create", "dsb_set_positions", "dsb_set_positions_dev", "dsb_run", "dsb_sync",
    "dsb_get_signal", "dsb_get_positions", "dsb_get_phases", "dsb_get_iter_exc",
    "dsb_get_rng_states", "dsb_get_run_stats", "dsb_timer_start", "dsb_timer_stop",
    "dsb_measure_fp64_peak", "dsb_stream", "dsb_signal_dev", "dsb_destroy",
    "dsb_simulate", "dsb_rng_states", "dsb_fill_mesh", "dsb_mesh_subdivide",
    "dsb_mesh_subdivide_fetch", "dsb_triangle_box_overlap", "dsb_interval_sv_overlap",
    "dsb_host_fill", "dsb_device_count", "dsb_last_error", "dsb_version",
    "dsb_rewind", "dsb_set_positions_part", "dsb_run_part", "dsb_finish", "dsb_release_cache",
    "dsb_set_rng_states", "dsb_fill_mesh_sim", "dsb_protocol_rank", "dsb_protocol_factor", "dsb_set_rng_part", "dsb_host_sampler_create", "dsb_host_sampler_next", "dsb_host_sampler_destroy",
    "dsb_fill_shard_begin", "dsb_fill_shard_round", "dsb_fill_shard_end",
    "dsb_copy_signal_dev", "dsb_simulate_multi", "dsb_fill_mesh_multi", "dsb_selftest_sqrt", "dsb_measure_l2_peak", "dsb_selftest_device_function", "dsb_format_traj_line",
    "dsb_nccl_unique_id", "dsb_nccl_init", "dsb_allreduce_signal", "dsb_allreduce_zeros", "dsb_nccl_destroy",
]

_lib = None


class DsbError(RuntimeError):
    pass


def lib():
    """Load the native library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). disimpy_b200 has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.dsb_last_error.restype = ctypes.c_char_p
        L.dsb_version.restype = ctypes.c_char_p
        L.dsb_stream.restype = ctypes.c_void_p
        L.dsb_stream.argtypes = [ctypes.c_void_p]
        L.dsb_signal_dev.restype = ctypes.c_void_p
        L.dsb_signal_dev.argtypes = [ctypes.c_void_p]
        L.dsb_create.argtypes = [ctypes.POINTER(DsbParams), ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_void_p)]
        for name in ("dsb_set_positions", "dsb_set_positions_dev", "dsb_get_positions",
                     "dsb_get_phases", "dsb_get_iter_exc", "dsb_get_rng_states"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.dsb_run.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        L.dsb_sync.argtypes = [ctypes.c_void_p]
        L.dsb_destroy.argtypes = [ctypes.c_void_p]
        L.dsb_get_signal.argtypes = [ctypes.c_void_p, ctypes.c_void_p, c_int64_p]
        L.dsb_get_run_stats.argtypes = [ctypes.c_void_p, c_double_p, c_int64_p]
        L.dsb_timer_start.argtypes = [ctypes.c_void_p]
        L.dsb_timer_stop.argtypes = [ctypes.c_void_p, c_double_p]
        L.dsb_measure_fp64_peak.argtypes = [ctypes.c_int32, c_double_p]
        L.dsb_measure_l2_peak.argtypes = [ctypes.c_int32, c_double_p]
        L.dsb_simulate.argtypes = [ctypes.POINTER(DsbParams)] + [ctypes.c_void_p] * 3 + [
            c_int64_p] + [ctypes.c_void_p] * 3
        L.dsb_rng_states.argtypes = [ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64,
                                     ctypes.c_int64, ctypes.c_void_p]
        L.dsb_fill_mesh.argtypes = [ctypes.c_int32, ctypes.POINTER(DsbMesh), ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_uint64, ctypes.c_int64,
                                    ctypes.c_int64, ctypes.c_void_p]
        L.dsb_mesh_subdivide.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         c_int64_p, ctypes.POINTER(ctypes.c_void_p)]
        L.dsb_mesh_subdivide_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.dsb_host_fill.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64,
                                    ctypes.c_void_p, ctypes.c_void_p]
        L.dsb_protocol_rank.argtypes = [ctypes.c_void_p]
        L.dsb_protocol_factor.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                          ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p, ctypes.c_void_p]
        L.dsb_set_rng_part.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]
        L.dsb_rewind.argtypes = [ctypes.c_void_p]
        L.dsb_finish.argtypes = [ctypes.c_void_p]
        L.dsb_set_positions_part.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        L.dsb_run_part.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        L.dsb_set_rng_states.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.dsb_fill_mesh_sim.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]
        L.dsb_fill_shard_begin.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64]
        L.dsb_fill_shard_round.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, c_int64_p]
        L.dsb_fill_shard_end.argtypes = [ctypes.c_void_p]
        L.dsb_host_sampler_create.argtypes = [ctypes.c_int32, ctypes.c_uint64, ctypes.c_void_p,
                                              ctypes.POINTER(ctypes.c_void_p)]
        L.dsb_host_sampler_next.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.dsb_host_sampler_destroy.argtypes = [ctypes.c_void_p]
        L.dsb_copy_signal_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.dsb_simulate_multi.argtypes = [ctypes.POINTER(DsbParams), ctypes.c_void_p, ctypes.c_int32] + [
            ctypes.c_void_p] * 3 + [c_int64_p] + [ctypes.c_void_p] * 3
        L.dsb_fill_mesh_multi.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_uint64, ctypes.c_int64]
        L.dsb_selftest_sqrt.argtypes = [ctypes.c_int32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                        c_int64_p, c_double_p]
        L.dsb_selftest_device_function.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p,
                                                   ctypes.c_void_p]
        L.dsb_format_traj_line.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, c_int64_p]
        L.dsb_nccl_unique_id.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
        L.dsb_nccl_init.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_void_p)]
        L.dsb_allreduce_signal.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, c_int64_p]
        L.dsb_allreduce_zeros.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, c_int64_p]
        L.dsb_nccl_destroy.argtypes = [ctypes.c_void_p]
        L.dsb_release_cache.argtypes = []
        L.dsb_device_count.argtypes = [ctypes.POINTER(ctypes.c_int32)]
        L.dsb_triangle_box_overlap.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.dsb_interval_sv_overlap.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                              ctypes.c_double, c_int64_p, c_int64_p]
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dsb_last_error()
        raise DsbError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def mesh_struct(substrate, vertices=None, faces=None, triangle_indices=None,
                subvoxel_indices=None):
    """Fill a DsbMesh from a mesh substrate; returns (struct, keep_alive list)."""
    keep = [
        f64(substrate.vertices if vertices is None else vertices),
        i64(substrate.faces if faces is None else faces),
        f64(substrate.xs), f64(substrate.ys), f64(substrate.zs),
        i64(substrate.subvoxel_indices if subvoxel_indices is None else subvoxel_indices),
        i64(substrate.triangle_indices if triangle_indices is None else triangle_indices),
    ]
    m = DsbMesh()
    m.vertices, m.n_vertices = keep[0].ctypes.data, keep[0].shape[0]
    m.faces, m.n_faces = keep[1].ctypes.data, keep[1].shape[0]
    m.xs, m.ys, m.zs = keep[2].ctypes.data, keep[3].ctypes.data, keep[4].ctypes.data
    m.subvoxel_indices = keep[5].ctypes.data
    m.triangle_indices, m.n_triangle_indices = keep[6].ctypes.data, keep[6].shape[0]
    m.n_sv[:] = [int(v) for v in substrate.n_sv]
    m.perm_prob = float(substrate.perm_prob)
    return m, keep
