"""Substrate objects: drop-in for disimpy/substrates.py:11-269.

``free() / sphere() / cylinder() / ellipsoid() / mesh()`` return a ``_Substrate`` with
the same ``.type`` and attributes as the reference, validated the same way (same
``ValueError`` texts).  The only heavy part, binning the mesh triangles into the
subvoxel grid (``_mesh_space_subdivision``, disimpy/substrates.py:467-536), runs in the
native library (csrc/dsb_subdivide.cpp) and returns the reference's arrays element for
element.
"""

import ctypes

import numpy as np

from . import _lib


class _Substrate:
    """Attribute bag describing the simulated microstructure (substrates.py:11-44)."""

    def __init__(self, substrate_type, **kwargs):
        self.type = substrate_type
        if substrate_type in ("sphere", "cylinder"):
            self.radius = kwargs["radius"]
        if substrate_type == "cylinder":
            self.orientation = kwargs["orientation"]
        if substrate_type == "ellipsoid":
            self.semiaxes = kwargs["semiaxes"]
            self.R = kwargs["R"]
        if substrate_type == "mesh":
            for name in ("vertices", "faces", "voxel_size", "periodic", "init_pos", "n_sv",
                         "perm_prob"):
                setattr(self, name, kwargs[name])
            if not kwargs["quiet"]:
                print("Dividing the mesh into subvoxels")
            (self.xs, self.ys, self.zs, self.triangle_indices,
             self.subvoxel_indices) = _mesh_space_subdivision(
                 self.vertices, self.faces, self.voxel_size, self.n_sv)
            if not kwargs["quiet"]:
                print("Finished dividing the mesh into subvoxels")


def _is_float_array(a, shape):
    return (isinstance(a, np.ndarray) and a.shape == shape
            and np.issubdtype(a.dtype, np.floating))


def free():
    """Substrate for free diffusion."""
    return _Substrate("free")


def sphere(radius):
    """Substrate for diffusion inside a sphere of the given radius (m)."""
    if not isinstance(radius, float) or radius <= 0:
        raise ValueError(f"Incorrect value ({radius}) for radius")
    return _Substrate("sphere", radius=radius)


def cylinder(radius, orientation):
    """Substrate for diffusion inside an infinite cylinder of the given radius whose axis
    points along ``orientation`` (float array of shape (3,), normalised here)."""
    if not isinstance(radius, float) or radius <= 0:
        raise ValueError(f"Incorrect value ({radius}) for radius")
    if not _is_float_array(orientation, (3,)):
        raise ValueError(f"Incorrect value ({orientation}) for orientation")
    return _Substrate("cylinder", radius=radius,
                      orientation=orientation / np.linalg.norm(orientation))


def ellipsoid(semiaxes, R=np.eye(3)):
    """Substrate for diffusion inside an ellipsoid with the given semi-axes, rotated by the
    rotation matrix ``R`` (ellipsoid frame -> lab frame)."""
    if not _is_float_array(semiaxes, (3,)):
        raise ValueError(f"Incorrect value ({semiaxes}) for semiaxes")
    if not _is_float_array(R, (3, 3)):
        raise ValueError(f"Incorrect value ({R}) for R")
    if not np.isclose(np.linalg.det(R), 1) or not np.all(np.isclose(R.T, np.linalg.inv(R))):
        raise ValueError(f"R ({R}) is not a valid rotation matrix")
    return _Substrate("ellipsoid", semiaxes=semiaxes, R=R)


def mesh(vertices, faces, periodic, padding=np.zeros(3), init_pos="uniform",
         n_sv=np.array([50, 50, 50]), quiet=False, perm_prob=0):
    """Substrate for diffusion restricted by a triangular mesh (substrates.py:143-269).

    The simulated voxel is the bounding box of the triangles plus ``padding`` on every side,
    moved so that its lower corner is the origin.  ``periodic=False`` closes the voxel with
    12 impermeable wall triangles; ``init_pos`` is an (n_walkers, 3) array or one of
    'uniform', 'intra', 'extra'; ``n_sv`` is the subvoxel grid used to accelerate collision
    checks; ``perm_prob`` is the probability that a walker passes through a triangle.
    """
    if not (isinstance(vertices, np.ndarray) and vertices.ndim == 2 and vertices.shape[1] == 3
            and np.issubdtype(vertices.dtype, np.floating)):
        raise ValueError(f"Incorrect value ({vertices}) for vertices.")
    if not (isinstance(faces, np.ndarray) and faces.ndim == 2 and faces.shape[1] == 3
            and np.issubdtype(faces.dtype, np.integer)):
        raise ValueError(f"Incorrect value ({faces}) for faces.")
    if not isinstance(periodic, bool):
        raise ValueError(f"Incorrect value ({periodic}) for periodic")
    if not _is_float_array(padding, (3,)):
        raise ValueError(f"Incorrect value ({padding}) for padding")
    if isinstance(init_pos, np.ndarray):
        if (init_pos.ndim != 2 or init_pos.shape[1] != 3
                or not np.issubdtype(init_pos.dtype, np.floating)):
            raise ValueError(f"Incorrect value ({init_pos}) for init_pos")
    elif not (isinstance(init_pos, str) and init_pos in ("uniform", "intra", "extra")):
        raise ValueError(f"Incorrect value ({init_pos}) for init_pos")
    if not (isinstance(n_sv, np.ndarray) and n_sv.shape == (3,)
            and np.issubdtype(n_sv.dtype, np.integer)):
        raise ValueError(f"Incorrect value ({n_sv}) for n_sv")
    if (perm_prob != 0 and not isinstance(perm_prob, float)) or perm_prob < 0 or perm_prob > 1:
        raise ValueError(f"Incorrect value ({perm_prob}) for perm_prob.")
    if not quiet:
        print("Aligning the corner of the simulated voxel with the origin")
    shift = -np.min(vertices, axis=0) + padding
    vertices = vertices + shift
    if not quiet:
        print(f"Moved the vertices by {shift}")
    voxel_size = np.max(vertices, axis=0) + padding
    if not periodic:
        wall_vertices, wall_faces = _aabb_to_mesh(np.zeros(3), voxel_size)
        faces = np.vstack((faces, wall_faces + len(vertices)))
        vertices = np.vstack((vertices, wall_vertices))
    return _Substrate("mesh", vertices=vertices, faces=faces, voxel_size=voxel_size, n_sv=n_sv,
                      periodic=periodic, init_pos=init_pos, quiet=quiet, perm_prob=perm_prob)


# Corner selector (0 -> a, 1 -> b per axis) and triangulation of the voxel walls.  The vertex
# and face ORDER is the reference's (substrates.py:539-570): wall triangles are the last 12
# faces / last 8 vertices of a non-periodic mesh and their indices break distance ties.
_WALL_CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1], [0, 1, 1], [0, 0, 1],
                          [0, 1, 0], [1, 0, 1]])
_WALL_FACES = np.array([[0, 1, 2], [0, 6, 2], [5, 7, 3], [5, 4, 3], [1, 2, 3], [1, 7, 3],
                        [0, 6, 4], [0, 5, 4], [0, 1, 7], [0, 5, 7], [6, 2, 3], [6, 4, 3]])


def _aabb_to_mesh(a, b):
    """Triangular mesh (8 vertices, 12 faces) of the axis-aligned box with corners a and b."""
    ab = np.stack([np.asarray(a, dtype=float), np.asarray(b, dtype=float)])
    vertices = ab[_WALL_CORNERS, np.arange(3)]
    return vertices, _WALL_FACES.copy()


def _mesh_space_subdivision(vertices, faces, voxel_size, n_sv):
    """Bin the triangles into the ``n_sv`` grid.  Returns ``xs, ys, zs, triangle_indices,
    subvoxel_indices`` with the reference's meaning and element order
    (substrates.py:467-536): the triangles overlapping subvoxel ``i`` are
    ``triangle_indices[subvoxel_indices[i, 0]:subvoxel_indices[i, 1]]``, ascending."""
    xs = np.linspace(0, voxel_size[0], n_sv[0] + 1)
    ys = np.linspace(0, voxel_size[1], n_sv[1] + 1)
    zs = np.linspace(0, voxel_size[2], n_sv[2] + 1)
    v = _lib.f64(vertices)
    f = _lib.i64(faces)
    nsv = _lib.i64(n_sv)
    subvoxel_indices = np.zeros((int(np.prod(nsv)), 2), dtype=np.int64)
    n_out = ctypes.c_int64(0)
    handle = ctypes.c_void_p()
    L = _lib.lib()
    rc = L.dsb_mesh_subdivide(_lib.ptr(v), v.shape[0], _lib.ptr(f), f.shape[0], _lib.ptr(xs),
                              _lib.ptr(ys), _lib.ptr(zs), _lib.ptr(nsv),
                              _lib.ptr(subvoxel_indices), ctypes.byref(n_out),
                              ctypes.byref(handle))
    if rc != 0:
        raise ValueError("mesh subdivision failed: faces must index into vertices and n_sv "
                         "must be positive")
    triangle_indices = np.zeros(n_out.value, dtype=np.int64)
    L.dsb_mesh_subdivide_fetch(handle, _lib.ptr(triangle_indices))
    return xs, ys, zs, triangle_indices.astype(int), subvoxel_indices.astype(int)


def _cross_product(a, b):
    """substrates.py:272-280 (the subdivision's own copies of these helpers live in csrc/dsb_subdivide.cpp)."""
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _dot_product(a, b):
    """substrates.py:283-287."""
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _triangle_aabb(triangle):
    """Corners of the triangle's bounding box closest to and furthest from the origin, (2, 3);
    substrates.py:422-441."""
    t = np.asarray(triangle, dtype=float)
    return np.vstack((t.min(axis=0), t.max(axis=0)))


def _box_subvoxel_overlap(box, xs, ys, zs):
    """Lowest and highest (exclusive) subvoxel index the box overlaps along each axis, (3, 2) int32;
    substrates.py:444-464."""
    subvoxels = np.zeros((3, 2), dtype=np.int32)
    for i, a in enumerate([xs, ys, zs]):
        subvoxels[i] = _interval_sv_overlap(a, box[0, i], box[1, i])
    return subvoxels


def _triangle_box_overlap(triangle, box):
    """True when the triangle (3, 3) overlaps the box (2, 3); substrates.py:290-368."""
    return bool(_lib.lib().dsb_triangle_box_overlap(_lib.ptr(_lib.f64(triangle)),
                                                    _lib.ptr(_lib.f64(box))))


def _interval_sv_overlap(xs, x1, x2):
    """Index range [ll, ul) of subvoxels overlapping [x1, x2]; substrates.py:371-419."""
    xs = _lib.f64(xs)
    ll, ul = ctypes.c_int64(), ctypes.c_int64()
    _lib.lib().dsb_interval_sv_overlap(_lib.ptr(xs), len(xs), float(x1), float(x2),
                                       ctypes.byref(ll), ctypes.byref(ul))
    return ll.value, ul.value
