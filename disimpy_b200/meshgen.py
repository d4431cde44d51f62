"""Synthetic triangular meshes for parity tests and for the BASELINE.json mesh
workloads.  Nothing here is on the hot path.

All generators return ``(vertices float64 (V, 3), faces int64 (F, 3))`` in
metres, in the same convention ``substrates.mesh`` expects.
"""

import numpy as np


def icosphere(radius, subdivisions=1, centre=(0.0, 0.0, 0.0)):
    """Closed sphere mesh: an icosahedron subdivided ``subdivisions`` times."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
         (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9),
         (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2),
         (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    verts = [np.asarray(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = list(f)
    for _ in range(subdivisions):
        cache = {}

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        new_faces = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    vertices = np.asarray(verts) * radius + np.asarray(centre, dtype=np.float64)
    return vertices, np.asarray(faces, dtype=np.int64)


def open_tube(radius, length, n_theta, n_z, centre_xy=(0.0, 0.0)):
    """Open-ended cylinder parallel to z from z=0 to z=length: ``n_z`` bands of
    ``n_theta`` quads, two triangles per quad (same topology as the reference's
    ``cylinder_mesh_open`` fixture, for periodic continuation along z)."""
    theta = 2.0 * np.pi * np.arange(n_theta) / n_theta
    ring = np.stack([centre_xy[0] + radius * np.cos(theta),
                     centre_xy[1] + radius * np.sin(theta)], axis=1)
    zs = np.linspace(0.0, length, n_z + 1)
    vertices = np.concatenate(
        [np.concatenate([ring, np.full((n_theta, 1), z)], axis=1) for z in zs])
    k = np.arange(n_theta)
    kn = (k + 1) % n_theta
    faces = []
    for j in range(n_z):
        a, b = j * n_theta, (j + 1) * n_theta
        faces.append(np.stack([a + k, a + kn, b + k], axis=1))
        faces.append(np.stack([a + kn, b + kn, b + k], axis=1))
    return vertices, np.concatenate(faces).astype(np.int64)


def tube_lattice(n_x, n_y, radius, pitch, length, n_theta, n_z):
    """``n_x`` × ``n_y`` open tubes ‖ z on a square lattice of spacing ``pitch``.

    Every tube lies wholly inside its lattice cell, so with
    ``padding = [pitch/2 - radius, pitch/2 - radius, 0]`` the voxel built by
    ``substrates.mesh(..., periodic=True)`` tiles space periodically.
    Returns ``(vertices, faces, padding, centres)`` where ``centres`` (n_x*n_y, 2)
    are the tube axes in the coordinates of the *shifted* mesh (voxel corner at
    the origin), for containment checks.
    """
    vs, fs, centres = [], [], []
    off = 0
    for ix in range(n_x):
        for iy in range(n_y):
            c = ((ix + 0.5) * pitch, (iy + 0.5) * pitch)
            v, f = open_tube(radius, length, n_theta, n_z, c)
            vs.append(v)
            fs.append(f + off)
            off += len(v)
            centres.append(c)
    padding = np.array([pitch / 2 - radius, pitch / 2 - radius, 0.0])
    return (np.concatenate(vs), np.concatenate(fs), padding,
            np.asarray(centres, dtype=np.float64))


def fibonacci_sphere(n):
    """``n`` deterministic, roughly uniform unit vectors (the reference ships no
    direction set; SURVEY.md §8d config 3 uses these as b-vectors)."""
    i = np.arange(n) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = np.pi * (1.0 + 5.0 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi),
                     np.cos(phi)], axis=1)
