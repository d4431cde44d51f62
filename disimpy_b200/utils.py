"""Host math shared by the substrate constructors (mirror of disimpy/utils.py:11-42; the
matplotlib viewers of the reference are out of scope)."""

import numpy as np


def vec2vec_rotmat(v, k):
    """Rotation matrix that aligns ``v`` with ``k`` (Rodrigues' formula).

    Same contract as disimpy/utils.py:11-42, including ``±eye(3)`` for (anti)parallel
    inputs; the evaluation order of the final sum is kept so that R is bit-identical
    (R enters the walk through the per-step frame change).
    """
    v = v / np.linalg.norm(v)
    k = k / np.linalg.norm(k)
    axis = np.cross(v, k)
    if np.linalg.norm(axis) < np.finfo(float).eps:
        if np.linalg.norm(v - k) > np.linalg.norm(v):
            return -np.eye(3)
        return np.eye(3)
    axis /= np.linalg.norm(axis)
    angle = np.arccos(np.dot(v, k))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * np.matmul(K, K)
