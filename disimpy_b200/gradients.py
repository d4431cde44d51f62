"""Gradient-array helpers needed to build inputs for the walk (host NumPy, off the hot path).

A gradient array has shape (n_measurements, n_time_points, 3) in T/m, exactly what
``simulations.simulation`` takes.  Same function names and argument meaning as
disimpy/gradients.py:13-173 so that scripts written for the reference keep working.
"""

import numpy as np

from . import utils

GAMMA = 267.513e6  # gyromagnetic ratio of the simulated spins (disimpy/gradients.py:13)


def interpolate_gradient(gradient, dt, n_t):
    """Linearly resample every waveform to ``n_t`` points; returns (gradient, new dt)."""
    n_old = gradient.shape[1]
    total = dt * (n_old - 1)
    t_new = np.linspace(0, total, n_t)
    t_old = np.linspace(0, total, n_old)
    out = np.zeros((gradient.shape[0], n_t, 3))
    for m in range(gradient.shape[0]):
        for axis in range(3):
            out[m, :, axis] = np.interp(t_new, t_old, gradient[m, :, axis])
    return out, total / (n_t - 1)


def calc_q(gradient, dt):
    """q(t) = gamma * integral of g (trapezoid rule), shape like ``gradient``."""
    mid = dt * (gradient[:, 1:, :] + gradient[:, :-1, :]) / 2
    zero = np.zeros((gradient.shape[0], 1, 3))
    return GAMMA * np.concatenate((zero, np.cumsum(mid, axis=1)), axis=1)


def calc_b(gradient, dt):
    """b-value of every measurement: integral of |q|^2 dt (trapezoid rule)."""
    q2 = np.linalg.norm(calc_q(gradient, dt), axis=2) ** 2
    # np.trapz's evaluation order (disimpy/gradients.py:89), so that set_b / pgse return arrays
    # bit-identical to the reference's
    return (dt * (q2[:, 1:] + q2[:, :-1]) / 2.0).sum(axis=1)


def set_b(gradient, dt, b):
    """Scale each waveform so that its b-value becomes ``b``."""
    b = np.asarray(b)
    current = calc_b(gradient, dt)
    if np.any(np.isclose(current, 0)):
        raise Exception("b-value can not be changed for measurements with b = 0")
    return gradient * np.sqrt(b / current)[:, np.newaxis, np.newaxis]


def rotate_gradient(gradient, Rs):
    """Apply rotation matrix ``Rs[m]`` to the waveform of measurement m."""
    out = np.zeros(gradient.shape)
    for m, R in enumerate(Rs):
        if not np.isclose(np.linalg.det(R), 1) or not np.all(np.isclose(R.T, np.linalg.inv(R))):
            raise ValueError(f"Rs[{m}] ({R}) is not a valid rotation matrix")
        out[m] = np.matmul(R, gradient[m].T).T
    return out


def pgse(delta, DELTA, n_t, bvals, bvecs):
    """Pulsed-gradient spin echo: two rectangular lobes of duration ``delta`` whose onsets are
    ``DELTA`` apart, sampled on ``n_t`` points, one measurement per (bval, bvec) pair.
    Returns (gradient, dt)."""
    bvals = np.atleast_1d(np.asarray(bvals, dtype=float))
    fine = np.zeros((1, int(1e6), 3))
    dt = (delta + DELTA) / (fine.shape[1] - 1)
    n_lobe = int(np.round(delta / dt))
    fine[0, 1:n_lobe, 0] = 1
    fine[0, -n_lobe:-1, 0] = -1
    wave, dt = interpolate_gradient(fine, dt, n_t)
    gradient = set_b(np.repeat(wave, len(bvals), axis=0), dt, bvals)
    Rs = np.stack([utils.vec2vec_rotmat(np.array([1.0, 0.0, 0.0]), np.asarray(v, dtype=float))
                   for v in bvecs])
    return rotate_gradient(gradient, Rs), dt


def load_camino_scheme_file(path):
    """Gradient array and time step from a Camino general-waveform scheme file
    (disimpy/gradients.py:182-212): first line 'VERSION: GRADIENT_WAVEFORM', then one row per
    measurement: K, dt, g_x1 g_y1 g_z1 g_x2 ...; all rows must share one time step."""
    with open(path, "r") as file:
        if file.readline().strip() != "VERSION: GRADIENT_WAVEFORM":
            raise Exception("The scheme file does not start with 'VERSION: GRADIENT_WAVEFORM'")
    scheme = np.loadtxt(path, skiprows=1, ndmin=2)
    dts = scheme[:, 1]
    if len(set(dts)) != 1:
        raise Exception(
            "Not all rows of the scheme file have the same time step duration. "
            "Disimpy does not support scheme files with multiple time step durations.")
    return scheme[:, 2:].reshape(len(scheme), -1, 3), dts[0]
