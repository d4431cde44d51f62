"""Closed-form PGSE signals in the Gaussian phase approximation (test infrastructure): free
diffusion, sphere (Murday & Cotts 1968) and cylinder with the gradient perpendicular to its
axis (van Gelderen et al. 1994).  Used to check the simulated signals against theory within
Monte Carlo error, like the reference's validation notebook does (docs/source/validation.ipynb)."""

import numpy as np
from scipy import optimize, special

GAMMA = 267.513e6


def _roots(fn, n, step=0.1):
    out, x = [], step
    while len(out) < n:
        if fn(x) * fn(x + step) < 0:
            out.append(optimize.brentq(fn, x, x + step))
        x += step
    return np.array(out)


def gradient_strength(b, delta, DELTA):
    return np.sqrt(b / (GAMMA ** 2 * delta ** 2 * (DELTA - delta / 3)))


def free(b, D):
    return np.exp(-np.asarray(b) * D)


def _bracket(a2D, delta, DELTA):
    return (2 * delta - (2 + np.exp(-a2D * (DELTA - delta)) - 2 * np.exp(-a2D * DELTA)
                         - 2 * np.exp(-a2D * delta) + np.exp(-a2D * (DELTA + delta))) / a2D)


def sphere(b, D, radius, delta, DELTA, n_terms=40):
    # alpha_m R are the roots of the derivative of the spherical Bessel function j1
    x = _roots(lambda x: special.spherical_jn(1, x, derivative=True), n_terms)
    alpha = x / radius
    G2 = gradient_strength(np.asarray(b, dtype=float), delta, DELTA) ** 2
    s = np.sum(alpha ** -4 / (alpha ** 2 * radius ** 2 - 2) * _bracket(alpha ** 2 * D, delta, DELTA))
    return np.exp(-2 * GAMMA ** 2 * G2 / D * s)


def cylinder(b, D, radius, delta, DELTA, n_terms=40):
    x = _roots(lambda x: special.jvp(1, x), n_terms)
    alpha = x / radius
    G2 = gradient_strength(np.asarray(b, dtype=float), delta, DELTA) ** 2
    s = np.sum(_bracket(alpha ** 2 * D, delta, DELTA) / (alpha ** 4 * (alpha ** 2 * radius ** 2 - 1)))
    return np.exp(-2 * GAMMA ** 2 * G2 / D * s)
