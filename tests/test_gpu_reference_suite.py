"""The reference's own end-to-end tests (disimpy/tests/test_simulations.py:460-683), run against
this package on the GPU with the reference's tolerances: analytic free diffusion, MISST signals
for sphere and cylinder (tests/golden/ref_misst_signals.npz holds the reference's fixture files),
trajectory files that stay inside the substrate and touch its wall, cylinder orientation
symmetries, ellipsoid(r, r, r) == sphere(r).  Their 100-measurement protocols are one waveform
scaled per b-value, i.e. rank 1: they run through the virtual-measurement path."""

import os

import numpy as np
import numpy.testing as npt
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

D = 2e-9


def example_gradient():  # test_simulations.py:460-466
    T = 80e-3
    gradient = np.zeros((1, 100, 3))
    gradient[0, 1:11, 0] = 1
    gradient[0, -11:-1, 0] = -1
    return gradient, T / (gradient.shape[1] - 1)


def scaled(gradient, dt, n_t, bs):
    from disimpy_b200 import gradients
    gradient = np.concatenate([gradient for _ in bs], axis=0)
    gradient, dt = gradients.interpolate_gradient(gradient, dt, n_t)
    return gradients.set_b(gradient, dt, bs), dt


def misst_protocol(n_on, n_total, T, n_t=1000):  # test_simulations.py:527-534, 546-553
    gradient = np.zeros((1, n_total, 3))
    gradient[0, 1:n_on, 0] = 1
    gradient[0, -n_on:-1, 0] = -1
    return scaled(gradient, T / (n_total - 1), n_t, np.linspace(1, 3e9, 100))


def trajectories(tmp_path, n_s, n_t, substrate):
    from disimpy_b200 import gradients, simulations
    gradient, dt = example_gradient()
    gradient, dt = gradients.interpolate_gradient(gradient, dt, n_t)
    path = str(tmp_path / "example_traj.txt")
    signals = simulations.simulation(n_s, D, gradient, dt, substrate, traj=path, quiet=True)
    tr = np.loadtxt(path)
    assert tr.shape == (n_t + 1, n_s * 3)
    return signals, tr.reshape((n_t + 1, n_s, 3)), gradient, dt


def test_free_diffusion(tmp_path):  # :469-500
    from disimpy_b200 import simulations, substrates
    bs = np.linspace(1, 2e9, 100)
    gradient, dt = scaled(*example_gradient(), 1000, bs)
    signals = simulations.simulation(int(1e5), D, gradient, dt, substrates.free(), quiet=True)
    npt.assert_almost_equal(signals / 1e5, np.exp(-bs * D), 2)
    _, tr, _, _ = trajectories(tmp_path, int(1e4), 100, substrates.free())
    assert np.all(tr[0] == 0)
    npt.assert_almost_equal(np.mean(tr[-1], axis=0), 0, 5)


def test_random_step_like_reference():  # :110-139 (test__cuda_random_step)
    """One time step of free diffusion from the origin = one `_cuda_random_step` per walker: same
    seed -> same steps, another seed -> every component differs, zero mean, unit length, and --
    beyond the reference's test -- the oracle's steps bit for bit."""
    from scipy.stats import normaltest
    from disimpy_b200 import simulations, substrates
    from oracle import oracle as O
    N, dt = int(1e5), 1e-3
    g = np.zeros((1, 1, 3))
    g[0, 0, 0] = 0.01
    step_l = np.sqrt(6 * D * dt)
    steps = np.zeros((3, N, 3))
    for i, seed in enumerate([1, 1, 12]):
        _, pos = simulations.simulation(N, D, g, dt, substrates.free(), seed=seed, final_pos=True, quiet=True)
        steps[i] = pos / step_l
    npt.assert_equal(steps[0], steps[1])
    assert np.all(steps[0] != steps[2])
    npt.assert_almost_equal(np.mean(np.sum(steps[1::], axis=1) / N), 0, 3)
    _, p = normaltest(steps[1::].ravel())
    npt.assert_almost_equal(p, 0)
    npt.assert_almost_equal(np.linalg.norm(steps, axis=2), np.ones((3, N)))
    ref = O.simulation(N, D, g, dt, substrates.free(), seed=12, n_threads=8)
    assert np.array_equal(steps[2] * step_l, ref["positions"])


@pytest.mark.parametrize("shape", ["sphere", "cylinder"])
def test_signal_against_misst(shape):  # :521-566, :602-654
    from disimpy_b200 import simulations, substrates
    misst = load_golden("ref_misst_signals")
    sub = substrates.sphere(5e-6) if shape == "sphere" else substrates.cylinder(5e-6, np.array([0, 0, 1.0]))
    for key, (n_on, n_total, T) in {"30ms": (300, 700, 70e-3), "1ms": (10, 410, 41e-3)}.items():
        gradient, dt = misst_protocol(n_on, n_total, T)
        signals = simulations.simulation(int(1e5), D, gradient, dt, sub, quiet=True)
        npt.assert_almost_equal(signals / 1e5, misst["%s_%s" % (shape, key)], 2)


def test_cylinder_trajectories_and_orientation(tmp_path):  # :503-519, :568-584
    from disimpy_b200 import simulations, substrates
    for radius in [1e-6, 5e-6, 1e-3]:
        _, tr, _, _ = trajectories(tmp_path, 100, 100, substrates.cylinder(radius, np.array([1.0, 0, 0])))
        max_pos = np.max(np.linalg.norm(tr[..., 1::], axis=2))
        assert max_pos < radius
        npt.assert_almost_equal(max_pos, radius)
    bs = np.linspace(1, 3e9, 100)
    gradient, dt = scaled(*example_gradient(), 1000, bs)
    n = int(1e5)
    s1 = simulations.simulation(n, D, gradient, dt, substrates.cylinder(5e-6, np.array([1.0, 0, 1.0])), quiet=True)
    s2 = simulations.simulation(n, D, gradient, dt, substrates.cylinder(5e-6, -np.array([1.0, 0, 1.0])), quiet=True)
    npt.assert_almost_equal(s1 / n, s2 / n)
    s3 = simulations.simulation(n, D, gradient, dt, substrates.cylinder(5e-6, -np.array([1.0, 0, 0])), quiet=True)
    npt.assert_almost_equal(s3 / n, np.exp(-bs * D), 2)


def test_sphere_and_ellipsoid_trajectories(tmp_path):  # :587-600, :657-683
    from disimpy_b200 import simulations, substrates
    radius = 5e-6
    sig_sphere, tr, gradient, dt = trajectories(tmp_path, 100, 100, substrates.sphere(radius))
    max_pos = np.max(np.linalg.norm(tr, axis=2))
    assert max_pos < radius
    npt.assert_almost_equal(max_pos, radius)
    sig_ell, tr, _, _ = trajectories(tmp_path, 100, 100, substrates.ellipsoid(np.ones(3) * radius))
    max_pos = np.max(np.linalg.norm(tr, axis=2))
    assert max_pos < radius
    npt.assert_almost_equal(max_pos, radius)
    npt.assert_almost_equal(sig_ell, sig_sphere)


def test_mesh_diffusion():  # :686-812 (the neuron-model part, :814-832, is test_neuron_model_no_leak below)
    from disimpy_b200 import simulations, substrates
    meshes = load_golden("ref_meshes")
    misst = load_golden("ref_misst_signals")["cylinder_30ms"]
    gradient, dt = misst_protocol(300, 700, 70e-3)
    n_s = int(1e4)
    vertices, faces = meshes["cylinder_mesh_closed_vertices"], meshes["cylinder_mesh_closed_faces"]
    for periodic in [True, False]:
        for padding in [np.zeros(3), np.zeros(3) + 1e-6]:
            for n_sv in [np.array([1, 1, 1]), np.array([1, 5, 20]), np.array([10, 10, 10])]:
                substrate = substrates.mesh(vertices, faces, periodic, padding=padding, init_pos="intra",
                                            n_sv=n_sv, quiet=True)
                signals, pos = simulations.simulation(n_s, D, gradient, dt, substrate, final_pos=True, quiet=True)
                npt.assert_almost_equal(signals / n_s, misst, 2)
                # no spins leaked
                r = np.max(np.linalg.norm(substrate.vertices[:, 0:2] - (substrate.voxel_size[0:2] - padding[0:2] * 2) / 2,
                                          axis=1))
                assert np.min(pos[:, 2]) > 0
                assert np.max(pos[:, 2]) < substrate.voxel_size[2]
                assert np.max(np.linalg.norm(pos[:, 0:2] - np.max(substrate.vertices, axis=0)[0:2] / 2, axis=1)) < r
    # open-ended tube, periodic: walkers leave through the ends into the next voxel, never through the wall
    vertices, faces = meshes["cylinder_mesh_open_vertices"], meshes["cylinder_mesh_open_faces"]
    init_pos = np.zeros((n_s, 3)) + np.array([5e-6, 5e-6, 12.5e-6])
    for padding in [np.zeros(3), np.array([1e-6, 1e-6, 0])]:
        for n_sv in [np.array([1, 1, 1]), np.array([1, 5, 20]), np.array([10, 10, 10])]:
            substrate = substrates.mesh(vertices, faces, init_pos=init_pos + padding, periodic=True, padding=padding,
                                        n_sv=n_sv, quiet=True)
            signals, pos = simulations.simulation(n_s, D, gradient, dt, substrate, final_pos=True, quiet=True)
            r = np.max(np.linalg.norm(substrate.vertices[:, 0:2] - (substrate.voxel_size[0:2] - padding[0:2] * 2) / 2,
                                      axis=1))
            assert np.min(pos[:, 2]) < 0
            assert np.max(pos[:, 2]) > substrate.voxel_size[2]
            assert np.max(np.linalg.norm(pos[:, 0:2] - np.max(substrate.vertices, axis=0)[0:2] / 2, axis=1)) < r


def _real_mesh(name):
    m = load_golden("ref_real_meshes")
    return m[name + "_vertices"], m[name + "_faces"].astype(np.int64)


@pytest.mark.parametrize("dt", [1e-5, 1e-3, 1e-1])
def test_neuron_model_no_leak(dt):  # :814-832
    """The reference's neuron-model test: 29 688 irregular triangles with float32 vertices, default
    n_sv, init_pos='intra', periodic.  dt = 1e-1 gives 35 um steps that span many grid cells (the
    per-lane search on a real mesh).  The reference's containment assertion, and -- beyond it --
    final positions bit for bit and the signal against the oracle."""
    from disimpy_b200 import simulations, substrates
    from oracle import oracle as O
    vertices, faces = _real_mesh("neuron_model")
    assert vertices.dtype == np.float32
    n_s, n_t = int(1e3), int(1e2)
    gradient = np.ones((1, n_t, 3))
    substrate = substrates.mesh(vertices, faces, init_pos="intra", periodic=True, quiet=True)
    signals, pos = simulations.simulation(n_s, D, gradient, dt, substrate, final_pos=True, quiet=True)
    assert np.all(np.max(pos, axis=0) < substrate.voxel_size)
    assert np.all(np.min(pos, axis=0) > 0)
    ref = O.simulation(n_s, D, gradient, dt, substrate, seed=123, n_threads=8)
    assert np.array_equal(pos, ref["positions"])
    assert np.allclose(signals, ref["signals"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name,periodic,init_pos,dt", [
    ("example_mesh", False, "uniform", 1e-4),
    ("example_mesh", True, "uniform", 1e-2),
    ("fibre_mesh", True, "uniform", 1e-3),
    ("fibre_mesh", False, "uniform", 1e-2),
    ("neuron_model", False, "extra", 1e-3),
])
def test_real_meshes_match_oracle(name, periodic, init_pos, dt):
    """The other irregular meshes the reference ships (example_mesh.pkl, fibre_mesh.pkl), and the
    neuron model closed by walls with walkers outside it: positions bit for bit against the oracle,
    short and long steps, two measurements."""
    from disimpy_b200 import simulations, substrates
    from oracle import oracle as O
    vertices, faces = _real_mesh(name)
    gradient = np.ones((2, 60, 3)) * np.array([1.0, 0.5])[:, None, None]
    substrate = substrates.mesh(vertices, faces, periodic, init_pos=init_pos, quiet=True,
                                n_sv=np.array([20, 20, 20]))
    n_s = 2000
    signals, pos = simulations.simulation(n_s, D, gradient, dt, substrate, final_pos=True, quiet=True, seed=5)
    ref = O.simulation(n_s, D, gradient, dt, substrate, seed=5, n_threads=8)
    assert np.array_equal(pos, ref["positions"])
    assert np.allclose(signals, ref["signals"], rtol=1e-12, atol=0)
    if not periodic:
        assert np.all(pos > 0) and np.all(pos < substrate.voxel_size)


def test_fill_mesh_sphere():  # test_simulations.py:428-456 (test__fill_mesh)
    from disimpy_b200 import simulations, substrates
    meshes = load_golden("ref_meshes")
    vertices, faces = meshes["sphere_mesh_vertices"], meshes["sphere_mesh_faces"]
    n_s = int(1e4)
    for n_sv in [np.array([1, 1, 1]), np.array([1, 5, 20]), np.array([10, 10, 10])]:
        for periodic in [True, False]:
            for padding in [np.zeros(3), np.zeros(3) + 1e-6]:
                substrate = substrates.mesh(vertices, faces, periodic, padding=padding, n_sv=n_sv, quiet=True)
                points = simulations._fill_mesh(n_s, substrate, True, seed=123)
                r = (substrate.voxel_size - padding * 2) / 2
                points -= r + padding
                assert np.max(np.linalg.norm(points, axis=1)) < np.min(r)
                npt.assert_almost_equal(np.mean(points, axis=0), np.zeros(3))
                points = simulations._fill_mesh(n_s, substrate, False, seed=123)
                points -= r + padding
                assert np.min(np.linalg.norm(points, axis=1)) > 0.9 * np.min(r)
                npt.assert_almost_equal(np.mean(points, axis=0), np.zeros(3))
