"""The CPU oracle against (a) outputs of the unmodified reference's Numba-CUDA kernels run on a
B200 (tests/golden/*.npz, made by tools/gen_golden_gpu.py) and (b) the reference's own golden
GPU trajectory disimpy/tests/test_traj.txt (tests/golden/ref_test_traj.npz, made by
tools/make_ref_fixtures.py).  Everything here is bit-exact: == on float64."""

import ctypes
import types

import numpy as np
import pytest

from conftest import SIM_CASES, golden_kwargs, load_golden, oracle_substrate
from oracle import oracle as O


@pytest.mark.parametrize("seed", [0, 123, 2 ** 31 + 12345])
def test_rng_states_and_draws(seed):
    r = load_golden("rng")
    st = O.rng_states(seed, 128)
    assert np.array_equal(st[:, 0], r["states_seed%d_s0" % seed])
    assert np.array_equal(st[:, 1], r["states_seed%d_s1" % seed])
    for i in range(128):
        normals, uniforms, after = O.draw(st[i], 24, 4)
        assert np.array_equal(normals, r["normals_seed%d" % seed][i])
        assert np.array_equal(uniforms, r["uniforms_seed%d" % seed][i])
        assert after[0] == r["after_seed%d_s0" % seed][i]
        assert after[1] == r["after_seed%d_s1" % seed][i]


def test_rng_known_answers():
    """state table and first outputs of walker 0, seed 123 / seed 0."""
    st = O.rng_states(123, 4)
    assert [hex(int(v)) for v in st.ravel()] == [
        "0xb4dc9bd462de412b", "0xb4dc9bd462de412b", "0x786604eda3ebc8b3", "0x697e7f19ac2a146d",
        "0x2875c3ea17b04bfe", "0xff6bf1ab383baf92", "0x714047bc8d6702aa", "0x16ce1678e431f65e"]
    st0 = O.rng_states(0, 2)
    assert hex(int(st0[0, 0])) == "0xe220a8397b1dcdaf" and st0[0, 0] == st0[0, 1]
    assert (hex(int(st0[1, 0])), hex(int(st0[1, 1]))) == ("0x12513ce25be05eb1", "0x70d189c276ba17a4")
    normals, _, _ = O.draw(st[0], 3, 0)
    assert np.allclose(normals, [-1.14317007, -1.13542584, 0.33488839], rtol=0, atol=5e-9)


def test_rng_subsequence_start():
    r = load_golden("rng")
    st = O.rng_states(123, 8, 1000003)
    assert np.array_equal(st[:, 0], r["states_seed123_off1000003_s0"])
    assert np.array_equal(st[:, 1], r["states_seed123_off1000003_s1"])


def test_reference_test_traj_replay():
    """disimpy/tests/test_traj.txt: free diffusion, seed 123, 10 walkers, 999 steps."""
    g = load_golden("ref_test_traj")
    tr, step_l = g["traj"], float(g["step_l"])
    sub = types.SimpleNamespace(type="free")
    p, _ = O.make_params(sub, 10, 1, 999, step_l, 1e-5, 123, 1000, 1e-13)
    grad = np.zeros((1, 999, 3))
    pos, ph, ie = np.zeros((10, 3)), np.zeros((1, 10)), np.zeros(10, np.uint8)
    out = np.zeros((1000, 10, 3))
    O.lib().oracle_simulate(ctypes.byref(p), O._ptr(grad), O._ptr(pos), O._ptr(ph), O._ptr(ie),
                            None, O._ptr(out))
    assert np.array_equal(out, tr)


@pytest.mark.parametrize("name", SIM_CASES)
def test_simulation_golden(name):
    g = load_golden(name)
    sub = oracle_substrate(name, g)
    res = O.simulation(int(g["n_walkers"]), float(g["diffusivity"]), g["gradient"], float(g["dt"]),
                       sub, seed=int(g["seed"]), traj=("traj" in g), n_threads=4,
                       **golden_kwargs(g))
    assert np.array_equal(res["positions"], g["positions"])
    assert np.array_equal(res["signals"], g["signals"])
    assert np.array_equal(O.signals_from_phases(res["phases"], res["iter_exc"], True),
                          g["all_signals"], equal_nan=True)
    assert bool(res["iter_exc"].any()) == (len(g["iter_exc_warning"]) > 0)
    if "traj" in g:
        assert np.array_equal(res["traj"], g["traj"])


def test_oracle_shards_compose():
    """Walker w uses subsequence w whatever the shard layout: two half runs == one full run."""
    g = load_golden("sphere_small")
    sub = oracle_substrate("sphere_small", g)
    n = int(g["n_walkers"])
    pos0 = O.initial_positions(sub, n, int(g["seed"]))
    full = O.run_walk(sub, g["gradient"], float(g["dt"]), float(g["diffusivity"]), pos0)
    a = O.run_walk(sub, g["gradient"], float(g["dt"]), float(g["diffusivity"]), pos0[:200])
    b = O.run_walk(sub, g["gradient"], float(g["dt"]), float(g["diffusivity"]), pos0[200:],
                   walker_offset=200)
    assert np.array_equal(np.vstack([a["positions"], b["positions"]]), full["positions"])
    assert np.array_equal(np.hstack([a["phases"], b["phases"]]), full["phases"])


def test_device_function_known_answers():
    """the oracle's restatements of the reference's device functions against the known
    answers of the reference's own unit tests (disimpy/tests/test_simulations.py:23-360)."""
    from conftest import check_device_function_known_answers
    check_device_function_known_answers(O.unit)


@pytest.mark.parametrize("name", ["ll_subvoxel_overlap", "ul_subvoxel_overlap", "ll_subvoxel_overlap_periodic",
                                  "ul_subvoxel_overlap_periodic"])
def test_subvoxel_overlap_functions(name):
    """The oracle's subvoxel range lookups against the reference's loops written out in Python
    (disimpy/simulations.py:616-679).  The periodic ones shift by fma(-voxel, n, x) in the compiled
    kernel and by x - n * voxel in plain Python: rows where those two round differently (the shifted
    coordinate lands on the other side of a boundary; a third of the rows sit on boundaries or their
    periodic images on purpose) must be few and off by one; the golden trajectories pin the fma form."""
    from conftest import device_function_random_rows, subvoxel_overlap_restated
    rows = device_function_random_rows(name, 4000)
    got = O.unit(name, rows)[:, 0]
    want = np.array([subvoxel_overlap_restated(name, r) for r in rows], dtype=float)
    differ = got != want
    if name.endswith("periodic"):
        assert differ.mean() < 0.05 and np.all(np.abs(got - want)[differ] <= 1)
    else:
        assert not differ.any()
