"""Same-seed A/B against the LIVE unmodified reference (its Numba-CUDA kernels) on the GPU box:
tools/ab_live_reference.py in --quick mode (a tenth of the BASELINE walker counts; the full-size
report is kept under profiles/).  Skipped where the reference install (oracle/_ref, git-ignored,
travels with gpurun) or Numba's CUDA target is not available."""

import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_live_reference_ab(tmp_path):
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "disimpy")):
        pytest.skip("oracle/_ref (pip install --target of the reference) is not present")
    probe = subprocess.run([sys.executable, "-c", "from numba import cuda; assert cuda.is_available()"],
                           capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("numba.cuda unavailable: " + probe.stderr[-200:])
    out = str(tmp_path / "ab.json")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_live_reference.py"), "--quick", "--out", out],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    rep = json.load(open(out))
    assert rep["ok"] and len(rep["cases"]) >= 7
    for name, c in rep["cases"].items():
        assert c["positions_array_equal"], name
        assert c["signals_max_rel_diff"] <= 1e-6, name
        assert c["iter_exc_warning_equal"], name
    assert rep["cases"]["sphere_small_iterexc"]["n_flagged_warning"] == 1
