"""Time the UNMODIFIED reference under NUMBA_ENABLE_CUDASIM=1 on host cores (north_star: "its
NUMBA_ENABLE_CUDASIM CPU path timed on the host cores, with core count stated"; a reported
baseline, not an optimisation target).  BASELINE.json configs[0] (free diffusion, PGSE b = 1e9
s/m^2, 1000 steps) on a bounded sample of walkers: the simulator runs Python threads under the
GIL, i.e. one effective core, at a few hundred walker-steps/s, so the full 1e4 walkers x 1000
steps would take hours.

Test/measurement infrastructure only: imports the reference from oracle/_ref (git-ignored pip
install of /root/reference).  Run:  python tools/bench_reference_cudasim.py [n_walkers] [n_t] [--json]
Writes profiles/r02_reference_cudasim.json, or with --json (bench.py's `baselines` leg) ONE JSON
line to stdout (free diffusion only).
"""
import json
import os
import sys
import time
import warnings

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from disimpy import gradients, simulations as S, substrates  # noqa: E402

# the simulator cannot call the reference's @numba.jit helpers from "device" code
S._cuda_reflection = S._cuda_reflection.py_func
S._cuda_crossing = S._cuda_crossing.py_func

JSON_ONLY = "--json" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(argv[0]) if len(argv) > 0 else 128
n_t = int(argv[1]) if len(argv) > 1 else 100
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    g, dt = gradients.pgse(10e-3, 30e-3, n_t, np.array([1e9]), np.array([[1.0, 0, 0]]))
out = {"numba_enable_cudasim": True, "host_cores": os.cpu_count(), "effective_cores": 1, "n_walkers": n, "n_t": n_t,
       "what": "the unmodified reference under NUMBA_ENABLE_CUDASIM=1 (Python threads under the GIL: one effective core) on "
               "a bounded sample of BASELINE config 1 (free diffusion, PGSE b = 1e9 s/m^2); the full 1e4 walkers x 1000 steps "
               "would take hours"}
cases = (("free", substrates.free()),) if JSON_ONLY else (("free", substrates.free()), ("sphere", substrates.sphere(10e-6)))
for name, sub in cases:
    t0 = time.time()
    sig = S.simulation(n, 2e-9, g, float(dt), sub, quiet=True)
    el = time.time() - t0
    out[name] = {"seconds": el, "walker_steps_per_s": n * n_t / el, "signal_over_n": float(sig[0]) / n}
    print(name, out[name], file=sys.stderr if JSON_ONLY else sys.stdout, flush=True)
if JSON_ONLY:
    print(json.dumps(out), flush=True)
else:
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_reference_cudasim.json"), "w"), indent=1)
