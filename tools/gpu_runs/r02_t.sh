#!/bin/bash
# round 2, last single-GPU check: tests, smoke, full bench with the final build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5
timeout 1500 python bench.py > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err; tail -c 300 gpurun_out/bench_r02_final2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_final2.json").read().strip().splitlines()[-1])
print("value %.4e frac %.3f e2e %.4e (%.1f ms) verbose %.4e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_default_verbose_call"]["value"]))
print(d["mesh_config4_value"], d["mesh_config5_value"])
PY
