#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_r02_y.json 2> gpurun_out/bench_r02_y.err; tail -c 300 gpurun_out/bench_r02_y.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_y.json").read().strip().splitlines()[-1])
print("value %.4e frac %.3f e2e %.4e (%.1f ms) verbose %.4e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_default_verbose_call"]["value"]))
for w in d["other_workloads"]:
    print(w["workload"][:50], "%.3e" % w["value"], w.get("l2"))
PY
