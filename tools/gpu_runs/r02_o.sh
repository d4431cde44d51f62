#!/bin/bash
# round 2, fifteenth GPU pass (1 GPU): uploaded-mesh cache -- parity, e2e of the mesh configurations, sanitizer.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DISIMPY_B200_DEVICE=0 timeout 300 python tools/e2e_mesh.py 2>&1 | tail -3 | tee gpurun_out/e2e_mesh_r02_o.txt
timeout 900 python bench.py --no-cpu-baseline --no-secondary --no-reference-baselines > gpurun_out/bench_r02_o.json 2> gpurun_out/bench_r02_o.err; tail -c 200 gpurun_out/bench_r02_o.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_o.json").read().strip().splitlines()[-1])
print("value %.4e e2e %.4e" % (d["value"], d["e2e"]["value"]))
for m in d["mesh"]:
    print(m["config"], "%.3e" % m["value"], "%.1f ms" % m["e2e_ms"])
PY
bash tools/gpu_runs/r02_sanitizer.sh
