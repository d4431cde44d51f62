#!/bin/bash
mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_N4.json 2> gpurun_out/bench_N4.err; tail -c 200 gpurun_out/bench_N4.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_N4.json").read().strip().splitlines()[-1])
print("value %.4e e2e %.4e (%.1f ms)" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
for m in d["mesh"]:
    print(m["config"], "%.3e" % m["value"], "%.1f ms" % m["e2e_ms"])
PY
