#!/bin/bash
# 4 GPUs after the cell-order sort: the bench line (sphere + configs 4 and 5 through simulation() on every rank)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --no-cpu-baseline --no-secondary --no-reference-baselines > gpurun_out/bench_r02_ap_4gpu.json 2> gpurun_out/bench_r02_ap_4gpu.err
tail -c 300 gpurun_out/bench_r02_ap_4gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_ap_4gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])
for m in d['mesh']: print(m['config'], m['value'], m['e2e_ms'])
"
