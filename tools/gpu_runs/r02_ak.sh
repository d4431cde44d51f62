#!/bin/bash
# the round's closing run on the final tree: all GPU tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
timeout 1500 python bench.py > gpurun_out/bench_r02_ak.json 2> gpurun_out/bench_r02_ak.err; tail -c 300 gpurun_out/bench_r02_ak.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ak_reference.json 2>> gpurun_out/bench_r02_ak.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_ak.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['mesh_config4_value'], d['mesh_config5_value'], d['gpu_launches'], d['clocks'])
r=json.loads(open('gpurun_out/bench_r02_ak_reference.json').read().strip().splitlines()[-1])
print(r['impl'], r['value'], r['cpu_baseline'])
"
