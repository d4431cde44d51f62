#!/bin/bash
# after the cell-order sort: all GPU tests, smoke, the default bench line, ncu of the config-5 shard walk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/bench_r02_ab.json 2> gpurun_out/bench_r02_ab.err; tail -c 300 gpurun_out/bench_r02_ab.err
export KBENCH_NT=1000 KBENCH_N=250000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_ab_config5 python tools/kbench.py config5_shard 2>&1 | tail -2
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_ab.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['roofline']['frac'], d['mesh_config4_value'], d['mesh_config5_value'], d['gpu_launches'])
for m in d['mesh']: print(m['config'], m['value'], m['e2e_ms'])
"
