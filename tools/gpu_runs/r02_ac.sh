#!/bin/bash
# cell-order sort, final policy (meshes above 64 MB, once per run, slabs across the longest voxel edge):
# its test, kernel times against index order, ncu of the config-5 shard walk, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cell_order or mesh" 2>&1 | tail -2
log=gpurun_out/kbench_r02_ac.log; : > $log
for S in 0 default; do
  echo "== DISIMPY_B200_RESORT=$S" >> $log
  if [ $S = default ]; then unset DISIMPY_B200_RESORT; else export DISIMPY_B200_RESORT=$S; fi
  timeout 300 python tools/kbench.py mesh mesh_big config5_shard 2>&1 | grep -v "mesh:" >> $log
done
unset DISIMPY_B200_RESORT
cat $log
timeout 1500 python bench.py > gpurun_out/bench_r02_ac.json 2> gpurun_out/bench_r02_ac.err; tail -c 300 gpurun_out/bench_r02_ac.err
export KBENCH_NT=1000 KBENCH_N=250000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_ac_config5 python tools/kbench.py config5_shard 2>&1 | tail -2
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_ac.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['roofline']['frac'], d['mesh_config4_value'], d['mesh_config5_value'], d['gpu_launches'])
for m in d['mesh']: print(m['config'], m['value'], m['e2e_ms'])
"
