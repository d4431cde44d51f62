#!/bin/bash
# round 2, sixteenth GPU pass (gpurun --gpus 2): the library's own NCCL all-reduce -- sharded == single GPU,
# trace, 2-GPU bench.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29575"
timeout 600 $TR tools/check_multi_gpu.py 2>&1 | grep "^ok\|Error\|error\|Traceback" | tee gpurun_out/multi_gpu_check_N${N}_c.txt
timeout 300 $TR tools/e2e_trace.py 2>&1 | grep "rank\|trace" | tee gpurun_out/e2e_trace_N${N}_nccl.txt
DISIMPY_B200_NCCL=torch PART_FIRST=32768 timeout 300 $TR tools/e2e_trace.py 2>&1 | grep "rank" | tee -a gpurun_out/e2e_trace_N${N}_nccl.txt
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_N${N}_c.json 2> gpurun_out/bench_N${N}_c.err; tail -c 300 gpurun_out/bench_N${N}_c.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_N${N}_c.json").read().strip().splitlines()[-1])
print("value %.4e e2e %.4e (%.1f ms) %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["collective"]))
for m in d["mesh"]:
    print(m["config"], "%.3e" % m["value"], "%.1f ms" % m["e2e_ms"])
PY
