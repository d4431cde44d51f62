timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r01_g_8gpu.json 2> gpurun_out/bench_r01_g_8gpu.err
tail -c 800 gpurun_out/bench_r01_g_8gpu.err; cut -c1-400 gpurun_out/bench_r01_g_8gpu.json
