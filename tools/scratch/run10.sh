WITH_TORCH=1 timeout 100 python tools/e2e_breakdown.py
timeout 100 python tools/e2e_breakdown.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_bench_c.log 2>&1
export KBENCH_NT=1000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_c_sphere python tools/kbench.py sphere 2>&1 | tail -2
