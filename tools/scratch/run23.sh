export KBENCH_NT=208 KBENCH_N=400000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_g_sphere180 python tools/kbench.py sphere180 2>&1 | tail -2
