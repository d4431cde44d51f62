export KBENCH_NT=300
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_mesh_c python tools/kbench.py mesh 2>&1 | tail -3
