for v in "" _mmb5 _mmb3; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 100 python tools/kbench.py mesh 2>&1 | grep -v "^$"
done
unset DISIMPY_B200_LIB
export KBENCH_NT=1000 KBENCH_N=500000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_mesh_i python tools/kbench.py mesh 2>&1 | tail -2
