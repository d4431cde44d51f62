export KBENCH_NT=1000
for v in "" _park1; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200$v.so; else unset DISIMPY_B200_LIB; fi
  ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_sphere$v python tools/kbench.py sphere 2>&1 | tail -3
done
