for v in "" _gr8; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200$v.so; else unset DISIMPY_B200_LIB; fi
  KBENCH_N=1000000 timeout 200 python tools/kbench.py sphere180 ellipsoid180 sphere8 2>&1 | grep -v "^$"
done
