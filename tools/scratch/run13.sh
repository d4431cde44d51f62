timeout 200 python -m pytest tests -m gpu -x -q -k "analytic or partwise" 2>&1 | tail -3
for v in "" _mb8 _mb6 _b256 _b64; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200$v.so; else unset DISIMPY_B200_LIB; fi
  timeout 100 python tools/kbench.py sphere_t1e4 cylinder ellipsoid free 2>&1 | grep -v "^$"
done
