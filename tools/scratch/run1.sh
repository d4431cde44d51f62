python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in "" _park1 _park2 _park3 _park6 _park8; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200$v.so; else unset DISIMPY_B200_LIB; fi
  python tools/kbench.py sphere cylinder ellipsoid 2>&1 | grep -v "^$"
done
