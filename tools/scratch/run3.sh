python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in "" _chunk4 _chunk16 _mb3 _mb5 _mb6; do
  if [ -n "$v" ]; then export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200$v.so; else unset DISIMPY_B200_LIB; fi
  python tools/kbench.py mesh 2>&1 | grep -v "^$"
done
