timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python tools/kbench.py mesh mesh180 sphere180 ellipsoid180 2>&1 | grep -v "^$"
export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200_mb5.so
timeout 100 python tools/kbench.py mesh 2>&1 | grep -v "^$"
export DISIMPY_B200_LIB=$PWD/disimpy_b200/libdisimpy_b200_tc16.so
timeout 200 python tools/kbench.py mesh180 sphere180 2>&1 | grep -v "^$"
