timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
KBENCH_N=1000000 timeout 200 python tools/kbench.py sphere180 ellipsoid180 sphere8 mesh180 2>&1 | grep -v "^$"
