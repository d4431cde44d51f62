timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 100 python tools/kbench.py sphere_t1e4 sphere cylinder ellipsoid free 2>&1 | grep -v "^$"
