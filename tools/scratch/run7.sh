timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/kbench.py mesh180 sphere180 ellipsoid180 sphere8 2>&1 | grep -v "^$"
