timeout 300 python -m pytest tests -m gpu -x -q -k "config4" 2>&1 | tail -3
timeout 300 python tools/kbench.py mesh_big 2>&1 | grep -v "^$"
