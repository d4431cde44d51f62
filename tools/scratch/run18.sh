timeout 300 python tools/kbench.py mesh mesh mesh_coarse 2>&1 | grep -v "^$"
