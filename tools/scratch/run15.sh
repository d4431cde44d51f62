timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/kbench.py mesh mesh180 2>&1 | grep -v "^$"
KBENCH_NT=300 timeout 100 python tools/kbench.py mesh 2>&1 | grep -v "^$"
