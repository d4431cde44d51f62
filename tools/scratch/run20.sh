timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/e2e_breakdown.py
