timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/kbench.py c5:8 c2 c1 c4
