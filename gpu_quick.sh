timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 120 python tools/kbench.py c5:8 c2
