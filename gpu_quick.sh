# one gpurun call that checks everything quickly: tests, smoke, small-config timings, a 96-observation bench step
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" | tail -1
timeout 300 python tools/kbench.py c1 c2 c4 c5:8
timeout 300 python bench.py --obs 96 --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-300
