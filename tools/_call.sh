python tools/phase_bench.py 1000000 2 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
