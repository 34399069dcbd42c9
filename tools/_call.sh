python tools/phase_bench.py 400000 2 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
