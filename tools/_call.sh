python tools/phase_bench.py 1024 4 1 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_realigner.py -x -q 2>&1 | tail -3
python tools/phase_bench.py 1000000 2 2>&1 | tail -1
