python -m pytest tests/test_gpu_reference_callers.py -x -q 2>&1 | tail -15
