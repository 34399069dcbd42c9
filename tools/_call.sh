timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-configs --cpu-sample 20000 > gpurun_out/r02am_bench.json 2> gpurun_out/r02am_bench.err; echo rc=$?
