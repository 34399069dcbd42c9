python -m pytest tests/test_gpu_parity.py -x -q -k "packed or golden or config1" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo rc=$?; tail -3 gpurun_out/r02q_bench.err
python tests/harness/zero_edit_cost.py 200 2 2>&1 | tail -1
