timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_realigner.py tests/test_gpu_pool.py -x -q 2>&1 | tail -3
python bench.py --config 3 --steps 3 --warmup 2 > gpurun_out/r02ar_c3_pipe.json 2> gpurun_out/r02ar_c3_pipe.err; echo rc=$?
MPN_NO_SPANS_PIPE=1 python bench.py --config 3 --steps 3 --warmup 2 > gpurun_out/r02ar_c3_nopipe.json 2> gpurun_out/r02ar_c3_nopipe.err; echo rc=$?
