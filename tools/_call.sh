timeout 900 python -m pytest tests/test_gpu_revband.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r02ad_band.json 2> gpurun_out/r02ad_band.err; echo rc=$?
MPN_NO_REVBAND=1 python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r02ad_noband.json 2> gpurun_out/r02ad_noband.err; echo rc=$?
