timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r02as_bench_full.json 2> gpurun_out/r02as_bench_full.err; echo rc=$?
