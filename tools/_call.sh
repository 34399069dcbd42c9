timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r02au_bench.json 2> gpurun_out/r02au_bench.err; echo rc=$?
