set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo rc=$?; tail -3 gpurun_out/r02h_bench.err
python bench.py --config 5 --pairs 300000 --warmup 2 > gpurun_out/r02h_c5.json 2> gpurun_out/r02h_c5.err; echo rc=$?; tail -3 gpurun_out/r02h_c5.err
python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/r02h_c4.json 2> gpurun_out/r02h_c4.err; echo rc=$?; tail -3 gpurun_out/r02h_c4.err
python bench.py --config 3 > gpurun_out/r02h_c3.json 2> gpurun_out/r02h_c3.err; echo rc=$?; tail -3 gpurun_out/r02h_c3.err
