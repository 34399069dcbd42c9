set -x
python -m pytest tests/test_gpu_pool.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02i_n2.json 2> gpurun_out/r02i_n2.err; echo rc=$?; tail -2 gpurun_out/r02i_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --pairs 1000000 --warmup 2 > gpurun_out/r02i_c5_n2.json 2> gpurun_out/r02i_c5_n2.err; echo rc=$?; tail -2 gpurun_out/r02i_c5_n2.err
python bench.py --config 5 --pairs 1000000 --warmup 2 > gpurun_out/r02i_c5_n1.json 2> gpurun_out/r02i_c5_n1.err; echo rc=$?; tail -2 gpurun_out/r02i_c5_n1.err
python tests/harness/pool_bench.py > gpurun_out/r02i_pool.json 2> gpurun_out/r02i_pool.err; echo rc=$?; tail -2 gpurun_out/r02i_pool.err
