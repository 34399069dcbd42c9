python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02ac_n8.json 2> gpurun_out/r02ac_n8.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02ac_n4.json 2> gpurun_out/r02ac_n4.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02ac_n2.json 2> gpurun_out/r02ac_n2.err; echo rc=$?
python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r02ac_n1.json 2> gpurun_out/r02ac_n1.err; echo rc=$?
python -m pytest tests/test_gpu_pool.py -q 2>&1 | tail -2
python tests/harness/pool_bench.py > gpurun_out/r02ac_pool.json 2> gpurun_out/r02ac_pool.err; echo rc=$?
