python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --no-configs > gpurun_out/r02ap_n8.json 2> gpurun_out/r02ap_n8.err; echo rc=$?
python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r02ap_n1.json 2> gpurun_out/r02ap_n1.err; echo rc=$?
