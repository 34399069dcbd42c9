N=$1
if [ "$N" = "1" ]; then
python bench.py --config 5 --warmup 3 > gpurun_out/r02t_c5_n$N.json 2> gpurun_out/r02t_c5_n$N.err; echo rc=$?
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --config 5 --warmup 3 > gpurun_out/r02t_c5_n$N.json 2> gpurun_out/r02t_c5_n$N.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --config 5 --pairs 1000000 --flag 1 --warmup 3 > gpurun_out/r02t_c5f1_n$N.json 2> gpurun_out/r02t_c5f1_n$N.err; echo rc=$?
fi
tail -2 gpurun_out/r02t_c5_n$N.err
grep -o '"value": [0-9.]*' gpurun_out/r02t_c5_n$N.json | head -2
