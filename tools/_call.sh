python bench.py --steps 1 --warmup 1 --no-configs --cpu-sample 2000 > gpurun_out/r02at_plain.json 2> gpurun_out/r02at_plain.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02at_launches.csv python bench.py --steps 1 --warmup 1 --no-configs --cpu-sample 2000 > gpurun_out/r02at_ncu1.log 2>&1; echo rc=$?
