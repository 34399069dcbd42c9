timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02aq_reference.json 2> gpurun_out/r02aq_reference.err; echo rc=$?
python bench.py --steps 5 --warmup 3 > gpurun_out/r02aq_bench_full.json 2> gpurun_out/r02aq_bench_full.err; echo rc=$?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02aq_launches.csv python bench.py --steps 1 --warmup 1 --no-configs --cpu-sample 2000 > gpurun_out/r02aq_ncu1.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:sw_revband_kernel --launch-skip 5 --launch-count 5 -o gpurun_out/r02aq_revband python bench.py --steps 1 --warmup 2 --no-configs --cpu-sample 2000 > gpurun_out/r02aq_ncu2.log 2>&1; echo rc=$?
