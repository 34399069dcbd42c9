set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02ab_pytest.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02ab_bench_reference.json 2> /dev/null; echo rc=$?
python bench.py --steps 5 --warmup 3 > gpurun_out/r02ab_bench_full.json 2> gpurun_out/r02ab_bench.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/r02ab_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02ab_launches.csv python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/r02ab_ncu1.log 2>&1; echo "ncu1 rc=$?"
python tools/phase_bench.py 400000 2 > gpurun_out/r02ab_phase.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:sw_strip16_kernelILi16ELi8ELb0ELb0ELb0 -s 2 -c 1 -o gpurun_out/r02ab_strip16 python tools/phase_bench.py 400000 2 > gpurun_out/r02ab_ncu2.log 2>&1; echo "ncu2 rc=$?"
