set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"; cat gpurun_out/r02a_bench.json
MPN_TIMING=1 python bench.py --steps 2 --warmup 3 > /dev/null 2> gpurun_out/r02a_timing.err; tail -12 gpurun_out/r02a_timing.err
python tools/phase_bench.py 400000 2 > gpurun_out/r02a_phase.log 2>&1 && cat gpurun_out/r02a_phase.log && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:sw_strip16_kernelILi16ELi8ELb0ELb0 -s 2 -c 1 -o gpurun_out/r02a_strip16 python tools/phase_bench.py 400000 2 > gpurun_out/r02a_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02a_ncu.log
