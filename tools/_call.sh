for n in 1024 2048 4096; do
python tools/phase_bench.py $n 4 0 2>&1 | tail -1
MPN_LONG_NWP=2 python tools/phase_bench.py $n 4 0 2>&1 | tail -1
done
MPN_LONG_NWP=4 python tools/phase_bench.py 1024 4 0 2>&1 | tail -1
