MPN_SSW_LIB=$PWD/megapath-nano_b200/libmpn_ssw_fin16.so python bench.py --steps 5 --warmup 3 --no-configs --cpu-sample 20000 > gpurun_out/r02aw_fin16.json 2> gpurun_out/r02aw_fin16.err; echo rc=$?
python bench.py --steps 5 --warmup 3 --no-configs --cpu-sample 20000 > gpurun_out/r02aw_fin8.json 2> gpurun_out/r02aw_fin8.err; echo rc=$?
