export MPCB_LIB=$PWD/build/v${V:-R}.so
timeout -s KILL 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/plain18.log 2>&1 && \
timeout -s KILL 800 ncu --set full --clock-control none --import-source on -k regex:admm_tma -s 4 -c 1 -f -o gpurun_out/prof_${V:-R} python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu18.log 2>&1
tail -2 gpurun_out/ncu18.log | cut -c1-200
