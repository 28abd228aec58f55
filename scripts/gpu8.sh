python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:admm_tma -s 1 -c 1 -o gpurun_out/prof_r1d python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu8.log 2>&1
tail -2 gpurun_out/ncu8.log | cut -c1-200
