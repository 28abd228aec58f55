timeout -s KILL 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/plain11.log 2>&1 && \
timeout -s KILL 800 ncu --set full --clock-control none --import-source on -k regex:admm_tma -s 4 -c 2 -o gpurun_out/prof_r1f python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu11.log 2>&1
tail -2 gpurun_out/ncu11.log | cut -c1-200
