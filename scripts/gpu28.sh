timeout -s KILL 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/plain28.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:admm_tma -s 4 -c 2 -f -o gpurun_out/prof_r1n python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu28.log 2>&1
tail -2 gpurun_out/ncu28.log | cut -c1-200
