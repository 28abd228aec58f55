timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:admm_wide -s 2 -c 1 -f -o gpurun_out/prof_wide python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu64.log 2>&1
tail -2 gpurun_out/ncu64.log | cut -c1-200
