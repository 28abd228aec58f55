timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain44.log 2>&1 && \
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu44.log 2>&1
tail -c 300 gpurun_out/ncu44.log
