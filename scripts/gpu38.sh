export MPCB_LIB=$PWD/build/vWide.so
timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain38.log 2>&1 && \
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_wide.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu38.log 2>&1
tail -c 200 gpurun_out/ncu38.log
