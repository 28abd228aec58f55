#!/bin/bash
# scripts/gpu.sh <tag> [what] [extra]   — parametrised recipe of the GPU runs summarised under profiles/.
#   what = tests     pytest -m gpu
#          bench     python bench.py (extra = more bench flags)
#          launches  ncu launch list (gpu__time_duration.sum) of one bench step
#          ncu       ncu --set full of the kernel whose name matches extra (regex), first launch
#          all       tests + bench + launches
# Every step writes gpurun_out/<tag>_*; a number printed under ncu is never a bench value.
tag=${1:?tag}; what=${2:-all}; extra=${3:-}
T=$'python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/'$tag$'_tests.log; cat gpurun_out/'$tag$'_tests.log'
B='python bench.py '$extra' > gpurun_out/'$tag'_bench.json 2> gpurun_out/'$tag'_bench.err; tail -c 400 gpurun_out/'$tag'_bench.json'
L='ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/'$tag'_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --records none > gpurun_out/'$tag'_ncu_launches.log 2>&1'
N='ncu --set full --clock-control none --import-source on -k regex:'$extra' -c 1 -o gpurun_out/'$tag'_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --records none > gpurun_out/'$tag'_ncu_full.log 2>&1; ncu -i gpurun_out/'$tag'_full.ncu-rep --page raw --csv > gpurun_out/'$tag'_full_raw.csv'
case $what in
  tests) cmd="$T";; bench) cmd="$B";; launches) cmd="$L";; ncu) cmd="$N";; all) cmd="$T; $B; $L";;
  *) echo "unknown step $what"; exit 1;;
esac
exec /usr/local/graft/bin/gpurun --timeout ${GPU_TIMEOUT:-1200} -- "$cmd"
