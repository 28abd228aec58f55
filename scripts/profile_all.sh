#!/bin/bash
# scripts/profile_all.sh <tag> — the ncu evidence of a round in one GPU call: launch list of a bench step and one
# `--set full` capture per kernel family (run each program once WITHOUT ncu first; numbers under ncu are never bench values)
tag=${1:?tag}
O=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --records none > $O/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches_batch65536.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --records none > $O/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:admm_tma -c 8 -o $O/${tag}_admm_tma python bench.py --steps 1 --warmup 1 \
    --no-cpu-baseline --records none > $O/${tag}_ncu_tma.log 2>&1
ncu -i $O/${tag}_admm_tma.ncu-rep --page raw --csv > $O/${tag}_admm_tma_ncu_raw.csv
ncu --set full --clock-control none -k regex:"scale_warp|FactorOp" -c 2 -o $O/${tag}_setup python bench.py --steps 1 --warmup 1 \
    --no-cpu-baseline --records none > $O/${tag}_ncu_setup.log 2>&1
ncu -i $O/${tag}_setup.ncu-rep --page raw --csv > $O/${tag}_setup_ncu_raw.csv
python scripts/dense_ncu.py > $O/${tag}_dense_plain.log 2>&1
ncu --set full --clock-control none -k regex:admm_dense -c 1 -o $O/${tag}_dense python scripts/dense_ncu.py > $O/${tag}_ncu_dense.log 2>&1
ncu -i $O/${tag}_dense.ncu-rep --page raw --csv > $O/${tag}_dense_ncu_raw.csv
python scripts/cfg3_ncu.py 1 > $O/${tag}_cfg3_plain.log 2>&1
# configs[3] as benched (chunked loop, unsolved QPs compacted between launches): DRAM bytes and duration of every admm_cta launch of one solve
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:admm_cta -c 64 --csv \
    --log-file $O/${tag}_cta_chunks.csv python scripts/cfg3_ncu.py > $O/${tag}_ncu_cta_chunks.log 2>&1
# one full capture of a main-phase launch (8192 QPs, two CTAs per SM) and one of a straggler launch (one CTA per SM, deep prefetch)
ncu --set full --clock-control none --import-source on -k regex:admm_cta --launch-skip 2 -c 1 -o $O/${tag}_cta python scripts/cfg3_ncu.py > $O/${tag}_ncu_cta.log 2>&1
ncu -i $O/${tag}_cta.ncu-rep --page raw --csv > $O/${tag}_cta_ncu_raw.csv
ncu --set full --clock-control none --import-source on -k regex:admm_cta --launch-skip 14 -c 1 -o $O/${tag}_cta_tail python scripts/cfg3_ncu.py > $O/${tag}_ncu_cta_tail.log 2>&1
ncu -i $O/${tag}_cta_tail.ncu-rep --page raw --csv > $O/${tag}_cta_tail_ncu_raw.csv
rm -f $O/${tag}_*.ncu-rep
ls -la $O | grep ${tag}
