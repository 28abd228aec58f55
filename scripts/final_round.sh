#!/bin/bash
# scripts/final_round.sh <tag> — end-of-round evidence in one GPU call: GPU tests, smoke, the bench line with every record,
# the reference arm, and the ncu captures of the CTA kernel on configs[3] (run without ncu first)
tag=${1:?tag}
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${tag}_tests.log; cat $O/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > $O/${tag}_bench_all_records.json 2> $O/${tag}_bench.err; tail -c 200 $O/${tag}_bench_all_records.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference_arm.json 2>> $O/${tag}_bench.err; tail -c 300 $O/${tag}_bench_reference_arm.json; echo
python scripts/cfg3_ncu.py 1 > $O/${tag}_cfg3_plain.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:admm_cta -c 64 --csv \
    --log-file $O/${tag}_cta_chunks.csv python scripts/cfg3_ncu.py > $O/${tag}_ncu_cta_chunks.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:admm_cta --launch-skip 2 -c 1 -o $O/${tag}_cta python scripts/cfg3_ncu.py > $O/${tag}_ncu_cta.log 2>&1
ncu -i $O/${tag}_cta.ncu-rep --page raw --csv > $O/${tag}_cta_ncu_raw.csv
ncu --set full --clock-control none --import-source on -k regex:admm_cta --launch-skip 14 -c 1 -o $O/${tag}_cta_tail python scripts/cfg3_ncu.py > $O/${tag}_ncu_cta_tail.log 2>&1
ncu -i $O/${tag}_cta_tail.ncu-rep --page raw --csv > $O/${tag}_cta_tail_ncu_raw.csv
rm -f $O/${tag}_*.ncu-rep
ls -la $O | grep ${tag}_
