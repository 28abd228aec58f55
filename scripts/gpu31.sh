timeout -s KILL 700 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4
timeout -s KILL 300 python bench.py > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/bench_r1k.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f frac %.3f ms/step %.2f clocks %s cpu %s launches %d'%(d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['clocks'], d.get('cpu_baseline',{}).get('value'), d['gpu_launches']))
PY
tail -3 gpurun_out/bench_r1k.err
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1k_ref.json 2>&1; tail -c 300 gpurun_out/bench_r1k_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
