timeout -s KILL 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 300 python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/bench_r1g.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f frac %.3f ms/step %.2f clocks %s cpu %s'%(d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['clocks'], d.get('cpu_baseline',{}).get('value')))
PY
tail -3 gpurun_out/bench_r1g.err
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1g_ref.json 2>&1; tail -c 600 gpurun_out/bench_r1g_ref.json
