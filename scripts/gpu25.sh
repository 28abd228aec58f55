timeout -s KILL 700 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -6
for i in 1 2; do
timeout -s KILL 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f admm_ms %.2f ms/step %.1f p50 %.1f launches %d steps %s'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['ms_per_step'], d['p50_batch_latency_ms'], d['gpu_launches'], d['step_ms']))
    elif 'rror' in l: print(l.strip())
"
done
MPCB_NO_CERT=1 timeout -s KILL 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('NO_CERT value %.0f admm_ms %.2f ms/step %.1f'%(d['value'], d['roofline']['avg_launch_ms'], d['ms_per_step']))
"
