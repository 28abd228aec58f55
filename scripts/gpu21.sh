run() { python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f admm_ms %.2f ms/step %.1f p50 %.1f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['ms_per_step'], d['p50_batch_latency_ms']))
    elif 'rror' in l: print(l.strip())
"; }
export MPCB_LIB=$PWD/build/vU.so
echo "== U certs on"; run
echo "== U certs off (code present, sweep skipped)"; MPCB_NO_CERT=1 run
