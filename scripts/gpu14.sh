for v in ${VARIANTS:-A B C E F}; do
echo "== variant $v"
MPCB_LIB=$PWD/build/v$v.so timeout -s KILL 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f admm_ms %.2f ms/step %.1f p50 %.1f launches %d steps %s'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['ms_per_step'], d['p50_batch_latency_ms'], d['gpu_launches'], d['step_ms']))
    elif 'rror' in l: print(l.strip())
"
done
