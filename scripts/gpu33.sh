for v in S2 S3 S4; do
echo "== variant $v"
MPCB_LIB=$PWD/build/v$v.so timeout -s KILL 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f admm_ms %.2f ms/step %.2f non-admm %.2f p50 %.1f'%(d['value'], d['roofline']['avg_launch_ms'], d['ms_per_step'], d['ms_per_step']-d['roofline']['avg_launch_ms'], d['p50_batch_latency_ms']))
    elif 'rror' in l: print(l.strip())
"
done
