timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f admm_ms %.2f frac %.3f ms/step %.1f iters %.1f solved %.3f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['ms_per_step'], d['config']['mean_admm_iterations'], d['config']['fraction_solved']))
    elif 'rror' in l: print(l.strip())
"
