export MPCB_LIB=$PWD/build/vWide4.so
timeout -s KILL 200 python scripts/wide_check.py 2>&1 | tail -8
for w in 0 1; do
if [ $w = 0 ]; then export MPCB_NO_WIDE=1; else unset MPCB_NO_WIDE; fi
timeout -s KILL 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('wide=$w value %.0f e2e %.0f admm_ms %.2f ms/step %.2f p50 %.2f launches %d solved %.4f iters %.2f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['ms_per_step'], d['p50_batch_latency_ms'], d['gpu_launches'], d['config']['fraction_solved'], d['config']['mean_admm_iterations']))
    elif 'rror' in l: print(l.strip())
"
done
