python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('DEFAULT value %.0f e2e %.0f admm_ms %.2f frac %.3f ms/step %.1f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['ms_per_step']))
"
for v in t128_b3 t128_b4 t64_b6 t64_b8; do
MPCB_LIB=$PWD/build_variants/lib_$v.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v value %.0f e2e %.0f admm_ms %.2f frac %.3f ms/step %.1f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['ms_per_step']))
"
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
