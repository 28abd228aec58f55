N=${N:-8}
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1n_${N}gpu.json 2> gpurun_out/bench_r1n_${N}gpu.err; echo rc=$?
tail -2 gpurun_out/bench_r1n_${N}gpu.err
python - <<PY
import json
for l in open('gpurun_out/bench_r1n_${N}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n_gpus %d value %.0f e2e %.0f ms/step %.2f clocks %s'%(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']))
PY
