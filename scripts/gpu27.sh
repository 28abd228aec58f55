timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1l_2gpu.json 2> gpurun_out/bench_r1l_2gpu.err; echo rc=$?
tail -3 gpurun_out/bench_r1l_2gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r1l_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n_gpus %d value %.0f e2e %.0f ms/step %.2f clocks %s'%(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']))
PY
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -c 400
