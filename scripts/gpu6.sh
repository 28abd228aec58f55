set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | grep -E "^\{|rror" | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep -E "^\{|rror" | cut -c1-900
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep -E "^\{|rror" | cut -c1-400
