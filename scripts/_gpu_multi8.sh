N=${1:-8}
echo "=== bench $N GPUs"; BENCH_BREAKDOWN=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -3 | cut -c1-1800
