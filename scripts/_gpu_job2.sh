nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 scripts/bench_collectives.py 2>&1 | grep -E '^\{|Error|error' | head -12
echo "=== bench N=4"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 4 --steps 6 --warmup 3 2>&1 | grep -E '^\{' | cut -c1-1100
