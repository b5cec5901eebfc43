# 2-GPU validation: fused optimizer (bucketed / overlapped / clip exchange), fused TP kernels, TP trainers, ZeRO-3, then the bench
echo "=== multigpu tests"; timeout 1500 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15
echo "=== bench 2 GPUs"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-1500
