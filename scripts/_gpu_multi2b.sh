# 2-GPU: in-place wgrad + overlap, fused TP (NVLS reduce-scatter) tests, bench
echo "=== kernel test (1 GPU)"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "wgrad or linear_autograd or adam8bit" 2>&1 | tail -5
echo "=== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -k "fused_tp or overlapped or fused_dp" 2>&1 | tail -8
echo "=== bench 2 GPUs"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-1500
echo "=== bench 1 GPU"; timeout 400 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-1500
echo "=== bench 1 GPU, wgrad accumulate off"; TRLX_B200_WGRAD_ACCUM=0 timeout 400 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
