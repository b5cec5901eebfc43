echo "=== kernel tests"; timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_short or cluster_split_k or test_gemm_matches_fp32 or lmhead or epilogue or folded" -p no:cacheprovider 2>&1 | tail -4
echo "=== chain"; timeout 200 python scripts/bench_chain.py 2>&1 | head -1 | cut -c1-600
echo "=== engine+trainer tests"; timeout 500 python -m pytest tests/test_engine_gpu.py tests/test_trainers_gpu.py -m gpu -q -x --timeout=300 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -12
echo "=== bench"; BENCH_BREAKDOWN=1 timeout 250 python bench.py --steps 6 --warmup 4 2>&1 | tail -2 | cut -c1-900
echo "=== bench sdpa+noprefetch"; TRLX_B200_ATTENTION=sdpa TRLX_B200_WEIGHT_PREFETCH=0 BENCH_BREAKDOWN=1 timeout 250 python bench.py --steps 6 --warmup 4 2>&1 | tail -2 | cut -c1-330
