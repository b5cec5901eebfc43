echo "=== csk tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "cluster_split_k or cta_pair or test_gemm_matches_fp32" -p no:cacheprovider 2>&1 | tail -5
echo "=== chain"; timeout 200 python scripts/bench_chain.py 2>&1 | tail -3
echo "=== bench"; BENCH_BREAKDOWN=1 timeout 250 python bench.py --steps 6 --warmup 4 2>&1 | tail -2 | cut -c1-900
echo "=== bench no csk"; B200_GEMM_NO_CSK=1 BENCH_BREAKDOWN=1 timeout 250 python bench.py --steps 6 --warmup 4 2>&1 | tail -2 | cut -c1-400
