echo "=== chain (default)"; timeout 200 python scripts/bench_chain.py 2>&1 | tail -2
echo "=== chain (bm64)"; B200_GEMM_BM64=1 timeout 200 python scripts/bench_chain.py 2>&1 | tail -2
echo "=== chain (direct store)"; B200_GEMM_DIRECT_STORE=1 timeout 200 python scripts/bench_chain.py 2>&1 | tail -2
echo "=== chain (8 weights: L2 resident)"; CHAIN_WEIGHTS=8 timeout 200 python scripts/bench_chain.py 2>&1 | tail -2
