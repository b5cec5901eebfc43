echo "=== sweep"; CHAIN_SWEEP=1 timeout 200 python scripts/bench_chain.py 2>&1 | tail -5
echo "=== bwd"; timeout 100 python scripts/bench_gemm.py bwd 2>&1 | tail -5 | cut -c1-330
echo "=== tests"; timeout 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "split_k or mn_major" -p no:cacheprovider 2>&1 | tail -2
