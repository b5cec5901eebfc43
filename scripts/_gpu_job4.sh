echo "=== bwd gemm"; timeout 200 python scripts/bench_gemm.py bwd train_fc train_proj 2>&1 | tail -8
echo "=== profile"; BENCH_PROFILE=1 BENCH_PROFILE_GRAPH=1 timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-300
