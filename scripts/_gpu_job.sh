for t in test_gemm_matches_fp32 test_gemm_split_k test_gemm_with_folded_norm_and_row_moments test_gemm_epilogue; do timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$t" --timeout=45 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head -6; done
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --timeout=200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -2
echo "=== bench (bm64 + splitk)"; BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-420
echo "=== bench (no bm64)"; B200_GEMM_NO_BM64=1 BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-420
echo "=== gemm microbench"; timeout 200 python scripts/bench_gemm.py decode_qkv decode_fc decode_proj train_proj 2>&1 | tail -4
echo "=== 2-CTA test"; timeout -s KILL 90 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k test_gemm_cta_pair --timeout=60 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -8; echo "2cta rc=$?"
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv | tail -1
