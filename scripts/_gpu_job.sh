for t in test_gemm_fp8_with_row_and_channel_scales test_norm test_decode_attention; do timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$t" --timeout=45 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head -6; done
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --timeout=250 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -8
echo "=== bench (uniform carveout)"; BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-420
echo "=== bench (default carveout)"; B200_NO_UNIFORM_CARVEOUT=1 BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-420
echo "=== bench (fp8 rollout)"; TRLX_B200_ROLLOUT_FP8=1 BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-420
