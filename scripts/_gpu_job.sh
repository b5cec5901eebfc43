for t in test_gemm_mn_major_operands test_linear_autograd test_fused_logprob_autograd test_lmhead_dlogits; do timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$t" --timeout=45 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -4; done
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --timeout=200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
echo "=== bench (tcgen05 backward)"; BENCH_BREAKDOWN=1 BENCH_PROFILE=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-700
echo "=== bench (cublas backward)"; TRLX_B200_BACKWARD_GEMM=cublas BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-400
