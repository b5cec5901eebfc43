echo "=== 1-CTA"; timeout 200 python scripts/bench_gemm.py square big_mlp lmhead_pad train_fc 2>&1 | tail -4
echo "=== 2-CTA"; B200_GEMM_2CTA=1 timeout 200 python scripts/bench_gemm.py square big_mlp lmhead_pad train_fc 2>&1 | tail -4
B200_GEMM_2CTA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_2cta_kernel -c 3 -o gpurun_out/prof_gemm_2cta python scripts/ncu_target.py gemm > gpurun_out/ncu_2cta.log 2>&1; tail -2 gpurun_out/ncu_2cta.log
for t in test_gemm_matches_fp32 test_gemm_persistent_many_tiles test_gemm_epilogue test_linear_autograd; do B200_GEMM_2CTA=1 timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$t" --timeout=45 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head -4; done
