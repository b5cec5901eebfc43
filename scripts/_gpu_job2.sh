# 1-GPU: new kernel tests, fresh kernel tables of the graphed train step / rollout, BASELINE config B4 in fp8
mkdir -p gpurun_out
echo "=== new tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "adam8bit or colsum or ln_" 2>&1 | tail -6
echo "=== profile"; BENCH_PROFILE=1 BENCH_PROFILE_GRAPH=1 timeout 400 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-300
ls gpurun_out | head -30
echo "=== configs llama_lora_fp8"; timeout 900 python scripts/bench_configs.py --config llama_lora_fp8 --steps 2 --warmup 1 2>&1 | tail -3 | cut -c1-900
