echo "=== new tests"; timeout 150 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider -k "quant_rows or fp8 or lora" 2>&1 | tail -6
echo "=== llama_lora_fp8 (all four GEMMs e4m3)"; timeout 150 python scripts/bench_configs.py --config llama_lora_fp8 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-600
