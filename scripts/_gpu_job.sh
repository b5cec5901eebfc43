bash scripts/gpu_test_groups.sh 2>&1 | grep -E "===|passed|failed|rc=[1-9]|Error" | head -60
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --timeout=200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
echo "=== bench"; BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-1500
