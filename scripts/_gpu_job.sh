bash scripts/gpu_test_groups.sh 2>&1 | grep -E "===|passed|failed|^E  " | paste - - | cut -c1-150 | head -40
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_trainers_gpu.py -m gpu -q -x --timeout=300 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -12
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench"; BENCH_BREAKDOWN=1 timeout 300 python bench.py 2>&1 | tail -2 | cut -c1-1800
echo "=== bench reference arm"; timeout 100 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-300
