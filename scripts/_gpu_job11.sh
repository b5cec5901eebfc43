echo "=== norm/colsum tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "training_norm or column_sum or linear_autograd or fused_logprob_autograd" -p no:cacheprovider 2>&1 | tail -4
echo "=== trainer tests"; timeout 400 python -m pytest tests/test_trainers_gpu.py tests/test_engine_gpu.py -m gpu -q -x --timeout=300 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -12
for v in "X=1" "TRLX_B200_NORM=torch"; do
echo "=== bench $v"; env $v BENCH_BREAKDOWN=1 timeout 250 python bench.py --steps 6 --warmup 4 2>&1 | tail -2 | cut -c1-330
done
