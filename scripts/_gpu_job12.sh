echo "=== norm tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "training_norm or column_sum" -p no:cacheprovider 2>&1 | tail -3
for v in "X=1" "TRLX_B200_NORM=torch"; do
echo "=== bench $v"; env $v BENCH_BREAKDOWN=1 timeout 250 python bench.py --steps 8 --warmup 4 2>&1 | tail -2 | cut -c1-330
done
