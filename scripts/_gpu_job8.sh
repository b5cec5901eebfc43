for v in "X=1" "TRLX_B200_ATTENTION=sdpa" "TRLX_B200_WEIGHT_PREFETCH=0" "TRLX_B200_FOLD_NORMS=1"; do
echo "=== bench $v"; env $v BENCH_BREAKDOWN=1 timeout 250 python bench.py --steps 6 --warmup 4 2>&1 | tail -2 | cut -c1-330
done
