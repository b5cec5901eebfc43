echo "=== pytest -m gpu (driver command)"; timeout 700 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench (defaults)"; BENCH_BREAKDOWN=1 timeout 300 python bench.py 2>&1 | tail -2 | cut -c1-1500
echo "=== bench reference arm"; timeout 100 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-300
echo "=== ncu csk"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_csk -c 2 -f -o gpurun_out/ncu_csk python scripts/ncu_target.py csk 2>&1 | tail -3
