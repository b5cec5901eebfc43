echo "=== bench"; BENCH_BREAKDOWN=1 timeout 200 python bench.py --steps 8 --warmup 4 2>&1 | tail -2 | cut -c1-1800
SAN_TIMEOUT=170 SAN_ONLY="gemm.memcheck plain.memcheck plain.racecheck" bash scripts/sanitize.sh gpurun_out/sanitize 2>&1 | tail -20
