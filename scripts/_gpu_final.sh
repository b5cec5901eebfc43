# final single-GPU pass: the driver's tiers (pytest -m gpu, smoke, bench + reference arm), then the sanitizer passes for the new kernels
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider 2>&1 | tail -6
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench"; BENCH_BREAKDOWN=1 timeout 300 python bench.py 2>&1 | tail -2 | cut -c1-1800
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-1800
echo "=== sanitizer (new kernels)"; SAN_ONLY="new.memcheck new.racecheck" SAN_TIMEOUT=240 bash scripts/sanitize.sh gpurun_out/sanitize 2>&1 | tail -8
