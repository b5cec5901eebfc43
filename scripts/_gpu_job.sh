# Round-end style validation on one B200 (what the driver runs, plus the reference arm):
#   gpurun --timeout 900 -- 'bash scripts/_gpu_job.sh > gpurun_out/validate.log 2>&1; tail -c 3000 gpurun_out/validate.log'
echo "=== pytest -m gpu"; timeout 700 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench"; BENCH_BREAKDOWN=1 timeout 300 python bench.py 2>&1 | tail -2 | cut -c1-1800
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference 2>&1 | tail -3 | cut -c1-1800
