# Round-end style validation on one B200 (what the driver runs, plus the reference arm):
#   gpurun --timeout 1500 -- 'bash scripts/_gpu_job.sh > gpurun_out/validate.log 2>&1; tail -c 3000 gpurun_out/validate.log'
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== decode bench"; TRLX_B200_MEGA_TIMING=1 timeout 300 python scripts/bench_decode.py 2>&1 | tail -3 | cut -c1-1500
echo "=== bench (megakernel off)"; TRLX_B200_DECODE_MEGA=0 BENCH_BREAKDOWN=1 timeout 300 python bench.py 2>&1 | tail -2 | cut -c1-1800
echo "=== bench (megakernel on)"; TRLX_B200_DECODE_MEGA=1 timeout 300 python bench.py 2>&1 | tail -1 | cut -c1-900
if [ -n "$WITH_REF" ]; then echo "=== bench reference arm"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-1800; fi
