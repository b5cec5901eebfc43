echo "=== engine tests"; timeout 500 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6
echo "=== decode bench"; TRLX_B200_MEGA_TIMING=1 timeout 300 python scripts/bench_decode.py 2>&1 | tail -4
