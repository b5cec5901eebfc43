echo "=== engine tests"; timeout 400 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -p no:cacheprovider -k "megakernel or fp32_oracle or greedy" 2>&1 | tail -5
echo "=== decode bench"; TRLX_B200_MEGA_TIMING=1 timeout 300 python scripts/bench_decode.py 2>&1 | tail -4
