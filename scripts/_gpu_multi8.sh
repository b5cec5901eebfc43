N=${1:-8}
echo "=== bench $N GPUs"; BENCH_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench$N.full.log 2>&1
grep -E "Error|error|Traceback|raise |File \"/" gpurun_out/bench$N.full.log | grep -v "errors.html\|error_file" | head -30 | cut -c1-300
tail -3 gpurun_out/bench$N.full.log | cut -c1-1800
