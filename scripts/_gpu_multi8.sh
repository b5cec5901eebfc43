N=${1:-8}
echo "=== bench $N GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench$N.log 2>&1
grep -E "Error|Traceback|File \"/root|File \"/tmp|trap" gpurun_out/bench$N.log | grep -v "errors.html\|error_file" | head -20 | cut -c1-260
grep -E "^\{\"metric" gpurun_out/bench$N.log | cut -c1-1900
