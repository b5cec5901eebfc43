N=${1:-4}
echo "=== bench $N GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench$N.log 2>&1
grep -E "Error|Traceback|File \"/root|File \"/tmp|trap" gpurun_out/bench$N.log | grep -v "errors.html\|error_file" | head -10 | cut -c1-260
grep -E "^\{\"metric" gpurun_out/bench$N.log | cut -c1-1900
echo "=== fused collectives TP=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 scripts/bench_collectives.py > gpurun_out/collectives_tp${N}_v3.jsonl 2> gpurun_out/collectives_tp${N}_v3.err
cat gpurun_out/collectives_tp${N}_v3.jsonl | cut -c1-400; tail -3 gpurun_out/collectives_tp${N}_v3.err | cut -c1-300
