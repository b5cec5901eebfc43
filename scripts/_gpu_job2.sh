nvidia-smi -L | wc -l
echo "=== bench N=8"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench8.log 2>&1; echo "rc=$?"; grep -E '^\{' gpurun_out/bench8.log | cut -c1-1000; grep -E "Error|error" gpurun_out/bench8.log | head -5; tail -2 gpurun_out/bench8.log | cut -c1-200
