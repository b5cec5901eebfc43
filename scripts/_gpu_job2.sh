nvidia-smi -L | head -2
echo "=== bench N=2 (no train graph)"
TRLX_B200_TRAIN_GRAPH=0 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench2_nograph.log 2>&1; echo "rc=$?"; grep -E '^\{' gpurun_out/bench2_nograph.log | cut -c1-700; grep -E "Error|error" gpurun_out/bench2_nograph.log | head -5; tail -5 gpurun_out/bench2_nograph.log | cut -c1-300
echo "=== bench N=2 (train graph)"
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench2_graph.log 2>&1; echo "rc=$?"; grep -E '^\{' gpurun_out/bench2_graph.log | cut -c1-700; grep -E "Error|error" gpurun_out/bench2_graph.log | head -5; tail -5 gpurun_out/bench2_graph.log | cut -c1-300
