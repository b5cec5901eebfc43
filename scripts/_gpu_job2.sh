nvidia-smi -L | wc -l
echo "=== bench N=4"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench4.log 2>&1; echo "rc=$?"; grep -E '^\{' gpurun_out/bench4.log | cut -c1-900; grep -E "Error|error" gpurun_out/bench4.log | head -5; tail -3 gpurun_out/bench4.log | cut -c1-300
echo "=== bench N=2"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2.log 2>&1; echo "rc=$?"; grep -E '^\{' gpurun_out/bench2.log | cut -c1-400; grep -E "Error|error" gpurun_out/bench2.log | head -5
echo "=== multigpu tests (4 ranks where supported)"
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q -x --timeout=250 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -10
