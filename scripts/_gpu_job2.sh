nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x --timeout=300 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -30
echo "=== bench N=2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 6 --warmup 3 2>&1 | grep -E "^\{|Error|error|Traceback" | tail -5 | cut -c1-900
