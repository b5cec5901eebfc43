nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -x --timeout=300 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 scripts/bench_collectives.py 2>&1 | grep -E '^\{|Error|error' | head -12
