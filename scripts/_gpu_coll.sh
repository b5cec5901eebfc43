# fused TP collective sweep on N GPUs: correctness tests first (2 ranks), then GEMM->RS variants
N=${1:-4}
echo "=== fused TP tests"; timeout 600 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -k "fused_tp" 2>&1 | tail -8
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 scripts/bench_collectives.py 2> gpurun_out/coll.err | cut -c1-400; tail -2 gpurun_out/coll.err | grep -i "error\|trap" | cut -c1-300; }
echo "=== full (NVLS split 2)"; BENCH_COLL_PARTS=opt,ag,rs run 29551 | tee gpurun_out/collectives_tp${N}_v4.jsonl
echo "=== rs split 1"; BENCH_COLL_PARTS=rs TRLX_B200_TP_RS_SPLIT=1 run 29552 | tee -a gpurun_out/collectives_tp${N}_v4.jsonl
echo "=== rs split 4"; BENCH_COLL_PARTS=rs TRLX_B200_TP_RS_SPLIT=4 run 29553 | tee -a gpurun_out/collectives_tp${N}_v4.jsonl
echo "=== rs staged (NVLS off)"; BENCH_COLL_PARTS=rs TRLX_B200_TP_RS_NVLS=0 run 29554 | tee -a gpurun_out/collectives_tp${N}_v4.jsonl
