# evidence refresh on one B200: GEMM family vs cuBLAS, in-graph decode-chain latency, kernel table of the graphed train step and rollout
mkdir -p gpurun_out
echo "=== gemm bench"; timeout 300 python scripts/bench_gemm.py > gpurun_out/gemm_bench_v6.jsonl 2>gpurun_out/gemm_bench_v6.err; tail -16 gpurun_out/gemm_bench_v6.jsonl | cut -c1-260
echo "=== chain"; timeout 200 python scripts/bench_chain.py > gpurun_out/decode_chain_v3.jsonl 2>/dev/null; tail -8 gpurun_out/decode_chain_v3.jsonl | cut -c1-300
echo "=== profile"; BENCH_PROFILE=1 BENCH_PROFILE_GRAPH=1 timeout 400 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-200
head -40 gpurun_out/profile_train_graph.txt | cut -c1-75,150-230
echo "=== configs"; timeout 900 python scripts/bench_configs.py --config llama_lora_fp8 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-600
