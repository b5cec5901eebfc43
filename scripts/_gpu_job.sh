for t in test_gemm_with_folded_norm_and_row_moments test_gemm_split_k test_gemm_mn_major_operands test_linear_autograd test_fused_logprob_autograd test_embed_rowdot; do timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$t" --timeout=45 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head -12; done
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --timeout=200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -4
echo "=== bench (fold norms)"; BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-500
echo "=== bench (no fold)"; TRLX_B200_FOLD_NORMS=0 BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-500
echo "=== graph-mode profile"; BENCH_PROFILE=1 BENCH_PROFILE_GRAPH=1 BENCH_PROFILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-200
python - <<'PY'
import json, collections
for phase in ("train", "rollout"):
    ev = json.load(open(f"gpurun_out/trace_{phase}_graph.json"))["traceEvents"]
    k = [e for e in ev if e.get("cat") == "kernel"]
    k.sort(key=lambda e: e["ts"])
    busy = sum(e["dur"] for e in k)
    span = k[-1]["ts"] + k[-1]["dur"] - k[0]["ts"]
    print(phase, "graph: kernels", len(k), "busy_us", round(busy), "span_us", round(span))
    gaps = sorted(((k[i+1]["ts"] - (k[i]["ts"] + k[i]["dur"]), k[i]["name"][:50], k[i+1]["name"][:50]) for i in range(len(k)-1)), reverse=True)[:10]
    for g in gaps: print("  gap", round(g[0],1), g[1], "->", g[2])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in k:
        agg[e["name"][:70]][0] += 1; agg[e["name"][:70]][1] += e["dur"]
    for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print(f"  {d:9.1f}us {c:5d}x {n}")
PY
rm -f gpurun_out/trace_*.json
