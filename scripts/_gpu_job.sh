timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_trainers_gpu.py -m gpu -q -x --timeout=250 --timeout-method=thread -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error" | head -12
echo "=== bench"; BENCH_BREAKDOWN=1 timeout 300 python bench.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-1700
echo "=== rollout trace without PDL"; TRLX_B200_PDL=0 BENCH_PROFILE=1 BENCH_PROFILE_GRAPH=1 BENCH_PROFILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-160
python - <<'PY'
import json, collections
ev = json.load(open("gpurun_out/trace_rollout_graph.json"))["traceEvents"]
k = [e for e in ev if e.get("cat") == "kernel"]
k.sort(key=lambda e: e["ts"])
# find the decode region: after the last cudnn attention fprop (prefill) … take the longest run of decode_step_kernel intervals
idx = [i for i, e in enumerate(k) if "decode_step_kernel" in e["name"]]
print("decode steps seen", len(idx))
if len(idx) > 10:
    a, b = idx[5], idx[6]
    seg = k[a + 1:b + 1]
    span = seg[-1]["ts"] + seg[-1]["dur"] - seg[0]["ts"]
    busy = sum(e["dur"] for e in seg)
    print("one decode step: kernels", len(seg), "span_us", round(span, 1), "sum_dur_us", round(busy, 1))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in seg:
        agg[e["name"][:60]][0] += 1; agg[e["name"][:60]][1] += e["dur"]
    for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]: print(f"   {d:8.1f}us {c:4d}x avg {d/c:6.2f}  {n}")
    gaps = [seg[i+1]["ts"] - (seg[i]["ts"] + seg[i]["dur"]) for i in range(len(seg)-1)]
    pos = [g for g in gaps if g > 0]
    print("   gaps>0:", len(pos), "sum", round(sum(pos), 1), "max", round(max(pos) if pos else 0, 1))
PY
rm -f gpurun_out/trace_*.json
