echo "=== graph-mode profile"; BENCH_PROFILE=1 BENCH_PROFILE_GRAPH=1 BENCH_PROFILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-200
python - <<'PY'
import json
ev = json.load(open("gpurun_out/trace_train_graph.json"))["traceEvents"]
k = [e for e in ev if e.get("cat") == "kernel"]
k.sort(key=lambda e: e["ts"])
busy = sum(e["dur"] for e in k)
span = k[-1]["ts"] + k[-1]["dur"] - k[0]["ts"]
print("train graph: kernels", len(k), "busy_us", round(busy), "span_us", round(span))
gaps = sorted(((k[i+1]["ts"] - (k[i]["ts"] + k[i]["dur"]), k[i]["name"][:50], k[i+1]["name"][:50]) for i in range(len(k)-1)), reverse=True)[:12]
for g in gaps: print(round(g[0],1), g[1], "->", g[2])
# longest kernels
for e in sorted(k, key=lambda e: -e["dur"])[:12]: print(round(e["dur"],1), e["name"][:90])
PY
rm -f gpurun_out/trace_*.json
