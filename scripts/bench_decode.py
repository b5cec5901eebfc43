"""Decode latency of the rollout engine: milliseconds per generated token (whole decode graph: embed → blocks → LM-head
sampling → value head → bookkeeping) for GPT-2 124M at batch 128, megakernel vs the kernel-per-op graph.

    python scripts/bench_decode.py [--model gpt2] [--batch 128] [--new 40] > gpurun_out/decode_latency.jsonl
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--prompt", type=int, default=8)
    ap.add_argument("--new", type=int, default=40)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from trlx_b200 import ops
    from trlx_b200.engine.rollout import RolloutEngine
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx_b200.utils.modeling import freeze_bottom_causal_layers

    torch.manual_seed(0)
    m = AutoModelForCausalLMWithHydraValueHead.from_pretrained(args.model, num_layers_unfrozen=2)
    freeze_bottom_causal_layers(m.base_model, 2)
    m = m.cuda().to(torch.bfloat16).eval()
    V = m.base_model.config.vocab_size
    gen = dict(max_new_tokens=args.new, do_sample=True, eos_token_id=V - 1, pad_token_id=V - 1, top_k=0, top_p=1.0,
               min_new_tokens=args.new)  # no early EOS: every step decodes the full batch
    ids = torch.randint(0, V - 2, (args.batch, args.prompt), device="cuda")
    mask = torch.ones_like(ids)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mega in ("1", "0"):
        os.environ["TRLX_B200_DECODE_MEGA"] = mega
        eng = RolloutEngine(m, V - 1, V - 1, gen, seed=0)
        for _ in range(3):
            eng.rollout(ids, mask)
        st = eng._state
        torch.cuda.synchronize()
        # (a) full rollout (prefill + decode loop + deferred reference scoring), (b) decode graph replays only
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(args.iters):
            flush.fill_(1)
            a.record()
            eng.rollout(ids, mask)
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        rollout_ms = tot / args.iters
        eng._reset(st, mask.sum(1), ids[:, -1])
        tot = 0.0
        n = 0
        for _ in range(args.iters):
            eng._reset(st, mask.sum(1), ids[:, -1])
            flush.fill_(1)
            a.record()
            for _ in range(args.new):
                st["graph"].replay()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
            n += args.new
        out = {"model": args.model, "batch": args.batch, "new_tokens": args.new, "megakernel": eng.mega,
               "ms_per_token": round(tot / n, 4), "rollout_ms": round(rollout_ms, 3), "kernels_per_step": eng.launches_per_step}
        if eng.mega:
            spec = eng.spec
            groups = (args.batch + 15) // 16
            cs = int(ops.C.decode_mega_cluster_size(spec.hidden_size, spec.ffn_size, groups))
            out["cluster_size"] = cs
            out["max_active_clusters"] = {str(c): int(ops.C.decode_mega_max_clusters(spec.hidden_size, spec.ffn_size, c)) for c in (16, 14, 12, 10, 8)}
            out["ring_stages"] = int(ops.C.decode_mega_stages(spec.hidden_size, spec.ffn_size))
            # streamed weight bytes per token (policy stack) / time → achieved fraction of the measured HBM copy bandwidth
            wbytes = sum(W.qkv_w.numel() + W.out_w.numel() + W.up_w.numel() + W.down_w.numel() for W in eng.layers) * 2
            out["stack_weight_mb"] = round(wbytes / 2 ** 20, 1)
        tm = (st.get("mega") or {}).get("timing")
        if tm is not None:
            L = len(eng.layers)
            t = tm.view(16, L, 16).cpu()
            mhz = 1965.0
            lay = min(5, L - 1)
            per_rank = {}
            for r in (0, 5, 11, 12, 15):
                row = t[r, lay]
                nz = row[row > 0]
                per_rank[str(r)] = [round(float(x - nz[0]) / mhz, 2) for x in nz]
            out["phase_stamps_us_layer%d" % lay] = per_rank
            out["layer_us_rank0"] = [round(float(t[0, l + 1, 0] - t[0, l, 0]) / mhz, 2) for l in range(L - 1)]
        print(json.dumps(out), flush=True)
        del eng


if __name__ == "__main__":
    main()
