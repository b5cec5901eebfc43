"""Per-kernel latency of dependent small GEMMs inside one CUDA graph (the decode regime): a chain y = gemm(y, W_i) over many
distinct weight matrices (cold in L2, like consecutive layers), timed per kernel for several tile widths."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trlx_b200 import ops  # noqa: E402

C = ops.C
dev = "cuda"
torch.manual_seed(0)
M, H = 128, 768
n_w = int(os.environ.get("CHAIN_WEIGHTS", 48))
Ws = [(torch.randn(H, H, device=dev) * H ** -0.5).to(torch.bfloat16) for _ in range(n_w)]          # square: chainable
Wq = [(torch.randn(2304, H, device=dev) * H ** -0.5).to(torch.bfloat16) for _ in range(n_w)]        # QKV-shaped (not chained)
big = [(torch.randn(3072, H, device=dev) * H ** -0.5).to(torch.bfloat16) for _ in range(n_w)]
x0 = (torch.randn(M, H, device=dev)).to(torch.bfloat16)
g = torch.ones(H, device=dev, dtype=torch.bfloat16)


def timed_graph(fn, iters=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def chain_square(bn):
    def fn():
        y = x0
        for W in Ws:
            y = C.gemm(y, W, None, None, "none", force_bn=bn)
        return y
    return fn


def chain_qkv(bn, weights):
    def fn():
        y = x0
        outs = []
        for W in weights:  # dependent through a cheap slice so the kernels serialise like layers do
            o = C.gemm(y, W, None, None, "none", force_bn=bn)
            y = o[:, :H]
            if y.stride(0) % 8:
                y = y.contiguous()
        return y
    return fn


def chain_norm():
    def fn():
        y = x0
        for _ in range(n_w):
            y = C.norm(y, g, None, 1e-5, False)
        return y
    return fn


down = [(torch.randn(H, 3072, device=dev) * 3072 ** -0.5).to(torch.bfloat16) for _ in range(n_w)]


def chain_mlp(bn):
    """fc (768 -> 3072, GELU) then fc2 (3072 -> 768): the two GEMMs of a decoder MLP, properly dependent."""
    def fn():
        y = x0
        for Wu, Wd in zip(big, down):
            y = C.gemm(C.gemm(y, Wu, None, None, "gelu_new", force_bn=bn), Wd, None, None, "none", force_bn=bn)
        return y
    return fn


def chain_one(weights, act="none"):
    """Dependent chain over one weight shape [N, 768] -> feeds the first 768 columns on (auto plan / env override)."""
    def fn():
        y = x0
        for W in weights:
            o = C.gemm(y, W, None, None, act)
            y = o[:, :H] if o.shape[1] > H else o
        return y
    return fn


def chain_fc2():
    """fc pinned to the plain 128x32 kernel (force_bn > 0), fc2 ([128, 3072] x [768, 3072]^T) on the plan under test."""
    def fn():
        y = x0
        for Wu, Wd in zip(big, down):
            y = C.gemm(C.gemm(y, Wu, None, None, "gelu_new", force_bn=32), Wd, None, None, "none")
        return y
    return fn


if os.environ.get("CHAIN_SWEEP"):
    # cluster split-K tile-width / cluster-size sweep (B200_CSK_FORCE is read at every launch)
    C.set_pdl(True)
    sweep = {"sq768": (Ws, [(32, 6), (32, 4), (32, 3), (32, 2), (64, 6), (64, 4)]),
             "qkv2304": (Wq, [(32, 2), (64, 4), (64, 3), (64, 2)]),
             "fc3072": (big, [(64, 3), (64, 2)])}
    for name, (weights, cfgs) in sweep.items():
        rec = {"shape": name}
        os.environ["B200_GEMM_NO_CSK"] = "1"
        rec["plain"] = round(timed_graph(chain_one(weights)) / n_w, 2)
        del os.environ["B200_GEMM_NO_CSK"]
        for bn, S in cfgs:
            os.environ["B200_CSK_FORCE"] = f"{bn},{S}"
            try:
                rec[f"bn{bn}_S{S}"] = round(timed_graph(chain_one(weights)) / n_w, 2)
            except RuntimeError as e:
                rec[f"bn{bn}_S{S}"] = "n/a"
        os.environ.pop("B200_CSK_FORCE", None)
        print(json.dumps(rec), flush=True)
    rec = {"shape": "fc_plain+fc2"}
    os.environ["B200_GEMM_NO_CSK"] = "1"
    rec["plain"] = round(timed_graph(chain_fc2()) / n_w, 2)
    del os.environ["B200_GEMM_NO_CSK"]
    for bn, S in [(64, 8), (64, 6), (64, 4), (32, 6), (32, 4), (32, 3)]:
        os.environ["B200_CSK_FORCE"] = f"{bn},{S}"
        try:
            rec[f"bn{bn}_S{S}"] = round(timed_graph(chain_fc2()) / n_w, 2)
        except RuntimeError:
            rec[f"bn{bn}_S{S}"] = "n/a"
    os.environ.pop("B200_CSK_FORCE", None)
    print(json.dumps(rec), flush=True)
    sys.exit(0)

for pdl in (1, 0):
    C.set_pdl(bool(pdl))
    rec = {"pdl": pdl, "norm_us": round(timed_graph(chain_norm()) / n_w, 2)}
    for bn in (32, 64, 128):
        rec[f"sq768_bn{bn}_us"] = round(timed_graph(chain_square(bn)) / n_w, 2)
    for bn in (32, 64, 128):
        rec[f"qkv2304_bn{bn}_us"] = round(timed_graph(chain_qkv(bn, Wq)) / n_w, 2)
    rec["fc3072_bn32_us"] = round(timed_graph(chain_qkv(32, big)) / n_w, 2)
    # cluster split-K (force_bn = -3) and the automatic choice (0) against the plain 128x32 kernel
    for tag, bn in (("csk", -3), ("auto", 0)):
        rec[f"sq768_{tag}_us"] = round(timed_graph(chain_square(bn)) / n_w, 2)
        rec[f"qkv2304_{tag}_us"] = round(timed_graph(chain_qkv(bn, Wq)) / n_w, 2)
        rec[f"mlp_pair_{tag}_us"] = round(timed_graph(chain_mlp(bn)) / n_w, 2)
    rec["mlp_pair_bn32_us"] = round(timed_graph(chain_mlp(32)) / n_w, 2)
    print(json.dumps(rec), flush=True)
