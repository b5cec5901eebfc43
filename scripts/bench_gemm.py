"""Device-timed (CUDA events, warm, L2 flushed between iterations) throughput of the tcgen05 GEMM family against
cuBLAS (`torch.matmul`) on the shapes the PPO benchmark uses.  Prints one JSON line per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trlx_b200 import ops  # noqa: E402

C = ops.C
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def t(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).to(torch.bfloat16)


shapes = [("square", 8192, 8192, 8192), ("lmhead_bwd_logits", 1280, 50257, 768), ("lmhead_pad", 1280, 50304, 768),
          ("decode_qkv", 128, 2304, 768), ("decode_fc", 128, 3072, 768), ("decode_proj", 128, 768, 3072),
          ("train_fc", 1792, 3072, 768), ("train_proj", 1792, 768, 3072), ("big_mlp", 8192, 16384, 4096)]
only = sys.argv[1:] or None
for name, M, N, K in shapes:
    if only and name not in only:
        continue
    x, w = t(M, K), t(N, K)
    ours = timeit(lambda: C.gemm(x, w, None, None, "none"))
    lib = timeit(lambda: torch.matmul(x, w.t()))
    fl = 2.0 * M * N * K
    rec = dict(shape=name, M=M, N=N, K=K, ours_us=round(ours, 1), cublas_us=round(lib, 1),
               ours_tflops=round(fl / ours / 1e6, 1), cublas_tflops=round(fl / lib / 1e6, 1))
    if name.startswith("lmhead"):
        lab = torch.randint(0, N, (M,), device=dev)
        fused = timeit(lambda: C.lmhead(x, w, None, lab))
        rec["fused_lmhead_us"] = round(fused, 1)
        rec["fused_lmhead_tflops"] = round(fl / fused / 1e6, 1)
    print(json.dumps(rec), flush=True)

# ---- backward shapes of the PPO update (MN-major operands, split-K): dX = dY·W and dW = dYᵀ·X per linear of a GPT-2 block at
# 1792 tokens, plus the LM head at 1280 scored positions.  Run with `bwd` as the only argument (or no argument).
if not only or "bwd" in only:
    T = 1792
    for name, n_out, k_in, rows in [("qkv", 2304, 768, T), ("proj", 768, 768, T), ("fc", 3072, 768, T), ("fc2", 768, 3072, T),
                                    ("lmhead", 50304, 768, 1280)]:
        dy, w, x = t(rows, n_out), t(n_out, k_in), t(rows, k_in)
        dx_ours = timeit(lambda: C.gemm_ex(dy, w, False, True))
        dx_lib = timeit(lambda: dy @ w)
        dw_ours = timeit(lambda: C.gemm_ex(dy, x, True, True))
        dw_lib = timeit(lambda: dy.t() @ x)
        fl = 2.0 * rows * n_out * k_in
        print(json.dumps(dict(shape="bwd_" + name, rows=rows, n_out=n_out, k_in=k_in,
                              dx_splits=C.gemm_splitk_plan(rows, k_in, n_out) if hasattr(C, "gemm_splitk_plan") else None,
                              dx_ours_us=round(dx_ours, 1), dx_cublas_us=round(dx_lib, 1), dw_ours_us=round(dw_ours, 1),
                              dw_cublas_us=round(dw_lib, 1), dx_ours_tflops=round(fl / dx_ours / 1e6, 1),
                              dw_ours_tflops=round(fl / dw_ours / 1e6, 1))), flush=True)
