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
