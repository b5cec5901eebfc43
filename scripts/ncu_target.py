"""Short kernel driver for `ncu --set full` captures (profiles/README.md): a few launches of each hot kernel at the
shapes the PPO benchmark uses, plus a large square GEMM for the roofline comparison."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from trlx_b200 import ops

C = ops.C
dev = "cuda"
torch.manual_seed(0)


def t(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).to(torch.bfloat16)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    a, w = t(8192, 8192), t(8192, 8192)          # roofline shape
    for _ in range(3):
        C.gemm(a, w, None, None, "none")
    a, w, b = t(128, 768), t(2304, 768), t(2304)  # decode QKV (launch/latency bound)
    for _ in range(3):
        C.gemm(a, w, b, None, "none")
    a, w, b = t(1792, 768), t(3072, 768), t(3072)  # training MLP up-projection + GELU
    for _ in range(3):
        C.gemm(a, w, b, None, "gelu_tanh")
if which in ("all", "csk"):
    # decode shapes on the cluster split-K kernel: out-proj (768 -> 768) and MLP down (3072 -> 768) at batch 128
    for n, k in ((768, 768), (768, 3072)):
        a, w, b, r = t(128, k), t(n, k), t(n), t(128, n)
        for _ in range(3):
            C.gemm(a, w, b, r, "none")
if which in ("all", "dlogits"):
    a, w = t(1280, 768), t(50304, 768)            # LM-head backward recompute shape (short K, huge N)
    for _ in range(2):
        C.gemm(a, w, None, None, "none")
if which in ("all", "lmhead"):
    h, w = t(1280, 768), t(50257, 768)
    lab = torch.randint(0, 50257, (1280,), device=dev)
    for _ in range(3):
        C.lmhead(h, w, None, lab)
if which in ("all", "optim"):
    # HBM-bound optimizer kernels: flat AdamW (bf16 param + fp32 master / moments) and the 8-bit-state variant, 32 M parameters
    n = 32 * 1024 * 1024
    master = torch.randn(n, device=dev)
    param, grad = master.to(torch.bfloat16), (torch.randn(n, device=dev) * 0.01).to(torch.bfloat16)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    hyper = torch.tensor([1e-3, 0.1, 0.05, 1.0], device=dev)
    for _ in range(2):
        C.adamw_flat(param, master, grad, m, v, 0.9, 0.95, 1e-8, 0.01, True, hyper)
    mq, vq = torch.zeros(n, dtype=torch.int8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    ms, vs = torch.full((n // 256,), 1e-12, device=dev), torch.full((n // 256,), 1e-12, device=dev)
    for step in (1, 2):
        C.adam8bit(param, grad, mq, ms, vq, vs, 1e-3, 0.9, 0.95, 1e-8, 0.01, True, step)
torch.cuda.synchronize()
