"""Device-timed forward attention on the shapes of the PPO benchmark: CUDA-core kernel vs tcgen05 kernel vs library SDPA."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trlx_b200 import ops  # noqa: E402

C = ops.C


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


for name, B, H, T, d in [("prefill", 128, 12, 24, 64), ("ref_scoring", 128, 12, 63, 64), ("train", 32, 12, 56, 64),
                         ("long", 16, 32, 128, 128)]:
    qkv = (torch.randn(B, T, 3 * H * d, device="cuda") * 0.5).to(torch.bfloat16)
    q, k, v = (t.view(B, T, H, d).transpose(1, 2) for t in qkv.split(H * d, dim=-1))
    mask = torch.ones(T, T, dtype=torch.bool, device="cuda").tril()
    bias = torch.zeros(B, 1, T, T, device="cuda").masked_fill(~mask, torch.finfo(torch.float32).min)
    rec = dict(shape=name, B=B, H=H, T=T, d=d,
               cuda_core_us=round(timeit(lambda: C.attn_short_fwd(q, k, v, bias, False, d ** -0.5)), 2),
               tcgen05_us=round(timeit(lambda: C.attn_tc_fwd(q, k, v, bias, False, d ** -0.5)), 2),
               sdpa_us=round(timeit(lambda: F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype), scale=d ** -0.5)), 2))
    print(json.dumps(rec), flush=True)
