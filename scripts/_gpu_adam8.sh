timeout 100 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "adam8bit" 2>&1 | tail -4
timeout 60 python - <<'PY'
import os, torch
from trlx_b200 import ops
C = ops.C
n = 32 * 1024 * 1024
p = torch.randn(n, device="cuda").to(torch.bfloat16); g = (torch.randn(n, device="cuda") * 0.01).to(torch.bfloat16)
mq = torch.zeros(n, dtype=torch.int8, device="cuda"); vq = torch.zeros(n, dtype=torch.uint8, device="cuda")
ms = torch.full((n // 256,), 1e-12, device="cuda"); vs = ms.clone()
for warp in ("0", "1"):
    os.environ["TRLX_B200_ADAM8BIT_WARP"] = warp
    for s in range(1, 4): C.adam8bit(p, g, mq, ms, vq, vs, 1e-3, 0.9, 0.95, 1e-8, 0.01, True, s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(4, 14): C.adam8bit(p, g, mq, ms, vq, vs, 1e-3, 0.9, 0.95, 1e-8, 0.01, True, s)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
    print({"kernel": "adam8bit", "warp_variant": warp, "params": n, "us": round(us, 1), "GBps": round(n * 10 / us / 1e3, 1)})
PY
