"""Issue / completion cost of streams of small tcgen05.mma atoms (see csrc/umma_probe.cu) → JSON lines."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trlx_b200 import ops  # noqa: E402

n = 512
for M in (64, 128):
    for N in (16, 128, 256):
        if M == 128 and N % 16:
            continue
        for nacc in (1, 4, 14, 24, 101, 104, 114, 124):
            if (1 if nacc % 100 == 1 else 4) * N > 512:
                continue
            ops.C.umma_probe(M, N, n, nacc)  # warm-up
            torch.cuda.synchronize()
            t = ops.C.umma_probe(M, N, n, nacc).cpu().tolist()
            print(json.dumps({"M": M, "N": N, "K": 16, "n_mma": n, "mode": ("warp-uniform, " if nacc >= 100 else "one thread, ") + {1: "1 acc", 4: "4 acc", 14: "4 acc + commit/4", 24: "4 acc + commit/4 + try_wait/4"}[nacc % 100], "issue_cycles_per_mma": round(t[0] / n, 1),
                              "cycles_per_mma": round(t[1] / n, 1)}), flush=True)
