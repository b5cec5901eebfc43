"""Fused compute+collective kernels vs their NCCL(+cuBLAS) equivalents on N GPUs of one node (launch with torch.distributed.run).
Device-timed with CUDA events, max over ranks.  Prints one JSON line per measurement with the achieved fraction of the
NVLink roofline (bytes moved per rank / 900 GB/s per direction) or of the GEMM roofline, whichever is slower."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trlx_b200 import ops  # noqa: E402
from trlx_b200.parallel.fused_tp import FusedTP  # noqa: E402
from trlx_b200.parallel.optim import FusedAdamW  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
NVLINK = 770e9  # bytes/s per direction per GPU: measured peer-copy bandwidth on this pool (B200_PROFILING.md; nominal 900e9)
GEMM_PEAK = 1.46e15  # sustained cuBLAS bf16 on this part (MEASURED_PEAKS.json)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e-3  # seconds, max over ranks


def report(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


PARTS = os.environ.get("BENCH_COLL_PARTS", "opt,ag,rs").split(",")
TAG = {k: os.environ[k] for k in ("TRLX_B200_TP_RS_NVLS", "TRLX_B200_TP_RS_SPLIT") if k in os.environ}

# ---- 1. fused reduce-scatter + AdamW + all-gather vs NCCL RS + torch AdamW + NCCL AG -------------------------------------
n = 64 * 1024 * 1024 if "opt" in PARTS else 1 << 20  # 64 M bf16 parameters
p = torch.nn.Parameter(torch.randn(n, device=dev).to(torch.bfloat16))
opt = FusedAdamW([p], lr=1e-4, process_group=None).prepare()
p.grad.copy_(torch.randn_like(p))
t_fused = timed(lambda: opt.step())
moved = 2 * n * 2 * (world - 1) / world  # RS reads + AG writes, bytes per rank over NVLink
shard = n // world
q = torch.randn(n, device=dev).to(torch.bfloat16)
gq = torch.randn_like(q)
master = q[:shard].float()
m1, m2 = torch.zeros_like(master), torch.zeros_like(master)
gs = torch.empty(shard, device=dev, dtype=torch.bfloat16)


def nccl_step():
    dist.reduce_scatter_tensor(gs, gq)
    g32 = gs.float() / world
    m1.mul_(0.9).add_(g32, alpha=0.1)
    m2.mul_(0.999).addcmul_(g32, g32, value=0.001)
    master.mul_(1 - 1e-4 * 0.01).addcdiv_(m1, m2.sqrt().add_(1e-8), value=-1e-4)
    dist.all_gather_into_tensor(q, master.to(torch.bfloat16))


t_nccl = timed(nccl_step)
if "opt" in PARTS:
    report(op="rs+adamw+ag", params=n, world=world, fused_ms=round(t_fused * 1e3, 3), nccl_ms=round(t_nccl * 1e3, 3),
           nvlink_bytes_per_rank=int(moved), fused_nvlink_fraction=round(moved / NVLINK / t_fused, 3),
           speedup_vs_nccl=round(t_nccl / t_fused, 2))

# ---- 2. all-gather -> GEMM and GEMM -> reduce-scatter (TP = world), GPT-NeoX-20B-like block shapes -----------------------------
tp = FusedTP(None, rank, world, dev)
for (name, tokens, K, N) in ((("qkv 20B", 8192, 6144, 3 * 6144), ("mlp-up 20B", 8192, 6144, 4 * 6144)) if "ag" in PARTS else ()):
    m = tokens // world  # sequence shard per rank
    n_loc = N // world
    x = (torch.randn(m, K, device=dev) * 0.1).to(torch.bfloat16)
    w = (torch.randn(n_loc, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    t_f = timed(lambda: tp.allgather_gemm(x, w, None))
    full = torch.empty(tokens, K, device=dev, dtype=torch.bfloat16)

    def nccl_ag_gemm():
        dist.all_gather_into_tensor(full, x)
        return full @ w.t()

    t_n = timed(nccl_ag_gemm)
    flops = 2.0 * tokens * K * n_loc
    link = (world - 1) * m * K * 2  # bytes this rank pulls from its peers
    roof = max(flops / GEMM_PEAK, link / NVLINK)
    report(op="allgather->gemm", shape=name, world=world, fused_ms=round(t_f * 1e3, 3), nccl_cublas_ms=round(t_n * 1e3, 3),
           tflops=round(flops / t_f / 1e12, 1), roofline_fraction=round(roof / t_f, 3), speedup=round(t_n / t_f, 2))
for (name, tokens, Kfull, N) in ((("attn-out 20B", 8192, 6144, 6144), ("mlp-down 20B", 8192, 4 * 6144, 6144)) if "rs" in PARTS else ()):
    k_loc = Kfull // world
    x = (torch.randn(tokens, k_loc, device=dev) * 0.1).to(torch.bfloat16)
    w = (torch.randn(N, k_loc, device=dev) * Kfull ** -0.5).to(torch.bfloat16)
    t_f = timed(lambda: tp.gemm_reduce_scatter(x, w))
    out = torch.empty(tokens // world, N, device=dev, dtype=torch.bfloat16)

    def nccl_gemm_rs():
        y = x @ w.t()
        dist.reduce_scatter_tensor(out, y)
        return out

    t_n = timed(nccl_gemm_rs)
    flops = 2.0 * tokens * k_loc * N
    link = (world - 1) * (tokens // world) * N * 2  # bf16 partials of the other ranks arriving at the owner
    roof = max(flops / GEMM_PEAK, link / NVLINK)
    report(op="gemm->reduce_scatter", shape=name, world=world, fused_ms=round(t_f * 1e3, 3), nccl_cublas_ms=round(t_n * 1e3, 3),
           tflops=round(flops / t_f / 1e12, 1), roofline_fraction=round(roof / t_f, 3), speedup=round(t_n / t_f, 2), **TAG)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
