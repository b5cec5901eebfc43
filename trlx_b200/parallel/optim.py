"""Optimizers.

``FusedAdamW`` / ``FusedAdam`` keep every trainable parameter of a rank in ONE flat bf16 buffer (parameters and their
``.grad`` become views into flat storage), hold fp32 master weights + Adam moments only for the shard the rank owns,
and update with a single sm_100a kernel.  Under data parallelism the flat gradient / parameter buffers live in
symmetric memory and the same kernel performs the gradient reduce-scatter (P2P loads from every peer), the AdamW update
of the owned shard and the parameter all-gather (P2P stores to every peer) — see ``csrc/optim.cu``.  This is
the B200 replacement for DDP / DeepSpeed ZeRO-1/2 + ``torch.optim.AdamW`` used by the reference
(``trlx/trainer/accelerate_base_trainer.py:173-193,574-587``, ``configs/accelerate/zero2-bf16.yaml``).

On CPU / fp32 parameters the classes fall back to the equivalent ``torch.optim`` implementation (with an explicit
gradient all-reduce across ranks), so the trainers behave identically in gloo tests.

``Adam8bit`` / ``AdamW8bit`` keep the moments block-quantised to 8 bits (256-element blocks, absmax scales; on a GPU one
``adam8bit_kernel`` launch per tensor decodes, updates and re-encodes them) — an
in-repo stand-in for ``bitsandbytes`` (reference: ``trlx/utils/__init__.py:104-123``).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist
from torch.optim import Optimizer

from trlx_b200.utils import logging

logger = logging.get_logger(__name__)

_ALIGN = 8  # elements; keeps every shard 16-byte aligned for the vectorised kernels


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _FlatGroup:
    """Flat storage for one param group, cut into gradient BUCKETS.

    A bucket is a contiguous run of parameters (in ``model.parameters()`` order, ≈ ``bucket_elems`` elements) padded to a
    multiple of ``8 * world``; every rank owns the ``1 / world`` slice ``[lo + rank * shard, lo + (rank + 1) * shard)`` of
    every bucket and keeps fp32 master weights + Adam moments only for those slices (concatenated, bucket after bucket).
    Backward completes buckets last-to-first, so the reduce-scatter + update + all-gather of a finished bucket runs on a side
    stream while the rest of the backward is still computing (the per-layer buckets + overlap hooks of the reference's
    distributed optimizer, ``trlx/models/modeling_nemo_ppo.py:586-609,689-701``, and DDP's bucketed reducer)."""

    def __init__(self, params: List[torch.nn.Parameter], world: int, rank: int, group, symmetric: bool,
                 bucket_elems: int = 8 << 20):
        self.params = params
        self.device = params[0].device
        self.world, self.rank, self.group = world, rank, group
        unit = _ALIGN * max(world, 1)
        self.offsets: List[int] = []
        self.buckets: List[Dict[str, Any]] = []
        off = start = 0
        members: List[int] = []
        for i, p in enumerate(params):
            self.offsets.append(off)
            off += _round_up(p.numel(), _ALIGN)
            members.append(i)
            if off - start >= bucket_elems or i == len(params) - 1:
                off = start + _round_up(off - start, unit)
                self.buckets.append(dict(lo=start, hi=off, params=members))
                start, members = off, []
        self.numel = off
        moff = 0
        for b in self.buckets:
            b["shard"] = (b["hi"] - b["lo"]) // max(world, 1)
            b["mine"] = b["lo"] + b["shard"] * rank
            b["moff"] = moff
            moff += b["shard"]
        self.shard = moff  # elements of optimizer state this rank holds
        self.bucket_of = {}
        for k, b in enumerate(self.buckets):
            for i in b["params"]:
                self.bucket_of[i] = k
        self.symm_param = self.symm_grad = self.symm_flags = self.symm_sq = None
        self.mc_grad = self.mc_param = 0
        nb = len(self.buckets)
        if symmetric:
            import torch.distributed._symmetric_memory as symm

            self.flat_param = symm.empty(self.numel, dtype=torch.bfloat16, device=self.device)
            self.flat_grad = symm.empty(self.numel, dtype=torch.bfloat16, device=self.device)
            # rows 0..nb-1: "bucket gradients final" flags, row nb: end-of-step barrier, row nb+1: clip-norm exchange
            self.flags = symm.empty((nb + 2) * world, dtype=torch.int32, device=self.device)
            self.sqbuf = symm.empty(2 * world, dtype=torch.float64, device=self.device)
            for t in (self.flat_param, self.flat_grad, self.flags, self.sqbuf):
                t.zero_()
            name = group.group_name if group is not None else dist.group.WORLD.group_name
            self.symm_param = symm.rendezvous(self.flat_param, name)
            self.symm_grad = symm.rendezvous(self.flat_grad, name)
            self.symm_flags = symm.rendezvous(self.flags, name)
            self.symm_sq = symm.rendezvous(self.sqbuf, name)
            import os

            # NVLS (NVSwitch multicast mapping of the symmetric buffers): multimem.ld_reduce / multimem.st in the fused kernel —
            # one in-switch reduction / broadcast instead of `world` peer loads / stores per element (TRLX_B200_NVLS=0 disables)
            if os.environ.get("TRLX_B200_NVLS", "1") == "1":
                self.mc_grad = int(getattr(self.symm_grad, "multicast_ptr", 0) or 0)
                self.mc_param = int(getattr(self.symm_param, "multicast_ptr", 0) or 0)
                if not (self.mc_grad and self.mc_param):
                    self.mc_grad = self.mc_param = 0
        else:
            self.flat_param = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.device)
            self.flat_grad = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.device)
        for p, o in zip(params, self.offsets):
            n = p.numel()
            self.flat_param[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + n].view(p.shape)
            p.grad = self.flat_grad[o:o + n].view(p.shape)
        self.master = torch.empty(self.shard, dtype=torch.float32, device=self.device)
        self.resync_master()
        self.exp_avg = torch.zeros(self.shard, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(self.shard, dtype=torch.float32, device=self.device)
        self.gshard: Optional[torch.Tensor] = None
        self.hyper = torch.ones(4, dtype=torch.float32, device=self.device)
        self.hyper_host = torch.ones(4, dtype=torch.float32).pin_memory() if self.device.type == "cuda" else torch.ones(4)
        self.sq = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.epochs = torch.zeros(nb + 2, dtype=torch.int32, device=self.device)  # device-side epochs (buckets, barrier, clip)
        self.done = torch.zeros(nb + 1, dtype=torch.int32, device=self.device)    # per-bucket finished-block counters
        self.norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.pending = [0] * nb       # overlap: parameters of the bucket still waiting for their gradient
        self.launched = [False] * nb

    def resync_master(self):
        for b in self.buckets:
            self.master[b["moff"]:b["moff"] + b["shard"]].copy_(self.flat_param[b["mine"]:b["mine"] + b["shard"]].float())

    def layout(self):
        return [(b["lo"], b["hi"]) for b in self.buckets]

    def rebind_grads(self):
        """(Re)attach ``.grad`` views — needed after anything set them to ``None``."""
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + o * 2:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)


class FusedAdamW(Optimizer):
    """AdamW (``decoupled=True``) with flat bf16 storage + fp32 master shard and a fused (optionally cross-GPU) update.

    Data-parallel runs (``zero_stage >= 1``): gradients are reduced, the owned shard updated and the new parameters
    written to every peer by ONE kernel per gradient bucket over NVLink peer memory; the cross-rank "gradients of this
    bucket are final" handshake is folded into that kernel.  With ``overlap`` (default) the bucket kernels are launched from
    post-accumulate-grad hooks on a side stream, so all but the first layers' traffic hides behind the backward pass.
    Parameters that receive no gradient in a step are updated with a zero gradient (moment / weight decay still apply) —
    unlike ``torch.optim.AdamW``, which skips them; freeze what must not move."""

    decoupled = True

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 grad_clip: Optional[float] = None, process_group=None, zero_stage: int = 2, overlap: Optional[bool] = None,
                 bucket_mb: Optional[float] = None, local_only: bool = False, **unused):
        import os

        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_clip = grad_clip
        # zero_stage >= 1: optimizer state + gradient reduction sharded (fused RS + AdamW + AG over NVLink peer memory);
        # zero_stage == 0: DDP semantics — NCCL all-reduce of the flat gradient, replicated update
        self.zero_stage = int(zero_stage)
        self.process_group = process_group
        # local_only: the parameters handed in are already this rank's partition and their gradients arrive reduced
        # (ZeRO-3, parallel/zero3.py) — no cross-rank traffic here, only the global-norm exchange when clipping
        self.local_only = bool(local_only)
        self.overlap = (os.environ.get("TRLX_B200_OVERLAP_GRAD", "1") == "1") if overlap is None else bool(overlap)
        mb = float(os.environ.get("TRLX_B200_BUCKET_MB", bucket_mb if bucket_mb is not None else 16))
        self.bucket_elems = max(int(mb * (1 << 20) / 2), 1024)
        # CTAs per overlapped bucket kernel: two per SM.  (32 CTAs — chosen at first so that a kernel waiting for its peers' flags
        # would hold few SMs — made every bucket NVLink-latency bound: 54 dependent peer round trips per thread, +0.7 ms per
        # optimizer step at 2 GPUs; the flag wait only lasts as long as the ranks are out of step.)
        self.overlap_blocks = int(os.environ.get("TRLX_B200_OVERLAP_BLOCKS", "296"))
        self._step_count_fused = 0
        self._flat: Optional[List[Optional[_FlatGroup]]] = None
        self._fallback: Optional[Optimizer] = None
        self._armed = False
        self._side: Optional[torch.cuda.Stream] = None
        self._hooks: List[Any] = []
        self.last_grad_norm: Optional[torch.Tensor] = None

    # -- setup ------------------------------------------------------------------------------------------------------------
    def _world(self):
        if self.local_only:
            return 1, 0
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group), dist.get_rank(self.process_group)
        return 1, 0

    def _use_kernels(self) -> bool:
        from trlx_b200 import ops

        ps = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        return bool(ps) and all(p.is_cuda and p.dtype == torch.bfloat16 for p in ps) and ops.available()

    def _lazy_init(self):
        if self._flat is not None or self._fallback is not None:
            return
        world, rank = self._world()
        if not self._use_kernels():
            cls = torch.optim.AdamW if self.decoupled else torch.optim.Adam
            groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": [p for p in g["params"] if p.requires_grad]}
                      for g in self.param_groups]
            groups = [g for g in groups if g["params"]]
            known = ("lr", "betas", "eps", "weight_decay", "params")
            self._fallback = cls([{k: v for k, v in g.items() if k in known} for g in groups])
            return
        self._flat = []
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            # a param group may name its own reduction group and divisor: `reduce_group` (default: the optimizer's data-
            # parallel group) and `grad_divisor` (default: its size).  Tensor-parallel trainers use it for the parameters that
            # sequence parallelism leaves replicated inside the TP group: reduce over TP x DP, divide by DP only
            pg = g.get("reduce_group", self.process_group)
            if "reduce_group" in g and not self.local_only and dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(pg), dist.get_rank(pg)
            else:
                world, rank = self._world()
            symmetric = world > 1 and self.zero_stage >= 1
            if world > 1 and not symmetric:
                self._flat.append(_FlatGroup(ps, 1, 0, None, False, 1 << 62))
                self._flat[-1].needs_allreduce = True
                self._flat[-1].reduce_group, self._flat[-1].reduce_world = pg, world
                self._flat[-1].grad_divisor = float(g.get("grad_divisor", world))
                continue
            try:
                self._flat.append(_FlatGroup(ps, world, rank, pg, symmetric,
                                             self.bucket_elems if symmetric else 1 << 62))
                self._flat[-1].grad_divisor = float(g.get("grad_divisor", world))
            except Exception as err:  # symmetric memory unavailable → replicated update after an NCCL all-reduce
                if not symmetric:
                    raise
                logger.warning(f"symmetric memory unavailable ({err}); falling back to all-reduce + replicated update")
                self._flat.append(_FlatGroup(ps, 1, 0, None, False, 1 << 62))
                self._flat[-1].needs_allreduce = True
        if self.overlap and not self.grad_clip and any(fg is not None and fg.world > 1 for fg in self._flat):
            self._side = torch.cuda.Stream(device=next(fg for fg in self._flat if fg is not None).device)
            for gi, fg in enumerate(self._flat):
                if fg is None or fg.world == 1:
                    continue
                for pi, p in enumerate(fg.params):
                    self._hooks.append(p.register_post_accumulate_grad_hook(
                        lambda _p, gi=gi, pi=pi: self._on_grad(gi, pi)))

        # weight-gradient GEMMs may add straight into the flat gradient buffer (ops/functional.py: _Linear.backward) instead of
        # returning a tensor for autograd to add (Apex `gradient_accumulation_fusion`, megatron_20b.yaml:72): the parameter
        # carries the callback that stands in for its post-accumulate hook
        if not self.local_only:
            for gi, fg in enumerate(self._flat):
                if fg is None:
                    continue
                for pi, p in enumerate(fg.params):
                    if p.dim() == 2:
                        p._b200_grad_sink = (lambda gi=gi, pi=pi: self._on_grad(gi, pi))

    def prepare(self):
        """Flatten storage now (call once after the model is on its device, before the first backward)."""
        self._lazy_init()
        return self

    # -- overlap ------------------------------------------------------------------------------------------------------------
    @property
    def can_overlap(self) -> bool:
        self._lazy_init()
        return self._side is not None

    def arm_overlap(self):
        """Call right before the backward whose gradients complete an optimizer step (the last micro-batch), after
        :meth:`host_prepare`: from then on every bucket whose gradients are final is reduced / updated / gathered on a side
        stream while the rest of the backward runs.  :meth:`device_step` joins."""
        if not self.can_overlap:
            return
        for fg in self._flat:
            if fg is None or fg.world == 1:
                continue
            fg.hyper.copy_(fg.hyper_host, non_blocking=True)
            for k, b in enumerate(fg.buckets):
                fg.pending[k] = len(b["params"])
                fg.launched[k] = False
        self._armed = True

    def _on_grad(self, gi: int, pi: int):
        if not self._armed:
            return
        fg = self._flat[gi]
        k = fg.bucket_of[pi]
        fg.pending[k] -= 1
        if fg.pending[k] == 0 and not fg.launched[k]:
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self._bucket_kernel(self.param_groups[gi], fg, k, 0, self.overlap_blocks)
            fg.launched[k] = True

    def _bucket_kernel(self, g, fg: _FlatGroup, k: int, mode: int, max_blocks: int = 0):
        from trlx_b200 import ops

        b = fg.buckets[k]
        b1, b2 = g["betas"]
        sl = slice(b["moff"], b["moff"] + b["shard"])
        gsh = fg.gshard[sl] if (mode != 0 and fg.gshard is not None) else None
        ops.C.rs_adamw_ag_bucket(list(fg.symm_grad.buffer_ptrs), list(fg.symm_param.buffer_ptrs), fg.rank, b["mine"],
                                 b["shard"], fg.master[sl], fg.exp_avg[sl], fg.exp_avg_sq[sl], gsh, mode, b1, b2, g["eps"],
                                 g["weight_decay"], self.decoupled, fg.hyper, fg.sq if mode == 1 else None,
                                 list(fg.symm_flags.buffer_ptrs), k * fg.world, fg.epochs[k:k + 1], fg.done[k:k + 1],
                                 max_blocks, fg.mc_grad, fg.mc_param, 1.0 / getattr(fg, "grad_divisor", fg.world))

    def _end_barrier(self, fg: _FlatGroup):
        """Every rank's new parameters are visible everywhere (and nobody still reads this rank's gradients)."""
        from trlx_b200 import ops

        nb = len(fg.buckets)
        pads = [int(p) + nb * fg.world * 4 for p in fg.symm_flags.buffer_ptrs]
        ops.C.signal_barrier(pads, fg.rank, fg.epochs[nb:nb + 1])

    # -- step -------------------------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        self._lazy_init()
        if self._fallback is not None:
            return self._fallback.zero_grad(set_to_none=True)
        for fg in self._flat:
            if fg is not None:
                fg.flat_grad.zero_()
                fg.rebind_grads()
                for p in fg.params:  # in-place wgrad bookkeeping: forwards that were never backpropagated must not linger
                    if getattr(p, "_b200_uses", 0):
                        p._b200_uses = 0

    @torch.no_grad()
    def step(self, closure=None, graph=None):
        """``graph``: a captured CUDA graph whose tail is :meth:`device_step` (whole-train-step graphs) — the host half runs
        here and the graph replay stands in for the device half, so ``scheduler.step()`` keeps following ``optimizer.step()``."""
        self._lazy_init()
        if graph is not None:
            self.host_prepare()
            graph.replay()
            return None
        loss = closure() if closure is not None else None
        world, rank = self._world()
        if self._fallback is not None:
            for g_self, g_fb in zip([g for g in self.param_groups if any(p.requires_grad for p in g["params"])],
                                    self._fallback.param_groups):
                g_fb["lr"] = g_self["lr"]
            params = [p for g in self._fallback.param_groups for p in g["params"] if p.grad is not None]
            live = [g for g in self.param_groups if any(p.requires_grad for p in g["params"])]
            for g_self in live:  # every param group reduces over its own group / divisor (see _lazy_init)
                pg = g_self.get("reduce_group", self.process_group)
                if self.local_only or not (dist.is_available() and dist.is_initialized()):
                    break
                gw = dist.get_world_size(pg)
                gp = [p for p in g_self["params"] if p.requires_grad and p.grad is not None]
                if gw > 1 and gp:
                    flat = torch.cat([p.grad.reshape(-1).float() for p in gp])
                    dist.all_reduce(flat, group=pg)
                    flat /= float(g_self.get("grad_divisor", gw))
                    off = 0
                    for p in gp:
                        p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                        off += p.numel()
            if self.grad_clip:
                if self.local_only and dist.is_available() and dist.is_initialized():
                    sq = torch.stack([p.grad.float().pow(2).sum() for p in params]).sum() if params else torch.zeros(())
                    dist.all_reduce(sq, group=self.process_group)
                    norm = sq.sqrt()
                    coef = (self.grad_clip / (norm + 1e-6)).clamp(max=1.0)
                    for p in params:
                        p.grad.mul_(coef.to(p.grad.dtype))
                    self.last_grad_norm = norm
                else:
                    self.last_grad_norm = torch.nn.utils.clip_grad_norm_(params, self.grad_clip)
            self._fallback.step()
            return loss

        if not self._armed:
            self.host_prepare()
        self.device_step()
        return loss

    def host_prepare(self):
        """Host half of a fused step: advance the step count and stage (lr, bias corrections) in pinned memory."""
        self._lazy_init()
        self._step_count_fused += 1
        t = self._step_count_fused
        for g, fg in zip(self.param_groups, self._flat or []):
            if fg is None:
                continue
            b1, b2 = g["betas"]
            fg.hyper_host[0] = float(g["lr"])
            fg.hyper_host[1] = 1.0 - b1 ** t
            fg.hyper_host[2] = 1.0 - b2 ** t
            fg.hyper_host[3] = 1.0

    def device_step(self):
        """Device half: everything here is stream work only (H2D of the staged scalars + kernels), so it can be
        captured in a CUDA graph and replayed after :meth:`host_prepare`."""
        from trlx_b200 import ops

        C = ops.C
        world, rank = self._world()
        armed, self._armed = self._armed, False
        for g, fg in zip(self.param_groups, self._flat):
            if fg is None:
                continue
            b1, b2 = g["betas"]
            args = (b1, b2, g["eps"], g["weight_decay"], self.decoupled)
            if fg.world == 1:
                fg.hyper.copy_(fg.hyper_host, non_blocking=True)
                if getattr(fg, "needs_allreduce", False) and getattr(fg, "reduce_world", world) > 1:
                    dist.all_reduce(fg.flat_grad, group=getattr(fg, "reduce_group", self.process_group))
                    fg.flat_grad.div_(getattr(fg, "grad_divisor", world))
                if self.grad_clip:
                    fg.sq.zero_()
                    C.sqnorm_(fg.flat_grad, fg.sq)
                    if self.local_only and dist.is_available() and dist.is_initialized():
                        dist.all_reduce(fg.sq, group=self.process_group)  # the partitions' norms add up to the global one
                    C.clip_coef_(fg.sq, float(self.grad_clip), fg.hyper, fg.norm)
                    self.last_grad_norm = fg.norm
                C.adamw_flat(fg.flat_param, fg.master, fg.flat_grad, fg.exp_avg, fg.exp_avg_sq, *args, fg.hyper)
                continue
            # ---- data-parallel: fused reduce-scatter + AdamW + all-gather over NVLink peer memory, bucket by bucket
            nb = len(fg.buckets)
            if armed and self._side is not None:
                main = torch.cuda.current_stream()
                late = [k for k in range(nb) if not fg.launched[k]]  # buckets holding parameters that got no gradient
                if late:
                    self._side.wait_stream(main)
                    with torch.cuda.stream(self._side):
                        for k in reversed(late):
                            self._bucket_kernel(g, fg, k, 0)
                main.wait_stream(self._side)
            elif self.grad_clip:
                fg.hyper.copy_(fg.hyper_host, non_blocking=True)
                if fg.gshard is None:
                    fg.gshard = torch.empty(fg.shard, dtype=torch.float32, device=fg.device)
                fg.sq.zero_()
                for k in reversed(range(nb)):
                    self._bucket_kernel(g, fg, k, 1)           # reduce into the parked fp32 shard + Σg² partial
                C.clip_exchange(list(fg.symm_sq.buffer_ptrs), list(fg.symm_flags.buffer_ptrs), (nb + 1) * fg.world, fg.rank,
                                fg.sq, fg.epochs[nb + 1:nb + 2], float(self.grad_clip), fg.hyper, fg.norm)
                self.last_grad_norm = fg.norm
                for k in reversed(range(nb)):
                    self._bucket_kernel(g, fg, k, 2)           # clipped update + all-gather
            else:
                fg.hyper.copy_(fg.hyper_host, non_blocking=True)
                for k in reversed(range(nb)):
                    self._bucket_kernel(g, fg, k, 0)
            self._end_barrier(fg)

    @property
    def graph_capturable(self) -> bool:
        """True when ``device_step`` is pure stream work on the fused kernels (no torch.optim fallback)."""
        self._lazy_init()
        return self._flat is not None

    def snapshot(self):
        """Clone of all mutable optimizer/parameter storage (used to undo CUDA-graph warm-up steps)."""
        self._lazy_init()
        snap = {"step": self._step_count_fused, "groups": []}
        for fg in self._flat or []:
            snap["groups"].append(None if fg is None else [t.clone() for t in (fg.flat_param, fg.master, fg.exp_avg, fg.exp_avg_sq)])
        return snap

    def restore(self, snap):
        self._step_count_fused = snap["step"]
        for fg, saved in zip(self._flat or [], snap["groups"]):
            if fg is None:
                continue
            for dst, src in zip((fg.flat_param, fg.master, fg.exp_avg, fg.exp_avg_sq), saved):
                dst.copy_(src)
            fg.flat_grad.zero_()

    @torch.no_grad()
    def resync_master(self, reset_moments: bool = False) -> None:
        """Re-derive the fp32 master shard from the bf16 parameters.  MUST be called after any write to the parameters
        that did not go through :meth:`step` (checkpoint import, ``load_from_pretrained``, manual surgery): the update
        kernel reads ``master`` and rewrites the flat parameters from it, so a stale master would silently undo the
        write at the next step."""
        self._lazy_init()
        for fg in self._flat or []:
            if fg is None:
                continue
            fg.resync_master()
            if reset_moments:
                fg.exp_avg.zero_()
                fg.exp_avg_sq.zero_()
        if reset_moments:
            self._step_count_fused = 0

    # -- checkpointing ------------------------------------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, Any]:
        self._lazy_init()
        if self._fallback is not None:
            return {"kind": "torch", "state": self._fallback.state_dict(), "step": self._step_count_fused}
        shards = []
        for fg in self._flat:
            shards.append(None if fg is None else dict(master=fg.master.cpu(), exp_avg=fg.exp_avg.cpu(),
                                                       exp_avg_sq=fg.exp_avg_sq.cpu(), layout=fg.layout(), rank=fg.rank,
                                                       world=fg.world))
        return {"kind": "flat", "step": self._step_count_fused, "shards": shards,
                "lrs": [g["lr"] for g in self.param_groups]}

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        self._lazy_init()
        mine = "torch" if self._fallback is not None else "flat"
        if sd.get("kind") != mine:
            raise ValueError(f"optimizer state was saved by the '{sd.get('kind')}' backend but this run uses '{mine}' "
                             "(CPU/fp32 ↔ CUDA/bf16 switch?); load the weights only, or resume on the same backend")
        self._step_count_fused = sd.get("step", 0)
        if mine == "torch":
            self._fallback.load_state_dict(sd["state"])
            return
        if len(sd["shards"]) != len(self._flat):
            raise ValueError("optimizer state has a different number of parameter groups")
        for fg, sh in zip(self._flat, sd["shards"]):
            if fg is None or sh is None:
                continue
            if (sh.get("layout") != fg.layout() or sh.get("rank") != fg.rank or sh.get("world") != fg.world
                    or sh["master"].numel() != fg.shard):
                raise ValueError("optimizer shard layout changed (different world size or bucket size?)")
            fg.master.copy_(sh["master"])
            fg.exp_avg.copy_(sh["exp_avg"])
            fg.exp_avg_sq.copy_(sh["exp_avg_sq"])
        for g, lr in zip(self.param_groups, sd.get("lrs", [])):
            g["lr"] = lr


class FusedAdam(FusedAdamW):
    """Adam with L2 regularisation folded into the gradient (``torch.optim.Adam`` semantics)."""

    decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, **kw):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)


# ---- 8-bit state optimizers ---------------------------------------------------------------------------------------------
_BLOCK = 256


def _quantize(x: torch.Tensor, signed: bool):
    n = x.numel()
    pad = (-n) % _BLOCK
    xf = torch.nn.functional.pad(x.reshape(-1).float(), (0, pad)).view(-1, _BLOCK)
    scale = xf.abs().amax(dim=1, keepdim=True).clamp_min(1e-12)
    if signed:
        # sign + log-magnitude (1/4-octave steps, 2^-31 … 1 of the block max): Adam's update m/sqrt(v) is scale free per
        # element, so small entries need *relative* precision, which linear int8 cannot give
        lg = torch.log2((xf.abs() / scale).clamp_min(2.0 ** -40))
        mag = torch.round(127 + 4 * lg).clamp_(1, 127)
        q = torch.where(xf == 0, torch.zeros_like(mag), mag * torch.sign(xf)).to(torch.int8)
    else:
        # second moments span many orders of magnitude inside a block: store log2(v / blockmax) with 1/8-octave steps
        # (covers 2^-31 … 1; code 0 means exactly zero) so that tiny entries never collapse to 0 and blow up m/sqrt(v)
        lg = torch.log2((xf / scale).clamp_min(2.0 ** -40))
        q = torch.round(255 + 8 * lg).clamp_(1, 255)
        q = torch.where(xf <= 0, torch.zeros_like(q), q).to(torch.uint8)
    return q, scale.squeeze(1)


def _dequantize(q: torch.Tensor, scale: torch.Tensor, n: int, signed: bool, shape):
    if signed:
        qf = q.float()
        x = torch.where(qf != 0, torch.exp2((qf.abs() - 127) / 4) * torch.sign(qf), torch.zeros_like(qf)) * scale[:, None]
    else:
        qf = q.float()
        x = torch.where(qf > 0, torch.exp2((qf - 255) / 8), torch.zeros_like(qf)) * scale[:, None]
    return x.reshape(-1)[:n].view(shape)


class AdamW8bit(Optimizer):
    """AdamW whose moments are stored block-wise in 8 bits (math in fp32 per step)."""

    decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, min_8bit_size: int = 4096, **unused):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.min_8bit_size = min_8bit_size

    @staticmethod
    def _kernel_ok(p) -> bool:
        if not (p.is_cuda and p.is_contiguous() and p.grad.is_contiguous() and p.grad.dtype == p.dtype
                and p.dtype in (torch.bfloat16, torch.float32)):
            return False
        from trlx_b200 import ops

        return ops.available() and hasattr(ops.C, "adam8bit")

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for g in self.param_groups:
            b1, b2 = g["betas"]
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                grad = p.grad.float()
                if not st:
                    st["step"] = 0
                    st["q8"] = p.numel() >= self.min_8bit_size and not getattr(p, "_optim_32bit", False)
                    z = torch.zeros_like(p, dtype=torch.float32)
                    if st["q8"]:
                        st["m"], st["ms"] = _quantize(z, True)
                        st["v"], st["vs"] = _quantize(z, False)
                    else:
                        st["m"], st["v"] = z, z.clone()
                st["step"] += 1
                if st["q8"] and self._kernel_ok(p):
                    # one launch per tensor: decode, update, block maxima, re-encode (csrc/optim.cu: adam8bit_kernel)
                    from trlx_b200 import ops

                    ops.C.adam8bit(p.data, p.grad, st["m"].view(-1), st["ms"], st["v"].view(-1), st["vs"], float(g["lr"]), float(b1),
                                   float(b2), float(g["eps"]), float(g["weight_decay"]), bool(self.decoupled), int(st["step"]))
                    continue
                if st["q8"]:
                    m = _dequantize(st["m"], st["ms"], p.numel(), True, p.shape)
                    v = _dequantize(st["v"], st["vs"], p.numel(), False, p.shape)
                else:
                    m, v = st["m"], st["v"]
                w = p.data.float()
                if not self.decoupled and g["weight_decay"]:
                    grad = grad + g["weight_decay"] * w
                m = m.mul(b1).add_(grad, alpha=1 - b1)
                v = v.mul(b2).addcmul_(grad, grad, value=1 - b2)
                bc1, bc2 = 1 - b1 ** st["step"], 1 - b2 ** st["step"]
                if self.decoupled and g["weight_decay"]:
                    w.mul_(1 - g["lr"] * g["weight_decay"])
                w.addcdiv_(m / bc1, (v / bc2).sqrt_().add_(g["eps"]), value=-g["lr"])
                p.data.copy_(w)
                if st["q8"]:
                    st["m"], st["ms"] = _quantize(m, True)
                    st["v"], st["vs"] = _quantize(v, False)
                else:
                    st["m"], st["v"] = m, v
        return loss


class Adam8bit(AdamW8bit):
    decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, **kw):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
