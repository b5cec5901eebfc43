"""ZeRO-3 style parameter partitioning (``train.parallel.zero_stage: 3``).

Reference: DeepSpeed ZeRO-3 behind ``configs/accelerate/zero3.yaml:8-10`` (parameters, gradients and optimizer state all
partitioned over the data-parallel ranks; ``deepspeed.zero.GatheredParameters`` at the touch points
``trlx/models/modeling_ppo.py:527-533`` and ``trlx/models/modeling_ilql.py:221-227``).

Design here: the model is cut into *units*: every transformer block (policy stack, frozen hydra branch, T5 stacks) is one
unit, gathered right before its forward / backward and dropped right after; everything else (embeddings, final norm, LM
head, value / Q heads — modules whose weights the fused kernels read directly rather than through ``forward``) forms the
"outer" unit, gathered when a step first touches the model and dropped when its gradients are reduced or the step ends.
A unit's parameters live, between uses, only as this rank's ``1 / world`` slice of the unit's flat bf16 buffer.  On CUDA the slices are NVLink **symmetric memory**, so a gather is ``world - 1`` peer copies
(copy engines, P2P over NVSwitch) into a pooled full-size buffer — no collective library on the path; elsewhere (gloo / CPU
tests) it is ``all_gather_into_tensor``.

* forward pre-hook  → gather the unit (parameters become views of the pooled buffer), forward post-hook → release;
* backward pre-hook → gather again; when every parameter of the unit has its gradient, the full gradient is
  reduce-scattered into the rank's fp32 gradient slice (averaged over ranks) and both full buffers are dropped;
* the optimizer only ever sees the slices (``FusedAdamW(..., local_only=True)`` — one fused kernel per step, no collective:
  the gradients it reads are already reduced), so optimizer state is partitioned the same way;
* ``summon_full_params()`` gathers everything for checkpoint export / generation engines that need stable addresses.

Per-rank persistent memory for P parameters: ``2P / world`` (bf16 slices) + ``4P / world`` (gradient slices) +
``12P / world`` (fp32 master + moments) — plus two pooled unit-sized buffers.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from trlx_b200.utils import logging

logger = logging.get_logger(__name__)

_ALIGN = 8


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _Unit:
    def __init__(self, name: str, module: Optional[nn.Module], params: List[nn.Parameter], world: int, rank: int):
        self.name, self.module, self.params = name, module, params
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += _round_up(p.numel(), _ALIGN)
        self.numel = _round_up(max(off, _ALIGN), _ALIGN * world)
        self.shard_numel = self.numel // world
        self.lo = self.shard_numel * rank
        self.shapes = [p.shape for p in params]
        self.numels = [p.numel() for p in params]
        self.dtype = params[0].dtype
        self.trainable = any(p.requires_grad for p in params)
        self.full: Optional[torch.Tensor] = None       # gathered parameters (pooled)
        self.full_grad: Optional[torch.Tensor] = None  # gradients of the gathered parameters while the unit's backward runs
        self.users = 0                                  # nested gathers (forward inside summon_full_params, …)
        self.pending = 0
        self.shard: Optional[nn.Parameter] = None
        self.symm = None


class Zero3ParamSharder:
    """Partition the parameters of ``model`` over ``group`` and gather them per unit on demand (module docstring)."""

    def __init__(self, model: nn.Module, units: List[nn.Module], group=None, use_symmetric: Optional[bool] = None):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = next(model.parameters()).device
        self.units: List[_Unit] = []
        seen = set()

        def add(name, module, ps):
            # trainable and frozen tensors never share a flat slice: the optimizer updates whole slices (weight decay would
            # otherwise move frozen entries)
            for flag, tag in ((True, ""), (False, ":frozen")):
                part = [p for p in ps if p.requires_grad == flag]
                if part:
                    self.units.append(_Unit(name + tag, module, part, self.world, self.rank))

        for i, m in enumerate(units):
            ps = [p for p in m.parameters() if id(p) not in seen]
            seen.update(id(p) for p in ps)
            add(f"unit{i}:{type(m).__name__}", m, ps)
        rest = [p for p in model.parameters() if id(p) not in seen]
        add("outer", None, rest)
        self.outer_units = [u for u in self.units if u.module is None]
        if use_symmetric is None:
            use_symmetric = self.device.type == "cuda" and self.world > 1
        self.symmetric = False
        if use_symmetric:
            try:
                self._alloc_symmetric()
                self.symmetric = True
            except Exception as err:  # pragma: no cover - depends on the platform
                logger.warning(f"ZeRO-3: symmetric memory unavailable ({err}); gathers use all_gather_into_tensor")
        for u in self.units:
            self._partition(u)
        self._pool: Dict[int, List[torch.Tensor]] = {}
        self._grad_armed = True
        self._install_hooks()
        total = sum(u.numel for u in self.units)
        logger.info(f"ZeRO-3: {len(self.units)} units, {total / 1e6:.1f}M parameters partitioned over {self.world} ranks "
                    f"({'NVLink peer copies' if self.symmetric else 'collective'} gathers)")

    # ---- storage ----------------------------------------------------------------------------------------------------------
    def _alloc_symmetric(self):
        import torch.distributed._symmetric_memory as symm

        name = self.group.group_name if self.group is not None else dist.group.WORLD.group_name
        for u in self.units:
            buf = symm.empty(u.shard_numel, dtype=u.dtype, device=self.device)
            u.symm = (buf, symm.rendezvous(buf, name))

    def _partition(self, u: _Unit):
        flat = torch.zeros(u.numel, dtype=u.dtype, device=self.device)
        for p, o in zip(u.params, u.offsets):
            flat[o:o + p.numel()].copy_(p.data.reshape(-1))
        mine = flat[u.lo:u.lo + u.shard_numel]
        u.shard = nn.Parameter(mine.clone(), requires_grad=u.trainable)
        u.shard._zero3_unit = u.name
        for p in u.params:
            p.data = torch.empty(0, dtype=u.dtype, device=self.device)
            p._zero3 = True

    def shard_parameters(self) -> List[nn.Parameter]:
        """What the optimizer updates: one flat slice per trainable unit."""
        return [u.shard for u in self.units if u.trainable]

    def _take(self, numel: int, dtype) -> torch.Tensor:
        key = (numel, dtype)
        pool = self._pool.setdefault(key, [])
        return pool.pop() if pool else torch.empty(numel, dtype=dtype, device=self.device)

    def _give(self, t: torch.Tensor):
        self._pool.setdefault((t.numel(), t.dtype), []).append(t)

    # ---- gather / release -------------------------------------------------------------------------------------------------
    def _gather(self, u: _Unit):
        u.users += 1
        if u.full is not None:
            return
        full = self._take(u.numel, u.dtype)
        if self.world == 1:
            full.copy_(u.shard.data)
        elif self.symmetric:
            hdl = u.symm[1]
            u.symm[0].copy_(u.shard.data)  # the slice may live in the optimizer's flat buffer: stage it in the peer-visible window
            hdl.barrier()  # every rank's slice holds this step's values before anyone pulls it
            for r in range(self.world):
                src = u.symm[0] if r == self.rank else hdl.get_buffer(r, (u.shard_numel,), u.dtype)
                full[r * u.shard_numel:(r + 1) * u.shard_numel].copy_(src, non_blocking=True)
            hdl.barrier()  # nobody overwrites a slice (optimizer step) while a peer still reads it
        else:
            dist.all_gather_into_tensor(full, u.shard.data.contiguous(), group=self.group)
        u.full = full
        for p, o, n, shape in zip(u.params, u.offsets, u.numels, u.shapes):
            p.data = full[o:o + n].view(shape)

    def _release(self, u: _Unit, force: bool = False):
        u.users = 0 if force else max(u.users - 1, 0)
        if u.users or u.full is None:
            return
        for p in u.params:
            p.data = torch.empty(0, dtype=u.dtype, device=self.device)
        self._give(u.full)
        u.full = None

    # ---- gradients --------------------------------------------------------------------------------------------------------
    def _on_grad(self, u: _Unit):
        u.pending -= 1
        if u.pending > 0:
            return
        self._reduce_grads(u)

    def _reduce_grads(self, u: _Unit):
        full_grad = self._take(u.numel, torch.float32)
        full_grad.zero_()
        for p, o, n in zip(u.params, u.offsets, u.numels):
            if p.grad is not None:
                full_grad[o:o + n].copy_(p.grad.reshape(-1))
                p.grad = None
        if self.world > 1:
            out = torch.empty(u.shard_numel, dtype=torch.float32, device=self.device)
            try:
                dist.reduce_scatter_tensor(out, full_grad, op=dist.ReduceOp.SUM, group=self.group)
            except (RuntimeError, NotImplementedError):  # gloo has no reduce-scatter
                dist.all_reduce(full_grad, group=self.group)
                out.copy_(full_grad[u.lo:u.lo + u.shard_numel])
            out.div_(self.world)
        else:
            out = full_grad[: u.shard_numel].clone()
        self._give(full_grad)
        if u.shard.grad is None:
            u.shard.grad = out.to(u.shard.dtype) if u.shard.dtype != torch.float32 else out
        else:
            u.shard.grad.add_(out.to(u.shard.grad.dtype))
        u.pending = sum(1 for p in u.params if p.requires_grad)
        self._release(u, force=True)

    # ---- hooks ------------------------------------------------------------------------------------------------------------
    def _install_hooks(self):
        for u in self.units:
            u.pending = sum(1 for p in u.params if p.requires_grad)
            for p in u.params:
                if p.requires_grad:
                    p.register_post_accumulate_grad_hook(lambda _p, u=u: self._on_grad(u))
            if u.module is None:
                continue
            u.module.register_forward_pre_hook(lambda m, a, u=u: self._gather(u))
            u.module.register_forward_hook(lambda m, a, out, u=u: self._release(u))
            # the backward of a block needs its weights again (dX = dY . W), trainable or not
            u.module.register_full_backward_pre_hook(lambda m, g, u=u: self._gather(u))
            if not u.trainable:
                u.module.register_full_backward_hook(lambda m, gi, go, u=u: self._release(u))
        if self.outer_units:
            # embeddings / norms / heads: their weights are also read directly by fused kernels (`lm_head.weight` in the
            # LM-head kernel, the value / Q MLPs), so they are gathered when a step first enters the model and stay until
            # their gradients are reduced or the trainer ends the step / phase (`finish_step`, `release_all`)
            def enter(m, a):
                self.ensure_outer()

            self.model.register_forward_pre_hook(enter)
            for m in self.model.modules():
                if m is not self.model and hasattr(m, "transformer") and hasattr(m, "lm_head"):
                    m.register_forward_pre_hook(enter)

    def ensure_outer(self):
        for u in self.outer_units:
            if u.full is None:
                self._gather(u)

    def reduce_pending(self):
        """Before ``optimizer.step()``: reduce-scatter whatever gradients are still sitting on gathered parameters (units in
        which some trainable tensor received no gradient never hit their countdown)."""
        for u in self.units:
            if u.trainable and any(p.grad is not None for p in u.params):
                self._reduce_grads(u)

    def release_all(self):
        for u in self.units:
            self._release(u, force=True)

    @contextlib.contextmanager
    def summon_full_params(self, writeback: bool = False):
        """All parameters materialised (checkpoint export, ``state_dict``, engines).  ``writeback`` copies this rank's slice
        of any in-place modification back into its partition on exit."""
        for u in self.units:
            self._gather(u)
        try:
            yield
        finally:
            for u in self.units:
                if writeback and u.full is not None:
                    with torch.no_grad():
                        u.shard.data.copy_(u.full[u.lo:u.lo + u.shard_numel])
                self._release(u, force=True)

    def finish_step(self):
        """After ``optimizer.step()``: every gathered copy is stale now — drop them all."""
        for u in self.units:
            self._release(u, force=True)
            u.pending = sum(1 for p in u.params if p.requires_grad)


def default_units(model: nn.Module) -> List[nn.Module]:
    """Every transformer block (policy stack, frozen hydra branch, T5 encoder / decoder stacks) is one unit."""
    units: List[nn.Module] = []
    for name, m in model.named_modules():
        leaf = name.rsplit(".", 1)[-1]
        if isinstance(m, nn.ModuleList) and leaf in ("h", "block", "decoder_blocks", "layers"):
            units.extend(list(m))
    return units
