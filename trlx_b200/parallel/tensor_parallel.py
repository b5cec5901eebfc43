"""Megatron-style tensor (and sequence) parallelism for the generic transformer.

Reference counterpart: Apex ``ColumnParallelLinear`` / ``RowParallelLinear`` + NeMo ``sequence_parallel`` as used by
``trlx/models/modeling_nemo_ppo.py:93-120,820-870`` (external, not importable in the reference snapshot).

Sharding of one block over ``tp`` ranks (heads and FFN columns are split; everything is ``[out, in]``):

=====================  ==========================  =======================================
module                 kind                        local shape
=====================  ==========================  =======================================
``attn.qkv``           column-parallel (by head)   ``[(nq + 2·nkv)/tp · d, H]``
``attn.out``           row-parallel                ``[H, nq·d/tp]``   (bias on rank 0)
``mlp.up``             column-parallel             ``[F/tp (·2 gated), H]``
``mlp.down``           row-parallel                ``[H, F/tp]``      (bias on rank 0)
=====================  ==========================  =======================================

Communication per block (forward): without sequence parallelism one all-reduce after each row-parallel GEMM; with it
the activations between blocks are sharded along the sequence, the column-parallel GEMM is preceded by an
all-gather and the row-parallel GEMM followed by a reduce-scatter.  On CUDA those pairs run on the fused kernels of
``parallel/fused_tp.py``, forward and backward: *all-gather→GEMM* — the peers' shards arrive in a local
gathered buffer by NVLink copies on a side stream while ONE tcgen05 GEMM is already running, its TMA producer gated per shard
by device-side ready flags; *GEMM→reduce-scatter* — the partial product is written locally into a symmetric buffer and the
owner of a row block sums the ranks' copies inside the NVSwitch with ``multimem.ld_reduce`` (two column windows, the
reduction of one under the GEMM of the next).  Elsewhere (CPU / gloo tests, shapes the kernels do not cover) the same maths
runs on ``torch.distributed`` collectives through the autograd functions below.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200.models.modeling_base import base_lm
from trlx_b200.nn.transformer import AttnContext, Attention, Block, MLP, apply_rotary
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


# ---- autograd-aware collectives (Megatron's f / g and their sequence-parallel variants) ---------------------------------
class _CopyToTP(torch.autograd.Function):
    """identity forward, all-reduce backward (input of a column-parallel layer)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, group=ctx.group)
        return g, None


class _ReduceFromTP(torch.autograd.Function):
    """all-reduce forward, identity backward (output of a row-parallel layer)."""

    @staticmethod
    def forward(ctx, x, group):
        x = x.contiguous()
        dist.all_reduce(x, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


class _GatherSeq(torch.autograd.Function):
    """all-gather along the sequence (dim 1) forward, reduce-scatter backward."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=1)

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        chunks = [c.contiguous() for c in g.chunk(world, dim=1)]
        out = torch.empty_like(chunks[0])
        dist.reduce_scatter(out, chunks, group=ctx.group)
        return out, None


class _GatherSeqReplicated(torch.autograd.Function):
    """all-gather along the sequence forward; backward keeps this rank's slice.  For consumers whose computation is
    *replicated* across the TP group (final norm → heads → loss): every rank already holds the full, identical gradient,
    so reducing would multiply it by the group size (Megatron's ``tensor_parallel_output_grad=False``)."""

    @staticmethod
    def forward(ctx, x, group, rank):
        ctx.group, ctx.rank = group, rank
        world = dist.get_world_size(group)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=1)

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        return g.chunk(world, dim=1)[ctx.rank].contiguous(), None, None


class _ScatterSeq(torch.autograd.Function):
    """reduce-scatter along the sequence forward, all-gather backward."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        chunks = [c.contiguous() for c in x.chunk(world, dim=1)]
        out = torch.empty_like(chunks[0])
        dist.reduce_scatter(out, chunks, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g.contiguous(), group=ctx.group)
        return torch.cat(parts, dim=1), None


class _SplitSeqReplicated(torch.autograd.Function):
    """keep this rank's sequence slice forward, all-gather the gradient slices backward.  For producers whose computation is
    *replicated* across the group (the embedding lookup): every rank then holds the complete gradient of the full-length
    activation, so the replicated parameters behind it get full — not partial — gradients (mirror of ``_GatherSeqReplicated``)."""

    @staticmethod
    def forward(ctx, x, group, rank):
        ctx.group = group
        world = dist.get_world_size(group)
        return x.chunk(world, dim=1)[rank].contiguous()

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g.contiguous(), group=ctx.group)
        return torch.cat(parts, dim=1), None, None


def split_sequence(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Keep this rank's slice of the sequence dimension (entering a sequence-parallel region)."""
    return x.chunk(world, dim=1)[rank].contiguous()


class TPContext:
    def __init__(self, group, rank: int, size: int, sequence_parallel: bool):
        self.group, self.rank, self.size, self.sequence_parallel = group, rank, size, sequence_parallel
        self.fused = None  # set by enable_fused_kernels()
        self.sp_active = True  # cleared for forwards whose sequence cannot be sharded (see _install_sequence_parallel)

    @property
    def sp(self) -> bool:
        """Sequence parallelism applies to the forward that is running."""
        return self.sequence_parallel and self.sp_active

    def enter_column(self, x):
        """Activation entering a column-parallel GEMM."""
        if self.sp:
            return _GatherSeq.apply(x, self.group)
        return _CopyToTP.apply(x, self.group)

    def exit_row(self, y):
        """Partial sums leaving a row-parallel GEMM."""
        if self.sp:
            return _ScatterSeq.apply(y, self.group)
        return _ReduceFromTP.apply(y, self.group)

    # ---- GEMM + collective pairs (fused into single NVLink kernels on CUDA with sequence parallelism) ----------------
    def _fusable(self, x: torch.Tensor, linear: nn.Linear, rows_per_rank: int) -> bool:
        if self.fused is None or not (x.is_cuda and x.dtype == torch.bfloat16 and type(linear) is nn.Linear):
            return False
        from trlx_b200.parallel.fused_tp import FusedTP

        return FusedTP.usable(rows_per_rank, linear.in_features, linear.out_features)

    def column_linear(self, linear: nn.Linear, x: torch.Tensor) -> torch.Tensor:
        """``all_gather_seq(x) · Wᵀ + b`` (SP) / ``x · Wᵀ + b`` with an all-reduced input gradient (no SP)."""
        if self.sp and self._fusable(x, linear, x.shape[0] * x.shape[1]):
            from trlx_b200.parallel.fused_tp import column_linear

            return column_linear(self.fused, linear, x)
        return linear(self.enter_column(x))

    def row_linear(self, linear: nn.Linear, x: torch.Tensor) -> torch.Tensor:
        """``reduce_scatter_seq(x · Wᵀ) + b`` (SP) / ``all_reduce(x · Wᵀ) + b`` (no SP); ``b`` lives on rank 0."""
        if self.sp and x.shape[1] % self.size == 0 and \
                self._fusable(x, linear, x.shape[0] * x.shape[1] // self.size):
            from trlx_b200.parallel.fused_tp import row_linear

            return row_linear(self.fused, linear, x)
        return self.exit_row(linear(x))

    def enable_fused_kernels(self, device: torch.device) -> bool:
        """Switch the SP GEMM↔collective pairs to the fused NVLink kernels (needs symmetric memory)."""
        from trlx_b200 import ops

        if not (self.sequence_parallel and device.type == "cuda" and ops.available()):
            return False
        try:
            from trlx_b200.parallel.fused_tp import FusedTP

            self.fused = FusedTP(self.group, self.rank, self.size, device)
            return True
        except Exception as err:  # pragma: no cover
            logger.warning(f"fused TP kernels unavailable: {err}")
            return False


def _shard_rows(linear: nn.Linear, row_index: torch.Tensor) -> nn.Linear:
    new = nn.Linear(linear.in_features, len(row_index), bias=linear.bias is not None, device=linear.weight.device,
                    dtype=linear.weight.dtype)
    with torch.no_grad():
        new.weight.copy_(linear.weight[row_index])
        if linear.bias is not None:
            new.bias.copy_(linear.bias[row_index])
    new.weight.requires_grad_(linear.weight.requires_grad)
    if new.bias is not None:
        new.bias.requires_grad_(linear.bias.requires_grad)
    return new


def _shard_cols(linear: nn.Linear, col_index: torch.Tensor, keep_bias: bool) -> nn.Linear:
    new = nn.Linear(len(col_index), linear.out_features, bias=linear.bias is not None, device=linear.weight.device,
                    dtype=linear.weight.dtype)
    with torch.no_grad():
        new.weight.copy_(linear.weight[:, col_index])
        if linear.bias is not None:
            new.bias.copy_(linear.bias if keep_bias else torch.zeros_like(linear.bias))
    new.weight.requires_grad_(linear.weight.requires_grad)
    if new.bias is not None:
        new.bias.requires_grad_(linear.bias.requires_grad and keep_bias)
    return new


def qkv_row_index(spec, rank: int, size: int) -> torch.Tensor:
    """Rows of the fused ``[Q|K|V]`` weight owned by ``rank`` (contiguous head ranges of each part)."""
    d = spec.head_dim
    nq, nkv = spec.num_heads // size, spec.num_kv_heads // size
    q = torch.arange(rank * nq * d, (rank + 1) * nq * d)
    k = spec.q_size + torch.arange(rank * nkv * d, (rank + 1) * nkv * d)
    v = spec.q_size + spec.kv_size + torch.arange(rank * nkv * d, (rank + 1) * nkv * d)
    return torch.cat([q, k, v])


def up_row_index(spec, rank: int, size: int) -> torch.Tensor:
    f = spec.ffn_size // size
    idx = torch.arange(rank * f, (rank + 1) * f)
    return torch.cat([idx, spec.ffn_size + idx]) if spec.gated_mlp else idx


class TPAttention(nn.Module):
    """Attention over this rank's heads; ``qkv`` column-parallel, ``out`` row-parallel."""

    def __init__(self, attn: Attention, tp: TPContext):
        super().__init__()
        spec = attn.spec
        if spec.num_heads % tp.size or spec.num_kv_heads % tp.size:
            raise ValueError(f"heads ({spec.num_heads}/{spec.num_kv_heads}) must be divisible by tp={tp.size}")
        self.spec, self.tp, self.scale, self.is_local = spec, tp, attn.scale, attn.is_local
        self.nq, self.nkv = spec.num_heads // tp.size, spec.num_kv_heads // tp.size
        self.qkv = _shard_rows(attn.qkv, qkv_row_index(spec, tp.rank, tp.size).to(attn.qkv.weight.device))
        cols = torch.arange(tp.rank * self.nq * spec.head_dim, (tp.rank + 1) * self.nq * spec.head_dim)
        self.out = _shard_cols(attn.out, cols.to(attn.out.weight.device), keep_bias=tp.rank == 0)

    def forward(self, x, ctx: AttnContext, past=None, use_cache=False):
        s, tp = self.spec, self.tp
        qkv = tp.column_linear(self.qkv, x)
        B, T, _ = qkv.shape
        d = s.head_dim
        q, k, v = qkv.split([self.nq * d, self.nkv * d, self.nkv * d], dim=-1)
        q = q.view(B, T, self.nq, d).transpose(1, 2)
        k = k.view(B, T, self.nkv, d).transpose(1, 2)
        v = v.view(B, T, self.nkv, d).transpose(1, 2)
        if s.pos == "rotary":
            q = apply_rotary(q, ctx.cos, ctx.sin, s.rotary_dim, s.rotary_interleaved)
            k = apply_rotary(k, ctx.cos, ctx.sin, s.rotary_dim, s.rotary_interleaved)
        if past is not None:
            k, v = torch.cat([past[0], k], 2), torch.cat([past[1], v], 2)
        present = (k, v) if use_cache else None
        if self.nkv != self.nq:
            k, v = k.repeat_interleave(self.nq // self.nkv, 1), v.repeat_interleave(self.nq // self.nkv, 1)
        bias = ctx.local_bias if self.is_local else ctx.bias
        if bias is not None and bias.shape[1] > 1:  # per-head bias (alibi): keep this rank's heads
            bias = bias[:, tp.rank * self.nq:(tp.rank + 1) * self.nq]
        if bias is None:
            o = F.scaled_dot_product_attention(q, k, v, is_causal=(T > 1), scale=self.scale)
        else:
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype), scale=self.scale)
        o = o.transpose(1, 2).reshape(B, T, self.nq * d)
        return tp.row_linear(self.out, o), present


class TPMLP(nn.Module):
    def __init__(self, mlp: MLP, spec, tp: TPContext):
        super().__init__()
        if spec.ffn_size % tp.size:
            raise ValueError(f"ffn size {spec.ffn_size} must be divisible by tp={tp.size}")
        self.tp, self.gated, self.act = tp, mlp.gated, mlp.act
        self.up = _shard_rows(mlp.up, up_row_index(spec, tp.rank, tp.size).to(mlp.up.weight.device))
        f = spec.ffn_size // tp.size
        cols = torch.arange(tp.rank * f, (tp.rank + 1) * f)
        self.down = _shard_cols(mlp.down, cols.to(mlp.down.weight.device), keep_bias=tp.rank == 0)

    def forward(self, x):
        tp = self.tp
        h = tp.column_linear(self.up, x)
        if self.gated:
            g, u = h.chunk(2, dim=-1)
            h = self.act(g) * u
        else:
            h = self.act(h)
        return tp.row_linear(self.down, h)


def shard_block(block: Block, tp: TPContext) -> None:
    block.attn = TPAttention(block.attn, tp)
    block.mlp = TPMLP(block.mlp, block.spec, tp)


def apply_tensor_parallel(model, group, rank: int, size: int, sequence_parallel: bool = False) -> TPContext:
    """Shard every transformer block of ``model`` (wrapper or bare LM, incl. hydra / value branches) in place.

    All ranks must hold identical full weights when this is called (same seed / same checkpoint)."""
    tp = TPContext(group, rank, size, sequence_parallel)
    lm = base_lm(getattr(model, "base_model", model))
    if getattr(lm.config, "post_norm", False) or not getattr(lm.config, "plain_tail", True):
        raise NotImplementedError("tensor parallelism does not cover the OPT-350m layout (post-LN blocks, project_in / project_out)")
    blocks = list(lm.transformer.h)
    for extra in ("frozen_head", "v_head"):
        branch = getattr(model, extra, None)
        if branch is not None and hasattr(branch, "decoder_blocks"):
            blocks += list(branch.decoder_blocks)
    for b in blocks:
        shard_block(b, tp)
    lm._tp_context = tp  # pipeline stages size their boundary activations by the sequence-parallel layout
    if sequence_parallel:
        _install_sequence_parallel(lm, tp)
        for extra in ("frozen_head", "v_head"):
            branch = getattr(model, extra, None)
            if branch is not None and hasattr(branch, "final_norm"):
                _gather_before(branch.final_norm, tp)
    model.tp_context = tp
    if group is not None and size > 1:  # sampled tokens are taken from the group's first rank (generation.sync_tokens)
        lm._sample_sync = getattr(lm, "_sample_sync", []) + [(group, dist.get_process_group_ranks(group)[0])]
    tp.enable_fused_kernels(next(lm.parameters()).device)
    return tp


def _gather_before(norm: nn.Module, tp: "TPContext") -> None:
    orig = norm.forward
    norm.forward = lambda x: orig(_GatherSeqReplicated.apply(x, tp.group, tp.rank) if tp.sp else x)


def _install_sequence_parallel(lm, tp: TPContext) -> None:
    """Shard the residual stream along the sequence between blocks: scatter after the embedding, gather before the
    final norm's consumers (heads are evaluated on the full sequence, like the reference which gathers before its heads,
    ``modeling_nemo_ppo.py:162-163``)."""
    trunk = lm.transformer
    orig_embed = trunk.embed

    def embed(input_ids, position_ids, inputs_embeds=None):
        x = orig_embed(input_ids, position_ids, inputs_embeds)
        if not tp.sp:
            return x
        return _SplitSeqReplicated.apply(x, tp.group, tp.rank) if x.requires_grad else split_sequence(x, tp.rank, tp.size)

    trunk.embed = embed
    _gather_before(trunk.ln_f, tp)

    # A sequence can only be sharded when its length is a multiple of the group size, and an incremental decode step (KV cache,
    # one new token) cannot be sharded at all.  Such forwards run with replicated activations (plain tensor parallelism) — the
    # reference switches sequence parallelism off around inference for the same reason (``modeling_nemo_ppo.py:820-870``).
    # With autograd the replicated block norms would then hold full instead of partial gradients while the optimizer step still
    # sums them over the group (``allreduce_sequence_parallel_grads``): their gradients are scaled by 1 / tp for such forwards.
    from trlx_b200.nn.transformer import Norm

    for name, module in lm.named_modules():
        if isinstance(module, Norm) and ".h." in f".{name}.":
            module._grad_scale = lambda: (1.0 / tp.size) if (not tp.sp_active and torch.is_grad_enabled()) else None
    orig_forward = lm.forward

    def forward(*args, **kw):
        ids = kw.get("input_ids", args[0] if args else None)
        hidden_in = kw.get("hidden_in")
        cached = bool(kw.get("past_key_values")) or bool(kw.get("use_cache"))
        if kw.get("inputs_embeds") is not None or (ids is None and hidden_in is None):
            return orig_forward(*args, **kw)
        if hidden_in is not None:
            mask, pos = kw.get("attention_mask"), kw.get("position_ids")
            full = mask.shape[1] if mask is not None else (pos.shape[1] if pos is not None else hidden_in.shape[1])
            if hidden_in.shape[1] != full:  # already a sequence shard (produced by a sharded forward)
                tp.sp_active = True
                return orig_forward(*args, **kw)
            length = full
        else:
            length = ids.shape[1]
        # The decision is sticky until the next LM forward: the frozen reference branch (``ModelBranch.run_blocks``) and the heads
        # that consume this forward's activations run right after it, outside this wrapper, and must see the same layout.
        tp.sp_active = (not cached) and length % tp.size == 0
        if tp.sp_active and hidden_in is not None:  # a replicated (cached) activation enters a sharded region: keep our slice
            kw = dict(kw, hidden_in=split_sequence(hidden_in, tp.rank, tp.size))
        return orig_forward(*args, **kw)  # sp_active False: replicated activations, block norms scale their gradients by 1 / tp

    lm.forward = forward


# ---- checkpoint resharding (HF → TP shards and back) ------------------------------------------------------------------------
def shard_state_dict(spec, canonical_sd, rank: int, size: int):
    """Slice a full canonical state dict into the tensors rank ``rank`` of ``size`` holds (port of the logic of the
    reference's ``examples/llama_nemo/convert_llama_to_nemo.py:55-105`` to this framework's layout)."""
    out = {}
    qi, ui = qkv_row_index(spec, rank, size), up_row_index(spec, rank, size)
    d = spec.head_dim
    oc = torch.arange(rank * (spec.num_heads // size) * d, (rank + 1) * (spec.num_heads // size) * d)
    f = spec.ffn_size // size
    dc = torch.arange(rank * f, (rank + 1) * f)
    for k, v in canonical_sd.items():
        if k.endswith("attn.qkv.weight") or k.endswith("attn.qkv.bias"):
            out[k] = v[qi].clone()
        elif k.endswith("mlp.up.weight") or k.endswith("mlp.up.bias"):
            out[k] = v[ui].clone()
        elif k.endswith("attn.out.weight"):
            out[k] = v[:, oc].clone()
        elif k.endswith("mlp.down.weight"):
            out[k] = v[:, dc].clone()
        elif k.endswith("attn.out.bias") or k.endswith("mlp.down.bias"):
            out[k] = v.clone() if rank == 0 else torch.zeros_like(v)
        else:
            out[k] = v.clone()
    return out


def unshard_state_dicts(spec, shards):
    """Inverse of :func:`shard_state_dict` for a list of per-rank state dicts."""
    size = len(shards)
    out = {}
    for k in shards[0]:
        parts = [s[k] for s in shards]
        if k.endswith("attn.qkv.weight") or k.endswith("attn.qkv.bias"):
            d = spec.head_dim
            nq, nkv = spec.num_heads // size * d, spec.num_kv_heads // size * d
            q = torch.cat([p[:nq] for p in parts]); kk = torch.cat([p[nq:nq + nkv] for p in parts]); vv = torch.cat([p[nq + nkv:] for p in parts])
            out[k] = torch.cat([q, kk, vv])
        elif k.endswith("mlp.up.weight") or k.endswith("mlp.up.bias"):
            if spec.gated_mlp:
                f = spec.ffn_size // size
                out[k] = torch.cat([torch.cat([p[:f] for p in parts]), torch.cat([p[f:] for p in parts])])
            else:
                out[k] = torch.cat(parts)
        elif k.endswith("attn.out.weight") or k.endswith("mlp.down.weight"):
            out[k] = torch.cat(parts, dim=1)
        elif k.endswith("attn.out.bias") or k.endswith("mlp.down.bias"):
            out[k] = sum(parts)
        else:
            out[k] = parts[0]
    return out


def sequence_parallel_grad_params(model):
    """Parameters that are replicated across the TP group but see sequence-sharded activations (the norms inside the
    blocks): with sequence parallelism each rank holds a partial gradient for them, which has to be summed over the TP
    group before the optimizer step (reference: ``modeling_nemo_ppo.py:627-645``)."""
    from trlx_b200.nn.transformer import Norm

    out = []
    for name, module in model.named_modules():
        if isinstance(module, Norm) and (".h." in name or "decoder_blocks" in name):
            out += [p for p in module.parameters() if p.requires_grad]
    return out


def allreduce_sequence_parallel_grads(model, group) -> None:
    params = [p for p in sequence_parallel_grad_params(model) if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1).float() for p in params])
    dist.all_reduce(flat, group=group)
    off = 0
    for p in params:
        p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
        off += p.numel()
