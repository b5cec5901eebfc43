"""Pipeline parallelism for the generic decoder (:class:`trlx_b200.nn.transformer.CausalLM`).

Reference counterpart: the NeMo/Megatron path (``trlx/models/modeling_nemo_ppo.py:560-700`` drives Apex's
``forward_backward_pipelining_without_interleaving``; configs ``configs/nemo_configs/*.yaml`` carry
``pipeline_model_parallel_size``).  That path cannot be imported in the reference snapshot; this is a from-scratch design
around the one-process-per-GPU runtime:

* :func:`apply_pipeline_parallel` keeps a contiguous slice of the blocks on each stage (embeddings on the first, final
  norm + LM head on the last; a tied embedding lives on both and its gradient is summed across the two before the
  optimizer step) and installs a :class:`PipelineStage` on the LM.  The wrappers (value head, ILQL heads) and the
  trainers' ``loss`` functions are unchanged.
* **Inference / scoring** (no schedule active): a forward call relays the activation stage to stage with point-to-point
  NCCL transfers over NVLink and the last stage broadcasts ``(final hidden, logits)`` to its pipeline group, so every
  rank sees complete outputs (generation and ``make_experience`` run as-is).  With autograd enabled the relay is
  differentiable: the send records a node whose backward *receives* the activation gradient from the next stage, so a
  plain ``loss.backward()`` on every stage trains correctly (one micro-batch in flight).
* **Training** (:func:`run_1f1b`): the non-interleaved one-forward-one-backward schedule over the micro-batches of an
  optimizer step.  A stage runs the trainer's own ``loss(microbatch)``; on every stage but the last the LM raises
  :class:`StageBoundary` as soon as its blocks are done, which hands the boundary activation to the schedule.  Steady
  state uses grouped send/recv pairs (``batch_isend_irecv``) so that neighbouring stages, which are sending to each other
  at the same moment, cannot deadlock.  The shape of the activation to receive is found by *probing* the loss function
  (it is run up to the LM call with the stage in probe mode), so no shape handshake is exchanged.
"""
from __future__ import annotations

from collections import deque
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


class StageBoundary(Exception):
    """Raised by a non-final stage's LM forward while a schedule is active (control flow, not an error)."""


class _Hole(nn.Module):
    """Stands in for a block owned by another stage (keeps ``transformer.h`` indices stable)."""

    def forward(self, *a, **k):  # pragma: no cover - never called
        raise RuntimeError("this block lives on another pipeline stage")


def partition_layers(num_layers: int, stages: int) -> List[Tuple[int, int]]:
    """Balanced contiguous split; earlier stages take the remainder (the last stage also carries the LM head)."""
    base, rem = divmod(num_layers, stages)
    out, lo = [], 0
    for s in range(stages):
        n = base + (1 if s < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


# ---- differentiable relay ops (sequential mode) ---------------------------------------------------------------------------
class _RecvFromPrev(torch.autograd.Function):
    """forward: receive the activation from the previous stage; backward: send its gradient back."""

    @staticmethod
    def forward(ctx, anchor: torch.Tensor, stage: "PipelineStage", shape, dtype):
        ctx.stage = stage
        buf = torch.empty(shape, dtype=dtype, device=anchor.device)
        dist.recv(buf, src=stage.ring_prev, group=stage.group)
        return buf

    @staticmethod
    def backward(ctx, grad):
        dist.send(grad.contiguous(), dst=ctx.stage.ring_prev, group=ctx.stage.group)
        return None, None, None, None


class _SendToNext(torch.autograd.Function):
    """forward: send the activation on, return a zero scalar tied to the graph; backward: ignore the incoming (zero)
    gradient, *receive* the real activation gradient from the next stage and return it."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, stage: "PipelineStage"):
        ctx.stage, ctx.shape, ctx.dtype = stage, x.shape, x.dtype
        dist.send(x.contiguous(), dst=stage.ring_next, group=stage.group)
        return x.new_zeros(())

    @staticmethod
    def backward(ctx, _):
        g = torch.empty(ctx.shape, dtype=ctx.dtype, device=_.device)
        dist.recv(g, src=ctx.stage.ring_next, group=ctx.stage.group)
        return g, None


class PipelineStage:
    """Per-LM pipeline state + the stage-local forward."""

    def __init__(self, lm, group, rank: int, size: int, virtual: int = 1):
        self.lm, self.group, self.rank, self.size = lm, group, rank, size
        self.first, self.last = rank == 0, rank == size - 1
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(size))
        self.global_ranks = ranks
        self.prev_global = ranks[rank - 1] if rank > 0 else None
        self.next_global = ranks[rank + 1] if rank < size - 1 else None
        # with virtual stages the activation travels round the ring `virtual` times: the last rank hands chunk c to chunk c + 1
        # of the first rank
        self.ring_prev, self.ring_next = ranks[(rank - 1) % size], ranks[(rank + 1) % size]
        self.last_global = ranks[-1]
        L = len(lm.transformer.h)
        self.virtual = max(1, min(int(virtual or 1), L // size)) if size > 1 else 1
        parts = partition_layers(L, size * self.virtual)
        # virtual stage s = c * size + rank owns `parts[s]`: every rank holds `virtual` model chunks (Megatron's interleaved layout)
        self.chunks = [parts[c * size + rank] for c in range(self.virtual)]
        self.lo, self.hi = self.chunks[0][0], self.chunks[-1][1]
        self.chunk = 0  # the chunk a scheduled forward runs
        # schedule hand-off
        self.mode = "relay"  # "relay" | "schedule" | "probe"
        self.input_tensor: Optional[torch.Tensor] = None
        self.output_tensor: Optional[torch.Tensor] = None
        self.probe_shape: Optional[Tuple[int, ...]] = None

    # -- forward ----------------------------------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                use_cache=False, output_hidden_states=False, labels=None, compute_logits: bool = True,
                hidden_in=None, start_layer: int = 0, stop_layer: Optional[int] = None, **_ignored):
        from trlx_b200.nn.transformer import CausalLMOutput, build_attn_context, project

        if hidden_in is not None or start_layer or stop_layer is not None:
            raise NotImplementedError("layer-sliced forwards (hydra / trunk cache) are not available under pipeline parallelism")
        lm, spec, trunk = self.lm, self.lm.config, self.lm.transformer
        ref = inputs_embeds if inputs_embeds is not None else input_ids
        B, T = ref.shape[0], ref.shape[1]
        device, dtype = ref.device, self.param_dtype()
        oshape = (B, T, spec.final_hidden_size)
        # with sequence parallelism the residual stream between blocks — and therefore every stage boundary — is this tensor-
        # parallel rank's 1 / tp slice of the sequence: each TP rank relays its own shard to the same TP rank of the next stage
        # (the reference runs TP x PP with SP on, configs/nemo_configs/megatron_65b.yaml:47-50,80)
        tpc = getattr(lm, "_tp_context", None)
        sp_on = tpc is not None and tpc.sp
        hshape = (B, T // tpc.size, spec.hidden_size) if sp_on else (B, T, spec.hidden_size)
        if self.mode == "probe":
            self.probe_shape = hshape
            raise StageBoundary()
        past_len = past_key_values[0][0].shape[2] if past_key_values else 0
        if position_ids is None:
            if attention_mask is not None:
                position_ids = (attention_mask.long().cumsum(-1) - 1).clamp_min(0)[:, -T:]
            else:
                position_ids = torch.arange(past_len, past_len + T, device=device).unsqueeze(0).expand(B, T)
        grad_mode = torch.is_grad_enabled()

        V = self.virtual
        scheduled = self.mode == "schedule"
        ctx, x, j = None, None, 0
        presents = [] if use_cache else None
        phantoms = []
        for c in ([self.chunk] if scheduled else range(V)):
            vfirst, vlast = self.first and c == 0, self.last and c == V - 1
            if vfirst:
                x = trunk.embed(input_ids, position_ids, inputs_embeds)
            elif scheduled:
                x = self.input_tensor
                assert x is not None and tuple(x.shape) == hshape, "pipeline schedule handed over a mismatched activation"
            elif grad_mode:
                anchor = torch.zeros((), device=device, requires_grad=True)
                x = _RecvFromPrev.apply(anchor, self, hshape, dtype)
            else:
                x = torch.empty(hshape, dtype=dtype, device=device)
                dist.recv(x, src=self.ring_prev, group=self.group)
            if ctx is None:
                ctx = build_attn_context(spec, attention_mask, position_ids, T, past_len, x.dtype, device)
            lo, hi = self.chunks[c]
            for i in range(lo, hi):
                past = past_key_values[j] if past_key_values else None
                x, present = trunk.h[i](x, ctx, past, use_cache)
                if presents is not None:
                    presents.append(present)
                j += 1
            if vlast:
                break
            if scheduled:
                self.output_tensor = x
                raise StageBoundary()
            # relay: hand the activation to the next rank of the ring
            if grad_mode:
                if not x.requires_grad:  # fully frozen chunk: the next stage still returns a gradient, consume it
                    x = x.detach().requires_grad_(True)
                phantoms.append(_SendToNext.apply(x, self))
            else:
                dist.send(x.contiguous(), dst=self.ring_next, group=self.group)
        if use_cache and not presents:  # a stage without blocks still has to report the cache length
            presents = [(x.new_zeros(B, 1, past_len + T, 1), x.new_zeros(B, 1, past_len + T, 1))]

        logits = None
        if self.last:
            x = trunk.ln_f(x)
            if compute_logits:
                logits = project(lm.lm_head, x)
            if self.mode == "relay":
                self._broadcast_outputs(x, logits, compute_logits)
        else:  # relay, a rank that does not hold the final chunk
            x = torch.empty(oshape, dtype=dtype, device=device)
            logits = torch.empty((B, T, spec.vocab_size), dtype=dtype, device=device) if compute_logits else None
            self._broadcast_outputs(x, logits, compute_logits)
        if phantoms:  # ties the outputs to the graphs of the chunks whose activation left this rank
            phantom = phantoms[0] if len(phantoms) == 1 else torch.stack(phantoms).sum()
            x = x + phantom.to(x.dtype)
            if logits is not None:
                logits = logits + phantom.to(logits.dtype)
        loss = None
        if labels is not None and logits is not None:
            import torch.nn.functional as F

            loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1),
                                   ignore_index=-100)
        return CausalLMOutput(logits=logits, past_key_values=presents,
                              hidden_states=(x,) if output_hidden_states else None, loss=loss, last_hidden_state=x)

    def param_dtype(self):
        for p in self.lm.parameters():
            return p.dtype
        return torch.float32

    def _broadcast_outputs(self, x, logits, with_logits: bool):
        dist.broadcast(x.detach() if self.last else x, src=self.last_global, group=self.group)
        if with_logits:
            dist.broadcast(logits.detach() if self.last else logits, src=self.last_global, group=self.group)


# ---- installation ----------------------------------------------------------------------------------------------------------
def _find_lm(model):
    from trlx_b200.nn.transformer import CausalLM

    if isinstance(model, CausalLM):
        return model
    for m in model.modules():
        if isinstance(m, CausalLM):
            return m
    raise TypeError("pipeline parallelism needs a decoder-only CausalLM inside the model")


def apply_pipeline_parallel(model, group, rank: int, size: int, virtual: int = 1) -> PipelineStage:
    """Keep this stage's slice(s) of ``model``'s decoder, free the rest, install the stage forward.  ``virtual`` > 1 gives
    every rank that many non-adjacent model chunks (``virtual_pipeline_model_parallel_size`` in the NeMo recipes)."""
    lm = _find_lm(model)
    if getattr(lm, "_pp", None) is not None:
        return lm._pp
    stage = PipelineStage(lm, group, rank, size, virtual)
    trunk = lm.transformer
    for i in range(len(trunk.h)):
        if not any(lo <= i < hi for lo, hi in stage.chunks):
            trunk.h[i] = _Hole()
    tied = lm.lm_head.weight is trunk.wte.weight
    stage.tied = tied
    if not stage.first and not (tied and stage.last):
        trunk.wte.weight.requires_grad_(False)
        trunk.wte.weight.data = trunk.wte.weight.data.new_empty(0)
    if not stage.first:
        if trunk.wpe is not None:
            trunk.wpe.weight.requires_grad_(False)
            trunk.wpe.weight.data = trunk.wpe.weight.data.new_empty(0)
        if trunk.emb_norm is not None:
            for p in trunk.emb_norm.parameters():
                p.requires_grad_(False)
    if not stage.last:
        for p in trunk.ln_f.parameters():
            p.requires_grad_(False)
        if not tied:
            lm.lm_head.weight.requires_grad_(False)
            lm.lm_head.weight.data = lm.lm_head.weight.data.new_empty(0)
            if lm.lm_head.bias is not None:
                lm.lm_head.bias.requires_grad_(False)
    lm._pp = stage
    if size > 1:  # every stage continues a generation with the token the last stage sampled (generation.sync_tokens)
        lm._sample_sync = getattr(lm, "_sample_sync", []) + [(group, stage.last_global)]
    logger.info(f"pipeline stage {rank}/{size}: blocks {' + '.join(f'[{lo}, {hi})' for lo, hi in stage.chunks)} of {len(trunk.h)}"
                f"{' +embeddings' if stage.first else ''}{' +lm_head' if stage.last else ''}")
    return stage


def allreduce_tied_embedding_grads(stage: PipelineStage):
    """A tied embedding is used by the first stage (lookup) and the last (LM head): sum the two gradients."""
    if not getattr(stage, "tied", False) or stage.size == 1:
        return
    w = stage.lm.transformer.wte.weight
    if stage.first or stage.last:
        g = w.grad if w.grad is not None else torch.zeros_like(w)
    else:
        g = None
    # the whole pipeline group takes part (cheap, and avoids building a {first,last} sub-group): middle stages add zeros
    if g is None:
        H = stage.lm.config.hidden_size
        g = torch.zeros(stage.lm.config.vocab_size, H, dtype=stage.param_dtype(), device=next(stage.lm.parameters()).device)
    dist.all_reduce(g, group=stage.group)
    if stage.first or stage.last:
        w.grad = g


# ---- point-to-point helpers for the schedule --------------------------------------------------------------------------------
def _exchange(stage: PipelineStage, send_next=None, send_prev=None, recv_prev_shape=None, recv_next_shape=None,
              dtype=None, device=None):
    """One grouped round of point-to-point transfers with the neighbours.  Returns ``(from_prev, from_next)``."""
    ops, from_prev, from_next = [], None, None
    if send_prev is not None:
        ops.append(dist.P2POp(dist.isend, send_prev.contiguous(), stage.prev_global, stage.group))
    if recv_prev_shape is not None:
        from_prev = torch.empty(recv_prev_shape, dtype=dtype, device=device)
        ops.append(dist.P2POp(dist.irecv, from_prev, stage.prev_global, stage.group))
    if send_next is not None:
        ops.append(dist.P2POp(dist.isend, send_next.contiguous(), stage.next_global, stage.group))
    if recv_next_shape is not None:
        from_next = torch.empty(recv_next_shape, dtype=dtype, device=device)
        ops.append(dist.P2POp(dist.irecv, from_next, stage.next_global, stage.group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return from_prev, from_next


def run_1f1b(stage: PipelineStage, microbatches: Sequence[Any], loss_fn: Callable[[Any], Tuple[torch.Tensor, Dict]],
             device, loss_scale: float = 1.0, before_backward: Optional[Callable] = None,
             after_backward: Optional[Callable] = None) -> List[Optional[Dict]]:
    """Non-interleaved 1F1B over ``microbatches``.  ``loss_fn(mb) → (loss, stats)`` is the trainer's loss; gradients
    accumulate into ``.grad``.  Returns the per-micro-batch stats on the last stage (``None`` elsewhere)."""
    M, P, r = len(microbatches), stage.size, stage.rank
    dtype = stage.param_dtype()
    warm = min(P - 1 - r, M)
    steady = M - warm
    pending: deque = deque()  # (input leaf, output) in forward order
    stats: List[Optional[Dict]] = []

    def probe(mb):
        if stage.first:
            return None
        stage.mode = "probe"
        try:
            with torch.no_grad():
                loss_fn(mb)
            raise RuntimeError("the loss function never called the language model")
        except StageBoundary:
            return stage.probe_shape
        finally:
            stage.mode = "relay"

    def forward(mb, x):
        if x is not None:
            x.requires_grad_(True)
        stage.mode, stage.input_tensor, stage.output_tensor = "schedule", x, None
        try:
            loss, st = loss_fn(mb)
            if not stage.last:
                raise RuntimeError("a non-final stage completed its loss function; the LM was not called")
            out = loss * loss_scale
            stats.append(st)
        except StageBoundary:
            out = stage.output_tensor
            stats.append(None)
        finally:
            stage.mode, stage.input_tensor = "relay", None
        pending.append((x, out))
        return out

    def backward(gy):
        x, y = pending.popleft()
        if before_backward is not None:
            before_backward()
        if stage.last:
            y.backward()
        elif y.requires_grad:  # (a fully frozen first stage has nothing to differentiate)
            torch.autograd.backward(y, gy)
        if after_backward is not None:
            after_backward()
        return None if x is None else x.grad

    kw = dict(dtype=dtype, device=device)
    # warm-up: forwards only
    for k in range(warm):
        x, _ = _exchange(stage, recv_prev_shape=probe(microbatches[k]), **kw)
        y = forward(microbatches[k], x)
        _exchange(stage, send_next=y.detach(), **kw)
    x = None
    if steady > 0:
        x, _ = _exchange(stage, recv_prev_shape=probe(microbatches[warm]), **kw)
    # steady state: one forward, one backward
    for i in range(steady):
        k = warm + i
        y = forward(microbatches[k], x)
        gy = None
        if not stage.last:
            _, gy = _exchange(stage, send_next=y.detach(), recv_next_shape=tuple(pending[0][1].shape), **kw)
        gx = backward(gy)
        nxt_shape = probe(microbatches[k + 1]) if i < steady - 1 else None
        if stage.first:
            x = None
        else:
            x, _ = _exchange(stage, send_prev=gx, recv_prev_shape=nxt_shape, **kw)
    # cool-down: remaining backwards
    for _ in range(warm):
        y = pending[0][1]
        _, gy = _exchange(stage, recv_next_shape=tuple(y.shape), **kw)
        gx = backward(gy)
        if not stage.first:
            _exchange(stage, send_prev=gx, **kw)
    return stats


# ---- interleaved (virtual-stage) schedule -------------------------------------------------------------------------------------
def interleaved_order(P: int, V: int, M: int, r: int) -> List[Tuple[str, int, int]]:
    """The order in which rank ``r`` of ``P`` runs its ``2 * M * V`` operations ``(kind, microbatch, chunk)``.

    Micro-batches advance in groups of ``P``: a rank runs chunk 0 for the whole group, then chunk 1, ... so that by the time
    it returns to the first member of the group the activation has been round the ring once.  Backwards mirror the forwards
    (last chunk first).  A rank starts alternating forward / backward after ``2 (P - r - 1) + (V - 1) P`` warm-up forwards,
    which is what shrinks the bubble by ``V`` compared with plain 1F1B (reference recipes:
    ``virtual_pipeline_model_parallel_size``, ``configs/nemo_configs/megatron_20b.yaml:53``)."""
    seq_f: List[Tuple[str, int, int]] = []
    seq_b: List[Tuple[str, int, int]] = []
    for m0 in range(0, M, P):
        group = range(m0, min(m0 + P, M))
        for c in range(V):
            seq_f += [("F", m, c) for m in group]
            seq_b += [("B", m, V - 1 - c) for m in group]
    total = M * V
    warm = min(2 * (P - r - 1) + (V - 1) * P, total)
    ops = seq_f[:warm]
    for i in range(total - warm):
        ops += [seq_f[warm + i], seq_b[i]]
    return ops + seq_b[total - warm:]


def interleaved_rounds(P: int, V: int, M: int) -> List[List[Optional[Tuple[str, int, int]]]]:
    """Lock-step plan of the whole pipeline: ``rounds[t][r]`` is the operation rank ``r`` runs in round ``t`` (``None`` = idle).

    Every rank derives the same plan by simulating all ranks: a rank runs the next operation of its
    :func:`interleaved_order` as soon as the tensor it consumes was produced in an earlier round.  All transfers of a round
    then form one matched set — rank ``a`` sends to ``b`` in round ``t`` exactly when ``b`` posts the receive in round ``t`` —
    so each rank can issue them as a single grouped ``batch_isend_irecv`` and no ordering between neighbours can deadlock,
    whatever ``P``, ``V`` and ``M`` are."""
    S = P * V
    orders = [interleaved_order(P, V, M, r) for r in range(P)]
    ptr = [0] * P
    have = set()  # (kind, m, virtual stage) whose input has been delivered / whose forward has run
    rounds: List[List[Optional[Tuple[str, int, int]]]] = []
    while any(ptr[r] < len(orders[r]) for r in range(P)):
        acts: List[Optional[Tuple[str, int, int]]] = [None] * P

        def ready(r, op):
            kind, m, c = op
            s = c * P + r
            if kind == "F":
                return s == 0 or ("F", m, s) in have
            return ("done", m, s) in have and (s == S - 1 or ("B", m, s) in have)

        for r in range(P):
            if ptr[r] < len(orders[r]) and ready(r, orders[r][ptr[r]]):
                acts[r] = orders[r][ptr[r]]
        if not any(a is not None for a in acts):
            # a micro-batch count that is not a multiple of P can leave every rank waiting on an operation that sits deeper
            # in a neighbour's list: let each rank pull its first runnable operation forward (the dependency graph is acyclic,
            # so one always exists)
            for r in range(P):
                for i in range(ptr[r], len(orders[r])):
                    if ready(r, orders[r][i]):
                        orders[r].insert(ptr[r], orders[r].pop(i))
                        acts[r] = orders[r][ptr[r]]
                        break
        if not any(a is not None for a in acts):
            raise RuntimeError(f"pipeline schedule cannot make progress (P={P}, V={V}, M={M})")
        for r, a in enumerate(acts):
            if a is None:
                continue
            ptr[r] += 1
            kind, m, c = a
            s = c * P + r
            if kind == "F":
                have.add(("done", m, s))
                if s < S - 1:
                    have.add(("F", m, s + 1))
            elif s > 0:
                have.add(("B", m, s - 1))
        rounds.append(acts)
    return rounds


def run_interleaved(stage: PipelineStage, microbatches: Sequence[Any], loss_fn: Callable[[Any], Tuple[torch.Tensor, Dict]],
                    device, loss_scale: float = 1.0, before_backward: Optional[Callable] = None,
                    after_backward: Optional[Callable] = None) -> List[Optional[Dict]]:
    """Interleaved 1F1B: every rank owns ``stage.virtual`` model chunks and a micro-batch visits each rank that many times.
    Same contract as :func:`run_1f1b`."""
    M, P, V, r = len(microbatches), stage.size, stage.virtual, stage.rank
    S = P * V
    dtype = stage.param_dtype()
    plan = interleaved_rounds(P, V, M)
    shapes: Dict[int, Tuple[int, ...]] = {}
    inbox: Dict[Tuple[str, int, int], torch.Tensor] = {}
    saved: Dict[Tuple[int, int], Tuple[Optional[torch.Tensor], torch.Tensor]] = {}
    stats: List[Optional[Dict]] = [None] * M

    def shape_of(m):
        if m not in shapes:
            stage.mode = "probe"
            try:
                with torch.no_grad():
                    loss_fn(microbatches[m])
                raise RuntimeError("the loss function never called the language model")
            except StageBoundary:
                shapes[m] = stage.probe_shape
            finally:
                stage.mode = "relay"
        return shapes[m]

    def forward(m, c):
        s = c * P + r
        x = inbox.pop(("F", m, s)) if s > 0 else None
        if x is not None:
            x.requires_grad_(True)
        stage.mode, stage.chunk, stage.input_tensor, stage.output_tensor = "schedule", c, x, None
        try:
            loss, st = loss_fn(microbatches[m])
            if s != S - 1:
                raise RuntimeError("a non-final stage completed its loss function; the LM was not called")
            out = loss * loss_scale
            stats[m] = st
        except StageBoundary:
            out = stage.output_tensor
        finally:
            stage.mode, stage.chunk, stage.input_tensor = "relay", 0, None
        saved[(m, c)] = (x, out)
        return None if s == S - 1 else out.detach()

    def backward(m, c):
        s = c * P + r
        x, y = saved.pop((m, c))
        if before_backward is not None:
            before_backward()
        if s == S - 1:
            y.backward()
        else:
            gy = inbox.pop(("B", m, s))
            if y.requires_grad:  # (a fully frozen chunk has nothing to differentiate)
                torch.autograd.backward(y, gy)
        if after_backward is not None:
            after_backward()
        if x is None:
            return None
        return x.grad if x.grad is not None else torch.zeros_like(x)

    for acts in plan:
        mine = acts[r]
        out = None
        if mine is not None:
            kind, m, c = mine
            out = forward(m, c) if kind == "F" else backward(m, c)
        ops = []
        if out is not None:
            dst = stage.ring_next if mine[0] == "F" else stage.ring_prev
            ops.append(dist.P2POp(dist.isend, out.contiguous(), dst, stage.group))
        for q, a in enumerate(acts):  # what the other ranks produced for this one in the same round
            if a is None or q == r:
                continue
            kind, m, c = a
            s = c * P + q
            if kind == "F" and s < S - 1 and (q + 1) % P == r:
                key, src = ("F", m, s + 1), stage.global_ranks[q]
            elif kind == "B" and s > 0 and (q - 1) % P == r:
                key, src = ("B", m, s - 1), stage.global_ranks[q]
            else:
                continue
            buf = torch.empty(shape_of(m), dtype=dtype, device=device)
            inbox[key] = buf
            ops.append(dist.P2POp(dist.irecv, buf, src, stage.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    assert not saved and not inbox, "pipeline schedule finished with tensors in flight"
    return stats if stage.last else [None] * M


def run_schedule(stage: PipelineStage, microbatches, loss_fn, device, **kw) -> List[Optional[Dict]]:
    """1F1B, interleaved when the stage holds more than one model chunk."""
    return (run_interleaved if stage.virtual > 1 else run_1f1b)(stage, microbatches, loss_fn, device, **kw)


def broadcast_stats(stage: PipelineStage, stats: Optional[Dict[str, Any]], device) -> Dict[str, Any]:
    """Ship the last stage's (scalar) statistics to the whole pipeline group."""
    keys = [sorted(stats.keys())] if stage.last else [None]
    dist.broadcast_object_list(keys, src=stage.last_global, group=stage.group)
    names = keys[0]
    if stage.last:
        vals = torch.stack([torch.as_tensor(stats[k], dtype=torch.float32, device=device).reshape(()) for k in names])
    else:
        vals = torch.empty(len(names), dtype=torch.float32, device=device)
    dist.broadcast(vals, src=stage.last_global, group=stage.group)
    return {k: vals[i] for i, k in enumerate(names)}


def broadcast_module_from_last(stage: PipelineStage, modules: Sequence[nn.Module]):
    """Heads (value / Q heads) are trained on the last stage only; refresh the replicas the other stages use when they
    score samples."""
    for m in modules:
        for p in m.parameters():
            if p.numel():
                dist.broadcast(p.data, src=stage.last_global, group=stage.group)
