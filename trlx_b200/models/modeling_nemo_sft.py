"""Model-parallel supervised fine-tuning model (reference: ``trlx/models/modeling_nemo_sft.py`` — ``SFTGPT`` with Apex's
``vocab_parallel_cross_entropy`` as its loss, ``:433-457``).

:func:`vocab_parallel_cross_entropy` is the loss for logits whose *vocabulary* dimension is sharded over the
tensor-parallel group: three small all-reduces (row max, sum of exponentials, the target logit) instead of gathering
``[B, T, V]`` logits; its backward is local (softmax − one-hot on the owning shard).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from trlx_b200.models.megatron_api import MegatronModelMixin
from trlx_b200.parallel import state as parallel_state


class _VocabParallelCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, group, rank, world):
        part = logits.shape[-1]
        lo = rank * part
        x = logits.float()
        mx = x.max(dim=-1).values
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        x = x - mx.unsqueeze(-1)
        ex = x.exp()
        denom = ex.sum(-1)
        local = (target >= lo) & (target < lo + part)
        idx = (target - lo).clamp(0, part - 1)
        tgt = torch.where(local, x.gather(-1, idx.unsqueeze(-1)).squeeze(-1), torch.zeros_like(denom))
        if world > 1:
            dist.all_reduce(denom, group=group)
            dist.all_reduce(tgt, group=group)
        loss = denom.log() - tgt
        ctx.save_for_backward(ex / denom.unsqueeze(-1), local, idx)
        ctx.in_dtype = logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        softmax, local, idx = ctx.saved_tensors
        grad = softmax.clone()
        grad.scatter_add_(-1, idx.unsqueeze(-1), -local.to(grad.dtype).unsqueeze(-1))
        return (grad * g.unsqueeze(-1)).to(ctx.in_dtype), None, None, None, None


def vocab_parallel_cross_entropy(vocab_parallel_logits: torch.Tensor, target: torch.Tensor,
                                 tp: Optional[parallel_state.ModelParallelState] = None) -> torch.Tensor:
    """Per-token cross entropy ``[...]`` from logits ``[..., V / tp]`` that hold this rank's vocabulary slice
    ``[rank · V/tp, (rank + 1) · V/tp)`` and full-vocabulary ``target`` ids ``[...]``."""
    st = tp or parallel_state.get_model_parallel()
    world = st.tp_size if st.tp_group is not None else 1
    return _VocabParallelCrossEntropy.apply(vocab_parallel_logits, target, st.tp_group, st.tp_rank if world > 1 else 0, world)


class SFTGPT(MegatronModelMixin, nn.Module):
    """Causal LM (tensor-parallel blocks) trained with next-token cross entropy over the positions selected by
    ``loss_mask`` (reference ``SFTGPT``).  With ``vocab_parallel=True`` the LM head weight is sharded by vocabulary rows and
    the loss is :func:`vocab_parallel_cross_entropy`; otherwise the (replicated) head feeds the fused log-prob kernel."""

    def __init__(self, config=None, language_model: Optional[nn.Module] = None, vocab_parallel: bool = False,
                 dtype: torch.dtype = torch.bfloat16, metric_fn=None):
        super().__init__()
        self.config, self.metric_fn = config, metric_fn
        st = parallel_state.get_model_parallel()
        if language_model is None:
            from trlx_b200.models.modeling_base import build_base_model
            from trlx_b200.parallel.tensor_parallel import apply_tensor_parallel

            language_model = build_base_model(config.model.model_path, "causal", dtype=dtype)
            if st.tp_group is not None and st.tp_size > 1:
                apply_tensor_parallel(language_model, st.tp_group, st.tp_rank, st.tp_size,
                                      sequence_parallel=bool(getattr(config.train.parallel, "sequence_parallel", False)))
        self.language_model = language_model
        self.vocab_parallel = bool(vocab_parallel and st.tp_group is not None and st.tp_size > 1)
        self._tp = st

    def _head_weight(self) -> torch.Tensor:
        w = self.language_model.lm_head.weight
        if not self.vocab_parallel:
            return w
        if w.shape[0] % self._tp.tp_size:
            raise ValueError("vocab_parallel needs a vocabulary divisible by the tensor-parallel size")
        return w.chunk(self._tp.tp_size, dim=0)[self._tp.tp_rank]

    def forward(self, input_ids, attention_mask=None, position_ids=None, loss_mask: Optional[torch.Tensor] = None):
        """Mean next-token loss over ``loss_mask[:, 1:]`` (all non-masked positions when omitted) and the logits."""
        out = self.language_model(input_ids, attention_mask=attention_mask, position_ids=position_ids, output_hidden_states=True)
        labels = input_ids[:, 1:]
        mask = (loss_mask if loss_mask is not None else (attention_mask if attention_mask is not None
                                                         else torch.ones_like(input_ids)))[:, 1:].float()
        if self.vocab_parallel:
            from trlx_b200.parallel.tensor_parallel import _CopyToTP

            hidden = _CopyToTP.apply(out.hidden_states[-1][:, :-1], self._tp.tp_group)
            local_logits = torch.nn.functional.linear(hidden, self._head_weight())
            per_token = vocab_parallel_cross_entropy(local_logits.float(), labels, self._tp)
        else:
            per_token = torch.nn.functional.cross_entropy(out.logits[:, :-1].float().transpose(1, 2), labels, reduction="none")
        loss = (per_token * mask).sum() / mask.sum().clamp_min(1.0)
        return loss, out.logits

    def _loss(self, batch):
        """``batch``: mapping with ``input_ids`` and optionally ``attention_mask`` / ``loss_mask`` (reference loss closure
        ``:433-457``)."""
        loss, _ = self(batch["input_ids"], batch.get("attention_mask"), batch.get("position_ids"), batch.get("loss_mask"))
        return loss, {"loss": float(loss.detach())}

    def build_attention_mask_and_position_ids(self, input_ids, pad_token_id: int = 0):
        """``(attention_mask, position_ids)`` of a right- or left-padded batch (reference ``:394-406`` builds Megatron's
        lower-triangular mask explicitly; causality is handled inside the attention kernels here)."""
        am = input_ids.ne(pad_token_id).long()
        return am, (am.cumsum(-1) - 1).clamp_min(0)

    def generate(self, *args, **kwargs):
        from trlx_b200.models.generation import generate

        return generate(self.language_model, *args, **kwargs)
