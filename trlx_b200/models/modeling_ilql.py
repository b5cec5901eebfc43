"""ILQL method config, heads and model wrappers.

Parity map (reference ``trlx/models/modeling_ilql.py``): ``topk_mask`` ``:29-33``, ``batched_index_select`` ``:36-45``,
``ILQLConfig`` + ``loss`` ``:48-166``, ``ILQLHeads`` (V head, 1–2 Q heads, frozen Polyak-averaged target Q heads)
``:169-227``, ``AutoModelForCausalLMWithILQLHeads`` incl. the advantage-shifted sampler ``:262-442`` and the seq2seq
variant ``:481-666``.

B200 design: every vocabulary-wide quantity in the loss is a *(gathered value, logsumexp)* pair —
``Q_i = q_i[a]``, ``CQL_i = lse(q_i) − q_i[a]``, ``AWAC = lse(logits) − logits[a]`` — so the training path evaluates
the second linear of each Q head and the LM head through the fused tcgen05 GEMM + online-logsumexp epilogue
(:func:`trlx_b200.ops.fused_logprob`) and never materialises a ``[B, A, V]`` tensor (the reference allocates five).
"""
from __future__ import annotations

import gc
import os
from copy import deepcopy
from dataclasses import dataclass
from functools import reduce
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200 import ops
from trlx_b200.data.ilql_types import ILQLBatch
from trlx_b200.data.method_configs import MethodConfig, register_method
from trlx_b200.models.generation import sample_sync_groups, sync_tokens
from trlx_b200.models.modeling_base import PreTrainedModelWrapper, base_lm, export_base_state_dict
from trlx_b200.models.peft import PeftModel
from trlx_b200.utils.modeling import flatten_dict, get_tensor_stats, make_head


def topk_mask(xs: torch.Tensor, k: int) -> torch.Tensor:
    """Keep the k largest entries of each row, ``-inf`` elsewhere."""
    if k > xs.shape[-1]:
        return xs
    kth = torch.topk(xs, k)[0][:, -1].unsqueeze(-1)
    return torch.where(xs < kth, torch.full_like(xs, float("-inf")), xs)


def batched_index_select(x: torch.Tensor, idxs: torch.Tensor, dim: int) -> torch.Tensor:
    """``out[b, i, :] = x[b, idxs[b, i], :]`` (for ``dim == 1``)."""
    return x.gather(dim=dim, index=idxs.unsqueeze(-1).expand(idxs.shape[0], idxs.shape[1], x.shape[-1]))


@dataclass
class ILQLParts:
    """Loss inputs in gathered form (each ``[B, A]``; ``V`` is ``[B, A+1]``)."""

    q_taken: List[torch.Tensor]
    q_lse: List[torch.Tensor]
    target_q_taken: List[torch.Tensor]
    values: torch.Tensor
    policy_logprob: torch.Tensor


@dataclass
@register_method
class ILQLConfig(MethodConfig):
    """
    :param tau: expectile for the value loss (0.5 = MSE, →1 = max over Q)
    :param gamma: discount
    :param cql_scale: weight of the conservative (CQL) regulariser
    :param awac_scale: weight of the advantage-weighted behaviour-cloning term
    :param alpha: Polyak coefficient for target-Q sync
    :param beta: AWAC temperature (0 = plain cross entropy)
    :param steps_for_target_q_sync: sync period in optimizer steps
    :param two_qs: use two Q heads and take their minimum
    :param gen_kwargs: generation kwargs (``beta``, ``top_k``, ``temperature``, ``max_new_tokens`` …)
    """

    tau: float
    gamma: float
    cql_scale: float
    awac_scale: float
    alpha: float
    beta: float
    steps_for_target_q_sync: int
    two_qs: bool
    gen_kwargs: dict

    # -- gathered-form loss (what the kernels produce) ------------------------------------------------------------------
    def loss_from_parts(self, parts: ILQLParts, rewards: torch.Tensor, dones: torch.Tensor):
        mask = dones[:, :-1].to(parts.values.dtype)
        n = mask.sum().clamp_min(1)
        V = parts.values[:, :-1]
        Vnext = parts.values[:, 1:] * dones[:, 1:].to(parts.values.dtype)
        target = rewards + self.gamma * Vnext.detach()
        loss_q = sum((((Qi - target) * mask) ** 2).sum() / n for Qi in parts.q_taken)
        tQ = reduce(torch.minimum, [t.detach() for t in parts.target_q_taken])
        diff2 = (tQ - V) ** 2
        loss_v = (((tQ >= V).to(V.dtype) * self.tau + (tQ < V).to(V.dtype) * (1 - self.tau)) * diff2 * mask).sum() / n
        loss_cql = sum(((lse - Qi) * mask).sum() / n for Qi, lse in zip(parts.q_taken, parts.q_lse))
        with torch.no_grad():
            awac_weight = torch.exp(self.beta * (tQ - V))
        loss_awac = (-parts.policy_logprob * awac_weight * mask).sum() / n
        loss = loss_q + loss_v + self.cql_scale * loss_cql + self.awac_scale * loss_awac
        with torch.no_grad():
            stats = dict(
                losses=dict(loss=loss.detach(), loss_q=loss_q.detach(), loss_v=loss_v.detach(), loss_cql=loss_cql.detach(),
                            loss_awac=loss_awac.detach()),
                values=get_tensor_stats(V.detach(), mask, n),
                qvalues={str(i): get_tensor_stats(q.detach(), mask, n) for i, q in enumerate(parts.q_taken)},
                awac_weight=get_tensor_stats(awac_weight, mask, n),
            )
        return loss, flatten_dict(stats)

    # -- reference-shaped entry point -------------------------------------------------------------------------------------
    def loss(self, outputs, labels):
        """``outputs = (logits, (qs, target_qs, vs))`` with full ``[B, ·, V]`` tensors (reference signature), or an
        :class:`ILQLParts`."""
        if isinstance(outputs, ILQLParts):
            return self.loss_from_parts(outputs, labels.rewards, labels.dones)
        logits, (qs, target_qs, vs) = outputs
        if isinstance(labels, ILQLBatch):
            actions = labels.input_ids[:, 1:].gather(dim=1, index=labels.actions_ixs).unsqueeze(-1)
        else:
            actions = labels.decoder_input_ids[:, 1:].unsqueeze(-1)
        action_logits = batched_index_select(logits, labels.actions_ixs, dim=1) if logits.shape[1] != actions.shape[1] or isinstance(labels, ILQLBatch) else logits
        lse_pi = torch.logsumexp(action_logits.float(), -1)
        parts = ILQLParts(
            q_taken=[q.gather(-1, actions).squeeze(-1) for q in qs],
            q_lse=[torch.logsumexp(q.float(), -1) for q in qs],
            target_q_taken=[q.gather(-1, actions).squeeze(-1).detach() for q in target_qs],
            values=vs[..., 0] if vs.dim() == 3 else vs,
            policy_logprob=action_logits.gather(-1, actions).squeeze(-1).float() - lse_pi,
        )
        return self.loss_from_parts(parts, labels.rewards, labels.dones)


class ILQLHeads(nn.Module):
    """V head + (1|2) Q heads + frozen target Q heads (``H → 2H → {1, V}`` MLPs)."""

    def __init__(self, hidden_size: int, vocab_size: int, two_qs: bool, alpha: float, dtype: torch.dtype):
        super().__init__()
        self.hidden_size, self.vocab_size, self.two_qs, self.alpha = hidden_size, vocab_size, two_qs, alpha
        self.v_head = make_head(hidden_size, 1, dtype)
        n_qs = 2 if two_qs else 1
        self.q_heads = nn.ModuleList(make_head(hidden_size, vocab_size, dtype) for _ in range(n_qs))
        self.target_q_heads = nn.ModuleList(deepcopy(q) for q in self.q_heads)
        for t in self.target_q_heads:
            t.requires_grad_(False)

    def forward(self, hs: torch.Tensor, states_ixs: Optional[torch.Tensor] = None,
                actions_ixs: Optional[torch.Tensor] = None, **kwargs):
        """Full-width outputs ``(qs, target_qs, vs)``.  Hidden states are gathered at the state / action indices
        *before* the heads so the vocabulary-wide GEMMs only see the rows that matter."""
        if states_ixs is not None:
            states_hs = batched_index_select(hs, states_ixs, 1)
            actions_hs = batched_index_select(hs, actions_ixs, 1)
        else:
            states_hs = actions_hs = hs
        qs = tuple(q(actions_hs) for q in self.q_heads)
        target_qs = tuple(q(actions_hs) for q in self.target_q_heads)
        return qs, target_qs, self.v_head(states_hs)

    def gathered(self, hs, states_ixs, actions_ixs, actions):
        """Gathered-form outputs for the loss: per Q head ``(q[a], lse(q))``, target ``q[a]``, ``V`` — the V-wide
        second linear of each head runs inside the fused GEMM+logsumexp kernel."""
        states_hs = batched_index_select(hs, states_ixs, 1)
        actions_hs = batched_index_select(hs, actions_ixs, 1)
        q_taken, q_lse, tq_taken = [], [], []
        for head in self.q_heads:
            mid = F.relu(ops.linear(actions_hs, head[0].weight, head[0].bias))
            lp, lse = ops.fused_logprob(mid, head[2].weight, head[2].bias, actions)
            q_taken.append(lp + lse)
            q_lse.append(lse)
        with torch.no_grad():
            for head in self.target_q_heads:
                mid = ops.linear(actions_hs, head[0].weight, head[0].bias, "relu")
                lp, lse = ops.fused_logprob(mid, head[2].weight, head[2].bias, actions)
                tq_taken.append(lp + lse)
        values = self.v_head(states_hs).squeeze(-1)
        return q_taken, q_lse, tq_taken, values

    def _sync_target_q_heads(self, alpha: float):
        for tgt, src in zip(self.target_q_heads, self.q_heads):
            for tp, sp in zip(tgt.parameters(), src.parameters()):
                if tp.is_cuda and tp.dtype == torch.bfloat16 and ops.available() and tp.is_contiguous():
                    ops.C.lerp_(tp.data, sp.data, alpha)
                else:
                    tp.data.copy_(alpha * sp.data + (1.0 - alpha) * tp.data)

    def sync_target_q_heads(self):
        """Polyak update ``target ← α·q + (1−α)·target`` (one fused kernel per tensor on CUDA)."""
        self._sync_target_q_heads(self.alpha)


@dataclass
class CausalILQLOutput:
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    value: Optional[torch.Tensor] = None
    qs: Optional[Tuple[torch.Tensor, ...]] = None
    target_qs: Optional[Tuple[torch.Tensor, ...]] = None


@dataclass
class Seq2SeqILQLOutput:
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    value: Optional[torch.Tensor] = None
    qs: Optional[Tuple[torch.Tensor, ...]] = None
    target_qs: Optional[Tuple[torch.Tensor, ...]] = None
    encoder_outputs: Optional[Tuple[Any]] = None


def _ilql_sample(logits, qs, vs, beta, top_k, temperature, logit_mask_row=None):
    """π ∝ softmax(topk(log_softmax(logits) + β·(Q − V)) / T)  (SURVEY A.6)."""
    logits = logits.float()
    if logit_mask_row is not None:
        if logit_mask_row.shape[-1] < logits.shape[-1]:  # masks may cover only the task's own tokens
            logit_mask_row = F.pad(logit_mask_row, (0, logits.shape[-1] - logit_mask_row.shape[-1]), value=False)
        logits = logits.masked_fill(logit_mask_row, float("-inf"))
    pi_beta = F.log_softmax(logits, -1)
    shifted = topk_mask(pi_beta + beta * (qs.float() - vs.float()), top_k)
    if temperature == 0.0:
        return shifted.argmax(dim=-1, keepdim=True)
    return torch.multinomial(F.softmax(shifted / temperature, -1), num_samples=1)


class AutoModelForCausalLMWithILQLHeads(PreTrainedModelWrapper):
    """Causal LM + ILQL heads (Snell et al. 2022)."""

    _supported_modules = ["ilql_heads"]
    _supported_args = ["two_qs", "alpha", "peft_config"]
    arch_type = "causal"

    def __init__(self, base_model: nn.Module, *, two_qs: bool = True, alpha: float = 0.99, peft_config=None):
        super().__init__(base_model, peft_config=peft_config)
        lm = base_lm(base_model)
        self.two_qs, self.alpha = two_qs, alpha
        self.ilql_heads = ILQLHeads(getattr(lm.config, "final_hidden_size", None) or lm.config.hidden_size, lm.config.vocab_size, two_qs, alpha, dtype=lm.dtype).to(lm.device)

    def _base(self, bypass_adapter: bool = False, **kw):
        if self.peft_type == "PREFIX_TUNING" and not bypass_adapter:
            kw.pop("past_key_values", None)
        if bypass_adapter and isinstance(self.base_model, PeftModel):
            return self.base_model.base_model(**kw)
        return self.base_model(**kw)

    def forward(self, input_ids, attention_mask=None, position_ids=None, past_key_values=None, actions_ixs=None,
                states_ixs=None, return_dict=False, bypass_peft_prompt_adapter=False):
        out = self._base(bypass_peft_prompt_adapter, input_ids=input_ids, attention_mask=attention_mask,
                         position_ids=position_ids, past_key_values=past_key_values, use_cache=True,
                         output_hidden_states=True)
        qs, target_qs, vs = self.ilql_heads(out.hidden_states[-1], states_ixs=states_ixs, actions_ixs=actions_ixs)
        if return_dict:
            return CausalILQLOutput(out.logits, out.past_key_values, out.hidden_states, vs, qs, target_qs)
        return out.logits, qs, target_qs, vs, out.past_key_values

    def loss_parts(self, batch: ILQLBatch) -> ILQLParts:
        """Gathered-form forward for training (no vocabulary-wide tensors)."""
        lm = base_lm(self.base_model)
        out = self.base_model(input_ids=batch.input_ids, attention_mask=batch.attention_mask, output_hidden_states=False,
                              compute_logits=False) if not self.peft_type else None
        if out is None:
            full = self.base_model(input_ids=batch.input_ids, attention_mask=batch.attention_mask, output_hidden_states=True)
            hs = full.hidden_states[-1]
        else:
            hs = out.last_hidden_state
        actions = batch.input_ids[:, 1:].gather(dim=1, index=batch.actions_ixs)
        q_taken, q_lse, tq_taken, values = self.ilql_heads.gathered(hs, batch.states_ixs, batch.actions_ixs, actions)
        action_hs = batched_index_select(hs, batch.actions_ixs, 1)
        pol_lp, _ = ops.fused_logprob(action_hs, lm.lm_head.weight, lm.lm_head.bias, actions)
        return ILQLParts(q_taken, q_lse, tq_taken, values, pol_lp)

    @torch.no_grad()
    def generate(self, input_ids, attention_mask=None, position_ids=None, past_key_values=None, beta=1,
                 max_new_tokens=32, max_length=1024, temperature=1, top_k=20, logit_mask=None, pad_token_id=None,
                 eos_token_id=None):
        """Sampling with ``log π_β + β·(min target-Q − V)`` re-weighting; finished rows keep emitting EOS."""
        cfg = base_lm(self.base_model).config
        pad_token_id = pad_token_id if pad_token_id is not None else cfg.pad_token_id
        eos_token_id = eos_token_id if eos_token_id is not None else cfg.eos_token_id
        if attention_mask is None:
            attention_mask = input_ids.not_equal(pad_token_id)
        attention_mask = attention_mask.long()
        if position_ids is None:
            position_ids = (attention_mask.cumsum(-1) - 1).masked_fill(attention_mask.eq(0), 0)
        samples = input_ids.clone()
        max_new_tokens = min(max_new_tokens, max_length - input_ids.shape[1])
        finished = torch.zeros(input_ids.shape[0], 1, dtype=torch.long, device=input_ids.device)
        bypass = False
        for step in range(max_new_tokens):
            logits, _, target_qs, vs, past_key_values = self.forward(
                input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                past_key_values=past_key_values, bypass_peft_prompt_adapter=bypass)
            qs = torch.minimum(target_qs[0][:, -1], target_qs[1][:, -1]) if self.two_qs else target_qs[0][:, -1]
            mask_row = None
            if logit_mask is not None:
                last = input_ids[:, -1].to(logit_mask.device)
                rows = logit_mask[last.clamp_max(logit_mask.shape[0] - 1)].to(logits.device).bool()
                mask_row = rows & (last < logit_mask.shape[0]).to(rows.device).unsqueeze(-1)
            nxt = _ilql_sample(logits[:, -1], qs, vs[:, -1], beta, top_k, temperature, mask_row)
            nxt = sync_tokens(nxt, sample_sync_groups(self))
            nxt = (1 - finished) * nxt + finished * eos_token_id
            finished = (nxt == eos_token_id).long()
            samples = torch.hstack((samples, nxt))
            attention_mask = torch.hstack((attention_mask, (nxt != eos_token_id).long()))
            position_ids = (position_ids[:, -1] + 1).view(-1, 1)
            input_ids = nxt
            if self.peft_type and step == 0 and "LORA" not in self.peft_type:
                bypass = True
                n = self.peft_config.num_virtual_tokens
                attention_mask = torch.cat((torch.ones(nxt.shape[0], n, dtype=attention_mask.dtype, device=nxt.device),
                                            attention_mask), dim=1)
                if self.peft_type == "PROMPT_TUNING":
                    position_ids = position_ids + n
            if bool(torch.all(finished)):
                break
        return samples

    def sync_target_q_heads(self):
        self.ilql_heads.sync_target_q_heads()

    def state_dict(self, *args, heads_only: bool = False, **kwargs):
        sd = {"ilql_heads." + k: v for k, v in self.ilql_heads.state_dict().items()}
        if not heads_only:
            sd.update(export_base_state_dict(self.base_model, prefix="" if self.peft_type else "base_model."))
        return sd

    def post_init(self, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        state_dict = state_dict or {}
        sub = {k[len("ilql_heads."):]: v for k, v in state_dict.items() if k.startswith("ilql_heads.")}
        if sub:
            self.ilql_heads.load_state_dict(sub, strict=not self.peft_type)
        gc.collect()


class AutoModelForSeq2SeqLMWithILQLHeads(PreTrainedModelWrapper):
    """Encoder-decoder LM + ILQL heads on the decoder states."""

    _supported_modules = ["ilql_heads"]
    _supported_args = ["two_qs", "alpha", "peft_config"]
    arch_type = "seq2seq"

    def __init__(self, base_model: nn.Module, *, two_qs: bool = True, alpha: float = 0.99, peft_config=None):
        super().__init__(base_model, peft_config=peft_config)
        lm = base_lm(base_model)
        self.two_qs, self.alpha = two_qs, alpha
        self.ilql_heads = ILQLHeads(lm.config.d_model, lm.config.vocab_size, two_qs, alpha, dtype=lm.dtype).to(lm.device)

    def sync_target_q_heads(self):
        self.ilql_heads.sync_target_q_heads()

    def state_dict(self, *args, heads_only: bool = False, **kwargs):
        sd = {"ilql_heads." + k: v for k, v in self.ilql_heads.state_dict().items()}
        if not heads_only:
            sd.update(export_base_state_dict(self.base_model, prefix="" if self.peft_type else "base_model."))
        return sd

    def post_init(self, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        state_dict = state_dict or {}
        sub = {k[len("ilql_heads."):]: v for k, v in state_dict.items() if k.startswith("ilql_heads.")}
        if sub:
            self.ilql_heads.load_state_dict(sub, strict=not self.peft_type)
        gc.collect()

    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, past_key_values=None,
                encoder_outputs=None, actions_ixs=None, states_ixs=None, output_attentions=None,
                output_hidden_states=True, return_dict=False, bypass_peft_prompt_adapter=False, position_ids=None,
                decoder_attention_mask=None, decoder_inputs_embeds=None):
        model = self.base_model
        if bypass_peft_prompt_adapter and isinstance(model, PeftModel):
            model = model.base_model
        extra = {}
        if decoder_attention_mask is not None:
            extra["decoder_attention_mask"] = decoder_attention_mask
        if decoder_inputs_embeds is not None:
            extra["decoder_inputs_embeds"] = decoder_inputs_embeds
        out = model(input_ids=input_ids, attention_mask=attention_mask, decoder_input_ids=decoder_input_ids,
                    past_key_values=past_key_values, encoder_outputs=encoder_outputs, use_cache=True,
                    output_hidden_states=True, **extra)
        hs = out.decoder_hidden_states[-1]
        qs, target_qs, vs = self.ilql_heads(hs, states_ixs=states_ixs, actions_ixs=actions_ixs)
        enc = (out.encoder_last_hidden_state,)
        if return_dict:
            return Seq2SeqILQLOutput(out.logits, out.past_key_values, out.decoder_hidden_states, vs, qs, target_qs, enc)
        return out.logits, qs, target_qs, vs, out.past_key_values, enc

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, decoder_input_ids=None, past_key_values=None,
                 encoder_outputs=None, beta=1, max_new_tokens=32, max_length=1024, temperature=1, top_k=20,
                 logit_mask=None, pad_token_id=None, eos_token_id=None, decoder_attention_mask=None):
        # (`decoder_attention_mask` is accepted for signature parity; like the reference, ``:622-666``, generation starts from the
        # decoder start token and every generated position is attended)
        cfg = base_lm(self.base_model).config
        pad_token_id = pad_token_id if pad_token_id is not None else cfg.pad_token_id
        eos_token_id = eos_token_id if eos_token_id is not None else cfg.eos_token_id
        if attention_mask is None:
            attention_mask = input_ids.not_equal(pad_token_id)
        if decoder_input_ids is None:
            decoder_input_ids = input_ids.new_full((input_ids.shape[0], 1), cfg.decoder_start_token_id)
        samples = decoder_input_ids.clone()
        max_new_tokens = min(max_new_tokens, max_length - decoder_input_ids.shape[1])
        finished = torch.zeros(input_ids.shape[0], 1, dtype=torch.long, device=input_ids.device)
        bypass = False
        step_in = decoder_input_ids
        for step in range(max_new_tokens):
            logits, _, target_qs, vs, past_key_values, encoder_outputs = self.forward(
                input_ids=input_ids, attention_mask=attention_mask, decoder_input_ids=step_in,
                past_key_values=past_key_values, encoder_outputs=encoder_outputs, bypass_peft_prompt_adapter=bypass)
            qs = torch.minimum(target_qs[0][:, -1], target_qs[1][:, -1]) if self.two_qs else target_qs[0][:, -1]
            mask_row = None
            if logit_mask is not None:
                mask_row = logit_mask[step_in[:, -1].to(logit_mask.device)].to(logits.device).bool()
            nxt = _ilql_sample(logits[:, -1], qs, vs[:, -1], beta, top_k, temperature, mask_row)
            nxt = (1 - finished) * nxt + finished * eos_token_id
            finished = (nxt == eos_token_id).long()
            samples = torch.hstack((samples, nxt))
            step_in = nxt
            if self.peft_type and step == 0 and "LORA" not in self.peft_type:
                bypass = True
                n = self.peft_config.num_virtual_tokens
                attention_mask = torch.cat((torch.ones(nxt.shape[0], n, dtype=attention_mask.dtype, device=nxt.device),
                                            attention_mask.long()), dim=1)
            if bool(torch.all(finished)):
                break
        return samples
