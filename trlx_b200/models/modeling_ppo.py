"""PPO method config, KL controllers and the value-head / hydra model wrappers.

Parity map (reference ``trlx/models/modeling_ppo.py``): KL controllers ``:35-67``; ``PPOConfig`` ``:117-134`` with
``get_advantages_and_returns`` ``:136-173`` and ``loss`` ``:175-238``; ``CausalLMOutputWithValue`` ``:244-252``;
``make_value_branch`` ``:255-263``; ``AutoModelForCausalLMWithValueHead`` ``:266-382``;
``AutoModelForCausalLMWithHydraValueHead`` ``:385-499``; ``ModelBranch`` and the per-family branches ``:502-1222``;
seq2seq variants ``:1228-1592``; ``hf_get_branch_class`` ``:1598-1637``.

B200 design: the hydra is the *native* shape of the model.  Because the base LM can be entered at any block
(``hidden_in`` / ``start_layer``), the frozen trunk is evaluated ONCE and its branch-point activation feeds the policy
branch, the frozen reference branch and the value head (``forward_hydra`` in the reference re-runs the whole trunk,
``modeling_ppo.py:442``).  ``score()`` additionally computes label log-probs through the fused LM-head kernel, so no
``[B,T,V]`` logits tensor exists on the rollout/scoring path.
"""
from __future__ import annotations

import copy
import gc
import re
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from trlx_b200 import ops
from trlx_b200.data.method_configs import MethodConfig, register_method
from trlx_b200.models.modeling_base import PreTrainedModelWrapper, base_lm, export_base_state_dict
from trlx_b200.models.peft import PeftModel
from trlx_b200.nn import hf_compat
from trlx_b200.nn.transformer import CausalLM, build_attn_context
from trlx_b200.utils.modeling import make_head, whiten


# ---- KL controllers ---------------------------------------------------------------------------------------------------
class AdaptiveKLController:
    """Proportional controller on the KL coefficient (Ziegler et al. 2019, §2.2):
    ``β ← β · (1 + clip(kl/target − 1, ±0.2) · n_steps / horizon)``."""

    def __init__(self, init_kl_coef: float, target: float, horizon: int):
        self.value = init_kl_coef
        self.target = target
        self.horizon = horizon

    def update(self, current: float, n_steps: int):
        error = float(np.clip(current / self.target - 1, -0.2, 0.2))
        self.value *= 1 + error * n_steps / self.horizon


class FixedKLController:
    """Constant KL coefficient."""

    def __init__(self, kl_coef: float):
        self.value = kl_coef

    def update(self, current: float, n_steps: int):
        pass


# ---- method config ----------------------------------------------------------------------------------------------------
@dataclass
@register_method
class PPOConfig(MethodConfig):
    """PPO hyper-parameters.

    :param ppo_epochs: passes over each batch of rollouts
    :param num_rollouts: rollouts collected per outer iteration
    :param chunk_size: prompts generated per rollout chunk
    :param init_kl_coef: initial KL penalty coefficient
    :param target: target KL for the adaptive controller (``None`` → fixed coefficient)
    :param horizon: adaptive-controller horizon
    :param gamma / lam: discount and GAE λ
    :param cliprange / cliprange_value: PPO clip ranges for the ratio and for value predictions
    :param vf_coef: value-loss weight
    :param scale_reward: ``"running"``, ``"ref"`` or anything else for no scaling
    :param ref_mean / ref_std: fixed reward statistics for ``scale_reward == "ref"``
    :param cliprange_reward: rewards are clipped to ±this value (falsy → no clipping)
    :param gen_kwargs: generation kwargs for rollouts and evaluation
    :param gen_experience_kwargs: if set, used instead of ``gen_kwargs`` while collecting experience
    :param num_value_layers_unfrozen: >0 gives the value function its own trainable copy of the top-k blocks
    """

    ppo_epochs: int
    num_rollouts: int
    chunk_size: int
    init_kl_coef: float
    target: Optional[float]
    horizon: int
    gamma: float
    lam: float
    cliprange: float
    cliprange_value: float
    vf_coef: float
    scale_reward: Optional[str]
    ref_mean: Optional[float]
    ref_std: Optional[float]
    cliprange_reward: float
    gen_kwargs: dict
    gen_experience_kwargs: Optional[dict] = None
    num_value_layers_unfrozen: int = 0

    def get_advantages_and_returns(self, values: torch.Tensor, rewards: torch.Tensor, response_length: int,
                                   use_whitening: Optional[bool] = True, width_tensor: Optional[torch.Tensor] = None
                                   ) -> Tuple[torch.Tensor, torch.Tensor]:
        """GAE: ``δ_t = r_t + γV_{t+1} − V_t``, ``A_t = δ_t + γλA_{t+1}``, ``returns = A + V``; advantages are whitened
        (globally across ranks) and detached.  Runs as one scan kernel + one whiten kernel on CUDA."""
        return ops.gae_and_whiten(values, rewards, response_length, self.gamma, self.lam, bool(use_whitening),
                                  width_tensor=width_tensor)

    def loss(self, logprobs, values, old_logprobs, old_values, advantages, returns, mask, width_tensor=None):
        """Clipped PPO objective → ``(loss, flat stats dict)``.  On CUDA the loss, its ~20 statistics and
        both gradients come out of one fused pass; stats stay on the device (no ``.item()`` syncs here)."""
        return ops.ppo_loss(logprobs, values, old_logprobs, old_values, advantages, returns, mask, self.cliprange,
                            self.cliprange_value, self.vf_coef, width_tensor=width_tensor)


# ---- outputs ----------------------------------------------------------------------------------------------------------
class _Output:
    """Dataclass-like output that also unpacks like HF's tuple form (``logits, *_, value = out``)."""

    _fields: Tuple[str, ...] = ()

    def to_tuple(self):
        return tuple(getattr(self, f) for f in self._fields if getattr(self, f) is not None)

    def __iter__(self):
        return iter(self.to_tuple())

    def __getitem__(self, i):
        if isinstance(i, str):
            return getattr(self, i)
        return self.to_tuple()[i]

    def keys(self):
        return [f for f in self._fields if getattr(self, f) is not None]


@dataclass
class CausalLMOutputWithValue(_Output):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Any] = None
    cross_attentions: Optional[Any] = None
    value: Optional[torch.Tensor] = None
    _fields = ("loss", "logits", "past_key_values", "hidden_states", "attentions", "cross_attentions", "value")


@dataclass
class Seq2SeqLMOutputWithValue(_Output):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    decoder_hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    decoder_attentions: Optional[Any] = None
    cross_attentions: Optional[Any] = None
    encoder_last_hidden_state: Optional[torch.Tensor] = None
    encoder_hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    encoder_attentions: Optional[Any] = None
    value: Optional[torch.Tensor] = None
    _fields = ("loss", "logits", "past_key_values", "decoder_hidden_states", "decoder_attentions", "cross_attentions",
               "encoder_last_hidden_state", "encoder_hidden_states", "encoder_attentions", "value")


# ---- branches ---------------------------------------------------------------------------------------------------------
class ModelBranch(nn.Module):
    """Frozen deep copy of the top ``num_layers_unfrozen`` blocks + final norm + LM head of a decoder-only model.
    One implementation serves every family because the base model is generic; the family-named subclasses below
    exist for import parity."""

    def __init__(self, base_model: nn.Module, *, num_layers_unfrozen: int, frozen: bool = True):
        super().__init__()
        lm = base_lm(base_model)
        self.config = lm.config
        blocks = list(lm.transformer.h)[-num_layers_unfrozen:] if num_layers_unfrozen > 0 else []
        self.decoder_blocks = nn.ModuleList(copy.deepcopy(b) for b in blocks)
        self.final_norm = copy.deepcopy(lm.transformer.ln_f)
        head = nn.Linear(lm.lm_head.in_features, lm.lm_head.out_features, bias=lm.lm_head.bias is not None,
                         device=lm.lm_head.weight.device, dtype=lm.lm_head.weight.dtype)
        head.load_state_dict(lm.lm_head.state_dict())
        self.lm_head = head
        self.hidden_size = self.config.hidden_size
        self.model_parallel = False
        self.device_map = None
        self.last_device = None
        self.gradient_checkpointing = False
        if frozen:
            for p in self.parameters():
                p.requires_grad_(False)

    def run_blocks(self, hidden_states, attention_mask=None, position_ids=None):
        B, T = hidden_states.shape[:2]
        if attention_mask is not None:  # hidden states may be sequence-sharded (sequence parallelism)
            T = attention_mask.shape[1]
        elif position_ids is not None:
            T = position_ids.shape[1]
        if position_ids is None:
            if attention_mask is not None:
                position_ids = (attention_mask.long().cumsum(-1) - 1).clamp_min(0)
            else:
                position_ids = torch.arange(T, device=hidden_states.device).unsqueeze(0).expand(B, T)
        ctx = build_attn_context(self.config, attention_mask, position_ids, T, 0, hidden_states.dtype, hidden_states.device)
        hiddens = []
        x = hidden_states
        for blk in self.decoder_blocks:
            hiddens.append(x)
            x, _ = blk(x, ctx)
        x = self.final_norm(x)
        hiddens.append(x)
        return x, tuple(hiddens)

    def forward(self, hidden_states: torch.Tensor, output_shape: Optional[torch.Size] = None, past_key_values=None,
                attention_mask=None, position_ids=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, use_cache=False, output_attentions=False, output_hidden_states=False,
                return_dict: Optional[bool] = True, **_):
        x, hiddens = self.run_blocks(hidden_states, attention_mask, position_ids)
        out = self.lm_head(x)
        if isinstance(self.lm_head, nn.Linear):  # LM branch → logits in fp32 like the reference (:684)
            out = out.float()
        if not return_dict:
            return (out,) + ((hiddens,) if output_hidden_states else ())
        return CausalLMOutputWithValue(logits=out, hidden_states=hiddens if output_hidden_states else None)


class GPTModelBranch(ModelBranch):
    """GPT-2 / GPT-J / GPT-Neo / GPT-NeoX branch."""


class OPTModelBranch(ModelBranch):
    pass


class BloomModelBranch(ModelBranch):
    pass


class LlamaModelBranch(ModelBranch):
    pass


class GPTBigCodeModelBranch(ModelBranch):
    pass


def hf_get_branch_class(config) -> type:
    """Branch class for a model config (``ArchSpec`` or HF-style object / dict with ``model_type``)."""
    mt = getattr(config, "model_type", None) or (config.get("model_type") if isinstance(config, dict) else None)
    table = {
        "gpt2": GPTModelBranch, "gptj": GPTModelBranch, "gpt_neo": GPTModelBranch, "gpt_neox": GPTModelBranch,
        "opt": OPTModelBranch, "bloom": BloomModelBranch, "llama": LlamaModelBranch, "mistral": LlamaModelBranch,
        "gpt_bigcode": GPTBigCodeModelBranch,
    }
    if mt in ("t5", "mt5"):
        return T5Branch
    if mt in table:
        return table[mt]
    raise ValueError(
        f"Unsupported architecture: `{mt}`. The following architectures are available for model branching:\n"
        f"{sorted(list(table) + ['t5'])}"
    )


def make_value_branch(base_model: nn.Module, num_value_layers_unfrozen: int) -> nn.Module:
    """Value function: an MLP on the last hidden state, or (k>0) a *trainable* copy of the top-k blocks whose
    ``lm_head`` is that MLP."""
    lm = base_lm(base_model)
    hidden = getattr(lm.config, "final_hidden_size", None) or (
        lm.config.hidden_size if hasattr(lm.config, "hidden_size") else lm.config.d_model)
    dtype, device = lm.dtype, lm.device
    head = make_head(hidden, 1, dtype).to(device)
    if num_value_layers_unfrozen == 0:
        return head
    branch_cls = hf_get_branch_class(lm.config)
    branch = branch_cls(base_model, num_layers_unfrozen=num_value_layers_unfrozen, frozen=False)
    branch.lm_head = head
    for p in branch.parameters():
        p.requires_grad_(True)
    return branch


# ---- decoder-only wrappers --------------------------------------------------------------------------------------------
class AutoModelForCausalLMWithValueHead(PreTrainedModelWrapper):
    """Causal LM + scalar value head ``v_head``."""

    _supported_modules = ["v_head"]
    _supported_args = ["peft_config", "num_value_layers_unfrozen"]
    arch_type = "causal"

    def __init__(self, base_model: nn.Module, peft_config=None, num_value_layers_unfrozen: int = 0):
        super().__init__(base_model, peft_config=peft_config)
        self.num_value_layers_unfrozen = num_value_layers_unfrozen
        self.v_head = make_value_branch(base_model, num_value_layers_unfrozen)

    # -- core ------------------------------------------------------------------------------------------------------------
    def _run_base(self, ignore_peft_adapter: bool = False, **kw):
        model = self.base_model
        if self.peft_type and ignore_peft_adapter and isinstance(model, PeftModel):
            with model.disable_adapter():
                return model(**kw)
        return model(**kw)

    def _value_from(self, hidden_states, attention_mask, position_ids):
        k = self.num_value_layers_unfrozen
        h = hidden_states[-(k + 1)]
        if k > 0:
            return self.v_head(h, attention_mask=attention_mask, position_ids=position_ids).logits.squeeze(-1)
        return self.v_head(h.to(self.v_head[0].weight.dtype)).squeeze(-1)

    def forward(self, input_ids=None, attention_mask=None, past_key_values=None, position_ids=None, head_mask=None,
                inputs_embeds=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict: Optional[bool] = None, ignore_peft_adapter: Optional[bool] = None):
        kw = dict(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                  past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=bool(use_cache),
                  output_hidden_states=True)
        if self.peft_type == "PREFIX_TUNING":
            kw.pop("past_key_values")
        out = self._run_base(bool(ignore_peft_adapter), **kw)
        value = self._value_from(out.hidden_states, attention_mask, position_ids)
        res = CausalLMOutputWithValue(loss=out.loss, logits=out.logits, past_key_values=out.past_key_values,
                                      hidden_states=out.hidden_states, value=value)
        if not return_dict:
            return res.to_tuple()
        return res

    def generate(self, *args, **kwargs) -> torch.Tensor:
        from trlx_b200.models.generation import generate

        return generate(self.base_model, *args, **kwargs)

    # -- persistence -----------------------------------------------------------------------------------------------------
    def _head_state(self, prefix: str, module: nn.Module) -> Dict[str, torch.Tensor]:
        sd = module.state_dict()
        if isinstance(module, ModelBranch):
            sd = hf_compat.branch_to_hf(module.config, sd)
        return {prefix + k: v for k, v in sd.items()}

    def state_dict(self, *args, heads_only: bool = False, **kwargs):
        """``v_head.*`` (+ ``base_model.<hf keys>`` unless ``heads_only``; with an adapter the base prefix is empty).
        Unlike the reference (dead code after an early return, ``modeling_ppo.py:363-366``), ``heads_only`` is honoured."""
        sd = self._head_state("v_head.", self.v_head)
        if not heads_only:
            sd.update(export_base_state_dict(self.base_model, prefix="" if self.peft_type else "base_model."))
        return sd

    def _load_heads(self, state_dict: Dict[str, torch.Tensor], strict: bool) -> None:
        for name in ("v_head", "frozen_head", "ilql_heads"):
            module = getattr(self, name, None)
            sub = {k[len(name) + 1:]: v for k, v in state_dict.items() if k.startswith(name + ".")}
            if module is None or not sub:
                continue
            if isinstance(module, ModelBranch):
                sub = hf_compat.branch_from_hf(module.config, sub)
            module.load_state_dict(sub, strict=strict)

    def post_init(self, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        state_dict = state_dict or {}
        strict = not self.peft_type and any(k.startswith("v_head.") for k in state_dict)
        self._load_heads(state_dict, strict)
        gc.collect()


class AutoModelForCausalLMWithHydraValueHead(AutoModelForCausalLMWithValueHead):
    """Causal LM + value head + frozen *reference branch* sharing the bottom layers (``frozen_head``)."""

    _supported_modules = ["v_head", "frozen_head"]
    _supported_args = ["num_layers_unfrozen", "peft_config", "num_value_layers_unfrozen"]

    def __init__(self, base_model: nn.Module, *, num_layers_unfrozen: int = -1, peft_config=None,
                 num_value_layers_unfrozen: int = 0):
        super().__init__(base_model, peft_config=peft_config, num_value_layers_unfrozen=num_value_layers_unfrozen)
        self.num_layers_unfrozen = num_layers_unfrozen
        self.frozen_head: Optional[ModelBranch] = None
        if self.num_layers_unfrozen > 0 and not self.peft_type:
            self._build_frozen_head()

    def _build_frozen_head(self):
        branch_class = hf_get_branch_class(base_lm(self.base_model).config)
        self.frozen_head = branch_class(self.base_model, num_layers_unfrozen=self.num_layers_unfrozen).eval()

    @property
    def branch_layer(self) -> int:
        """Index of the first block that differs between policy and reference."""
        L = base_lm(self.base_model).config.num_layers
        return L - self.num_layers_unfrozen if self.num_layers_unfrozen > 0 else 0

    def forward_hydra(self, input_ids=None, attention_mask=None, past_key_values=None, position_ids=None,
                      head_mask=None, inputs_embeds=None, use_cache=None, output_attentions=None,
                      output_hidden_states=None, return_dict: Optional[bool] = None):
        """Reference-policy forward.  ``return_dict`` falsy ⇒ logits only (reference default)."""
        want_dict = True if return_dict is None else bool(return_dict)
        if self.peft_type or self.frozen_head is None:
            out = self.forward(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                               inputs_embeds=inputs_embeds, return_dict=True, ignore_peft_adapter=True)
        else:
            lm = base_lm(self.base_model)
            with torch.no_grad():
                trunk = lm(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                           inputs_embeds=inputs_embeds, stop_layer=self.branch_layer, output_hidden_states=True)
            out = self.frozen_head(trunk.last_hidden_state, attention_mask=attention_mask, position_ids=position_ids,
                                   output_hidden_states=True, return_dict=True)
            out.hidden_states = tuple(trunk.hidden_states[:-1]) + tuple(out.hidden_states)
        if not want_dict:
            return out.logits
        return out

    # -- single-pass scoring (B200 path) ---------------------------------------------------------------------------------
    def trunk_hidden(self, input_ids, attention_mask, position_ids):
        """Activation entering the first unfrozen block (no grad: the trunk below it is frozen)."""
        lm = base_lm(self.base_model)
        with torch.no_grad():
            return lm(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                      stop_layer=self.branch_layer).last_hidden_state

    def can_share_trunk(self) -> bool:
        # a value branch deeper than the policy branch needs an activation below the shared one: no sharing then
        # (``score`` and ``forward`` must compute the same value function)
        import os

        return (self.frozen_head is not None and not self.peft_type and self.num_layers_unfrozen > 0
                and self.num_value_layers_unfrozen <= self.num_layers_unfrozen
                and os.environ.get("TRLX_B200_SHARE_TRUNK", "1") != "0")

    def freeze_trunk_parameters(self) -> int:
        """Mark every parameter below the branch layer as frozen (returns how many tensors changed).

        The shared-trunk paths run the trunk under ``no_grad`` — which is what ``num_layers_unfrozen`` promises — but
        the reference's ``freeze_bottom_causal_layers`` leaves *learned position embeddings* (GPT-2 ``wpe``, OPT
        ``embed_positions``) trainable, so there they still receive a gradient through the frozen blocks
        (``trlx/utils/modeling.py:22-38``).  Here they are frozen explicitly, so the optimizer does not hold state for,
        decay, or reduce tensors that never get a gradient.  ``train.trainer_kwargs.cache_trunk=False`` +
        ``TRLX_B200_SHARE_TRUNK=0`` restores the reference behaviour (full forward/backward)."""
        if not self.can_share_trunk():
            return 0
        lm = base_lm(self.base_model)
        keep = set()
        blocks = list(lm.transformer.h)
        for b in blocks[self.branch_layer:]:
            keep.update(id(p) for p in b.parameters())
        for name in ("ln_f",):
            m = getattr(lm.transformer, name, None)
            if m is not None:
                keep.update(id(p) for p in m.parameters())
        if getattr(lm, "lm_head", None) is not None:
            keep.update(id(p) for p in lm.lm_head.parameters())
        n = 0
        for p in lm.parameters():
            if id(p) not in keep and p.requires_grad:
                p.requires_grad_(False)
                n += 1
        return n

    def policy_from_trunk(self, trunk_hidden, attention_mask, position_ids, need_value_hidden: bool = True):
        """Run the trainable top blocks on a cached trunk activation → ``(final hidden, hidden_states tail)``."""
        lm = base_lm(self.base_model)
        out = lm(hidden_in=trunk_hidden, attention_mask=attention_mask, position_ids=position_ids,
                 start_layer=self.branch_layer, output_hidden_states=need_value_hidden, compute_logits=False)
        return out.last_hidden_state, out.hidden_states

    def score(self, input_ids, attention_mask, position_ids, labels, trunk_hidden=None, with_ref: bool = True,
              rows: Optional[Tuple[int, int]] = None):
        """One pass → ``(logprobs, values, ref_logprobs, trunk_hidden)`` for ``labels`` aligned with positions
        (``labels[:, t]`` is the token predicted from position ``t``; negative = ignore).  ``rows=(lo, hi)`` restricts
        heads to sequence positions ``[lo, hi)``.  Falls back to two forwards when no trunk can be shared."""
        lm = base_lm(self.base_model)
        lo, hi = rows if rows is not None else (0, labels.shape[1])
        lab = labels[:, lo:hi]
        if self.can_share_trunk():
            if trunk_hidden is None:
                trunk_hidden = self.trunk_hidden(input_ids, attention_mask, position_ids)
            final, tail = self.policy_from_trunk(trunk_hidden, attention_mask, position_ids)
            k = self.num_value_layers_unfrozen
            if k > 0:
                vh = tail[-(k + 1)] if k < len(tail) else trunk_hidden
                values = self.v_head(vh, attention_mask=attention_mask, position_ids=position_ids).logits.squeeze(-1)[:, lo:hi]
            else:
                values = _mlp_head(self.v_head, final[:, lo:hi])
            logprobs, _ = ops.fused_logprob(final[:, lo:hi], lm.lm_head.weight, lm.lm_head.bias, lab)
            ref_logprobs = None
            if with_ref:
                with torch.no_grad():
                    ref_final, _ = self.frozen_head.run_blocks(trunk_hidden, attention_mask, position_ids)
                    ref_logprobs, _ = ops.fused_logprob(ref_final[:, lo:hi], self.frozen_head.lm_head.weight,
                                                        self.frozen_head.lm_head.bias, lab)
            return logprobs, values, ref_logprobs, trunk_hidden
        out = self.forward(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids, return_dict=True)
        final = out.hidden_states[-1]
        logprobs, _ = ops.fused_logprob(final[:, lo:hi], lm.lm_head.weight, lm.lm_head.bias, lab)
        values = out.value[:, lo:hi]
        ref_logprobs = None
        if with_ref:
            with torch.no_grad():
                ref = self.forward_hydra(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                         return_dict=True)
                ref_final = ref.hidden_states[-1]
                head = self.frozen_head.lm_head if self.frozen_head is not None else lm.lm_head
                ref_logprobs, _ = ops.fused_logprob(ref_final[:, lo:hi], head.weight, head.bias, lab)
        return logprobs, values, ref_logprobs, None

    # -- persistence -----------------------------------------------------------------------------------------------------
    def state_dict(self, *args, heads_only: bool = False, **kwargs):
        sd = super().state_dict(*args, heads_only=heads_only, **kwargs)
        if not heads_only and self.frozen_head is not None:
            sd.update(self._head_state("frozen_head.", self.frozen_head))
        return sd

    def post_init(self, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        """Loads heads; when the checkpoint carries ``frozen_head.*`` and no branch exists yet, the number of
        unfrozen layers is inferred from the key indices and the branch is rebuilt before loading."""
        state_dict = state_dict or {}
        if not self.peft_type and self.frozen_head is None:
            for k in state_dict:
                m = re.search(r"^frozen_head\.decoder_blocks\.(\d+)\.", k)
                if m:
                    self.num_layers_unfrozen = max(self.num_layers_unfrozen, int(m.group(1)) + 1)
            if self.num_layers_unfrozen > 0 and any(k.startswith("frozen_head.") for k in state_dict):
                self._build_frozen_head()
        super().post_init(state_dict)


def _mlp_head(head: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """``Linear → ReLU → Linear(·,1)`` value MLP; first GEMM on the tcgen05 kernel (ReLU fused when no grad)."""
    l1, l2 = head[0], head[2]
    if x.dtype != l1.weight.dtype:
        x = x.to(l1.weight.dtype)
    h = ops.linear(x, l1.weight, l1.bias, "relu")
    return torch.nn.functional.linear(h, l2.weight, l2.bias).squeeze(-1)


# ---- seq2seq wrappers -------------------------------------------------------------------------------------------------
class T5Branch(nn.Module):
    """Frozen copy of the top-k T5 decoder blocks + final norm + LM head."""

    def __init__(self, base_model: nn.Module, *, num_layers_unfrozen: int, frozen: bool = True):
        super().__init__()
        lm = base_lm(base_model)
        self.config = lm.config
        self.first_layer = len(lm.decoder.block) - num_layers_unfrozen
        self.decoder_blocks = nn.ModuleList(copy.deepcopy(b) for b in list(lm.decoder.block)[-num_layers_unfrozen:])
        self.final_norm = copy.deepcopy(lm.decoder.final_layer_norm)
        self.lm_head = copy.deepcopy(lm.lm_head)
        self.dropout = nn.Identity()
        self.hidden_size = self.config.d_model
        self.rel_bias_owner = [copy.deepcopy(lm.decoder.block[0].self_attn.relative_attention_bias)]
        self.relative_attention_bias = self.rel_bias_owner[0]
        if frozen:
            for p in self.parameters():
                p.requires_grad_(False)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                output_hidden_states=False, return_dict=True, **_):
        from trlx_b200.nn.t5 import run_decoder_blocks

        x, hiddens = run_decoder_blocks(self.config, self.decoder_blocks, self.final_norm, self.relative_attention_bias,
                                        hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask)
        if self.config.scale_logits and isinstance(self.lm_head, nn.Linear) and self.lm_head.out_features > 1:
            x_out = x * (self.config.d_model ** -0.5)
        else:
            x_out = x
        logits = self.lm_head(x_out)
        if not return_dict:
            return (logits,)
        return Seq2SeqLMOutputWithValue(logits=logits, decoder_hidden_states=hiddens if output_hidden_states else None)


class AutoModelForSeq2SeqLMWithValueHead(PreTrainedModelWrapper):
    """Encoder-decoder LM + value head on the decoder states."""

    _supported_modules = ["v_head"]
    _supported_args = ["peft_config", "num_value_layers_unfrozen"]
    arch_type = "seq2seq"

    def __init__(self, base_model: nn.Module, peft_config=None, num_value_layers_unfrozen: int = 0):
        super().__init__(base_model, peft_config=peft_config)
        self.num_value_layers_unfrozen = num_value_layers_unfrozen
        lm = base_lm(base_model)
        if num_value_layers_unfrozen > 0:
            self.v_head = T5Branch(base_model, num_layers_unfrozen=num_value_layers_unfrozen, frozen=False)
            self.v_head.lm_head = make_head(lm.config.d_model, 1, lm.dtype).to(lm.device)
            for p in self.v_head.parameters():
                p.requires_grad_(True)
        else:
            self.v_head = make_head(lm.config.d_model, 1, lm.dtype).to(lm.device)

    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, decoder_attention_mask=None,
                encoder_outputs=None, past_key_values=None, inputs_embeds=None, decoder_inputs_embeds=None,
                head_mask=None, decoder_head_mask=None, cross_attn_head_mask=None, use_cache=None,
                output_attentions=None, output_hidden_states=True, return_dict: Optional[bool] = None,
                ignore_peft_adapter: Optional[bool] = None, labels=None):
        model = self.base_model
        kw = dict(input_ids=input_ids, attention_mask=attention_mask, decoder_input_ids=decoder_input_ids,
                  decoder_attention_mask=decoder_attention_mask, encoder_outputs=encoder_outputs,
                  past_key_values=past_key_values, use_cache=bool(use_cache), output_hidden_states=True, labels=labels)
        if self.peft_type and ignore_peft_adapter and isinstance(model, PeftModel):
            with model.disable_adapter():
                out = model(**kw)
        else:
            out = model(**kw)
        k = self.num_value_layers_unfrozen
        h = out.decoder_hidden_states[-(k + 1)]
        if k > 0:
            value = self.v_head(h, attention_mask=decoder_attention_mask, encoder_hidden_states=out.encoder_last_hidden_state,
                                encoder_attention_mask=attention_mask).logits.squeeze(-1)
        else:
            value = self.v_head(h.to(self.v_head[0].weight.dtype)).squeeze(-1)
        res = Seq2SeqLMOutputWithValue(loss=out.loss, logits=out.logits, past_key_values=out.past_key_values,
                                       decoder_hidden_states=out.decoder_hidden_states,
                                       encoder_last_hidden_state=out.encoder_last_hidden_state,
                                       encoder_hidden_states=out.encoder_hidden_states, value=value)
        if not return_dict:
            return res.to_tuple()
        return res

    def generate(self, *args, **kwargs):
        from trlx_b200.models.generation import generate

        return generate(self.base_model, *args, **kwargs)

    def state_dict(self, *args, heads_only: bool = False, **kwargs):
        sd = {"v_head." + k: v for k, v in self.v_head.state_dict().items()}
        if not heads_only:
            sd.update(export_base_state_dict(self.base_model, prefix="" if self.peft_type else "base_model."))
        return sd

    def post_init(self, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        state_dict = state_dict or {}
        strict = not self.peft_type and any(k.startswith("v_head.") for k in state_dict)
        for name in ("v_head", "frozen_head", "ilql_heads"):
            module = getattr(self, name, None)
            sub = {k[len(name) + 1:]: v for k, v in state_dict.items() if k.startswith(name + ".")}
            if module is not None and sub:
                module.load_state_dict(sub, strict=strict)
        gc.collect()


class AutoModelForSeq2SeqLMWithHydraValueHead(AutoModelForSeq2SeqLMWithValueHead):
    _supported_modules = ["v_head", "frozen_head"]
    _supported_args = ["num_layers_unfrozen", "peft_config", "num_value_layers_unfrozen"]

    def __init__(self, base_model: nn.Module, *, num_layers_unfrozen: int = -1, peft_config=None,
                 num_value_layers_unfrozen: int = 0):
        super().__init__(base_model, peft_config=peft_config, num_value_layers_unfrozen=num_value_layers_unfrozen)
        self.num_layers_unfrozen = num_layers_unfrozen
        self.frozen_head = None
        if self.num_layers_unfrozen > 0 and not self.peft_type:
            self.frozen_head = T5Branch(self.base_model, num_layers_unfrozen=self.num_layers_unfrozen).eval()

    def forward_hydra(self, input_ids=None, attention_mask=None, decoder_input_ids=None, decoder_attention_mask=None,
                      encoder_outputs=None, past_key_values=None, inputs_embeds=None, decoder_inputs_embeds=None,
                      head_mask=None, decoder_head_mask=None, cross_attn_head_mask=None, use_cache=None,
                      output_attentions=None, output_hidden_states=None, return_dict: Optional[bool] = None):
        want_dict = True if return_dict is None else bool(return_dict)
        if self.peft_type or self.frozen_head is None:
            out = self.forward(input_ids=input_ids, attention_mask=attention_mask, decoder_input_ids=decoder_input_ids,
                               decoder_attention_mask=decoder_attention_mask, return_dict=True, ignore_peft_adapter=True)
        else:
            with torch.no_grad():
                full = self.forward(input_ids=input_ids, attention_mask=attention_mask, decoder_input_ids=decoder_input_ids,
                                    decoder_attention_mask=decoder_attention_mask, return_dict=True)
                h = full.decoder_hidden_states[-(self.num_layers_unfrozen + 1)]
                out = self.frozen_head(h, attention_mask=decoder_attention_mask,
                                       encoder_hidden_states=full.encoder_last_hidden_state,
                                       encoder_attention_mask=attention_mask, output_hidden_states=True, return_dict=True)
        if not want_dict:
            return out.logits
        return out

    def state_dict(self, *args, heads_only: bool = False, **kwargs):
        sd = super().state_dict(*args, heads_only=heads_only, **kwargs)
        if not heads_only and self.frozen_head is not None:
            sd.update({"frozen_head." + k: v for k, v in self.frozen_head.state_dict().items()})
        return sd

    def post_init(self, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        state_dict = state_dict or {}
        if not self.peft_type and self.frozen_head is None:
            for k in state_dict:
                m = re.search(r"^frozen_head\.decoder_blocks\.(\d+)\.", k)
                if m:
                    self.num_layers_unfrozen = max(self.num_layers_unfrozen, int(m.group(1)) + 1)
            if self.num_layers_unfrozen > 0 and any(k.startswith("frozen_head.") for k in state_dict):
                self.frozen_head = T5Branch(self.base_model, num_layers_unfrozen=self.num_layers_unfrozen).eval()
        super().post_init(state_dict)
