"""Generic decoder-only transformer driven by an :class:`~trlx_b200.nn.arch.ArchSpec`.

This is the framework's own model code (the reference relies on HF ``transformers`` modeling files
and re-implements their forward passes per family in ``trlx/models/modeling_ppo.py:547-1222``).
Canonical layout — chosen for the kernels, not for any checkpoint format:

* every projection is ``[out, in]`` (K-major for the tcgen05 GEMM);
* Q, K, V live in ONE weight ``attn.qkv`` laid out ``[Q heads | K heads | V heads]``;
* gated MLPs keep ``[gate | up]`` in ONE weight ``mlp.up``;
* the trunk can be entered at any block with a cached activation (``hidden_in`` / ``start_layer``) —
  that is what lets the frozen trunk run once for policy, reference and value branches.

HF checkpoint key names are produced/consumed by :mod:`trlx_b200.nn.hf_compat`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200.nn.arch import ArchSpec


@dataclass
class CausalLMOutput:
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    loss: Optional[torch.Tensor] = None
    last_hidden_state: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return tuple(v for v in (self.loss, self.logits, self.past_key_values, self.hidden_states) if v is not None)[i]


def project(module: nn.Module, x: torch.Tensor, act: str = "none", residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(x @ W.T + b) (+ residual)`` for an ``nn.Linear`` — on CUDA/bf16 this is the tcgen05 GEMM with the bias /
    activation / residual fused into its epilogue; adapter-wrapped projections (LoRA) keep their own forward."""
    if type(module) is nn.Linear and x.is_cuda and x.dtype == torch.bfloat16:
        from trlx_b200 import ops

        return ops.linear(x, module.weight, module.bias, act, residual)
    y = module(x)
    if act not in ("none", "", None):
        y = activation_fn(act)(y)
    return y if residual is None else y + residual


def activation_fn(name: str):
    if name in ("gelu_new", "gelu_pytorch_tanh", "gelu_fast", "gelu_tanh"):
        return lambda x: F.gelu(x, approximate="tanh")
    if name == "gelu":
        return F.gelu
    if name == "relu":
        return F.relu
    if name in ("silu", "swish"):
        return F.silu
    raise ValueError(f"unknown activation {name}")


class Norm(nn.Module):
    def __init__(self, spec: ArchSpec, dtype=None):
        super().__init__()
        self.kind, self.eps = spec.norm, spec.norm_eps
        self.weight = nn.Parameter(torch.ones(spec.hidden_size, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(spec.hidden_size, dtype=dtype)) if spec.norm == "layernorm" else None

    #: optional ``() -> float | None`` installed by sequence parallelism: factor applied to the *gradients* of weight / bias
    #: (forward values unchanged) for forwards that run with replicated activations, so that the later sum over the
    #: tensor-parallel group yields the true gradient (parallel/tensor_parallel.py)
    _grad_scale = None

    def _params(self):
        factor = self._grad_scale() if self._grad_scale is not None else None
        if factor is None or factor == 1.0:
            return self.weight, self.bias
        scale = lambda p: None if p is None else p.detach() + (p - p.detach()) * factor  # noqa: E731
        return scale(self.weight), scale(self.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        weight, bias = self._params()
        if x.is_cuda and x.dtype == torch.bfloat16:
            from trlx_b200 import ops

            if ops.norm_ok(x, weight, bias):  # forward keeps (mean, rstd); one-pass backward (csrc/norm_train.cu)
                return ops.layer_norm(x, weight, bias, self.eps, self.kind != "layernorm")
        if self.kind == "layernorm":
            return F.layer_norm(x, (x.shape[-1],), weight, bias, self.eps)
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        return (xf.to(x.dtype)) * weight


def alibi_slopes(n_heads: int) -> torch.Tensor:
    """Bloom / ALiBi head slopes (power-of-two base with interleaved extras)."""
    def pow2(n):
        start = 2.0 ** (-(2.0 ** -(math.log2(n) - 3)))
        return [start * (start ** i) for i in range(n)]

    if math.log2(n_heads).is_integer():
        s = pow2(n_heads)
    else:
        closest = 2 ** math.floor(math.log2(n_heads))
        s = pow2(closest) + pow2(2 * closest)[0::2][: n_heads - closest]
    return torch.tensor(s, dtype=torch.float32)


def rotary_tables(spec: ArchSpec, position_ids: torch.Tensor, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos / sin of shape ``[B, T, rotary_dim/2]`` (fp32)."""
    half = spec.rotary_dim // 2
    inv = 1.0 / (spec.rotary_base ** (torch.arange(0, half, device=position_ids.device, dtype=torch.float32) / half))
    ang = position_ids.float().unsqueeze(-1) * inv
    return ang.cos(), ang.sin()


def apply_rotary(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, rot: int, interleaved: bool) -> torch.Tensor:
    """Rotate the first ``rot`` features of ``x`` ``[B, h, T, d]``; cos/sin ``[B, T, rot/2]``."""
    xr, xp = x[..., :rot].float(), x[..., rot:]
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    if interleaved:
        x1, x2 = xr[..., 0::2], xr[..., 1::2]
        out = torch.stack((x1 * c - x2 * s, x2 * c + x1 * s), dim=-1).flatten(-2)
    else:
        x1, x2 = xr[..., : rot // 2], xr[..., rot // 2:]
        out = torch.cat((x1 * c - x2 * s, x2 * c + x1 * s), dim=-1)
    return torch.cat((out.to(x.dtype), xp), dim=-1)


@dataclass
class AttnContext:
    """Per-forward tensors shared by every block."""

    bias: Optional[torch.Tensor]  # additive mask [B, 1|h, Tq, Tk] (fp32/-inf) or None for pure causal
    cos: Optional[torch.Tensor] = None
    sin: Optional[torch.Tensor] = None
    local_bias: Optional[torch.Tensor] = None  # GPT-Neo local-window variant of ``bias``
    causal_only: bool = False


class Attention(nn.Module):
    def __init__(self, spec: ArchSpec, layer_idx: int, dtype=None):
        super().__init__()
        self.spec, self.layer_idx = spec, layer_idx
        self.qkv = nn.Linear(spec.hidden_size, spec.q_size + 2 * spec.kv_size, bias=spec.qkv_bias, dtype=dtype)
        self.out = nn.Linear(spec.q_size, spec.hidden_size, bias=spec.attn_out_bias, dtype=dtype)
        self.scale = spec.attn_scale if spec.attn_scale is not None else 1.0 / math.sqrt(spec.head_dim)
        self.is_local = layer_idx in spec.local_layers

    def forward(self, x, ctx: AttnContext, past=None, use_cache=False):
        s = self.spec
        B, T, _ = x.shape
        qkv = project(self.qkv, x)
        q, k, v = qkv.split([s.q_size, s.kv_size, s.kv_size], dim=-1)
        q = q.view(B, T, s.num_heads, s.head_dim).transpose(1, 2)
        k = k.view(B, T, s.num_kv_heads, s.head_dim).transpose(1, 2)
        v = v.view(B, T, s.num_kv_heads, s.head_dim).transpose(1, 2)
        if s.pos == "rotary":
            q = apply_rotary(q, ctx.cos, ctx.sin, s.rotary_dim, s.rotary_interleaved)
            k = apply_rotary(k, ctx.cos, ctx.sin, s.rotary_dim, s.rotary_interleaved)
        if past is not None:
            k = torch.cat([past[0], k], dim=2)
            v = torch.cat([past[1], v], dim=2)
        present = (k, v) if use_cache else None
        if s.num_kv_heads != s.num_heads:
            rep = s.num_heads // s.num_kv_heads
            k = k.repeat_interleave(rep, dim=1)
            v = v.repeat_interleave(rep, dim=1)
        bias = ctx.local_bias if self.is_local else ctx.bias
        # short sequences: one-CTA-per-(batch, head) kernels (csrc/attention.cu); otherwise the library SDPA
        from trlx_b200 import ops

        o = ops.attention(q, k, v, bias, causal=(bias is None and T > 1), scale=self.scale)
        o = o.transpose(1, 2).reshape(B, T, s.q_size)
        return project(self.out, o), present


class MLP(nn.Module):
    def __init__(self, spec: ArchSpec, dtype=None):
        super().__init__()
        self.gated = spec.gated_mlp
        self.up = nn.Linear(spec.hidden_size, spec.ffn_size * (2 if spec.gated_mlp else 1), bias=spec.mlp_bias, dtype=dtype)
        self.down = nn.Linear(spec.ffn_size, spec.hidden_size, bias=spec.mlp_bias, dtype=dtype)
        self.act = activation_fn(spec.activation)
        self.act_name = spec.activation

    def forward(self, x):
        if self.gated:
            g, u = project(self.up, x).chunk(2, dim=-1)
            h = self.act(g) * u
        else:
            h = project(self.up, x, self.act_name)
        return project(self.down, h)


class Block(nn.Module):
    def __init__(self, spec: ArchSpec, layer_idx: int, dtype=None):
        super().__init__()
        self.spec = spec
        self.norm1 = Norm(spec, dtype)
        self.attn = Attention(spec, layer_idx, dtype)
        self.norm2 = None if spec.shared_parallel_norm else Norm(spec, dtype)
        self.mlp = MLP(spec, dtype)

    def forward(self, x, ctx: AttnContext, past=None, use_cache=False):
        if self.spec.post_norm:  # OPT-350m: attention / MLP read the raw stream, the norm follows the residual add
            a, present = self.attn(x, ctx, past, use_cache)
            x = self.norm1(x + a)
            return self.norm2(x + self.mlp(x)), present
        if self.spec.parallel_residual:
            n1 = self.norm1(x)
            a, present = self.attn(n1, ctx, past, use_cache)
            m = self.mlp(n1 if self.norm2 is None else self.norm2(x))
            return x + a + m, present
        a, present = self.attn(self.norm1(x), ctx, past, use_cache)
        x = x + a
        return x + self.mlp(self.norm2(x)), present


def build_attn_context(spec: ArchSpec, attention_mask: Optional[torch.Tensor], position_ids: torch.Tensor,
                       q_len: int, past_len: int, dtype, device) -> AttnContext:
    """Additive attention bias (causal ∧ padding [+ alibi] [+ local window]) and rotary tables."""
    k_len = past_len + q_len
    cos = sin = None
    if spec.pos == "rotary":
        cos, sin = rotary_tables(spec, position_ids, dtype)
    need_bias = attention_mask is not None or spec.pos == "alibi" or bool(spec.local_layers)
    if not need_bias:
        if q_len == 1:
            return AttnContext(bias=torch.zeros(1, 1, 1, k_len, device=device), cos=cos, sin=sin)
        if past_len == 0:
            return AttnContext(bias=None, cos=cos, sin=sin, causal_only=True)
    qi = torch.arange(past_len, k_len, device=device).view(1, 1, q_len, 1)
    ki = torch.arange(k_len, device=device).view(1, 1, 1, k_len)
    allowed = ki <= qi
    if attention_mask is not None:
        allowed = allowed & attention_mask.bool().view(attention_mask.shape[0], 1, 1, k_len)
    neg = torch.finfo(torch.float32).min
    bias = torch.zeros(allowed.shape, dtype=torch.float32, device=device).masked_fill(~allowed, neg)
    local_bias = None
    if spec.local_layers:
        local_ok = allowed & (ki > qi - spec.local_window)
        local_bias = torch.zeros(local_ok.shape, dtype=torch.float32, device=device).masked_fill(~local_ok, neg)
    if spec.pos == "alibi":
        slopes = alibi_slopes(spec.num_heads).to(device).view(1, spec.num_heads, 1, 1)
        if attention_mask is not None:
            kpos = ((attention_mask.long().cumsum(-1) - 1) * attention_mask.long()).view(-1, 1, 1, k_len).float()
        else:
            kpos = ki.float()
        bias = bias + slopes * kpos
    return AttnContext(bias=bias, cos=cos, sin=sin, local_bias=local_bias)


class Tail(nn.Module):
    """What follows the last block when it is not a plain norm: optional final norm, then ``project_out`` down to the
    word-embedding width (OPT-350m, HF ``OPTDecoder.final_layer_norm`` / ``project_out``)."""

    def __init__(self, spec: ArchSpec, dtype=None):
        super().__init__()
        self.norm = Norm(spec, dtype) if spec.final_norm else None
        E = spec.word_embed_dim
        self.project_out = nn.Linear(spec.hidden_size, E, bias=False, dtype=dtype) if E != spec.hidden_size else None

    def forward(self, x):
        if self.norm is not None:
            x = self.norm(x)
        return x if self.project_out is None else self.project_out(x)


class Trunk(nn.Module):
    """Embeddings + blocks + final norm (attribute names chosen so the generic ``hf_get_*`` getters find
    them: ``transformer.h`` / ``transformer.ln_f``)."""

    def __init__(self, spec: ArchSpec, dtype=None):
        super().__init__()
        self.spec = spec
        E = spec.word_embed_dim
        self.wte = nn.Embedding(spec.vocab_size, E, dtype=dtype)
        self.project_in = nn.Linear(E, spec.hidden_size, bias=False, dtype=dtype) if E != spec.hidden_size else None
        self.wpe = (nn.Embedding(spec.max_positions + spec.pos_offset, spec.hidden_size, dtype=dtype)
                    if spec.pos == "learned" else None)
        self.emb_norm = None
        if spec.embed_norm:
            self.emb_norm = Norm(spec, dtype)
        self.h = nn.ModuleList([Block(spec, i, dtype) for i in range(spec.num_layers)])
        self.ln_f = Norm(spec, dtype) if spec.plain_tail else Tail(spec, dtype)

    def embed(self, input_ids, position_ids, inputs_embeds=None):
        x = self.wte(input_ids) if inputs_embeds is None else inputs_embeds
        if self.project_in is not None:
            x = self.project_in(x)
        if self.wpe is not None:
            x = x + self.wpe(position_ids + self.spec.pos_offset)
        if self.emb_norm is not None:
            x = self.emb_norm(x)
        return x


class CausalLM(nn.Module):
    """Decoder-only LM.  ``forward`` accepts the HF-style keyword set the trainers use, plus
    ``hidden_in`` / ``start_layer`` / ``stop_layer`` to run a slice of the stack on a cached activation."""

    def __init__(self, spec: ArchSpec, dtype=None):
        super().__init__()
        self.config = spec
        self.transformer = Trunk(spec, dtype)
        self.lm_head = nn.Linear(spec.word_embed_dim, spec.vocab_size, bias=spec.lm_head_bias, dtype=dtype)
        self.reset_parameters()
        if spec.tie_word_embeddings:
            self.lm_head.weight = self.transformer.wte.weight

    # -- init -----------------------------------------------------------------------------------------
    def reset_parameters(self):
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, mean=0.0, std=std)

    gradient_checkpointing = False

    def gradient_checkpointing_enable(self, *_, **__):
        self.gradient_checkpointing = True

    def gradient_checkpointing_disable(self):
        self.gradient_checkpointing = False

    # -- HF-ish accessors --------------------------------------------------------------------------------
    def get_input_embeddings(self):
        return self.transformer.wte

    def get_output_embeddings(self):
        return self.lm_head

    @property
    def device(self):
        return self.transformer.wte.weight.device

    @property
    def dtype(self):
        return self.transformer.wte.weight.dtype

    # -- forward -------------------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, use_cache=False, output_hidden_states=False, return_dict=True, labels=None,
                hidden_in=None, start_layer: int = 0, stop_layer: Optional[int] = None, compute_logits: bool = True,
                **_ignored):
        pp = getattr(self, "_pp", None)
        if pp is not None:  # this LM holds one pipeline stage (parallel/pipeline_parallel.py)
            return pp.forward(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                              past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                              output_hidden_states=output_hidden_states, labels=labels, compute_logits=compute_logits,
                              hidden_in=hidden_in, start_layer=start_layer, stop_layer=stop_layer)
        spec, trunk = self.config, self.transformer
        ref = hidden_in if hidden_in is not None else (inputs_embeds if inputs_embeds is not None else input_ids)
        B, T = ref.shape[0], ref.shape[1]
        device = ref.device
        past_len = past_key_values[0][0].shape[2] if past_key_values else 0
        if hidden_in is not None:  # a (possibly sequence-sharded) cached activation: the true length comes from the mask
            if attention_mask is not None:
                T = attention_mask.shape[1] - past_len
            elif position_ids is not None:
                T = position_ids.shape[1]
        if position_ids is None:
            if attention_mask is not None:
                position_ids = (attention_mask.long().cumsum(-1) - 1).clamp_min(0)[:, -T:]
            else:
                position_ids = torch.arange(past_len, past_len + T, device=device).unsqueeze(0).expand(B, T)
        if attention_mask is not None and attention_mask.shape[1] != past_len + T:
            raise ValueError(f"attention_mask covers {attention_mask.shape[1]} keys, expected {past_len + T}")

        if hidden_in is not None:
            x = hidden_in
        else:
            x = trunk.embed(input_ids, position_ids, inputs_embeds)

        ctx = build_attn_context(spec, attention_mask, position_ids, T, past_len, x.dtype, device)
        stop = len(trunk.h) if stop_layer is None else stop_layer
        hiddens = [] if output_hidden_states else None
        presents = [] if use_cache else None
        recompute = self.gradient_checkpointing and torch.is_grad_enabled() and not use_cache
        for i in range(start_layer, stop):
            if hiddens is not None:
                hiddens.append(x)
            past = past_key_values[i - start_layer] if past_key_values else None
            if recompute and x.requires_grad:
                # activation checkpointing (reference: NeMo `activations_checkpoint_granularity`, megatron_20b.yaml:77):
                # keep only the block input, recompute the block in backward
                from torch.utils.checkpoint import checkpoint

                x, present = checkpoint(trunk.h[i], x, ctx, None, False, use_reentrant=False)
            else:
                x, present = trunk.h[i](x, ctx, past, use_cache)
            if presents is not None:
                presents.append(present)
        final = stop == len(trunk.h)
        if final:
            x = trunk.ln_f(x)
        if hiddens is not None:
            hiddens.append(x)
        logits = project(self.lm_head, x) if (final and compute_logits) else None
        loss = None
        if labels is not None and logits is not None:
            loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1),
                                   ignore_index=-100)
        return CausalLMOutput(logits=logits, past_key_values=presents,
                              hidden_states=tuple(hiddens) if hiddens is not None else None, loss=loss,
                              last_hidden_state=x)
