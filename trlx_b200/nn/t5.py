"""T5 encoder-decoder (t5 / t5-v1.1 / flan-t5 shapes), the framework's own implementation.

Module and parameter names follow the HuggingFace checkpoint layout (``shared``, ``encoder.block.N.layer.0.SelfAttention.q``
…) so state dicts interchange without renaming, and so the reference's seq2seq freezing rule
(``trlx/utils/modeling.py:41-60``) and ``T5Branch`` (``trlx/models/modeling_ppo.py:1483-1592``) have direct equivalents.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class T5Spec:
    vocab_size: int = 32128
    d_model: int = 512
    d_kv: int = 64
    d_ff: int = 2048
    num_layers: int = 6
    num_decoder_layers: int = 6
    num_heads: int = 8
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6
    feed_forward_proj: str = "relu"  # "relu" | "gated-gelu"
    tie_word_embeddings: bool = True
    scale_decoder_outputs: Optional[bool] = None  # None → follow tie_word_embeddings (classic T5)
    pad_token_id: int = 0
    eos_token_id: int = 1
    decoder_start_token_id: int = 0
    initializer_factor: float = 1.0
    model_type: str = "t5"
    is_encoder_decoder: bool = True
    extra: Dict[str, Any] = field(default_factory=dict)

    @property
    def hidden_size(self) -> int:
        return self.d_model

    @property
    def num_hidden_layers(self) -> int:
        return self.num_layers

    @property
    def scale_logits(self) -> bool:
        return self.tie_word_embeddings if self.scale_decoder_outputs is None else bool(self.scale_decoder_outputs)

    @property
    def is_gated_act(self) -> bool:
        return self.feed_forward_proj.startswith("gated")

    @property
    def bos_token_id(self):
        return None

    def to_dict(self):
        d = asdict(self)
        d.update(d.pop("extra"))
        return d


def t5_spec_from_config(cfg: Dict[str, Any]) -> T5Spec:
    if isinstance(cfg, T5Spec):
        return cfg
    if not isinstance(cfg, dict):
        cfg = cfg.to_dict()
    known = {f for f in T5Spec.__dataclass_fields__ if f != "extra"}
    kw = {k: v for k, v in cfg.items() if k in known and v is not None}
    kw["model_type"] = "t5"
    kw["is_encoder_decoder"] = True
    if "num_decoder_layers" not in kw or kw["num_decoder_layers"] is None:
        kw["num_decoder_layers"] = kw.get("num_layers", 6)
    if kw.get("decoder_start_token_id") is None:
        kw["decoder_start_token_id"] = kw.get("pad_token_id", 0)
    return T5Spec(**kw)


T5_PRESETS: Dict[str, Dict[str, Any]] = {
    "t5-small": dict(model_type="t5", d_model=512, d_kv=64, d_ff=2048, num_layers=6, num_heads=8),
    "t5-base": dict(model_type="t5", d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12),
    "t5-large": dict(model_type="t5", d_model=1024, d_kv=64, d_ff=4096, num_layers=24, num_heads=16),
    "flan-t5-small": dict(model_type="t5", d_model=512, d_kv=64, d_ff=1024, num_layers=8, num_heads=6,
                          feed_forward_proj="gated-gelu", tie_word_embeddings=False),
    "flan-t5-base": dict(model_type="t5", d_model=768, d_kv=64, d_ff=2048, num_layers=12, num_heads=12,
                         feed_forward_proj="gated-gelu", tie_word_embeddings=False),
    "flan-t5-large": dict(model_type="t5", d_model=1024, d_kv=64, d_ff=2816, num_layers=24, num_heads=16,
                          feed_forward_proj="gated-gelu", tie_word_embeddings=False),
    "flan-t5-xxl": dict(model_type="t5", d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                        feed_forward_proj="gated-gelu", tie_word_embeddings=False),
    "t5-efficient-tiny": dict(model_type="t5", d_model=256, d_kv=64, d_ff=1024, num_layers=4, num_heads=4),
}


@dataclass
class Seq2SeqOutput:
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    decoder_hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    encoder_last_hidden_state: Optional[torch.Tensor] = None
    encoder_hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    loss: Optional[torch.Tensor] = None
    last_hidden_state: Optional[torch.Tensor] = None


class T5LayerNorm(nn.Module):
    def __init__(self, d: int, eps: float, dtype=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d, dtype=dtype))
        self.variance_epsilon = eps

    def forward(self, x):
        var = x.float().pow(2).mean(-1, keepdim=True)
        return self.weight * (x.float() * torch.rsqrt(var + self.variance_epsilon)).to(x.dtype)


def relative_position_bucket(rel: torch.Tensor, bidirectional: bool, num_buckets: int, max_distance: int) -> torch.Tensor:
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).long() * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float().clamp_min(1) / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(is_small, rel, large)


def position_bias(table: nn.Embedding, spec: T5Spec, q_len: int, k_len: int, bidirectional: bool, device,
                  q_offset: int = 0) -> torch.Tensor:
    ctx = torch.arange(q_offset, q_offset + q_len, device=device)[:, None]
    mem = torch.arange(k_len, device=device)[None, :]
    buckets = relative_position_bucket(mem - ctx, bidirectional, spec.relative_attention_num_buckets,
                                       spec.relative_attention_max_distance)
    return table(buckets).permute(2, 0, 1).unsqueeze(0)  # [1, h, q, k]


class T5Attention(nn.Module):
    def __init__(self, spec: T5Spec, has_relative_bias: bool, dtype=None):
        super().__init__()
        inner = spec.num_heads * spec.d_kv
        self.spec = spec
        self.q = nn.Linear(spec.d_model, inner, bias=False, dtype=dtype)
        self.k = nn.Linear(spec.d_model, inner, bias=False, dtype=dtype)
        self.v = nn.Linear(spec.d_model, inner, bias=False, dtype=dtype)
        self.o = nn.Linear(inner, spec.d_model, bias=False, dtype=dtype)
        self.relative_attention_bias = (nn.Embedding(spec.relative_attention_num_buckets, spec.num_heads, dtype=dtype)
                                        if has_relative_bias else None)

    def forward(self, x, kv_source=None, bias=None, past=None, use_cache=False, static_kv=False):
        s = self.spec
        B, T, _ = x.shape

        def heads(t):
            return t.view(B, -1, s.num_heads, s.d_kv).transpose(1, 2)

        q = heads(self.q(x))
        if static_kv and past is not None:
            k, v = past
        else:
            src = x if kv_source is None else kv_source
            k, v = heads(self.k(src)), heads(self.v(src))
            if past is not None and not static_kv:
                k, v = torch.cat([past[0], k], 2), torch.cat([past[1], v], 2)
        present = (k, v) if use_cache else None
        scores = torch.matmul(q, k.transpose(-1, -2)).float()  # T5 does not scale by 1/sqrt(d)
        if bias is not None:
            scores = scores + bias
        attn = torch.softmax(scores, -1).to(q.dtype)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(B, T, s.num_heads * s.d_kv)
        return self.o(out), present


class T5DenseReluDense(nn.Module):
    def __init__(self, spec: T5Spec, dtype=None):
        super().__init__()
        self.gated = spec.is_gated_act
        if self.gated:
            self.wi_0 = nn.Linear(spec.d_model, spec.d_ff, bias=False, dtype=dtype)
            self.wi_1 = nn.Linear(spec.d_model, spec.d_ff, bias=False, dtype=dtype)
        else:
            self.wi = nn.Linear(spec.d_model, spec.d_ff, bias=False, dtype=dtype)
        self.wo = nn.Linear(spec.d_ff, spec.d_model, bias=False, dtype=dtype)

    def forward(self, x):
        if self.gated:
            return self.wo(F.gelu(self.wi_0(x), approximate="tanh") * self.wi_1(x))
        return self.wo(F.relu(self.wi(x)))


class _SelfAttnLayer(nn.Module):
    def __init__(self, spec, has_bias, dtype=None):
        super().__init__()
        self.SelfAttention = T5Attention(spec, has_bias, dtype)
        self.layer_norm = T5LayerNorm(spec.d_model, spec.layer_norm_epsilon, dtype)


class _CrossAttnLayer(nn.Module):
    def __init__(self, spec, dtype=None):
        super().__init__()
        self.EncDecAttention = T5Attention(spec, False, dtype)
        self.layer_norm = T5LayerNorm(spec.d_model, spec.layer_norm_epsilon, dtype)


class _FFLayer(nn.Module):
    def __init__(self, spec, dtype=None):
        super().__init__()
        self.DenseReluDense = T5DenseReluDense(spec, dtype)
        self.layer_norm = T5LayerNorm(spec.d_model, spec.layer_norm_epsilon, dtype)


class T5Block(nn.Module):
    def __init__(self, spec: T5Spec, is_decoder: bool, has_relative_bias: bool, dtype=None):
        super().__init__()
        self.is_decoder = is_decoder
        layers = [_SelfAttnLayer(spec, has_relative_bias, dtype)]
        if is_decoder:
            layers.append(_CrossAttnLayer(spec, dtype))
        layers.append(_FFLayer(spec, dtype))
        self.layer = nn.ModuleList(layers)

    @property
    def self_attn(self) -> T5Attention:
        return self.layer[0].SelfAttention

    def forward(self, x, self_bias, enc=None, cross_bias=None, past=None, use_cache=False):
        sa = self.layer[0]
        self_past = past[:2] if past is not None else None
        a, self_present = sa.SelfAttention(sa.layer_norm(x), bias=self_bias, past=self_past, use_cache=use_cache)
        x = x + a
        present = self_present
        if self.is_decoder:
            ca = self.layer[1]
            cross_past = past[2:] if past is not None else None
            c, cross_present = ca.EncDecAttention(ca.layer_norm(x), kv_source=enc, bias=cross_bias, past=cross_past,
                                                  use_cache=use_cache, static_kv=cross_past is not None)
            x = x + c
            if use_cache:
                present = self_present + cross_present
        ff = self.layer[-1]
        return x + ff.DenseReluDense(ff.layer_norm(x)), present


class T5Stack(nn.Module):
    def __init__(self, spec: T5Spec, is_decoder: bool, embed: nn.Embedding, dtype=None):
        super().__init__()
        self.spec, self.is_decoder = spec, is_decoder
        self.embed_tokens = embed
        n = spec.num_decoder_layers if is_decoder else spec.num_layers
        self.block = nn.ModuleList(T5Block(spec, is_decoder, i == 0, dtype) for i in range(n))
        self.final_layer_norm = T5LayerNorm(spec.d_model, spec.layer_norm_epsilon, dtype)


def _mask_bias(mask: Optional[torch.Tensor], dtype=torch.float32) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    return (1.0 - mask[:, None, None, :].to(dtype)) * torch.finfo(dtype).min


def run_decoder_blocks(spec: T5Spec, blocks, final_norm, rel_bias: nn.Embedding, hidden, decoder_attention_mask,
                       enc_hidden, enc_mask, start_index: int = 0):
    """Run a slice of decoder blocks without cache (used by :class:`T5Branch` and the full decoder)."""
    B, T, _ = hidden.shape
    dev = hidden.device
    causal = torch.tril(torch.ones(T, T, device=dev, dtype=torch.bool))[None, None]
    self_bias = position_bias(rel_bias, spec, T, T, False, dev).float()
    self_bias = self_bias.masked_fill(~causal, torch.finfo(torch.float32).min)
    if decoder_attention_mask is not None:
        self_bias = self_bias + _mask_bias(decoder_attention_mask)
    cross_bias = _mask_bias(enc_mask)
    hiddens = []
    x = hidden
    for blk in blocks:
        hiddens.append(x)
        x, _ = blk(x, self_bias, enc_hidden, cross_bias)
    x = final_norm(x)
    hiddens.append(x)
    return x, tuple(hiddens)


class T5Model(nn.Module):
    """``T5ForConditionalGeneration`` equivalent."""

    def __init__(self, spec: T5Spec, dtype=None):
        super().__init__()
        self.config = spec
        self.shared = nn.Embedding(spec.vocab_size, spec.d_model, dtype=dtype)
        self.encoder = T5Stack(spec, False, self.shared, dtype)
        self.decoder = T5Stack(spec, True, self.shared, dtype)
        self.lm_head = nn.Linear(spec.d_model, spec.vocab_size, bias=False, dtype=dtype)
        self.reset_parameters()
        if spec.tie_word_embeddings:
            self.lm_head.weight = self.shared.weight

    def reset_parameters(self):
        f = self.config.initializer_factor
        d = self.config.d_model
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0.0, f * (m.in_features ** -0.5))
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, 0.0, f * (1.0 if m is self.shared else d ** -0.5))

    def get_input_embeddings(self):
        return self.shared

    def get_output_embeddings(self):
        return self.lm_head

    @property
    def device(self):
        return self.shared.weight.device

    @property
    def dtype(self):
        return self.shared.weight.dtype

    # HF checkpoints carry the tied embedding under three names
    def to_hf_state_dict(self, sd):
        out = dict(sd)
        out.setdefault("encoder.embed_tokens.weight", sd["shared.weight"])
        out.setdefault("decoder.embed_tokens.weight", sd["shared.weight"])
        return out

    def from_hf_state_dict(self, sd):
        own = self.state_dict()
        return {k: v for k, v in sd.items() if k in own}

    def _shift_right(self, labels):
        start = self.config.decoder_start_token_id
        shifted = labels.new_zeros(labels.shape)
        shifted[:, 1:] = labels[:, :-1].clone()
        shifted[:, 0] = start
        return shifted.masked_fill(shifted == -100, self.config.pad_token_id)

    def encode(self, input_ids=None, attention_mask=None, inputs_embeds=None, output_hidden_states=False):
        spec = self.config
        x = inputs_embeds if inputs_embeds is not None else self.shared(input_ids)
        T = x.shape[1]
        bias = position_bias(self.encoder.block[0].self_attn.relative_attention_bias, spec, T, T, True, x.device).float()
        if attention_mask is not None:
            bias = bias + _mask_bias(attention_mask)
        hiddens = []
        for blk in self.encoder.block:
            hiddens.append(x)
            x, _ = blk(x, bias)
        x = self.encoder.final_layer_norm(x)
        hiddens.append(x)
        if output_hidden_states:
            return x, tuple(hiddens)
        return x

    def decode(self, decoder_input_ids=None, encoder_hidden_states=None, attention_mask=None, decoder_attention_mask=None,
               past_key_values=None, use_cache=False, output_hidden_states=False, decoder_inputs_embeds=None,
               prefix_kv=None):
        spec = self.config
        x = decoder_inputs_embeds if decoder_inputs_embeds is not None else self.shared(decoder_input_ids)
        B, T, _ = x.shape
        dev = x.device
        past_len = past_key_values[0][0].shape[2] if past_key_values else 0
        K = past_len + T
        rel = self.decoder.block[0].self_attn.relative_attention_bias
        self_bias = position_bias(rel, spec, T, K, False, dev, q_offset=past_len).float()
        qi = torch.arange(past_len, K, device=dev)[:, None]
        ki = torch.arange(K, device=dev)[None, :]
        self_bias = self_bias.masked_fill(~(ki <= qi)[None, None], torch.finfo(torch.float32).min)
        if decoder_attention_mask is not None and decoder_attention_mask.shape[1] == K:
            self_bias = self_bias + _mask_bias(decoder_attention_mask)
        cross_bias = _mask_bias(attention_mask)
        hiddens, presents = [], []
        for i, blk in enumerate(self.decoder.block):
            hiddens.append(x)
            past = past_key_values[i] if past_key_values else None
            x, present = blk(x, self_bias, encoder_hidden_states, cross_bias, past, use_cache)
            presents.append(present)
        x = self.decoder.final_layer_norm(x)
        hiddens.append(x)
        h = x * (spec.d_model ** -0.5) if spec.scale_logits else x
        return Seq2SeqOutput(logits=self.lm_head(h), past_key_values=presents if use_cache else None,
                             decoder_hidden_states=tuple(hiddens) if output_hidden_states else None, last_hidden_state=x)

    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, decoder_attention_mask=None,
                encoder_outputs=None, past_key_values=None, use_cache=False, output_hidden_states=False, labels=None,
                inputs_embeds=None, decoder_inputs_embeds=None, return_dict=True, **_):
        if encoder_outputs is None:
            enc, enc_hiddens = self.encode(input_ids, attention_mask, inputs_embeds, output_hidden_states=True)
        else:
            enc = encoder_outputs[0] if isinstance(encoder_outputs, (tuple, list)) else getattr(encoder_outputs, "last_hidden_state", encoder_outputs)
            enc_hiddens = None
        if decoder_input_ids is None and decoder_inputs_embeds is None and labels is not None:
            decoder_input_ids = self._shift_right(labels)
        out = self.decode(decoder_input_ids, enc, attention_mask, decoder_attention_mask, past_key_values, use_cache,
                          output_hidden_states, decoder_inputs_embeds)
        out.encoder_last_hidden_state = enc
        out.encoder_hidden_states = enc_hiddens if output_hidden_states else None
        if labels is not None:
            out.loss = F.cross_entropy(out.logits.reshape(-1, out.logits.shape[-1]).float(), labels.reshape(-1),
                                       ignore_index=-100)
        return out

    def forward_with_prompt(self, peft_model, input_ids=None, attention_mask=None, **kwargs):
        """Prompt / prefix tuning for the encoder side: virtual tokens are prepended to the encoder input."""
        n = peft_model.peft_config.num_virtual_tokens
        emb = self.shared(input_ids)
        B = emb.shape[0]
        if peft_model.peft_type == "PROMPT_TUNING":
            prompt = peft_model.prompt_embeddings.to(emb.dtype).unsqueeze(0).expand(B, -1, -1)
        else:  # prefix tuning on seq2seq: realised as learned encoder-side virtual embeddings of width d_model
            width = self.config.d_model
            prompt = peft_model.prompt_embeddings[:, :width].to(emb.dtype).unsqueeze(0).expand(B, -1, -1)
        emb = torch.cat([prompt, emb], 1)
        if attention_mask is None:
            attention_mask = torch.ones(B, input_ids.shape[1], dtype=torch.long, device=emb.device)
        attention_mask = torch.cat([torch.ones(B, n, dtype=attention_mask.dtype, device=emb.device), attention_mask], 1)
        kwargs.pop("inputs_embeds", None)
        return self.forward(attention_mask=attention_mask, inputs_embeds=emb, **kwargs)

    def generate(self, *args, **kwargs):
        from trlx_b200.models.generation import generate

        return generate(self, *args, **kwargs)


def inject_lora_t5(model: T5Model, cfg, lora_cls) -> None:
    """LoRA on T5 attention / FF projections addressed by their HF leaf names (default ``q`` and ``v``)."""
    targets = set(cfg.target_modules or ["q", "v"])
    found = False
    for mod_name, mod in list(model.named_modules()):
        for leaf, child in list(mod.named_children()):
            if leaf in targets and isinstance(child, nn.Linear):
                wrapped = lora_cls(child, cfg.r, cfg.lora_alpha, cfg.lora_dropout)
                path = f"{mod_name}.{leaf}" if mod_name else leaf
                wrapped.add_adapter(leaf, (0, -1), path)
                setattr(mod, leaf, wrapped)
                found = True
    if not found:
        raise ValueError(f"Target modules {sorted(targets)} not found in the base model")
