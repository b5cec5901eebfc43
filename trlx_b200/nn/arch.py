"""Architecture description shared by every decoder-only family.

One :class:`ArchSpec` drives one generic transformer implementation
(:mod:`trlx_b200.nn.transformer`); the per-family differences the reference inherits from eight
separate HF modeling files (GPT-2 / GPT-Neo / GPT-J / GPT-NeoX / OPT / Bloom / Llama / GPT-BigCode;
``trlx/models/modeling_ppo.py:1598-1637`` lists the supported set) reduce to the fields below.
"""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional


@dataclass
class ArchSpec:
    family: str = "gpt2"
    vocab_size: int = 50257
    hidden_size: int = 768
    num_layers: int = 12
    num_heads: int = 12
    num_kv_heads: int = 12
    head_dim: int = 64
    ffn_size: int = 3072
    max_positions: int = 1024
    # normalisation
    norm: str = "layernorm"  # "layernorm" | "rmsnorm"
    norm_eps: float = 1e-5
    # positions
    pos: str = "learned"  # "learned" | "rotary" | "alibi" | "none"
    pos_offset: int = 0  # OPT stores positions shifted by 2
    rotary_dim: int = 0
    rotary_base: float = 10000.0
    rotary_interleaved: bool = False  # GPT-J rotates (even, odd) pairs; NeoX/Llama rotate halves
    # block structure
    parallel_residual: bool = False  # x + attn(n1(x)) + mlp(n2(x))
    shared_parallel_norm: bool = False  # GPT-J: a single norm feeds both branches
    activation: str = "gelu_new"  # gelu_new | gelu | relu | silu | gelu_pytorch_tanh
    gated_mlp: bool = False  # SwiGLU-style: down(act(gate(x)) * up(x))
    attn_scale: Optional[float] = None  # None → 1/sqrt(head_dim); GPT-Neo uses 1.0
    local_window: int = 0  # GPT-Neo local attention window
    local_layers: List[int] = field(default_factory=list)
    # biases
    qkv_bias: bool = True
    attn_out_bias: bool = True
    mlp_bias: bool = True
    lm_head_bias: bool = False
    embed_norm: bool = False  # Bloom: LayerNorm right after the token embedding
    embed_dim: int = 0  # OPT-350m: token embeddings / LM head live in a narrower space (project_in / project_out); 0 = hidden_size
    post_norm: bool = False  # OPT-350m (`do_layer_norm_before=False`): norms follow the residual adds
    final_norm: bool = True  # post-norm OPT has no final LayerNorm
    tie_word_embeddings: bool = True
    # token ids
    bos_token_id: Optional[int] = None
    eos_token_id: Optional[int] = None
    pad_token_id: Optional[int] = None
    initializer_range: float = 0.02
    extra: Dict[str, Any] = field(default_factory=dict)

    # HF-config look-alikes so generic helpers (hf_get_hidden_size, …) work on an ArchSpec
    @property
    def n_embd(self) -> int:
        return self.hidden_size

    @property
    def word_embed_dim(self) -> int:
        return self.embed_dim or self.hidden_size

    @property
    def final_hidden_size(self) -> int:
        """Width of the final hidden state (what LM / value / Q heads consume)."""
        return self.word_embed_dim

    @property
    def plain_tail(self) -> bool:
        return self.final_norm and self.word_embed_dim == self.hidden_size

    @property
    def n_layer(self) -> int:
        return self.num_layers

    @property
    def num_hidden_layers(self) -> int:
        return self.num_layers

    @property
    def model_type(self) -> str:
        return self.family

    @property
    def is_encoder_decoder(self) -> bool:
        return False

    @property
    def q_size(self) -> int:
        return self.num_heads * self.head_dim

    @property
    def kv_size(self) -> int:
        return self.num_kv_heads * self.head_dim

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    def n_params(self) -> int:
        H, F, L, V = self.hidden_size, self.ffn_size, self.num_layers, self.vocab_size
        per_layer = H * (self.q_size + 2 * self.kv_size) + self.q_size * H + (3 if self.gated_mlp else 2) * H * F
        total = L * per_layer + V * H + (0 if self.tie_word_embeddings else V * H)
        if self.pos == "learned":
            total += (self.max_positions + self.pos_offset) * H
        return total


def _get(cfg, *names, default=None):
    for n in names:
        if isinstance(cfg, dict):
            if n in cfg and cfg[n] is not None:
                return cfg[n]
        elif getattr(cfg, n, None) is not None:
            return getattr(cfg, n)
    return default


def _rope(cfg, key, default):
    rp = _get(cfg, "rope_parameters", default=None)
    if isinstance(rp, dict) and key in rp and rp[key] is not None:
        return rp[key]
    legacy = {"rope_theta": "rope_theta", "partial_rotary_factor": "rotary_pct"}.get(key, key)
    return _get(cfg, key, legacy, default=default)


def spec_from_hf_config(cfg) -> ArchSpec:
    """Translate a ``transformers`` config object (or its dict) into an :class:`ArchSpec`."""
    if isinstance(cfg, ArchSpec):
        return cfg
    mt = _get(cfg, "model_type")
    H = _get(cfg, "hidden_size", "n_embd", "d_model")
    L = _get(cfg, "num_hidden_layers", "n_layer", "num_layers")
    nh = _get(cfg, "num_attention_heads", "n_head", "num_heads")
    common = dict(
        vocab_size=_get(cfg, "vocab_size"), hidden_size=H, num_layers=L, num_heads=nh,
        bos_token_id=_get(cfg, "bos_token_id"), eos_token_id=_get(cfg, "eos_token_id"),
        pad_token_id=_get(cfg, "pad_token_id"), initializer_range=_get(cfg, "initializer_range", "init_std", default=0.02),
    )
    d = H // nh
    if mt == "gpt2":
        return ArchSpec(family="gpt2", num_kv_heads=nh, head_dim=d, ffn_size=_get(cfg, "n_inner", default=4 * H),
                        max_positions=_get(cfg, "n_positions", default=1024), norm_eps=_get(cfg, "layer_norm_epsilon", default=1e-5),
                        activation=_get(cfg, "activation_function", default="gelu_new"),
                        tie_word_embeddings=_get(cfg, "tie_word_embeddings", default=True), **common)
    if mt == "gpt_bigcode":
        mq = _get(cfg, "multi_query", default=True)
        return ArchSpec(family="gpt_bigcode", num_kv_heads=1 if mq else nh, head_dim=d,
                        ffn_size=_get(cfg, "n_inner", default=4 * H), max_positions=_get(cfg, "n_positions", default=1024),
                        norm_eps=_get(cfg, "layer_norm_epsilon", default=1e-5),
                        activation=_get(cfg, "activation_function", default="gelu_pytorch_tanh"),
                        tie_word_embeddings=_get(cfg, "tie_word_embeddings", default=True), **common)
    if mt == "gpt_neo":
        layers = _get(cfg, "attention_layers")
        if layers is None:
            layers = []
            for pattern, repeat in (_get(cfg, "attention_types") or [[["global"], L]]):
                layers += list(pattern) * int(repeat)
            layers = (layers + ["global"] * L)[:L]
        return ArchSpec(family="gpt_neo", num_kv_heads=nh, head_dim=d, ffn_size=_get(cfg, "intermediate_size", default=4 * H),
                        max_positions=_get(cfg, "max_position_embeddings", default=2048),
                        norm_eps=_get(cfg, "layer_norm_epsilon", default=1e-5),
                        activation=_get(cfg, "activation_function", default="gelu_new"), attn_scale=1.0,
                        local_window=_get(cfg, "window_size", default=256),
                        local_layers=[i for i, t in enumerate(layers) if t == "local"], qkv_bias=False,
                        tie_word_embeddings=_get(cfg, "tie_word_embeddings", default=True), **common)
    if mt == "gptj":
        return ArchSpec(family="gptj", num_kv_heads=nh, head_dim=d, ffn_size=_get(cfg, "n_inner", default=4 * H),
                        max_positions=_get(cfg, "n_positions", default=2048), norm_eps=_get(cfg, "layer_norm_epsilon", default=1e-5),
                        pos="rotary", rotary_dim=_get(cfg, "rotary_dim", default=d), rotary_interleaved=True,
                        parallel_residual=True, shared_parallel_norm=True,
                        activation=_get(cfg, "activation_function", default="gelu_new"), qkv_bias=False, attn_out_bias=False,
                        lm_head_bias=True, tie_word_embeddings=False, **common)
    if mt == "gpt_neox":
        pct = _rope(cfg, "partial_rotary_factor", 0.25)
        return ArchSpec(family="gpt_neox", num_kv_heads=nh, head_dim=d, ffn_size=_get(cfg, "intermediate_size", default=4 * H),
                        max_positions=_get(cfg, "max_position_embeddings", default=2048), norm_eps=_get(cfg, "layer_norm_eps", default=1e-5),
                        pos="rotary", rotary_dim=int(d * pct), rotary_base=_rope(cfg, "rope_theta", 10000.0),
                        parallel_residual=_get(cfg, "use_parallel_residual", default=True),
                        activation=_get(cfg, "hidden_act", default="gelu"),
                        tie_word_embeddings=_get(cfg, "tie_word_embeddings", default=False), **common)
    if mt in ("llama", "mistral"):
        hd = _get(cfg, "head_dim", default=d)
        bias = _get(cfg, "attention_bias", default=False)
        return ArchSpec(family="llama", num_kv_heads=_get(cfg, "num_key_value_heads", default=nh), head_dim=hd,
                        ffn_size=_get(cfg, "intermediate_size"), max_positions=_get(cfg, "max_position_embeddings", default=2048),
                        norm="rmsnorm", norm_eps=_get(cfg, "rms_norm_eps", default=1e-6), pos="rotary", rotary_dim=hd,
                        rotary_base=_rope(cfg, "rope_theta", 10000.0), activation=_get(cfg, "hidden_act", default="silu"),
                        gated_mlp=True, qkv_bias=bias, attn_out_bias=bias, mlp_bias=_get(cfg, "mlp_bias", default=False),
                        tie_word_embeddings=_get(cfg, "tie_word_embeddings", default=False), **common)
    if mt == "opt":
        E = _get(cfg, "word_embed_proj_dim", default=H)
        pre = bool(_get(cfg, "do_layer_norm_before", default=True))
        bias = _get(cfg, "enable_bias", default=True)
        return ArchSpec(family="opt", num_kv_heads=nh, head_dim=d, ffn_size=_get(cfg, "ffn_dim", default=4 * H),
                        embed_dim=0 if E == H else E, post_norm=not pre,
                        final_norm=pre and not _get(cfg, "_remove_final_layer_norm", default=False),
                        max_positions=_get(cfg, "max_position_embeddings", default=2048), pos_offset=2,
                        activation=_get(cfg, "activation_function", default="relu"), qkv_bias=bias, attn_out_bias=bias,
                        mlp_bias=bias, tie_word_embeddings=_get(cfg, "tie_word_embeddings", default=True), **common)
    if mt == "bloom":
        return ArchSpec(family="bloom", num_kv_heads=nh, head_dim=d, ffn_size=4 * H, max_positions=1 << 20,
                        norm_eps=_get(cfg, "layer_norm_epsilon", default=1e-5), pos="alibi", activation="gelu_pytorch_tanh",
                        embed_norm=True, tie_word_embeddings=_get(cfg, "tie_word_embeddings", default=True), **common)
    raise NotImplementedError(f"unsupported model_type '{mt}'")


# ---- named presets (random-init; B200 boxes have no hub access) -------------------------------------
_PRESETS: Dict[str, Dict[str, Any]] = {
    "gpt2": dict(model_type="gpt2", vocab_size=50257, n_embd=768, n_layer=12, n_head=12, n_positions=1024,
                 bos_token_id=50256, eos_token_id=50256),
    "gpt2-medium": dict(model_type="gpt2", vocab_size=50257, n_embd=1024, n_layer=24, n_head=16, n_positions=1024,
                        bos_token_id=50256, eos_token_id=50256),
    "gpt2-large": dict(model_type="gpt2", vocab_size=50257, n_embd=1280, n_layer=36, n_head=20, n_positions=1024,
                       bos_token_id=50256, eos_token_id=50256),
    "gpt2-xl": dict(model_type="gpt2", vocab_size=50257, n_embd=1600, n_layer=48, n_head=25, n_positions=1024,
                    bos_token_id=50256, eos_token_id=50256),
    "gpt-j-6b": dict(model_type="gptj", vocab_size=50400, n_embd=4096, n_layer=28, n_head=16, rotary_dim=64,
                     n_positions=2048, bos_token_id=50256, eos_token_id=50256),
    "gpt-neox-20b": dict(model_type="gpt_neox", vocab_size=50432, hidden_size=6144, num_hidden_layers=44,
                         num_attention_heads=64, intermediate_size=24576, max_position_embeddings=2048,
                         bos_token_id=0, eos_token_id=0),
    "pythia-160m": dict(model_type="gpt_neox", vocab_size=50304, hidden_size=768, num_hidden_layers=12,
                        num_attention_heads=12, intermediate_size=3072, max_position_embeddings=2048,
                        bos_token_id=0, eos_token_id=0),
    "llama-2-7b": dict(model_type="llama", vocab_size=32000, hidden_size=4096, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=32, intermediate_size=11008,
                       max_position_embeddings=4096, rms_norm_eps=1e-5, bos_token_id=1, eos_token_id=2),
    "opt-125m": dict(model_type="opt", vocab_size=50272, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                     ffn_dim=3072, max_position_embeddings=2048, bos_token_id=2, eos_token_id=2, pad_token_id=1),
    "opt-350m": dict(model_type="opt", vocab_size=50272, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                     ffn_dim=4096, max_position_embeddings=2048, word_embed_proj_dim=512, do_layer_norm_before=False,
                     bos_token_id=2, eos_token_id=2, pad_token_id=1),
    "opt-6.7b": dict(model_type="opt", vocab_size=50272, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                     ffn_dim=16384, max_position_embeddings=2048, bos_token_id=2, eos_token_id=2, pad_token_id=1),
    "bloom-560m": dict(model_type="bloom", vocab_size=250880, hidden_size=1024, n_layer=24, n_head=16,
                       bos_token_id=1, eos_token_id=2),
    "gpt_bigcode-santacoder": dict(model_type="gpt_bigcode", vocab_size=49280, n_embd=2048, n_layer=24, n_head=16,
                                   n_positions=2048, bos_token_id=49152, eos_token_id=49152),
    "gpt-neo-125m": dict(model_type="gpt_neo", vocab_size=50257, hidden_size=768, num_layers=12, num_heads=12,
                         max_position_embeddings=2048, attention_layers=["global", "local"] * 6,
                         bos_token_id=50256, eos_token_id=50256),
}
_ALIASES = {
    "lvwerra/gpt2-imdb": "gpt2", "gpt2-imdb": "gpt2", "eleutherai/gpt-j-6b": "gpt-j-6b", "gptj": "gpt-j-6b",
    "eleutherai/gpt-neox-20b": "gpt-neox-20b", "eleutherai/pythia-160m": "pythia-160m",
    "meta-llama/llama-2-7b-hf": "llama-2-7b", "nousresearch/llama-2-7b-hf": "llama-2-7b",
    "facebook/opt-125m": "opt-125m", "facebook/opt-350m": "opt-350m", "facebook/opt-6.7b": "opt-6.7b", "bigscience/bloom-560m": "bloom-560m",
    "bigcode/gpt_bigcode-santacoder": "gpt_bigcode-santacoder", "eleutherai/gpt-neo-125m": "gpt-neo-125m",
}


def resolve_config(source) -> Dict[str, Any]:
    """Return a raw HF-style config dict (causal or seq2seq) for ``source``.

    ``source`` may be an :class:`ArchSpec`, a ``transformers`` config, a dict, a directory with
    ``config.json``, a path to a json file or a preset / hub name (resolved offline)."""
    if isinstance(source, ArchSpec):
        return {"model_type": "__spec__", "spec": source}
    if isinstance(source, dict):
        return dict(source)
    if hasattr(source, "to_dict") and hasattr(source, "model_type"):
        return source.to_dict()
    if isinstance(source, str):
        if os.path.isdir(source) and os.path.exists(os.path.join(source, "config.json")):
            with open(os.path.join(source, "config.json")) as fh:
                return json.load(fh)
        if os.path.isfile(source):
            with open(source) as fh:
                return json.load(fh)
        key = source.lower()
        key = _ALIASES.get(key, key)
        if key in _PRESETS:
            return dict(_PRESETS[key])
        from trlx_b200.nn.t5 import T5_PRESETS

        tail = key.split("/")[-1]
        if tail in T5_PRESETS:
            return dict(T5_PRESETS[tail])
        if tail in _PRESETS:
            return dict(_PRESETS[tail])
    raise ValueError(
        f"cannot resolve model config from {source!r}: not a directory with config.json, a config object, "
        f"or a known preset ({sorted(_PRESETS)})"
    )
