"""Canonical ⇄ HuggingFace checkpoint key conversion.

The checkpoint format is part of the public contract: wrappers save
``base_model.<hf keys>`` so that directories written by the reference load here and vice versa.
Internally weights use the kernel-friendly canonical layout of :mod:`trlx_b200.nn.transformer`
(fused ``attn.qkv`` as ``[Q|K|V]`` rows, fused ``[gate|up]``, all ``[out, in]``); this module
translates, including GPT-2's transposed ``Conv1D`` weights and NeoX/Bloom's per-head interleaved
``query_key_value``.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch

from trlx_b200.nn.arch import ArchSpec

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _t(w: Tensor) -> Tensor:
    return w.t().contiguous()


# ---- qkv packing helpers ----------------------------------------------------------------------------
def _interleave_qkv(spec: ArchSpec, w: Tensor) -> Tensor:
    """canonical ``[3·h·d, …]`` ([Q|K|V]) → per-head ``[h, 3, d, …]`` flattened (NeoX / Bloom)."""
    h, d = spec.num_heads, spec.head_dim
    rest = w.shape[1:]
    return w.view(3, h, d, *rest).transpose(0, 1).reshape(3 * h * d, *rest).contiguous()


def _deinterleave_qkv(spec: ArchSpec, w: Tensor) -> Tensor:
    h, d = spec.num_heads, spec.head_dim
    rest = w.shape[1:]
    return w.view(h, 3, d, *rest).transpose(0, 1).reshape(3 * h * d, *rest).contiguous()


def _split_qkv(spec: ArchSpec, w: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    return tuple(t.contiguous() for t in w.split([spec.q_size, spec.kv_size, spec.kv_size], dim=0))  # type: ignore


class FamilyMap:
    """Naming scheme of one HF family: block-relative key map + top-level key names."""

    def __init__(self, layer_prefix: str, wte: str, wpe: str, ln_f: str, lm_head: str, emb_norm: str = ""):
        self.layer_prefix, self.wte, self.wpe, self.ln_f, self.lm_head, self.emb_norm = (
            layer_prefix, wte, wpe, ln_f, lm_head, emb_norm)

    def block_to_hf(self, spec: ArchSpec, b: SD) -> SD:  # pragma: no cover - overridden
        raise NotImplementedError

    def block_from_hf(self, spec: ArchSpec, b: SD) -> SD:  # pragma: no cover - overridden
        raise NotImplementedError


def _pairs_map(pairs: List[Tuple[str, str]], fwd_tf: Dict[str, Callable] = None, bwd_tf: Dict[str, Callable] = None):
    """Build (to_hf, from_hf) for families whose block keys map 1:1 (with optional per-key transforms)."""
    fwd_tf, bwd_tf = fwd_tf or {}, bwd_tf or {}

    def to_hf(spec, b):
        out = {}
        for canon, hf in pairs:
            for suffix in ("weight", "bias"):
                k = f"{canon}.{suffix}"
                if k in b:
                    f = fwd_tf.get(k)
                    out[f"{hf}.{suffix}"] = f(spec, b[k]) if f else b[k]
        return out

    def from_hf(spec, b):
        out = {}
        for canon, hf in pairs:
            for suffix in ("weight", "bias"):
                k = f"{hf}.{suffix}"
                if k in b:
                    f = bwd_tf.get(f"{canon}.{suffix}")
                    out[f"{canon}.{suffix}"] = f(spec, b[k]) if f else b[k]
        return out

    return to_hf, from_hf


class _PairsFamily(FamilyMap):
    def __init__(self, pairs, fwd_tf=None, bwd_tf=None, **names):
        super().__init__(**names)
        self._to, self._from = _pairs_map(pairs, fwd_tf, bwd_tf)

    def block_to_hf(self, spec, b):
        return self._to(spec, b)

    def block_from_hf(self, spec, b):
        return self._from(spec, b)


class _SplitQKVFamily(FamilyMap):
    """Families that store q/k/v (and possibly gate/up) as separate matrices."""

    def __init__(self, pairs, q, k, v, gate_up=None, **names):
        super().__init__(**names)
        self._to, self._from = _pairs_map(pairs)
        self.q, self.k, self.v, self.gate_up = q, k, v, gate_up

    def block_to_hf(self, spec, b):
        out = self._to(spec, b)
        for suffix in ("weight", "bias"):
            key = f"attn.qkv.{suffix}"
            if key in b:
                q, k, v = _split_qkv(spec, b[key])
                out[f"{self.q}.{suffix}"], out[f"{self.k}.{suffix}"], out[f"{self.v}.{suffix}"] = q, k, v
            if self.gate_up and f"mlp.up.{suffix}" in b:
                g, u = b[f"mlp.up.{suffix}"].chunk(2, dim=0)
                out[f"{self.gate_up[0]}.{suffix}"], out[f"{self.gate_up[1]}.{suffix}"] = g.contiguous(), u.contiguous()
        return out

    def block_from_hf(self, spec, b):
        out = self._from(spec, b)
        for suffix in ("weight", "bias"):
            if f"{self.q}.{suffix}" in b:
                out[f"attn.qkv.{suffix}"] = torch.cat([b[f"{self.q}.{suffix}"], b[f"{self.k}.{suffix}"], b[f"{self.v}.{suffix}"]], 0)
            if self.gate_up and f"{self.gate_up[0]}.{suffix}" in b:
                out[f"mlp.up.{suffix}"] = torch.cat([b[f"{self.gate_up[0]}.{suffix}"], b[f"{self.gate_up[1]}.{suffix}"]], 0)
        return out


_conv1d = lambda spec, w: _t(w)  # noqa: E731
_GPT2_T = {k: _conv1d for k in ("attn.qkv.weight", "attn.out.weight", "mlp.up.weight", "mlp.down.weight")}
_ILV_F = {"attn.qkv.weight": _interleave_qkv, "attn.qkv.bias": _interleave_qkv}
_ILV_B = {"attn.qkv.weight": _deinterleave_qkv, "attn.qkv.bias": _deinterleave_qkv}

FAMILIES: Dict[str, FamilyMap] = {
    "gpt2": _PairsFamily(
        [("norm1", "ln_1"), ("attn.qkv", "attn.c_attn"), ("attn.out", "attn.c_proj"), ("norm2", "ln_2"),
         ("mlp.up", "mlp.c_fc"), ("mlp.down", "mlp.c_proj")], _GPT2_T, _GPT2_T,
        layer_prefix="transformer.h.", wte="transformer.wte", wpe="transformer.wpe", ln_f="transformer.ln_f",
        lm_head="lm_head"),
    "gpt_bigcode": _PairsFamily(
        [("norm1", "ln_1"), ("attn.qkv", "attn.c_attn"), ("attn.out", "attn.c_proj"), ("norm2", "ln_2"),
         ("mlp.up", "mlp.c_fc"), ("mlp.down", "mlp.c_proj")],
        layer_prefix="transformer.h.", wte="transformer.wte", wpe="transformer.wpe", ln_f="transformer.ln_f",
        lm_head="lm_head"),
    "gpt_neo": _SplitQKVFamily(
        [("norm1", "ln_1"), ("attn.out", "attn.attention.out_proj"), ("norm2", "ln_2"), ("mlp.up", "mlp.c_fc"),
         ("mlp.down", "mlp.c_proj")], "attn.attention.q_proj", "attn.attention.k_proj", "attn.attention.v_proj",
        layer_prefix="transformer.h.", wte="transformer.wte", wpe="transformer.wpe", ln_f="transformer.ln_f",
        lm_head="lm_head"),
    "gptj": _SplitQKVFamily(
        [("norm1", "ln_1"), ("attn.out", "attn.out_proj"), ("mlp.up", "mlp.fc_in"), ("mlp.down", "mlp.fc_out")],
        "attn.q_proj", "attn.k_proj", "attn.v_proj",
        layer_prefix="transformer.h.", wte="transformer.wte", wpe="", ln_f="transformer.ln_f", lm_head="lm_head"),
    "gpt_neox": _PairsFamily(
        [("norm1", "input_layernorm"), ("attn.qkv", "attention.query_key_value"), ("attn.out", "attention.dense"),
         ("norm2", "post_attention_layernorm"), ("mlp.up", "mlp.dense_h_to_4h"), ("mlp.down", "mlp.dense_4h_to_h")],
        _ILV_F, _ILV_B,
        layer_prefix="gpt_neox.layers.", wte="gpt_neox.embed_in", wpe="", ln_f="gpt_neox.final_layer_norm",
        lm_head="embed_out"),
    "bloom": _PairsFamily(
        [("norm1", "input_layernorm"), ("attn.qkv", "self_attention.query_key_value"), ("attn.out", "self_attention.dense"),
         ("norm2", "post_attention_layernorm"), ("mlp.up", "mlp.dense_h_to_4h"), ("mlp.down", "mlp.dense_4h_to_h")],
        _ILV_F, _ILV_B,
        layer_prefix="transformer.h.", wte="transformer.word_embeddings", wpe="", ln_f="transformer.ln_f",
        lm_head="lm_head", emb_norm="transformer.word_embeddings_layernorm"),
    "llama": _SplitQKVFamily(
        [("norm1", "input_layernorm"), ("attn.out", "self_attn.o_proj"), ("norm2", "post_attention_layernorm"),
         ("mlp.down", "mlp.down_proj")], "self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj",
        gate_up=("mlp.gate_proj", "mlp.up_proj"),
        layer_prefix="model.layers.", wte="model.embed_tokens", wpe="", ln_f="model.norm", lm_head="lm_head"),
    "opt": _SplitQKVFamily(
        [("norm1", "self_attn_layer_norm"), ("attn.out", "self_attn.out_proj"), ("norm2", "final_layer_norm"),
         ("mlp.up", "fc1"), ("mlp.down", "fc2")], "self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj",
        layer_prefix="model.decoder.layers.", wte="model.decoder.embed_tokens", wpe="model.decoder.embed_positions",
        ln_f="model.decoder.final_layer_norm", lm_head="lm_head"),
}


def family(spec: ArchSpec) -> FamilyMap:
    return FAMILIES[spec.family]


def _group_blocks(sd: SD, prefix: str) -> Tuple[Dict[int, SD], SD]:
    blocks: Dict[int, SD] = {}
    rest: SD = {}
    for k, v in sd.items():
        if k.startswith(prefix):
            idx, _, tail = k[len(prefix):].partition(".")
            blocks.setdefault(int(idx), {})[tail] = v
        else:
            rest[k] = v
    return blocks, rest


def _tail_keys(fam: FamilyMap) -> Dict[str, str]:
    """Canonical → HF module names of the OPT-350m extras (final norm inside ``Tail``, project_in / project_out)."""
    base = fam.ln_f.rpartition(".")[0]
    return {"transformer.ln_f.norm": fam.ln_f, "transformer.ln_f.project_out": f"{base}.project_out",
            "transformer.project_in": f"{base}.project_in"}


def to_hf(spec: ArchSpec, canonical: SD) -> SD:
    """Full-model canonical state dict → HF key names/layouts."""
    fam = family(spec)
    blocks, rest = _group_blocks(canonical, "transformer.h.")
    out: SD = {}
    top = {"transformer.wte": fam.wte, "transformer.wpe": fam.wpe, "transformer.ln_f": fam.ln_f,
           "lm_head": fam.lm_head, "transformer.emb_norm": fam.emb_norm}
    top.update(_tail_keys(fam))
    for k, v in rest.items():
        mod, _, suffix = k.rpartition(".")
        if mod in top and top[mod]:
            out[f"{top[mod]}.{suffix}"] = v
        else:
            out[k] = v
    for i, b in sorted(blocks.items()):
        for k, v in fam.block_to_hf(spec, b).items():
            out[f"{fam.layer_prefix}{i}.{k}"] = v
    return out


def from_hf(spec: ArchSpec, hf: SD) -> SD:
    """HF state dict → canonical.  Unknown keys (e.g. ``attn.bias`` causal-mask buffers) are dropped."""
    fam = family(spec)
    blocks, rest = _group_blocks(hf, fam.layer_prefix)
    out: SD = {}
    top = {fam.wte: "transformer.wte", fam.ln_f: "transformer.ln_f", fam.lm_head: "lm_head"}
    if fam.wpe:
        top[fam.wpe] = "transformer.wpe"
    if fam.emb_norm:
        top[fam.emb_norm] = "transformer.emb_norm"
    if not spec.plain_tail:
        top[fam.ln_f] = "transformer.ln_f.norm"
    if spec.word_embed_dim != spec.hidden_size:
        top.update({hf: ours for ours, hf in _tail_keys(fam).items() if ours != "transformer.ln_f.norm"})
    for k, v in rest.items():
        mod, _, suffix = k.rpartition(".")
        if mod in top:
            out[f"{top[mod]}.{suffix}"] = v
    for i, b in blocks.items():
        for k, v in fam.block_from_hf(spec, b).items():
            out[f"transformer.h.{i}.{k}"] = v
    return out


def branch_to_hf(spec: ArchSpec, branch_sd: SD) -> SD:
    """Keys of a hydra branch (``decoder_blocks.N.*``, ``final_norm.*``, ``lm_head.*``) → HF block naming."""
    fam = family(spec)
    blocks, rest = _group_blocks(branch_sd, "decoder_blocks.")
    out = dict(rest)
    for i, b in sorted(blocks.items()):
        for k, v in fam.block_to_hf(spec, b).items():
            out[f"decoder_blocks.{i}.{k}"] = v
    return out


def branch_from_hf(spec: ArchSpec, hf_branch_sd: SD) -> SD:
    fam = family(spec)
    blocks, rest = _group_blocks(hf_branch_sd, "decoder_blocks.")
    out = dict(rest)
    for i, b in blocks.items():
        for k, v in fam.block_from_hf(spec, b).items():
            out[f"decoder_blocks.{i}.{k}"] = v
    return out


# ---- adapter target resolution (LoRA on the fused canonical projections) -------------------------------------
# family -> {hf leaf name: [(canonical module path inside a block, row part, hf module path inside a block)]}
_T = lambda canon, part, hf: (canon, part, hf)  # noqa: E731
LORA_TARGETS: Dict[str, Dict[str, List[Tuple[str, str, str]]]] = {
    "gpt2": {"c_attn": [_T("attn.qkv", "all", "attn.c_attn")], "c_fc": [_T("mlp.up", "all", "mlp.c_fc")],
             "c_proj": [_T("attn.out", "all", "attn.c_proj"), _T("mlp.down", "all", "mlp.c_proj")]},
    "gpt_bigcode": {"c_attn": [_T("attn.qkv", "all", "attn.c_attn")], "c_fc": [_T("mlp.up", "all", "mlp.c_fc")],
                    "c_proj": [_T("attn.out", "all", "attn.c_proj"), _T("mlp.down", "all", "mlp.c_proj")]},
    "gpt_neo": {"q_proj": [_T("attn.qkv", "q", "attn.attention.q_proj")], "k_proj": [_T("attn.qkv", "k", "attn.attention.k_proj")],
                "v_proj": [_T("attn.qkv", "v", "attn.attention.v_proj")], "out_proj": [_T("attn.out", "all", "attn.attention.out_proj")],
                "c_fc": [_T("mlp.up", "all", "mlp.c_fc")], "c_proj": [_T("mlp.down", "all", "mlp.c_proj")]},
    "gptj": {"q_proj": [_T("attn.qkv", "q", "attn.q_proj")], "k_proj": [_T("attn.qkv", "k", "attn.k_proj")],
             "v_proj": [_T("attn.qkv", "v", "attn.v_proj")], "out_proj": [_T("attn.out", "all", "attn.out_proj")],
             "fc_in": [_T("mlp.up", "all", "mlp.fc_in")], "fc_out": [_T("mlp.down", "all", "mlp.fc_out")]},
    "gpt_neox": {"query_key_value": [_T("attn.qkv", "all_interleaved", "attention.query_key_value")],
                 "dense": [_T("attn.out", "all", "attention.dense")],
                 "dense_h_to_4h": [_T("mlp.up", "all", "mlp.dense_h_to_4h")], "dense_4h_to_h": [_T("mlp.down", "all", "mlp.dense_4h_to_h")]},
    "bloom": {"query_key_value": [_T("attn.qkv", "all_interleaved", "self_attention.query_key_value")],
              "dense": [_T("attn.out", "all", "self_attention.dense")],
              "dense_h_to_4h": [_T("mlp.up", "all", "mlp.dense_h_to_4h")], "dense_4h_to_h": [_T("mlp.down", "all", "mlp.dense_4h_to_h")]},
    "llama": {"q_proj": [_T("attn.qkv", "q", "self_attn.q_proj")], "k_proj": [_T("attn.qkv", "k", "self_attn.k_proj")],
              "v_proj": [_T("attn.qkv", "v", "self_attn.v_proj")], "o_proj": [_T("attn.out", "all", "self_attn.o_proj")],
              "gate_proj": [_T("mlp.up", "gate", "mlp.gate_proj")], "up_proj": [_T("mlp.up", "up", "mlp.up_proj")],
              "down_proj": [_T("mlp.down", "all", "mlp.down_proj")]},
    "opt": {"q_proj": [_T("attn.qkv", "q", "self_attn.q_proj")], "k_proj": [_T("attn.qkv", "k", "self_attn.k_proj")],
            "v_proj": [_T("attn.qkv", "v", "self_attn.v_proj")], "out_proj": [_T("attn.out", "all", "self_attn.out_proj")],
            "fc1": [_T("mlp.up", "all", "fc1")], "fc2": [_T("mlp.down", "all", "fc2")]},
}
DEFAULT_LORA_TARGETS = {"gpt2": ["c_attn"], "gpt_bigcode": ["c_attn"], "gpt_neo": ["q_proj", "v_proj"], "gptj": ["q_proj", "v_proj"],
                        "gpt_neox": ["query_key_value"], "bloom": ["query_key_value"], "llama": ["q_proj", "v_proj"],
                        "opt": ["q_proj", "v_proj"], "t5": ["q", "v"]}


def row_range(spec: ArchSpec, canon: str, part: str) -> Tuple[int, int]:
    """Rows of the fused canonical weight that belong to ``part`` of module ``canon``."""
    if part in ("all", "all_interleaved"):
        return 0, -1
    if canon == "attn.qkv":
        q, kv = spec.q_size, spec.kv_size
        return {"q": (0, q), "k": (q, q + kv), "v": (q + kv, q + 2 * kv)}[part]
    if canon == "mlp.up":
        f = spec.ffn_size
        return {"gate": (0, f), "up": (f, 2 * f)}[part]
    raise KeyError((canon, part))
