"""Megatron/NeMo-schema recipes → this framework's model + parallel configuration.

The reference's NeMo trainers take ``trainer_kwargs = {"megatron_cfg": <yaml name | path | OmegaConf>, "pretrained_model":
<dir>}`` (``trlx/trainer/nemo_ppo_trainer.py:35-70``, ``configs/nemo_configs/*.yaml``).  The same keywords are honoured here:
``parse_megatron_cfg`` reads a recipe (a file under ``configs/nemo_configs``, any YAML path, a plain dict, or one of our own
:class:`TRLConfig` objects) and returns the pieces a trainer needs: an architecture dict for
:func:`trlx_b200.nn.arch.spec_from_hf_config`, the ``ParallelConfig`` overrides, batch sizes and the optimizer/schedule block.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import yaml

_CFG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "configs", "nemo_configs")


def _load(cfg) -> Dict[str, Any]:
    if isinstance(cfg, dict):
        return cfg
    if hasattr(cfg, "to_dict") and not isinstance(cfg, str):
        return cfg.to_dict()
    path = str(cfg)
    for cand in (path, os.path.join(_CFG_DIR, path), os.path.join(_CFG_DIR, path + ".yaml")):
        if os.path.isfile(cand):
            with open(cand) as fh:
                return yaml.safe_load(fh)
    raise FileNotFoundError(f"megatron_cfg `{cfg}` is neither a mapping nor a YAML file (looked in {_CFG_DIR})")


def arch_from_megatron(model: Dict[str, Any]) -> Dict[str, Any]:
    """NeMo ``model:`` block → HF-style config dict (GPT-NeoX layout for rope models, GPT-2 layout otherwise)."""
    hidden = int(model["hidden_size"])
    ffn = model.get("ffn_hidden_size") or 4 * hidden
    ffn = 4 * hidden if isinstance(ffn, str) else int(ffn)  # "${multiply:4, ${.hidden_size}}" in NeMo's own files
    div = int(model.get("make_vocab_size_divisible_by", 128))
    vocab = int(model.get("vocab_size", 50257))
    vocab = (vocab + div - 1) // div * div
    common = dict(vocab_size=vocab, max_position_embeddings=int(model.get("max_position_embeddings", 2048)),
                  layer_norm_eps=float(model.get("layernorm_epsilon", 1e-5)),
                  tie_word_embeddings=bool(model.get("share_embeddings_and_output_weights", True)))
    if str(model.get("position_embedding_type", "learned_absolute")) == "rope":
        act = str(model.get("activation", "gelu"))
        if "swiglu" in act:
            return dict(model_type="llama", hidden_size=hidden, num_hidden_layers=int(model["num_layers"]),
                        num_attention_heads=int(model["num_attention_heads"]), intermediate_size=int(ffn),
                        rms_norm_eps=common["layer_norm_eps"], vocab_size=vocab,
                        max_position_embeddings=common["max_position_embeddings"],
                        tie_word_embeddings=common["tie_word_embeddings"],
                        partial_rotary_factor=float(model.get("rotary_percentage", 1.0)))
        return dict(model_type="gpt_neox", hidden_size=hidden, num_hidden_layers=int(model["num_layers"]),
                    num_attention_heads=int(model["num_attention_heads"]), intermediate_size=int(ffn),
                    rotary_pct=float(model.get("rotary_percentage", 1.0)), **common)
    return dict(model_type="gpt2", n_embd=hidden, n_layer=int(model["num_layers"]), n_head=int(model["num_attention_heads"]),
                n_inner=int(ffn), n_positions=common["max_position_embeddings"], vocab_size=vocab,
                layer_norm_epsilon=common["layer_norm_eps"])


def parse_megatron_cfg(cfg) -> Dict[str, Any]:
    """→ ``dict(arch=…, parallel=…, micro_batch_size=…, global_batch_size=…, seq_length=…, optim=…, precision=…)``."""
    raw = _load(cfg)
    if "train" in raw and "method" in raw:  # one of our TRLConfigs
        par = raw["train"].get("parallel", {}) or {}
        return dict(arch=raw["model"]["model_path"], parallel=dict(par), micro_batch_size=raw["train"].get("minibatch_size"),
                    global_batch_size=raw["train"].get("batch_size"), seq_length=raw["train"].get("seq_length"),
                    optim=None, precision=par.get("precision", "bf16"))
    model = raw.get("model", raw)
    parallel = dict(tensor_parallel=int(model.get("tensor_model_parallel_size", 1)),
                    pipeline_parallel=int(model.get("pipeline_model_parallel_size", 1)),
                    virtual_pipeline_parallel=int(model.get("virtual_pipeline_model_parallel_size") or 1),
                    sequence_parallel=bool(model.get("sequence_parallel", False)),
                    activation_checkpointing=model.get("activations_checkpoint_granularity") in ("full", "selective"))
    prec = str((raw.get("trainer") or {}).get("precision", model.get("precision", "bf16")))
    parallel["precision"] = {"bf16": "bf16", "16": "fp16", "32": "fp32", "bf16-mixed": "bf16"}.get(prec, "bf16")
    return dict(arch=arch_from_megatron(model), parallel=parallel, micro_batch_size=model.get("micro_batch_size"),
                global_batch_size=model.get("global_batch_size"), seq_length=model.get("encoder_seq_length"),
                optim=model.get("optim"), precision=parallel["precision"])


def debug_sized(arch: Dict[str, Any]) -> Dict[str, Any]:
    """A multi-billion-parameter recipe on a single CPU process can only be a plumbing run: keep the family, the head layout
    and the vocabulary, shrink depth and width (``TRLX_B200_FULL_SIZE=1`` keeps the recipe's size)."""
    import torch

    if torch.cuda.is_available() or int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("TRLX_B200_FULL_SIZE") == "1":
        return arch
    hidden = int(arch.get("n_embd", arch.get("hidden_size", 0)) or 0)
    layers = int(arch.get("n_layer", arch.get("num_hidden_layers", 0)) or 0)
    if 12 * hidden * hidden * layers < 1_000_000_000:
        return arch
    small = dict(arch)
    for key, val in (("n_embd", 64), ("hidden_size", 64), ("n_layer", 2), ("num_hidden_layers", 2), ("n_head", 4),
                     ("num_attention_heads", 4), ("num_key_value_heads", 4), ("n_inner", 256), ("intermediate_size", 256)):
        if key in small:
            small[key] = val
    import warnings

    warnings.warn(f"megatron_cfg: {layers} x {hidden} model on one CPU process — using a 2 x 64 model of the same family for this "
                  "run (set TRLX_B200_FULL_SIZE=1 to keep the recipe's size)")
    return small


def apply_megatron_cfg(config, megatron_cfg, pretrained_model: Optional[str] = None):
    """Fold a recipe into a :class:`TRLConfig` (returns a new config): parallel layout, architecture (unless a checkpoint
    directory / explicit model is given) and, when the recipe carries one, the optimizer + schedule."""
    rec = parse_megatron_cfg(megatron_cfg)
    par = {k: v for k, v in rec["parallel"].items() if v is not None}
    upd: Dict[str, Any] = dict(train=dict(parallel=par))
    if pretrained_model and os.path.isdir(str(pretrained_model)):
        upd["model"] = dict(model_path=str(pretrained_model))
    elif rec["arch"] is not None and not isinstance(config.model.model_path, dict):
        upd["model"] = dict(model_path=rec["arch"])
    optim = rec.get("optim")
    if optim and config.optimizer.name in ("distributed_fused_adam",):
        kw = {k: optim[k] for k in ("lr", "weight_decay", "betas", "eps") if k in optim}
        upd["optimizer"] = dict(kwargs={**kw, **config.optimizer.kwargs})
    return config.evolve(**upd)
