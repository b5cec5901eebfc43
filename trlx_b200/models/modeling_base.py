"""Base class of the model wrappers (value-head / hydra / ILQL).

Parity: ``trlx/models/modeling_base.py`` — ``from_config`` ``:95-121``, ``from_pretrained`` (adapter logic, sharded
checkpoints, ``post_init(state_dict)``) ``:123-326``, ``save_pretrained`` (adapter ⇒ heads-only ``pytorch_model.bin``)
``:328-355``.  The wrapped ``base_model`` is the framework's own :class:`~trlx_b200.nn.transformer.CausalLM` /
:class:`~trlx_b200.nn.t5.T5Model` (random-init from a config when no weights exist — B200 boxes are offline), and the
on-disk layout is the HF/PEFT one (``config.json``, ``pytorch_model.bin`` with ``base_model.*`` keys,
``adapter_config.json`` + ``adapter_model.bin``).
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from trlx_b200.models import checkpoint_io
from trlx_b200.models.peft import ADAPTER_CONFIG, PeftConfig, PeftModel, get_peft_config, get_peft_model
from trlx_b200.nn import hf_compat
from trlx_b200.nn.arch import resolve_config, spec_from_hf_config
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)

_HEAD_PREFIXES = ("v_head.", "ilql_heads.", "frozen_head.")


def build_base_model(config_source, arch_type: str = "causal", dtype=None) -> nn.Module:
    """Random-init base LM from anything :func:`resolve_config` understands."""
    raw = resolve_config(config_source)
    if raw.get("model_type") in ("t5", "mt5", "flan-t5") or raw.get("is_encoder_decoder"):
        from trlx_b200.nn.t5 import T5Model, t5_spec_from_config

        model = T5Model(t5_spec_from_config(raw), dtype=dtype)
    else:
        from trlx_b200.nn.transformer import CausalLM

        spec = raw["spec"] if raw.get("model_type") == "__spec__" else spec_from_hf_config(raw)
        model = CausalLM(spec, dtype=dtype)
    model.hf_config_dict = {k: v for k, v in raw.items() if k != "spec"} if raw.get("model_type") != "__spec__" else None
    return model


def base_lm(model: nn.Module) -> nn.Module:
    """Strip a :class:`PeftModel` shell."""
    return model.base_model if isinstance(model, PeftModel) else model


def export_base_state_dict(model: nn.Module, prefix: str = "") -> Dict[str, torch.Tensor]:
    """HF-named state dict of the base LM (adapters excluded; LoRA-wrapped bases export their frozen weights)."""
    lm = base_lm(model)
    sd = {}
    for k, v in lm.state_dict().items():
        if ".lora_A." in k or ".lora_B." in k:
            continue
        sd[k.replace(".base.weight", ".weight").replace(".base.bias", ".bias")] = v
    hf = lm.to_hf_state_dict(sd) if hasattr(lm, "to_hf_state_dict") else hf_compat.to_hf(lm.config, sd)
    return {prefix + k: v for k, v in hf.items()}


def _base_prefix(lm) -> str:
    """``transformer.`` / ``model.`` / ``gpt_neox.`` / ``model.decoder.``: what HF puts in front of base-model keys."""
    try:
        fam = hf_compat.family(lm.config)
        return fam.wte.rsplit(".", 1)[0] + "." if "." in fam.wte else ""
    except Exception:
        return ""


def import_base_state_dict(model: nn.Module, hf_sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
    """Load an HF-layout (or canonical) state dict into the base LM.

    Accepts checkpoints saved from the *base* model class (``wte.weight``, ``h.0.*`` — no ``transformer.`` prefix, no
    ``lm_head``; how the hub's ``gpt2`` weights are stored).  Never leaves the model silently random: raises when nothing
    matched, warns loudly about every non-head tensor that stayed at its initial value (``strict`` raises instead)."""
    lm = base_lm(model)

    def convert(sd):
        return lm.from_hf_state_dict(sd) if hasattr(lm, "from_hf_state_dict") else hf_compat.from_hf(lm.config, sd)

    target = lm.state_dict()
    canon = convert(hf_sd)
    if not any(k in target for k in canon) and hf_sd:
        prefix = _base_prefix(lm)
        if prefix and not any(k.startswith(prefix) for k in hf_sd):
            canon = convert({prefix + k: v for k, v in hf_sd.items()})  # base-model checkpoint: re-attach the prefix
        if not any(k in target for k in canon):
            canon = dict(hf_sd)  # already canonical
    remap = {}
    for k, v in canon.items():
        if k in target:
            remap[k] = v
        else:  # LoRA-wrapped projection: weights live under ``.base``
            alt = k.rsplit(".", 1)[0] + ".base." + k.rsplit(".", 1)[1] if "." in k else k
            if alt in target:
                remap[alt] = v
            elif strict:
                raise KeyError(f"unexpected key {k} in checkpoint")
    if hf_sd and not remap:
        raise ValueError("no tensor of the checkpoint matches the model (first checkpoint keys: "
                         f"{list(hf_sd)[:4]}, first model keys: {list(target)[:4]}); refusing to continue with "
                         "randomly initialised weights")
    missing = [k for k in target if k not in remap and ".lora_" not in k]
    tied = getattr(lm.config, "tie_word_embeddings", False)
    missing = [k for k in missing if not (tied and k == "lm_head.weight")]
    if "lm_head.weight" in missing and "transformer.wte.weight" in remap and \
            target["lm_head.weight"].shape == remap["transformer.wte.weight"].shape:
        remap["lm_head.weight"] = remap["transformer.wte.weight"]  # base-model checkpoint: the head is the embedding
        missing.remove("lm_head.weight")
    buffers = {k for k, _ in lm.named_buffers()}
    missing = [k for k in missing if k not in buffers]
    # tensors that alias a loaded one (T5's encoder / decoder `embed_tokens` are the `shared` embedding; newer HF checkpoints
    # store such duplicates once) receive their values through that alias
    loaded_storage = {target[k].data_ptr() for k in remap if k in target and target[k].numel()}
    missing = [k for k in missing if not (target[k].numel() and target[k].data_ptr() in loaded_storage)]
    if missing:
        msg = (f"{len(missing)} of {len(target)} model tensors are NOT in the checkpoint and keep their initial values: "
               f"{missing[:8]}{'…' if len(missing) > 8 else ''}")
        if strict:
            raise KeyError(msg)
        logger.warning(msg)
    lm.load_state_dict(remap, strict=False)


def hf_config_dict(model: nn.Module) -> Dict[str, Any]:
    lm = base_lm(model)
    raw = getattr(lm, "hf_config_dict", None)
    if raw:
        return dict(raw)
    d = lm.config.to_dict()
    d["model_type"] = "trlx_b200_spec"
    return d


class PreTrainedModelWrapper(nn.Module):
    """Wraps a base LM and adds heads.  Sub-classes define ``_supported_args`` (ctor kwargs split off from the
    base-model kwargs), ``state_dict`` and ``post_init``."""

    _auto_model_parent_class = None
    _supported_modules = None
    _supported_args = None
    arch_type = "causal"

    def __init__(self, base_model: Optional[nn.Module] = None, peft_config=None, **kwargs):
        super().__init__()
        self.base_model = base_model
        self.is_loaded_in_8bit = False  # 8-bit loading is not implemented (neither is it in the reference, :72-77)
        if isinstance(base_model, PeftModel) and peft_config is None:
            peft_config = base_model.peft_config
        self.peft_config = get_peft_config(peft_config) if peft_config is not None else None
        self.peft_type = self.peft_config.peft_type if self.peft_config else None
        self.forward_kwargs = ["input_ids", "attention_mask", "position_ids", "past_key_values", "inputs_embeds",
                               "use_cache", "output_hidden_states", "return_dict", "decoder_input_ids",
                               "decoder_attention_mask", "encoder_outputs", "decoder_inputs_embeds", "labels"]

    # ---- kwargs plumbing -----------------------------------------------------------------------------------------
    @classmethod
    def _split_kwargs(cls, kwargs: Dict[str, Any]):
        supported, unsupported = {}, {}
        for k, v in kwargs.items():
            (supported if k in (cls._supported_args or []) else unsupported)[k] = v
        return supported, unsupported

    def get_compatible_forward_kwargs(self, **kwargs) -> Dict[str, Any]:
        return {k: v for k, v in kwargs.items() if k in self.forward_kwargs}

    # ---- construction --------------------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config, peft_config=None, **kwargs):
        """Random-init model from a config (object, dict, preset name or directory)."""
        wrapped_kwargs, base_kwargs = cls._split_kwargs(kwargs)
        base = build_base_model(config, cls.arch_type, dtype=base_kwargs.get("torch_dtype"))
        if peft_config:
            base = get_peft_model(base, peft_config)
            wrapped_kwargs["peft_config"] = base.peft_config
        return cls(base, **wrapped_kwargs)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, revision=None, peft_config=None, *model_args, **kwargs):
        """Load from a directory (HF layout, with or without wrapper prefixes / adapters), a preset name
        (random-init, offline) or an ``nn.Module`` instance."""
        wrapped_kwargs, base_kwargs = cls._split_kwargs(kwargs)
        base_kwargs.pop("peft_from_pretrained_kwargs", None)
        base_kwargs.pop("peft_int8_kwargs", None)
        if base_kwargs.get("load_in_8bit"):
            raise NotImplementedError("8-bit loading is not supported")
        if peft_config is not None and not isinstance(peft_config, (dict, PeftConfig)) and not hasattr(peft_config, "to_dict"):
            raise ValueError("`peft_config` should be an instance of `PeftConfig` or a dict.")

        state_dict = None
        src = pretrained_model_name_or_path
        if isinstance(src, str):
            is_dir = os.path.isdir(src)
            adapter_here = is_dir and os.path.exists(os.path.join(src, ADAPTER_CONFIG))
            cfg_src = src
            if adapter_here and not os.path.exists(os.path.join(src, "config.json")):
                cfg_src = PeftConfig.from_pretrained(src).base_model_name_or_path or src
            base = build_base_model(cfg_src, cls.arch_type, dtype=base_kwargs.get("torch_dtype"))
            if is_dir and checkpoint_io.has_weights(src):
                state_dict = checkpoint_io.load_state_dict(src)
            elif not is_dir:
                logger.warning(f"'{src}' resolved to a config preset; weights are RANDOM-INITIALISED (no hub access)")
            if state_dict is not None:
                base_sd = {k[len("base_model."):] if k.startswith("base_model.") else k: v for k, v in state_dict.items()
                           if not k.startswith(_HEAD_PREFIXES)}
                if base_sd:
                    import_base_state_dict(base, base_sd, strict=False)
            if adapter_here and peft_config is None:
                base = PeftModel.from_pretrained(base, src)
                peft_config = base.peft_config
                logger.info("Trained peft adapter loaded")
            elif peft_config is not None:
                if adapter_here:
                    logger.warning(f"WARNING: peft config file detected in {src} but ignored since the argument "
                                   "`peft_config` is provided. Remove the argument `peft_config` to use the trained peft adapter.")
                base = get_peft_model(base, peft_config)
                peft_config = base.peft_config
                logger.info("peft adapter initialised")
        elif isinstance(src, nn.Module):
            base = src
            if not hasattr(base_lm(base), "transformer") and not hasattr(base_lm(base), "encoder"):
                base = _convert_foreign_model(src)
            if peft_config is not None and not isinstance(base, PeftModel):
                base = get_peft_model(base, peft_config)
                peft_config = base.peft_config
            elif isinstance(base, PeftModel):
                peft_config = base.peft_config
        else:
            raise ValueError(f"pretrained_model_name_or_path should be a string or a nn.Module, got {type(src)}")

        if peft_config is not None:
            wrapped_kwargs["peft_config"] = peft_config
        model = cls(base, **wrapped_kwargs)
        head_sd = {k: v for k, v in (state_dict or {}).items() if k.startswith(_HEAD_PREFIXES)}
        model.post_init(state_dict=head_sd)
        return model

    # ---- persistence ---------------------------------------------------------------------------------------------
    def save_pretrained(self, save_directory: str, state_dict: Optional[Dict[str, torch.Tensor]] = None, **kwargs):
        """Write ``config.json`` + ``pytorch_model.bin``.  With an adapter: heads-only ``pytorch_model.bin`` next to
        ``adapter_config.json`` / ``adapter_model.bin``."""
        os.makedirs(save_directory, exist_ok=True)
        checkpoint_io.save_config(save_directory, hf_config_dict(self.base_model))
        if self.peft_type:
            torch.save({k: v.detach().cpu() for k, v in self.state_dict(heads_only=True).items()},
                       os.path.join(save_directory, checkpoint_io.WEIGHTS_BIN))
            self.base_model.save_pretrained(save_directory)
            return
        if state_dict is None:
            state_dict = self.state_dict()
        checkpoint_io.save_state_dict(save_directory, state_dict, max_shard_bytes=kwargs.get("max_shard_bytes"),
                                      safe_serialization=bool(kwargs.get("safe_serialization", False)))

    def push_to_hub(self, repo_id: str, commit_message: str = "Upload model", private: Optional[bool] = None,
                    token: Optional[str] = None, revision: Optional[str] = None, create_pr: bool = False, **save_kwargs):
        """Serialise with :meth:`save_pretrained` into a temporary directory and upload it (the reference inherits this from
        ``transformers.utils.PushToHubMixin``, ``trlx/models/modeling_base.py:37``).  Needs ``huggingface_hub`` and network
        access; both are checked up front so an offline box fails with a clear message instead of half-way through."""
        import tempfile

        try:
            from huggingface_hub import HfApi
        except ImportError as err:  # pragma: no cover
            raise RuntimeError("push_to_hub needs the `huggingface_hub` package") from err
        if os.environ.get("HF_HUB_OFFLINE", "0") == "1" or os.environ.get("TRANSFORMERS_OFFLINE", "0") == "1":
            raise RuntimeError("push_to_hub: the Hugging Face Hub is disabled (HF_HUB_OFFLINE / TRANSFORMERS_OFFLINE)")
        api = HfApi(token=token)
        with tempfile.TemporaryDirectory() as tmp:
            self.save_pretrained(tmp, **save_kwargs)
            api.create_repo(repo_id, private=private, exist_ok=True)
            return api.upload_folder(repo_id=repo_id, folder_path=tmp, commit_message=commit_message, revision=revision,
                                     create_pr=create_pr)

    def post_init(self, *args, **kwargs):
        """Hook run after construction in ``from_pretrained`` (loads head weights)."""

    def state_dict(self, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def raw_state_dict(self, *args, **kwargs):
        """Canonical (un-renamed) parameters — used by the trainer's own optimizer/runtime checkpoints."""
        return nn.Module.state_dict(self, *args, **kwargs)

    @property
    def config(self):
        return base_lm(self.base_model).config

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return base_lm(self.base_model).dtype


def _convert_foreign_model(model: nn.Module) -> nn.Module:
    """Turn a HuggingFace ``PreTrainedModel`` instance into the in-repo equivalent (weights copied)."""
    cfg = getattr(model, "config", None)
    if cfg is None or not hasattr(cfg, "model_type"):
        raise ValueError("cannot wrap this module: it is neither an in-repo model nor a HuggingFace PreTrainedModel")
    ours = build_base_model(cfg.to_dict())
    import_base_state_dict(ours, {k: v.detach() for k, v in model.state_dict().items()}, strict=False)
    return ours
