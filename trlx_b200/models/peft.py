"""Parameter-efficient fine-tuning adapters — in-repo implementation, PEFT-compatible on disk.

The reference delegates to the external ``peft`` package (``trlx/models/modeling_base.py:34-41,114-118,206-241``)
and toggles LoRA off to obtain reference-policy logits (``trlx/models/modeling_ppo.py:318-324``).  Here:

* **LoRA** is applied to row-slices of the fused canonical projections (``attn.qkv``, ``mlp.up`` …) so HF target
  names such as ``q_proj`` / ``v_proj`` keep working although Q, K, V share one weight.  A LoRA layer can evaluate
  *both* the adapted (policy) and the frozen (reference) output from ONE base GEMM — ``y_ref = xWᵀ``,
  ``y_pol = y_ref + s·(xAᵀ)Bᵀ`` — see :meth:`LoRALinear.forward_both`.
* **Prompt tuning** prepends learned virtual-token embeddings; **prefix tuning** prepends learned per-layer K/V.
* ``save_pretrained`` writes ``adapter_config.json`` + ``adapter_model.bin`` with PEFT's key naming
  (``base_model.model.<hf module path>.lora_A.weight`` …, ``prompt_embeddings``) so adapters interchange with PEFT.
"""
from __future__ import annotations

import contextlib
import json
import math
import os
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200.nn import hf_compat

ADAPTER_CONFIG = "adapter_config.json"
ADAPTER_WEIGHTS = "adapter_model.bin"
ADAPTER_WEIGHTS_SAFE = "adapter_model.safetensors"


@dataclass
class PeftConfig:
    peft_type: str = "LORA"
    task_type: Optional[str] = "CAUSAL_LM"
    base_model_name_or_path: Optional[str] = None
    inference_mode: bool = False
    # LoRA
    r: int = 8
    lora_alpha: int = 8
    lora_dropout: float = 0.0
    target_modules: Optional[List[str]] = None
    fan_in_fan_out: bool = False
    bias: str = "none"
    modules_to_save: Optional[List[str]] = None
    # prompt / prefix tuning
    num_virtual_tokens: int = 0
    token_dim: Optional[int] = None
    num_transformer_submodules: Optional[int] = None
    num_attention_heads: Optional[int] = None
    num_layers: Optional[int] = None
    prompt_tuning_init: str = "RANDOM"
    prefix_projection: bool = False
    encoder_hidden_size: Optional[int] = None
    extra: Dict[str, Any] = field(default_factory=dict)

    def __post_init__(self):
        self.peft_type = str(getattr(self.peft_type, "value", self.peft_type)).upper()
        if self.task_type is not None:
            self.task_type = str(getattr(self.task_type, "value", self.task_type)).upper()
        if isinstance(self.target_modules, str):
            self.target_modules = [self.target_modules]

    def to_dict(self) -> Dict[str, Any]:
        d = asdict(self)
        extra = d.pop("extra")
        d.update(extra)
        return d

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "PeftConfig":
        known = {f for f in cls.__dataclass_fields__ if f != "extra"}
        kw = {k: v for k, v in d.items() if k in known}
        extra = {k: v for k, v in d.items() if k not in known}
        return cls(**kw, extra=extra)

    def save_pretrained(self, directory: str) -> None:
        os.makedirs(directory, exist_ok=True)
        with open(os.path.join(directory, ADAPTER_CONFIG), "w") as fh:
            json.dump(self.to_dict(), fh, indent=2, default=str)

    @classmethod
    def from_pretrained(cls, directory: str) -> "PeftConfig":
        with open(os.path.join(directory, ADAPTER_CONFIG)) as fh:
            return cls.from_dict(json.load(fh))


class TaskType:
    """Task identifiers (same strings as the `peft` package)."""

    CAUSAL_LM = "CAUSAL_LM"
    SEQ_2_SEQ_LM = "SEQ_2_SEQ_LM"


class PeftType:
    LORA = "LORA"
    PROMPT_TUNING = "PROMPT_TUNING"
    PREFIX_TUNING = "PREFIX_TUNING"


def LoraConfig(**kw) -> PeftConfig:
    """`peft.LoraConfig(...)`-style constructor."""
    return PeftConfig(peft_type=PeftType.LORA, **kw)


def PromptTuningConfig(**kw) -> PeftConfig:
    return PeftConfig(peft_type=PeftType.PROMPT_TUNING, **kw)


def PrefixTuningConfig(**kw) -> PeftConfig:
    return PeftConfig(peft_type=PeftType.PREFIX_TUNING, **kw)


def get_peft_config(config: Union[Dict[str, Any], PeftConfig, Any]) -> PeftConfig:
    """Accepts our :class:`PeftConfig`, a plain dict, or a foreign (``peft``) config object with ``to_dict``."""
    if isinstance(config, PeftConfig):
        return config
    if isinstance(config, dict):
        return PeftConfig.from_dict(config)
    if hasattr(config, "to_dict"):
        return PeftConfig.from_dict({k: (list(v) if isinstance(v, set) else v) for k, v in config.to_dict().items()})
    raise ValueError("`peft_config` should be a dict or a PeftConfig")


# ---- LoRA ---------------------------------------------------------------------------------------------------------
class LoRALinear(nn.Module):
    """A frozen ``nn.Linear`` plus low-rank updates on (slices of) its output rows."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, dropout: float):
        super().__init__()
        self.base = base
        self.r, self.scaling = r, alpha / r
        self.dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.lora_A = nn.ParameterDict()
        self.lora_B = nn.ParameterDict()
        self.slices: Dict[str, Tuple[int, int]] = {}
        self.hf_paths: Dict[str, str] = {}
        self.interleaved: Dict[str, bool] = {}
        self.enabled = True
        for p in base.parameters():
            p.requires_grad_(False)

    @property
    def weight(self):
        return self.base.weight

    @property
    def bias(self):
        return self.base.bias

    @property
    def in_features(self):
        return self.base.in_features

    @property
    def out_features(self):
        return self.base.out_features

    def add_adapter(self, key: str, rows: Tuple[int, int], hf_path: str, interleaved: bool = False) -> None:
        lo, hi = rows
        if hi < 0:
            lo, hi = 0, self.base.out_features
        dev, dt = self.base.weight.device, self.base.weight.dtype
        a = torch.empty(self.r, self.base.in_features, device=dev, dtype=dt)
        nn.init.kaiming_uniform_(a, a=math.sqrt(5))
        self.lora_A[key] = nn.Parameter(a)
        self.lora_B[key] = nn.Parameter(torch.zeros(hi - lo, self.r, device=dev, dtype=dt))
        self.slices[key], self.hf_paths[key], self.interleaved[key] = (lo, hi), hf_path, interleaved

    def delta(self, x: torch.Tensor) -> torch.Tensor:
        """Low-rank update for the full output width (zeros outside adapted row-slices)."""
        xd = self.dropout(x)
        full = None
        for key in self.lora_A:
            lo, hi = self.slices[key]
            upd = F.linear(F.linear(xd, self.lora_A[key]), self.lora_B[key]) * self.scaling
            if lo == 0 and hi == self.base.out_features:
                full = upd if full is None else full + upd
            else:
                if full is None:
                    full = x.new_zeros(*x.shape[:-1], self.base.out_features)
                full = torch.cat([full[..., :lo], full[..., lo:hi] + upd, full[..., hi:]], dim=-1)
        return full

    def _fused_ok(self, x: torch.Tensor) -> bool:
        if not (x.is_cuda and x.dtype == torch.bfloat16 and os.environ.get("TRLX_B200_LORA_FUSED", "1") == "1"):
            return False
        if self.training and isinstance(self.dropout, nn.Dropout) and self.dropout.p > 0:
            return False  # dropout decouples the adapter input from the base input
        from trlx_b200 import ops

        return ops.available() and (self.r * len(self.lora_A)) % 8 == 0 and self.base.out_features % 8 == 0

    def add_delta(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """``y + Δ(x)``.  On the CUDA path the update is the epilogue of one GEMM: all adapters of this projection
        are stacked into one down-projection ``A`` ``[R, in]`` and one block-structured up-projection ``B`` ``[out, R]`` (zero
        outside each adapter's row slice, scaling folded in), so ``y_pol = (x Aᵀ) Bᵀ + y_ref`` is a skinny GEMM followed by a GEMM
        whose residual operand is the frozen projection's output — no separate multiply / add / concatenate kernels."""
        if not self._fused_ok(x):
            return y + self.delta(x)
        from trlx_b200 import ops

        keys = list(self.lora_A)
        R, out = self.r * len(keys), self.base.out_features
        A = self.lora_A[keys[0]] if len(keys) == 1 else torch.cat([self.lora_A[k] for k in keys], 0)
        Bf = None
        for i, k in enumerate(keys):
            lo, hi = self.slices[k]
            blk = F.pad(self.lora_B[k] * self.scaling, (i * self.r, R - (i + 1) * self.r, lo, out - hi))
            Bf = blk if Bf is None else Bf + blk
        t = ops.linear(x, A)
        return ops.linear(t, Bf, None, "none", y)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from trlx_b200.nn.transformer import project

        y = project(self.base, x)  # the frozen projection runs on the tcgen05 GEMM like an unadapted one
        if self.enabled and len(self.lora_A):
            y = self.add_delta(x, y)
        return y

    def forward_both(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(adapted, frozen)`` outputs from a single base GEMM."""
        from trlx_b200.nn.transformer import project

        y_ref = project(self.base, x)
        return self.add_delta(x, y_ref) if len(self.lora_A) else y_ref, y_ref

    def merged_weight(self) -> torch.Tensor:
        w = self.base.weight.detach().clone()
        for key in self.lora_A:
            lo, hi = self.slices[key]
            w[lo:hi] += (self.lora_B[key] @ self.lora_A[key]) * self.scaling
        return w


def _resolve_modules(model: nn.Module, path: str):
    parent = model
    parts = path.split(".")
    for p in parts[:-1]:
        parent = getattr(parent, p)
    return parent, parts[-1]


class PeftModel(nn.Module):
    """Adapter wrapper around a :class:`trlx_b200.nn.transformer.CausalLM` (or T5)."""

    def __init__(self, model: nn.Module, config: PeftConfig):
        super().__init__()
        self.base_model = model
        self.peft_config = config
        self.peft_type = config.peft_type
        self.config = model.config
        self._adapters_enabled = True
        for p in model.parameters():
            p.requires_grad_(False)
        spec = model.config
        if self.peft_type == "LORA":
            self._inject_lora(spec)
        elif self.peft_type == "PROMPT_TUNING":
            n = config.num_virtual_tokens
            emb = model.get_input_embeddings().weight
            init = emb[torch.randint(0, emb.shape[0], (n,))].detach().clone() if config.prompt_tuning_init != "ZERO" else torch.zeros(n, emb.shape[1])
            self.prompt_embeddings = nn.Parameter(init.to(emb.dtype).to(emb.device))
        elif self.peft_type == "PREFIX_TUNING":
            n = config.num_virtual_tokens
            L, kv = self._num_prefix_layers(), self._kv_width()
            self.prompt_embeddings = nn.Parameter(torch.randn(n, L * 2 * kv, device=model.get_input_embeddings().weight.device) * 0.02)
        else:
            raise NotImplementedError(f"peft_type {self.peft_type} is not supported")
        self._mark_modules_to_save()

    # -- helpers --------------------------------------------------------------------------------------------------
    def _is_seq2seq(self) -> bool:
        return bool(getattr(self.base_model.config, "is_encoder_decoder", False))

    def _num_prefix_layers(self) -> int:
        cfg = self.base_model.config
        return cfg.num_layers if not self._is_seq2seq() else cfg.num_decoder_layers

    def _kv_width(self) -> int:
        cfg = self.base_model.config
        return cfg.kv_size if not self._is_seq2seq() else cfg.num_heads * cfg.d_kv

    def _inject_lora(self, spec) -> None:
        cfg = self.peft_config
        if self._is_seq2seq():
            from trlx_b200.nn.t5 import inject_lora_t5

            inject_lora_t5(self.base_model, cfg, LoRALinear)
            return
        table = hf_compat.LORA_TARGETS[spec.family]
        targets = cfg.target_modules or hf_compat.DEFAULT_LORA_TARGETS[spec.family]
        unknown = [t for t in targets if t not in table]
        if unknown:
            raise ValueError(f"Target modules {unknown} not found in the base model ({spec.family}: {sorted(table)})")
        fam = hf_compat.family(spec)
        for i, block in enumerate(self.base_model.transformer.h):
            for t in targets:
                for canon, part, hf_rel in table[t]:
                    parent, leaf = _resolve_modules(block, canon)
                    mod = getattr(parent, leaf)
                    if not isinstance(mod, LoRALinear):
                        mod = LoRALinear(mod, cfg.r, cfg.lora_alpha, cfg.lora_dropout)
                        setattr(parent, leaf, mod)
                    mod.add_adapter(f"{t}", hf_compat.row_range(spec, canon, part), f"{fam.layer_prefix}{i}.{hf_rel}",
                                    interleaved=(part == "all_interleaved"))

    def _mark_modules_to_save(self) -> None:
        self._saved_modules: Dict[str, nn.Module] = {}
        for name in self.peft_config.modules_to_save or []:
            for mod_name, mod in self.base_model.named_modules():
                if mod_name.split(".")[-1] == name or mod_name == name:
                    mod.requires_grad_(True)
                    self._saved_modules[mod_name] = mod

    def lora_layers(self):
        return [m for m in self.base_model.modules() if isinstance(m, LoRALinear)]

    # -- adapter toggling -----------------------------------------------------------------------------------------
    def disable_adapter_layers(self):
        self._adapters_enabled = False
        for m in self.lora_layers():
            m.enabled = False

    def enable_adapter_layers(self):
        self._adapters_enabled = True
        for m in self.lora_layers():
            m.enabled = True

    @contextlib.contextmanager
    def disable_adapter(self):
        self.disable_adapter_layers()
        try:
            yield
        finally:
            self.enable_adapter_layers()

    # -- forward / generate -----------------------------------------------------------------------------------------
    def get_input_embeddings(self):
        return self.base_model.get_input_embeddings()

    def get_output_embeddings(self):
        return self.base_model.get_output_embeddings()

    def _prefix_past(self, batch: int, dtype):
        cfg = self.base_model.config
        n, L = self.peft_config.num_virtual_tokens, self._num_prefix_layers()
        if self._is_seq2seq():
            nkv, d = cfg.num_heads, cfg.d_kv
        else:
            nkv, d = cfg.num_kv_heads, cfg.head_dim
        pkv = self.prompt_embeddings.to(dtype).view(n, L, 2, nkv, d).permute(1, 2, 3, 0, 4)  # [L, 2, nkv, n, d]
        return [(pkv[i, 0].unsqueeze(0).expand(batch, -1, -1, -1), pkv[i, 1].unsqueeze(0).expand(batch, -1, -1, -1))
                for i in range(L)]

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                **kwargs):
        model = self.base_model
        if self.peft_type == "LORA" or not self._adapters_enabled:
            return model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, **kwargs)
        if self._is_seq2seq():
            return model.forward_with_prompt(self, input_ids=input_ids, attention_mask=attention_mask,
                                             past_key_values=past_key_values, inputs_embeds=inputs_embeds, **kwargs)
        n = self.peft_config.num_virtual_tokens
        ref = input_ids if input_ids is not None else inputs_embeds
        B, T = ref.shape[0], ref.shape[1]
        dev = ref.device
        if self.peft_type == "PREFIX_TUNING":
            if past_key_values is None:
                past_key_values = self._prefix_past(B, model.dtype)
                if attention_mask is None:
                    attention_mask = torch.ones(B, T, dtype=torch.long, device=dev)
                attention_mask = torch.cat([torch.ones(B, n, dtype=attention_mask.dtype, device=dev), attention_mask], 1)
                if position_ids is None:
                    position_ids = (attention_mask[:, n:].long().cumsum(-1) - 1).clamp_min(0)
            return model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, **kwargs)
        # PROMPT_TUNING: virtual tokens occupy the first n positions of the sequence
        if past_key_values is not None:  # cached decoding: the prompt is already inside the cache
            if position_ids is not None:
                position_ids = position_ids + n
            return model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, **kwargs)
        if inputs_embeds is None:
            inputs_embeds = model.get_input_embeddings()(input_ids)
        prompt = self.prompt_embeddings.to(inputs_embeds.dtype).unsqueeze(0).expand(B, -1, -1)
        inputs_embeds = torch.cat([prompt, inputs_embeds], 1)
        if attention_mask is None:
            attention_mask = torch.ones(B, T, dtype=torch.long, device=dev)
        attention_mask = torch.cat([torch.ones(B, n, dtype=attention_mask.dtype, device=dev), attention_mask], 1)
        out = model(inputs_embeds=inputs_embeds, attention_mask=attention_mask, position_ids=None, **kwargs)
        # drop the virtual positions so callers see tensors aligned with their tokens
        if out.logits is not None:
            out.logits = out.logits[:, n:]
        if out.hidden_states is not None:
            out.hidden_states = tuple(h[:, n:] for h in out.hidden_states)
        if out.last_hidden_state is not None:
            out.last_hidden_state = out.last_hidden_state[:, n:]
        return out

    def generate(self, *args, **kwargs):
        from trlx_b200.models.generation import generate

        return generate(self, *args, **kwargs)

    # -- (de)serialisation ------------------------------------------------------------------------------------------
    def adapter_state_dict(self) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        if self.peft_type == "LORA":
            spec = self.base_model.config
            for m in self.lora_layers():
                for key in m.lora_A:
                    b = m.lora_B[key].detach()
                    if m.interleaved[key]:
                        b = hf_compat._interleave_qkv(spec, b)
                    out[f"base_model.model.{m.hf_paths[key]}.lora_A.weight"] = m.lora_A[key].detach()
                    out[f"base_model.model.{m.hf_paths[key]}.lora_B.weight"] = b
        else:
            out["prompt_embeddings"] = self.prompt_embeddings.detach()
        for name, mod in self._saved_modules.items():
            for k, v in mod.state_dict().items():
                out[f"base_model.model.{name}.{k}"] = v.detach()
        return out

    def load_adapter_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        own = self.adapter_state_dict()
        missing = [k for k in own if k not in sd]
        if strict and missing:
            raise KeyError(f"adapter weights missing: {missing[:5]}…")
        with torch.no_grad():
            if self.peft_type == "LORA":
                spec = self.base_model.config
                for m in self.lora_layers():
                    for key in m.lora_A:
                        ka = f"base_model.model.{m.hf_paths[key]}.lora_A.weight"
                        kb = f"base_model.model.{m.hf_paths[key]}.lora_B.weight"
                        if ka in sd:
                            m.lora_A[key].copy_(sd[ka])
                        if kb in sd:
                            b = sd[kb]
                            if m.interleaved[key]:
                                b = hf_compat._deinterleave_qkv(spec, b)
                            m.lora_B[key].copy_(b)
            elif "prompt_embeddings" in sd:
                self.prompt_embeddings.copy_(sd["prompt_embeddings"])
            for name, mod in self._saved_modules.items():
                sub = {k[len(f"base_model.model.{name}."):]: v for k, v in sd.items() if k.startswith(f"base_model.model.{name}.")}
                if sub:
                    mod.load_state_dict(sub, strict=False)

    def save_pretrained(self, directory: str, state_dict=None, safe_serialization: bool = False, **_):
        os.makedirs(directory, exist_ok=True)
        self.peft_config.save_pretrained(directory)
        sd = {k: v.cpu() for k, v in self.adapter_state_dict().items()}
        torch.save(sd, os.path.join(directory, ADAPTER_WEIGHTS))

    @classmethod
    def from_pretrained(cls, model: nn.Module, directory: str, **_) -> "PeftModel":
        cfg = PeftConfig.from_pretrained(directory)
        peft_model = cls(model, cfg)
        path = os.path.join(directory, ADAPTER_WEIGHTS)
        if os.path.exists(path):
            sd = torch.load(path, map_location="cpu", weights_only=True)
        else:
            from safetensors.torch import load_file

            sd = load_file(os.path.join(directory, ADAPTER_WEIGHTS_SAFE))
        peft_model.load_adapter_state_dict(sd, strict=False)
        return peft_model

    def trainable_parameters(self):
        return [p for p in self.parameters() if p.requires_grad]


def get_peft_model(model: nn.Module, config: Union[PeftConfig, Dict[str, Any]]) -> PeftModel:
    return PeftModel(model, get_peft_config(config))
