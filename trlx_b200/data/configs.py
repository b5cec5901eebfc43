"""Typed configuration tree.

Public names, fields and defaults follow ``trlx/data/configs.py`` (ModelConfig ``:37-72``,
TokenizerConfig ``:75-97``, OptimizerConfig ``:100-117``, SchedulerConfig ``:120-137``,
TrainConfig ``:140-236``, TRLConfig ``:239-335``) so that YAML files, ``evolve`` calls and
dotted-key ``update`` overrides written for the reference keep working.  On top of that the
train section grows one B200-specific block, ``TrainConfig.parallel``, that replaces the
out-of-tree Accelerate/DeepSpeed/NeMo YAMLs (``configs/accelerate/*.yaml``) with a single
in-tree description of the DP × TP × PP layout and optimizer sharding.
"""
from __future__ import annotations

import copy
import json
from dataclasses import dataclass, field, fields
from typing import Any, Dict, List, Optional, Set

import yaml

from trlx_b200.data.method_configs import MethodConfig, get_method


def merge(base: Dict, update: Dict, updated: Set) -> Dict:
    """In-place recursive overwrite of keys that already exist in ``base``.

    Every key that was found (at any depth) is recorded in ``updated`` so that callers can
    detect typos (keys in ``update`` that matched nothing).
    """
    for key in list(base.keys()):
        if key not in update:
            continue
        if isinstance(base[key], dict) and isinstance(update[key], dict):
            merge(base[key], update[key], updated)
        else:
            base[key] = update[key]
        updated.add(key)
    return base


_ATOMIC_KEYS = ("model_path", "tokenizer_path", "peft_config")  # dict-valued fields that are replaced, never merged


def _merge_dicts(base: Dict, update: Dict) -> Dict:
    """Pure recursive union (new keys allowed); returns a fresh dict."""
    out = copy.deepcopy(base)
    for key, val in update.items():
        if isinstance(val, dict) and key not in _ATOMIC_KEYS:
            prev = out.get(key)
            out[key] = _merge_dicts(prev if isinstance(prev, dict) else {}, val)
        else:
            out[key] = val
    return out


class _Section:
    """Mixin: ``from_dict`` / ``to_dict`` for flat dataclass sections."""

    @classmethod
    def from_dict(cls, config: Dict[str, Any]):
        return cls(**config)

    def to_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in fields(self)}  # type: ignore[arg-type]


@dataclass
class ModelConfig(_Section):
    """
    :param model_path: local directory / hub-style name of the model, or a config dict / object
        (random-init; there is no network on B200 boxes)
    :param model_arch_type: ``"causal"`` or ``"seq2seq"``
    :param num_layers_unfrozen: number of top transformer blocks to train (-1 = all)
    :param peft_config: dict (or object with ``to_dict``) describing a LoRA / prompt-tuning /
        prefix-tuning adapter, e.g. ``{"peft_type": "LORA", "r": 8, "lora_alpha": 32}``
    :param model_extra_configs: extra kwargs forwarded to ``from_pretrained`` / ``from_config``
    """

    model_path: Any
    model_arch_type: str = "causal"
    num_layers_unfrozen: int = -1
    peft_config: Any = None
    model_extra_configs: Dict[str, Any] = field(default_factory=dict)


@dataclass
class TokenizerConfig(_Section):
    """
    :param tokenizer_path: local directory / name of the tokenizer (``"toy://..."`` selects the
        in-repo tokenizer, see :mod:`trlx_b200.utils.tokenizer`)
    :param padding_side: ``"left"`` or ``"right"``
    :param truncation_side: ``"left"`` or ``"right"``
    """

    tokenizer_path: Any
    padding_side: str = "left"
    truncation_side: str = "right"
    tokenizer_extra_configs: Dict[str, Any] = field(default_factory=dict)


@dataclass
class OptimizerConfig(_Section):
    """:param name: one of :class:`trlx_b200.utils.OptimizerName`; ``kwargs`` go to its ctor."""

    name: str
    kwargs: Dict[str, Any] = field(default_factory=dict)


@dataclass
class SchedulerConfig(_Section):
    """:param name: one of :class:`trlx_b200.utils.SchedulerName`; ``kwargs`` go to its ctor."""

    name: str
    kwargs: Dict[str, Any] = field(default_factory=dict)


@dataclass
class ParallelConfig(_Section):
    """B200 runtime layout (no reference equivalent in-tree; stands in for
    ``configs/accelerate/{ddp,zero2-bf16,zero3}.yaml`` and the NeMo ``tensor_model_parallel_size``
    / ``pipeline_model_parallel_size`` / ``sequence_parallel`` keys,
    ``configs/nemo_configs/megatron_20b.yaml:51-55,82``).

    :param tensor_parallel: TP degree (ranks that share one replica along hidden dims)
    :param pipeline_parallel: PP degree
    :param virtual_pipeline_parallel: model chunks per pipeline rank (interleaved 1F1B when > 1;
        ``virtual_pipeline_model_parallel_size`` in the NeMo recipes)
    :param sequence_parallel: shard norm/dropout activations along sequence inside the TP group
    :param zero_stage: 0 = replicated optimizer (DDP-like), 1/2 = optimizer state + grads sharded
        across DP ranks (fused reduce-scatter + AdamW + all-gather), 3 = parameters sharded as well
    :param precision: compute dtype, ``"bf16"`` (default), ``"fp16"`` or ``"fp32"``
    :param rollout_dtype: ``"bf16"`` or ``"fp8"``: with fp8 the rollout engine runs the norm → QKV and norm → MLP-up GEMMs of
        every block in e4m3 x e4m3 (activations quantised per row inside the norm kernel, weights per output channel)
    :param cuda_graphs: capture decode steps / train steps in CUDA graphs when shapes are static
    :param bucket_mb: gradient bucket size for the fused reduce-scatter/AdamW kernel
    :param grad_clip: global-norm clip applied inside the fused optimizer (0/None = off;
        DeepSpeed configs of the reference use 1.0, ``configs/accelerate/zero2-bf16.yaml:5``)
    """

    tensor_parallel: int = 1
    pipeline_parallel: int = 1
    virtual_pipeline_parallel: int = 1
    sequence_parallel: bool = False
    zero_stage: int = 1
    precision: str = "bf16"
    rollout_dtype: str = "bf16"
    cuda_graphs: bool = True
    bucket_mb: float = 32.0
    grad_clip: Optional[float] = None
    activation_checkpointing: bool = False


@dataclass
class TrainConfig(_Section):
    """
    :param total_steps: total number of optimizer steps
    :param seq_length: context length (max tokenizer length)
    :param epochs: passes over the data / outer PPO iterations
    :param batch_size: per-rank batch size
    :param checkpoint_interval: save ``checkpoint_dir/checkpoint_{step}`` every N steps
    :param eval_interval: evaluate every N steps
    :param pipeline: registered pipeline name
    :param trainer: registered trainer name
    :param trainer_kwargs: extra kwargs for the trainer ctor
    :param project_name/run_name/entity_name/group_name/tags: tracker metadata
    :param checkpoint_dir: where checkpoints go
    :param rollout_logging_dir: if set, PPO rollouts are exported there as json
    :param save_best: keep ``best_checkpoint`` by mean eval reward
    :param save_optimizer: include optimizer/scheduler/RNG state in checkpoints
    :param resume_from_checkpoint: checkpoint directory to restore before training
    :param tracker: ``"wandb"``, ``"tensorboard"``, ``"jsonl"`` or ``None``
    :param logging_dir: directory for tensorboard / jsonl trackers
    :param seed: RNG seed (rank is added, as in the reference)
    :param minibatch_size: micro-batch size for gradient accumulation; must divide batch_size
    :param parallel: :class:`ParallelConfig` (dict accepted)
    """

    total_steps: int
    seq_length: int
    epochs: int
    batch_size: int
    checkpoint_interval: int
    eval_interval: int
    pipeline: str
    trainer: str
    trainer_kwargs: Dict[str, Any] = field(default_factory=dict)
    project_name: str = "trlx"
    run_name: Optional[str] = None
    entity_name: Optional[str] = None
    group_name: Optional[str] = None
    checkpoint_dir: str = "ckpts"
    rollout_logging_dir: Optional[str] = None
    save_best: bool = True
    save_optimizer: bool = True
    resume_from_checkpoint: Optional[str] = None
    tracker: Optional[str] = "wandb"
    logging_dir: Optional[str] = None
    tags: Optional[List[str]] = field(default_factory=list)
    seed: int = 1000
    minibatch_size: Optional[int] = None
    parallel: Any = field(default_factory=ParallelConfig)

    def __post_init__(self):
        if isinstance(self.parallel, dict):
            self.parallel = ParallelConfig.from_dict(self.parallel)
        elif self.parallel is None:
            self.parallel = ParallelConfig()

    def to_dict(self) -> Dict[str, Any]:
        d = super().to_dict()
        d["parallel"] = self.parallel.to_dict()
        return d


_SECTIONS = {
    "model": ModelConfig,
    "optimizer": OptimizerConfig,
    "scheduler": SchedulerConfig,
    "tokenizer": TokenizerConfig,
    "train": TrainConfig,
}


@dataclass
class TRLConfig:
    """Top-level config: ``method`` + the five sections above."""

    method: MethodConfig
    model: ModelConfig
    optimizer: OptimizerConfig
    scheduler: SchedulerConfig
    tokenizer: TokenizerConfig
    train: TrainConfig

    # ---- construction ------------------------------------------------------------------
    @classmethod
    def load_yaml(cls, yml_fp: str) -> "TRLConfig":
        with open(yml_fp, "r") as fh:
            return cls.from_dict(yaml.safe_load(fh))

    @classmethod
    def from_dict(cls, config: Dict) -> "TRLConfig":
        method_cls = get_method(config["method"]["name"])
        kwargs = {name: sec.from_dict(dict(config[name])) for name, sec in _SECTIONS.items()}
        return cls(method=method_cls.from_dict(dict(config["method"])), **kwargs)

    # ---- export ------------------------------------------------------------------------
    def to_dict(self) -> Dict[str, Any]:
        out = {"method": dict(self.method.__dict__)}
        for name in _SECTIONS:
            out[name] = getattr(self, name).to_dict()
        return out

    def evolve(self, **kwargs) -> "TRLConfig":
        """Functional nested update: ``cfg.evolve(method=dict(gamma=0.9), train=dict(seed=1))``."""
        return TRLConfig.from_dict(_merge_dicts(self.to_dict(), kwargs))

    @classmethod
    def update(cls, baseconfig, config: Dict) -> "TRLConfig":
        """Apply overrides given as nested dicts and/or dotted keys (``"train.batch_size": 8``).

        Raises ``ValueError`` when an override names a parameter that does not exist
        (reference behaviour: ``trlx/data/configs.py:303-329``).
        """
        nested: Dict[str, Any] = {}
        for dotted, value in config.items():
            *parents, leaf = dotted.split(".")
            if not parents and not isinstance(value, dict):
                continue  # a bare scalar at top level carries no section → ignored like the reference
            cursor = nested
            for p in parents:
                cursor = cursor.setdefault(p, {})
            if isinstance(value, dict) and isinstance(cursor.get(leaf), dict) and leaf not in _ATOMIC_KEYS:
                cursor[leaf] = _merge_dicts(cursor[leaf], value)
            else:
                cursor[leaf] = value

        base = baseconfig if isinstance(baseconfig, dict) else baseconfig.to_dict()
        base = copy.deepcopy(base)
        seen: Set[str] = set()
        merged = merge(base, nested, seen)
        _assert_all_consumed(nested, merged, path="")
        return cls.from_dict(merged)

    def __str__(self) -> str:
        return json.dumps(self.to_dict(), indent=4, default=str)


def _assert_all_consumed(update: Dict, merged: Dict, path: str) -> None:
    """Stricter typo check than the reference: verifies every *leaf* override landed."""
    for key, val in update.items():
        where = f"{path}{key}"
        if not isinstance(merged, dict) or key not in merged:
            raise ValueError(f"parameter {where} is not present in the config (typo or a wrong config)")
        if isinstance(val, dict) and isinstance(merged[key], dict):
            # free-form dict fields (kwargs, gen_kwargs, peft_config…) may gain new keys
            if key in ("kwargs", "gen_kwargs", "trainer_kwargs", "peft_config", "model_extra_configs",
                       "tokenizer_extra_configs") or key in _ATOMIC_KEYS:
                merged[key].update(val)
                continue
            _assert_all_consumed(val, merged[key], where + ".")
