"""Megatron-style (tensor / sequence / pipeline parallel) ILQL trainer.

Reference counterpart: ``trlx/trainer/nemo_ilql_trainer.py`` + ``trlx/models/modeling_nemo_ilql.py``, which build on
NeMo / Apex and cannot be imported in the reference snapshot.  Here the same trainer logic as
:class:`~trlx_b200.trainer.accelerate_ilql_trainer.AccelerateILQLTrainer` runs on a model whose blocks were sharded by
:func:`trlx_b200.parallel.tensor_parallel.apply_tensor_parallel` over the TP group described by
``config.train.parallel`` (ColumnParallel → RowParallel pairs with fused GEMM↔collective kernels, optional sequence
parallelism), with pipeline stages from :mod:`trlx_b200.parallel.pipeline_parallel`; data parallelism across the
remaining ranks uses the same fused reduce-scatter/AdamW optimizer.  Checkpoints use the
``mp_rank_XX/model_weights.ckpt`` layout (``modeling_nemo_ppo.py:445-495``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Sequence

import torch

from trlx_b200.parallel.megatron import MegatronMixin
from trlx_b200.trainer import register_trainer
from trlx_b200.trainer.accelerate_ilql_trainer import AccelerateILQLTrainer


@register_trainer
class NeMoILQLTrainer(MegatronMixin, AccelerateILQLTrainer):
    """ILQL with tensor/sequence/pipeline parallelism (``config.train.parallel``)."""


@dataclass
class MegatronTrainerPlan:
    """What the reference's ``megatron_trainer(cfg)`` (``trlx/trainer/nemo_ilql_trainer.py:31-82``) hands back as a PyTorch
    Lightning ``Trainer`` + ``NLPDDPStrategy`` + precision plugin + ``exp_manager``: here a plain record of the resolved
    run settings.  The training loop itself is the trainer's own ``learn()``."""

    devices: int = 1
    num_nodes: int = 1
    precision: str = "bf16"
    max_steps: Optional[int] = None
    max_time: Optional[str] = None            # wall-clock limit "DD:HH:MM:SS" (PTL ``StatelessTimer`` in the reference)
    val_check_interval: Optional[int] = None
    log_every_n_steps: Optional[int] = None
    parallel: Dict[str, Any] = field(default_factory=dict)   # tensor / pipeline / sequence parallel sizes, precision
    grad_scaler: Optional[Dict[str, Any]] = None              # fp16 only: init scale / growth interval / hysteresis
    distributed_optimizer: bool = False
    seed: int = 1000
    resume_from_checkpoint: Optional[str] = None
    exp_dir: Optional[str] = None

    def max_time_seconds(self) -> Optional[float]:
        if not self.max_time:
            return None
        d, h, m, s_ = (int(x) for x in str(self.max_time).split(":"))
        return ((d * 24 + h) * 60 + m) * 60 + s_


def megatron_trainer(cfg, seed_everything: bool = True) -> MegatronTrainerPlan:
    """Resolve a NeMo-style recipe (path / mapping with ``trainer``, ``model``, ``exp_manager`` sections) into a
    :class:`MegatronTrainerPlan`: seeds the process (``model.seed``, default 1000), picks the precision and — for fp16 — the
    loss-scaler settings, the model-parallel layout, the step / wall-clock limits and the checkpoint to resume from
    (``model.resume_from_checkpoint`` or the newest checkpoint under ``exp_manager.explicit_log_dir`` when
    ``resume_if_exists``)."""
    import glob
    import os

    from trlx_b200.parallel.megatron_cfg import _load, parse_megatron_cfg
    from trlx_b200.utils import set_seed

    raw = _load(cfg)
    model, tr, exp = raw.get("model", {}) or {}, raw.get("trainer", {}) or {}, raw.get("exp_manager", {}) or {}
    seed = int(model.get("seed", 1000))
    rec = parse_megatron_cfg(raw)
    if seed_everything:  # tensor / pipeline-parallel peers share a stream, data-parallel replicas differ
        from types import SimpleNamespace

        set_seed(seed, SimpleNamespace(**rec["parallel"]))
    precision = rec["precision"]
    scaler = None
    if precision == "fp16":
        scaler = dict(init_scale=model.get("native_amp_init_scale", 2 ** 32),
                      growth_interval=model.get("native_amp_growth_interval", 1000), hysteresis=model.get("hysteresis", 2))
    resume = model.get("resume_from_checkpoint")
    exp_dir = exp.get("explicit_log_dir") or exp.get("exp_dir")
    if resume is None and exp.get("resume_if_exists") and exp_dir and os.path.isdir(str(exp_dir)):
        found = sorted(glob.glob(os.path.join(str(exp_dir), "**", "checkpoint_*"), recursive=True), key=os.path.getmtime)
        resume = found[-1] if found else None
    return MegatronTrainerPlan(
        devices=int(tr.get("devices", 1) or 1), num_nodes=int(tr.get("num_nodes", 1) or 1), precision=precision,
        max_steps=tr.get("max_steps"), max_time=tr.get("max_time"), val_check_interval=tr.get("val_check_interval"),
        log_every_n_steps=tr.get("log_every_n_steps"), parallel=rec["parallel"], grad_scaler=scaler,
        distributed_optimizer=(model.get("optim") or {}).get("name") == "distributed_fused_adam", seed=seed,
        resume_from_checkpoint=resume, exp_dir=exp_dir)


class ShuffledCyclicSequence:
    """A fixed random permutation of ``new_length`` indices over ``data`` repeated cyclically (reference ``:85-98``): lets a
    small dataset fill exactly the number of samples a fixed-step schedule consumes."""

    def __init__(self, new_length: int, data: Sequence, seed: int):
        self.data, self.new_length = data, int(new_length)
        self.perm = torch.randperm(self.new_length, generator=torch.Generator().manual_seed(int(seed)), device="cpu")

    def __len__(self) -> int:
        return self.new_length

    def __getitem__(self, idx: int):
        return self.data[int(self.perm[idx]) % len(self.data)]
