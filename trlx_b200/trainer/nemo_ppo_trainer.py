"""Megatron-style (tensor / sequence / pipeline parallel) PPO trainer.

Reference counterpart: ``trlx/trainer/nemo_ppo_trainer.py`` + ``trlx/models/modeling_nemo_ppo.py``, which build on
NeMo / Apex and cannot be imported in the reference snapshot.  Here the same trainer logic as
:class:`~trlx_b200.trainer.accelerate_ppo_trainer.AcceleratePPOTrainer` runs on a model whose blocks were sharded by
:func:`trlx_b200.parallel.tensor_parallel.apply_tensor_parallel` over the TP group described by
``config.train.parallel`` (ColumnParallel → RowParallel pairs with fused GEMM↔collective kernels, optional sequence
parallelism), with pipeline stages from :mod:`trlx_b200.parallel.pipeline_parallel`; data parallelism across the
remaining ranks uses the same fused reduce-scatter/AdamW optimizer.  Checkpoints use the
``mp_rank_XX/model_weights.ckpt`` layout (``modeling_nemo_ppo.py:445-495``).
"""
from __future__ import annotations

from trlx_b200.parallel.megatron import MegatronMixin
from trlx_b200.trainer import register_trainer
from trlx_b200.trainer.accelerate_ppo_trainer import AcceleratePPOTrainer


@register_trainer
class NeMoPPOTrainer(MegatronMixin, AcceleratePPOTrainer):
    """PPO with tensor/sequence/pipeline parallelism (``config.train.parallel``)."""


def rank_0_tqdm(*args, **kwargs):
    """``tqdm`` that only draws on the first process of the job (reference ``nemo_ppo_trainer.py:44-49``)."""
    from trlx_b200.utils import logging, rank

    kwargs.setdefault("disable", rank() != 0)
    return logging.tqdm(*args, **kwargs)
