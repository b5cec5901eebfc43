"""Megatron-style (tensor / sequence / pipeline parallel) SFT trainer.

Reference counterpart: ``trlx/trainer/nemo_sft_trainer.py`` + ``trlx/models/modeling_nemo_sft.py``, which build on
NeMo / Apex and cannot be imported in the reference snapshot.  Here the same trainer logic as
:class:`~trlx_b200.trainer.accelerate_sft_trainer.AccelerateSFTTrainer` runs on a model whose blocks were sharded by
:func:`trlx_b200.parallel.tensor_parallel.apply_tensor_parallel` over the TP group described by
``config.train.parallel`` (ColumnParallel → RowParallel pairs with fused GEMM↔collective kernels, optional sequence
parallelism), with pipeline stages from :mod:`trlx_b200.parallel.pipeline_parallel`; data parallelism across the
remaining ranks uses the same fused reduce-scatter/AdamW optimizer.  Checkpoints use the
``mp_rank_XX/model_weights.ckpt`` layout (``modeling_nemo_ppo.py:445-495``).
"""
from __future__ import annotations

from trlx_b200.parallel.megatron import MegatronMixin
from trlx_b200.trainer import register_trainer
from trlx_b200.trainer.accelerate_sft_trainer import AccelerateSFTTrainer


@register_trainer
class NeMoSFTTrainer(MegatronMixin, AccelerateSFTTrainer):
    """SFT with tensor/sequence/pipeline parallelism (``config.train.parallel``)."""
