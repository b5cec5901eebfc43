"""Glue that turns an ``Accelerate*Trainer`` into its tensor/pipeline-parallel (``NeMo*Trainer``) variant."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


class MegatronMixin:
    """Overrides model setup (shard after construction) and checkpoint IO (``mp_rank_XX`` layout)."""

    def setup_model(self):
        model = super().setup_model()
        rt = self.runtime
        if rt.tp_size > 1:
            from trlx_b200.parallel.tensor_parallel import apply_tensor_parallel

            apply_tensor_parallel(model, rt.tp_group, rt.tp_rank, rt.tp_size,
                                  sequence_parallel=bool(self.config.train.parallel.sequence_parallel))
            logger.info(f"model sharded: tp={rt.tp_size} sp={self.config.train.parallel.sequence_parallel} dp={rt.dp_size}")
        if rt.pp_size > 1:
            from trlx_b200.parallel.pipeline_parallel import apply_pipeline_parallel

            apply_pipeline_parallel(model, rt.pp_group, rt.pp_rank, rt.pp_size)
        return model

    def _pre_optimizer_step(self):
        rt = self.runtime
        if rt.tp_size > 1 and self.config.train.parallel.sequence_parallel:
            from trlx_b200.parallel.tensor_parallel import allreduce_sequence_parallel_grads

            allreduce_sequence_parallel_grads(self.model, rt.tp_group)

    def save_pretrained(self, directory: Optional[str] = None, **kwargs):
        """``<dir>/mp_rank_XX/model_weights.ckpt`` per tensor-parallel rank (single file when TP == 1)."""
        rt = self.runtime
        if rt.tp_size == 1 and rt.pp_size == 1:
            return super().save_pretrained(directory, **kwargs)
        if rt.pp_size > 1:
            raise NotImplementedError("saving with pipeline parallelism > 1 is not supported (same as the reference, "
                                      "modeling_nemo_ppo.py:448-450); gather to PP=1 first")
        directory = directory or os.path.join(self.config.train.checkpoint_dir, "hf_model")
        rt.barrier()
        if rt.dp_rank == 0:
            sub = os.path.join(directory, f"mp_rank_{rt.tp_rank:02d}")
            os.makedirs(sub, exist_ok=True)
            torch.save({k: v.detach().cpu() for k, v in self.model.raw_state_dict().items()},
                       os.path.join(sub, "model_weights.ckpt"))
        rt.barrier()

    def load_from_pretrained(self, directory: str):
        rt = self.runtime
        sub = os.path.join(directory, f"mp_rank_{rt.tp_rank:02d}") if rt.tp_size > 1 else directory
        sd = torch.load(os.path.join(sub, "model_weights.ckpt"), map_location="cpu", weights_only=True)
        own = self.model.raw_state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                if k in own:
                    own[k].copy_(v.to(own[k].dtype))
