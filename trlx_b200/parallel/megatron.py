"""Glue that turns an ``Accelerate*Trainer`` into its tensor/pipeline-parallel (``NeMo*Trainer``) variant."""
from __future__ import annotations

import os
from typing import Optional

import torch

from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


class MegatronMixin:
    """Overrides model setup (shard after construction), the optimizer step under pipeline parallelism (1F1B over the
    micro-batches, interleaved when ``virtual_pipeline_parallel`` > 1) and checkpoint IO (``mp_rank_XX`` layout)."""

    def __init__(self, config, **kwargs):
        tk = dict(config.train.trainer_kwargs or {})
        for key in ("megatron_cfg", "pretrained_model"):  # reference-style trainer kwargs are consumed here
            if key in kwargs:
                tk.setdefault(key, kwargs.pop(key))
        if tk.get("megatron_cfg") is not None:  # NeMo-style recipe (reference trainer_kwargs): layout, architecture, optimizer
            from trlx_b200.parallel.megatron_cfg import apply_megatron_cfg

            config = apply_megatron_cfg(config, tk["megatron_cfg"], tk.get("pretrained_model"))
            if isinstance(config.model.model_path, dict):  # one CPU process cannot hold a 20B recipe: plumbing-run size
                from trlx_b200.parallel.megatron_cfg import debug_sized

                config = config.evolve(model=dict(model_path=debug_sized(config.model.model_path)))
            try:  # step / wall-clock limits of the recipe's ``trainer`` section (PTL Trainer + StatelessTimer in the reference)
                from trlx_b200.trainer.nemo_ilql_trainer import megatron_trainer

                plan = megatron_trainer(tk["megatron_cfg"], seed_everything=False)  # ``train()`` seeded already
                self._max_time = plan.max_time
                if plan.max_steps:
                    config = config.evolve(train=dict(total_steps=min(int(plan.max_steps), int(config.train.total_steps))))
            except (KeyError, ValueError, TypeError, OSError) as err:
                logger.warning(f"megatron_cfg: trainer section not applied ({err})")
        par = config.train.parallel
        if int(getattr(par, "tensor_parallel", 1) or 1) > 1 and bool(getattr(par, "sequence_parallel", False)):
            # with sequence parallelism the activation at the branch point is a per-rank sequence shard whose layout depends on
            # the length of the forward that produced it: it cannot be stored per rollout and re-sliced per minibatch
            config = config.evolve(train=dict(trainer_kwargs=dict(cache_trunk=False)))
        pp = int(getattr(config.train.parallel, "pipeline_parallel", 1) or 1)
        if pp > 1:
            if config.model.model_arch_type == "seq2seq":
                raise NotImplementedError("pipeline parallelism covers decoder-only models")
            if config.model.num_layers_unfrozen > 0:
                logger.warning("pipeline parallelism: the hydra reference branch is replaced by a separate, equally "
                               "partitioned reference model (num_layers_unfrozen → -1)")
                config = config.evolve(model=dict(num_layers_unfrozen=-1))
        super().__init__(config, **kwargs)
        self._pp_stage = getattr(self, "_pp_stage", None)

    def _shard_like_policy(self, model) -> None:
        """Give a second model (the separate PPO reference) the policy's tensor / pipeline layout; called before its weights are
        copied from the (already sharded) policy."""
        if getattr(model, "_model_parallel_applied", False):
            return
        rt = self.runtime
        if rt.tp_size > 1:
            from trlx_b200.parallel.tensor_parallel import apply_tensor_parallel

            apply_tensor_parallel(model, rt.tp_group, rt.tp_rank, rt.tp_size,
                                  sequence_parallel=bool(self.config.train.parallel.sequence_parallel))
        if rt.pp_size > 1:
            from trlx_b200.parallel.pipeline_parallel import apply_pipeline_parallel

            apply_pipeline_parallel(model, rt.pp_group, rt.pp_rank, rt.pp_size,
                                    int(getattr(self.config.train.parallel, "virtual_pipeline_parallel", 1) or 1))
        model._model_parallel_applied = True

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

            self._pp_stage = apply_pipeline_parallel(model, rt.pp_group, rt.pp_rank, rt.pp_size,
                                                     int(getattr(self.config.train.parallel, "virtual_pipeline_parallel", 1) or 1))
        return model

    def _optimizer_param_groups(self, params, optimizer_class):
        """With sequence parallelism the block norms are replicated inside the TP group but see sequence-sharded activations:
        every TP rank holds a PARTIAL gradient for them.  Instead of a separate flatten → all-reduce → copy pass before the
        optimizer (the reference pattern, ``modeling_nemo_ppo.py:627-645``), they form their own parameter group whose fused
        reduce-scatter + AdamW + all-gather runs over the whole stage (TP x DP ranks) and divides by DP only — the TP sum
        happens inside the optimizer kernel."""
        from trlx_b200.parallel.optim import FusedAdamW

        rt = self.runtime
        self._sp_grads_in_optimizer = False
        if not (issubclass(optimizer_class, FusedAdamW) and rt.tp_size > 1 and self.config.train.parallel.sequence_parallel
                and getattr(self, "zero3", None) is None and getattr(rt, "stage_group", None) is not None):
            return params
        from trlx_b200.parallel.tensor_parallel import sequence_parallel_grad_params

        shared = {id(p) for p in sequence_parallel_grad_params(self.model)}
        sp = [p for p in params if id(p) in shared]
        rest = [p for p in params if id(p) not in shared]
        if not sp:
            return params
        self._sp_grads_in_optimizer = True
        groups = [{"params": sp, "reduce_group": rt.stage_group, "grad_divisor": float(rt.dp_size)}]
        if rest:
            groups.insert(0, {"params": rest})
        return groups

    # ---- pipeline-parallel optimizer step -------------------------------------------------------------------------------
    def train_step(self, minibatch):
        stage = getattr(self, "_pp_stage", None)
        if stage is None:
            return super().train_step(minibatch)
        from trlx_b200.parallel import pipeline_parallel as pp

        microbatches = list(minibatch)
        for _ in microbatches:
            self.mb_count += 1
        per_mb = pp.run_schedule(stage, microbatches, self.loss, self.runtime.device,
                                 before_backward=self.model.train, after_backward=self.model.eval)
        stats = None
        if stage.last:
            stats = {k: sum(s[k] for s in per_mb) / self.num_mb for k in per_mb[0]}
        stats = pp.broadcast_stats(stage, stats, self.runtime.device)
        self._pre_optimizer_step()
        self.opt.step()
        self.opt.zero_grad()
        self.scheduler.step()
        self.iter_count += 1
        self._after_weights_changed()
        stats["time/forward"] = 0.0
        stats["time/backward"] = 0.0
        return stats

    def _after_weights_changed(self):
        super()._after_weights_changed()
        stage = getattr(self, "_pp_stage", None)
        if stage is not None:
            from trlx_b200.parallel.pipeline_parallel import broadcast_module_from_last

            heads = [getattr(self.model, n) for n in ("v_head", "ilql_heads") if getattr(self.model, n, None) is not None]
            broadcast_module_from_last(stage, heads)

    def _pre_optimizer_step(self):
        rt = self.runtime
        stage = getattr(self, "_pp_stage", None)
        if stage is not None:
            from trlx_b200.parallel.pipeline_parallel import allreduce_tied_embedding_grads

            allreduce_tied_embedding_grads(stage)
        if rt.tp_size > 1 and self.config.train.parallel.sequence_parallel and not getattr(self, "_sp_grads_in_optimizer", False):
            from trlx_b200.parallel.tensor_parallel import allreduce_sequence_parallel_grads

            allreduce_sequence_parallel_grads(self.model, rt.tp_group)

    def _mp_subdir(self) -> str:
        """Megatron naming: ``mp_rank_<tp>`` and, with pipeline stages, ``mp_rank_<tp>_<pp>``."""
        rt = self.runtime
        return f"mp_rank_{rt.tp_rank:02d}" + (f"_{rt.pp_rank:03d}" if rt.pp_size > 1 else "")

    def save_pretrained(self, directory: Optional[str] = None, **kwargs):
        """``<dir>/mp_rank_XX[_YYY]/model_weights.ckpt`` per model-parallel rank (single file when TP == PP == 1).
        The reference refuses to save with PP > 1 (``modeling_nemo_ppo.py:448-450``); here every stage writes its slice."""
        rt = self.runtime
        if rt.tp_size == 1 and rt.pp_size == 1:
            return super().save_pretrained(directory, **kwargs)
        from trlx_b200.utils import resolve_output_dir

        directory = resolve_output_dir(directory or os.path.join(self.config.train.checkpoint_dir, "hf_model"))
        rt.barrier()
        if rt.dp_rank == 0:
            sub = os.path.join(directory, self._mp_subdir())
            os.makedirs(sub, exist_ok=True)
            torch.save({k: v.detach().cpu() for k, v in self.model.raw_state_dict().items() if v.numel()},
                       os.path.join(sub, "model_weights.ckpt"))
        rt.barrier()

    def load_from_pretrained(self, directory: str):
        rt = self.runtime
        sub = os.path.join(directory, self._mp_subdir()) if (rt.tp_size > 1 or rt.pp_size > 1) else directory
        sd = torch.load(os.path.join(sub, "model_weights.ckpt"), map_location="cpu", weights_only=True)
        own = self.model.raw_state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                if k in own:
                    own[k].copy_(v.to(own[k].dtype))
        if hasattr(self.opt, "resync_master"):
            self.opt.resync_master(reset_moments=True)  # fresh weights: the fp32 master copy and moments restart from them
        self._after_weights_changed()
        rt.barrier()
