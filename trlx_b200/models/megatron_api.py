"""The Lightning-module surface of the reference's NeMo models (``PPOGPT`` / ``ILQLGPT`` / ``SFTGPT``, e.g.
``trlx/models/modeling_nemo_ppo.py:384-1222``) on this framework's parallel layer: per-replica RNG, Megatron-style
data loaders, ``mp_rank_XX/model_weights.ckpt`` checkpoints, optimizer construction, a micro-batched training step with the
sequence-parallel gradient all-reduce, an inference-mode context and forward closures.

The reference inherits all of this from ``MegatronGPTModel`` + PyTorch Lightning; here it is a mixin over a plain
``nn.Module`` that holds ``self.config`` (a :class:`~trlx_b200.data.configs.TRLConfig`) and implements ``_loss(batch)``.
The trainers (``NeMo*Trainer``) own the outer loop; these methods exist so model-level code written against the reference
(``model.training_step(batch)``, ``model.inference_mode()``, ``model.save_pretrained(dir)`` …) keeps working.
"""
from __future__ import annotations

import contextlib
import os
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch.utils.data import DataLoader

from trlx_b200.parallel import state as parallel_state
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


class MegatronBatchSampler:
    """Index batches for one data-parallel rank (NeMo ``MegatronPretrainingBatchSampler``): the sample stream is cut into
    global batches of ``global_batch_size``; rank ``r`` of ``dp_size`` receives the ``r``-th contiguous slice of each, which the
    training step then splits into micro-batches.  ``consumed_samples`` resumes mid-stream; incomplete tails are dropped."""

    def __init__(self, total_samples: int, consumed_samples: int, micro_batch_size: int, global_batch_size: int,
                 data_parallel_rank: int, data_parallel_size: int, drop_last: bool = True):
        if global_batch_size % (micro_batch_size * data_parallel_size):
            raise ValueError(f"global batch {global_batch_size} is not a multiple of micro batch {micro_batch_size} x "
                             f"data-parallel size {data_parallel_size}")
        self.total, self.consumed = int(total_samples), int(consumed_samples)
        self.micro, self.glob = int(micro_batch_size), int(global_batch_size)
        self.rank, self.size, self.drop_last = int(data_parallel_rank), int(data_parallel_size), drop_last
        self.per_rank = self.glob // self.size

    def __len__(self) -> int:
        n = self.total - self.consumed
        return n // self.glob if self.drop_last else -(-n // self.glob)

    def __iter__(self) -> Iterator[List[int]]:
        start = self.consumed
        while start + self.glob <= self.total:
            lo = start + self.rank * self.per_rank
            yield list(range(lo, lo + self.per_rank))
            start += self.glob
        if not self.drop_last and start < self.total:
            rest = list(range(start, self.total))
            mine = rest[self.rank::self.size]
            if mine:
                yield mine


def unwrap_float16_module(module: nn.Module) -> nn.Module:
    """Strip a half-precision wrapper (Megatron ``Float16Module`` in the reference, ``:315-318``); modules here are cast in
    place, so anything without a ``.module`` attribute is returned as is."""
    inner = getattr(module, "module", None)
    return inner if isinstance(inner, nn.Module) else module


def patch_attention_for_llama(module: nn.Module) -> None:
    """The reference flips NeMo's ``megatron_legacy`` QKV layout for LLaMA checkpoints (``:60-62``).  The QKV layout here is
    part of the architecture spec (``nn/arch.py``) and converted at load time (``nn/hf_compat.py``): nothing to patch."""
    return None


class MegatronModelMixin:
    """See the module docstring.  Expects ``self.config`` (TRLConfig) and ``self._loss(batch) -> (loss, stats)``."""

    metric_fn: Optional[Callable] = None

    # ---- bookkeeping -----------------------------------------------------------------------------------------------------------
    @classmethod
    def list_available_models(cls):
        return None

    def build_train_valid_test_datasets(self):
        """Datasets are injected with :meth:`set_train_dataset` / :meth:`set_valid_dataset` (as in the reference)."""

    def maybe_initalize_per_dp_rng(self, seed: Optional[int] = None) -> torch.Generator:  # (sic) reference spelling, ``:384-393``
        """Generator seeded with ``seed + data-parallel rank``: replicas draw different samples while the model-parallel
        peers of one replica (which must generate identical tokens) share the stream."""
        if getattr(self, "_dp_rng", None) is None:
            st = parallel_state.get_model_parallel()
            base = int(seed if seed is not None else getattr(self.config.train, "seed", 1000))
            self._dp_rng = torch.Generator(device="cpu").manual_seed(base + st.dp_rank)
        return self._dp_rng

    # ---- data ------------------------------------------------------------------------------------------------------------------
    def _batch_sizes(self) -> Tuple[int, int]:
        st = parallel_state.get_model_parallel()
        micro = int(self.config.train.minibatch_size or self.config.train.batch_size)
        return micro, int(self.config.train.batch_size) * max(st.dp_size, 1)

    def build_data_loader(self, dataset, collate_fn, consumed_samples: int = 0) -> DataLoader:
        st = parallel_state.get_model_parallel()
        micro, glob = self._batch_sizes()
        sampler = MegatronBatchSampler(len(dataset), consumed_samples, micro, glob, st.dp_rank, max(st.dp_size, 1))
        return DataLoader(dataset, batch_sampler=sampler, num_workers=0, pin_memory=torch.cuda.is_available(), collate_fn=collate_fn)

    def set_train_dataset(self, train_dataset, collate_fn):
        self._train_dataset, self._train_collate_fn = train_dataset, collate_fn

    def set_valid_dataset(self, valid_dataset, collate_fn):
        self._valid_dataset, self._valid_collate_fn = valid_dataset, collate_fn

    def setup_training_data(self, _=None):
        if hasattr(self, "_train_dataset"):
            self._train_dl = self.build_data_loader(self._train_dataset, self._train_collate_fn)

    def setup_validation_data(self, _=None):
        if hasattr(self, "_valid_dataset"):
            self._validation_dl = self.build_data_loader(self._valid_dataset, self._valid_collate_fn)

    # ---- checkpoints -----------------------------------------------------------------------------------------------------------
    @staticmethod
    def _rank_dir(directory: str) -> str:
        st = parallel_state.get_model_parallel()
        name = f"mp_rank_{st.tp_rank:02d}" + (f"_{st.pp_rank:03d}" if st.pp_size > 1 else "")
        return os.path.join(directory, name)

    def save_pretrained(self, directory: str) -> str:
        """``<directory>/mp_rank_XX[_YYY]/model_weights.ckpt`` written by the first data-parallel replica (reference ``:445-470``)."""
        st = parallel_state.get_model_parallel()
        path = os.path.join(self._rank_dir(directory), "model_weights.ckpt")
        if st.dp_rank == 0:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save(self.state_dict(), path)
        return path

    def load_from_pretrained(self, directory: str, strict: bool = False):
        """Load this rank's shard; a checkpoint written without pipeline parallelism is re-sharded for the current stage
        (reference ``:472-495`` + ``reshard_for_pipeline_parallelism``)."""
        from trlx_b200.models.modeling_nemo_ppo import reshard_for_pipeline_parallelism

        st = parallel_state.get_model_parallel()
        path = os.path.join(self._rank_dir(directory), "model_weights.ckpt")
        reshard = False
        if not os.path.exists(path) and st.pp_size > 1:
            path = os.path.join(directory, f"mp_rank_{st.tp_rank:02d}", "model_weights.ckpt")
            reshard = True
        sd = torch.load(path, map_location="cpu", weights_only=False)
        if reshard:
            n_layers = max((int(k.split(".h.")[1].split(".")[0]) for k in sd if ".h." in k), default=-1) + 1
            sd = reshard_for_pipeline_parallelism(n_layers, sd)
        return self.load_state_dict(sd, strict=strict)

    def model_provider_func(self, pre_process: bool = True, post_process: bool = True) -> nn.Module:
        """The wrapped (already sharded / frozen) module — construction happens in ``__init__`` here."""
        return self

    # ---- optimisation ----------------------------------------------------------------------------------------------------------
    def setup_optimizer_param_groups(self):
        self._optimizer_param_groups = [{"params": [p for p in self.parameters() if p.requires_grad]}]
        return self._optimizer_param_groups

    def configure_optimizers(self):
        """``(optimizer, scheduler)`` from ``config.optimizer`` / ``config.scheduler`` (reference ``:538-610`` configures Apex
        distributed Adam buckets; the flat-buffer fused AdamW of ``parallel/optim.py`` is what the trainers use)."""
        from trlx_b200.utils import get_optimizer_class, get_scheduler_class

        groups = self.setup_optimizer_param_groups()
        opt = get_optimizer_class(self.config.optimizer.name)(groups[0]["params"], **self.config.optimizer.kwargs)
        sched = get_scheduler_class(self.config.scheduler.name)(opt, **self.config.scheduler.kwargs)
        return opt, sched

    def allreduce_sequence_parallel_gradients(self) -> None:
        """Sum the gradients of parameters that are replicated inside a sequence-parallel region (norms) over the
        tensor-parallel group (reference ``:612-645``)."""
        st = parallel_state.get_model_parallel()
        if st.tp_group is None or st.tp_size == 1:
            return
        from trlx_b200.parallel.tensor_parallel import allreduce_sequence_parallel_grads

        allreduce_sequence_parallel_grads(self, st.tp_group)

    def activation_checkpointing_(self, enabled: bool) -> None:
        for m in self.modules():
            if hasattr(m, "gradient_checkpointing_enable") and m is not self:
                (m.gradient_checkpointing_enable if enabled else m.gradient_checkpointing_disable)()

    def sequence_parallel_(self, enabled: bool) -> None:
        """Sequence parallelism is a property of how the blocks were sharded (``apply_tensor_parallel``); generation runs with
        it in place, so unlike the reference (``:820-836``) there is nothing to switch off around inference."""
        self._sequence_parallel_requested = bool(enabled)

    def get_forward_output_and_loss_func(self):
        """``fwd(batch) -> (loss, stats)`` closure (reference ``:945-1026`` returns Megatron's fwd/loss pair)."""
        return lambda batch: self._loss(batch)

    def get_forward_output_only_func(self):
        def fwd(batch):
            with torch.no_grad():
                return self(**batch) if isinstance(batch, dict) else self(*batch)
        return fwd

    def training_step(self, batch, optimizer=None, micro_batches: Optional[Sequence[Any]] = None) -> Dict[str, Any]:
        """Gradient accumulation over the micro-batches of one optimizer step (``micro_batches``, default: ``batch`` as a
        single one), sequence-parallel gradient all-reduce, optional ``optimizer.step()``; returns the mean loss and the
        last micro-batch's statistics."""
        self.train()
        parts = list(micro_batches) if micro_batches is not None else [batch]
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        total, stats = 0.0, {}
        fwd = self.get_forward_output_and_loss_func()
        for mb in parts:
            loss, stats = fwd(mb)
            (loss / len(parts)).backward()
            total += float(loss.detach())
        self.allreduce_sequence_parallel_gradients()
        if optimizer is not None:
            optimizer.step()
        out = dict(stats or {})
        out["loss"] = total / len(parts)
        return out

    @contextlib.contextmanager
    def inference_mode(self):
        """Eval mode, no autograd, activation checkpointing off; restored afterwards (reference ``:838-870``)."""
        was_training = self.training
        ckpt = [m for m in self.modules() if getattr(m, "gradient_checkpointing", False)]
        self.eval()
        for m in ckpt:
            m.gradient_checkpointing_disable()
        try:
            with torch.no_grad():
                yield self
        finally:
            for m in ckpt:
                m.gradient_checkpointing_enable()
            self.train(was_training)

    def validation_step(self, batch, batch_idx: int = 0):
        """Generate continuations for a batch of prompts (``input_ids``, ``attention_mask``) in inference mode."""
        gen_kwargs = dict(getattr(self.config.method, "gen_kwargs", {}) or {})
        with self.inference_mode():
            ids = batch["input_ids"] if isinstance(batch, dict) else batch[0]
            am = batch.get("attention_mask") if isinstance(batch, dict) else None
            return self.generate(ids, attention_mask=am, **gen_kwargs)

    def validation_epoch_end(self, outputs: List[torch.Tensor], tokenizer=None) -> Dict[str, Any]:
        """Decode the generations and apply ``metric_fn`` (reference ``:905-927``)."""
        if tokenizer is None or self.metric_fn is None:
            return {"n_samples": sum(len(o) for o in outputs)}
        texts = [t for o in outputs for t in tokenizer.batch_decode(o, skip_special_tokens=True)]
        metrics = self.metric_fn(samples=texts)
        return {f"metrics/{k}": (sum(v) / max(len(v), 1) if hasattr(v, "__len__") else v) for k, v in metrics.items()}

    def free_kv_cache(self) -> None:
        """Generation caches live in the rollout engine's paged allocator, not in the module (reference ``:870``)."""
        if torch.cuda.is_available():
            torch.cuda.empty_cache()


class PipelineTensorAPI:
    """The two Megatron module hooks the reference's head wrappers forward to their language model
    (``trlx/models/modeling_nemo_ppo.py:198-203``): ``set_input_tensor`` (the activation handed over by the previous pipeline
    stage) and ``word_embeddings_weight`` (the tensor the first / last stage keep in sync when embeddings are tied).  Expects
    ``self.language_model``."""

    def _causal_lm(self):
        from trlx_b200.models.modeling_base import base_lm

        return base_lm(self.language_model)

    def set_input_tensor(self, input_tensor) -> None:
        stage = getattr(self._causal_lm(), "_pp", None)
        if stage is None:
            if input_tensor is not None:
                raise RuntimeError("set_input_tensor: this model holds no pipeline stage")
            return
        stage.input_tensor = input_tensor

    def word_embeddings_weight(self):
        return self._causal_lm().transformer.wte.weight
