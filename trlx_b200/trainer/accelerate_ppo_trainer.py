"""PPO trainer.

Parity: ``trlx/trainer/accelerate_ppo_trainer.py`` — ctor ``:42-106``, ``get_arch`` ``:108-125``, ``loss`` ``:127-204``,
rollout logging ``:206-217``, callbacks ``:219-231``, ``prepare_learning`` ``:233-243``, ``add_prompt_pipeline``
``:245-249``, ``make_experience`` ``:251-524``, ``save_pretrained`` ``:526-553``.  The index arithmetic of the rollout
(``start = prompt_len − 1``, ``ends``, EOS inclusion, score placement) is reproduced exactly, but
vectorised and device-resident:

* rollouts come from :class:`trlx_b200.engine.rollout.RolloutEngine` on CUDA (CUDA-graph decode over the sm_100a
  kernels that already emits log-prob, reference log-prob and value of every sampled token — the reference's second
  and third full forward passes disappear), or from the PyTorch sampler + one shared-trunk scoring pass elsewhere;
* every rank scores its own samples with ``reward_fn`` by default (the reference gathers everything to rank 0 and
  idles the others, ``:298-340``); ``trainer_kwargs={"rank0_reward": True}`` restores the funnel;
* experience is pushed as one dense :class:`~trlx_b200.pipeline.ppo_pipeline.RolloutBlock` (optionally carrying the
  frozen-trunk activations) — no ``.cpu()``, no per-sample Python slicing;
* during ``learn`` the cached trunk activation is reused across all ``ppo_epochs`` × minibatches (valid because the
  forward runs in eval mode and the trunk is frozen), so only the unfrozen top blocks + heads are recomputed.
"""
from __future__ import annotations

import json
import os
import uuid
from time import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from trlx_b200 import ops
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.ppo_types import PPORLBatch
from trlx_b200.models.modeling_ppo import (AdaptiveKLController, AutoModelForCausalLMWithHydraValueHead,
                                           AutoModelForSeq2SeqLMWithHydraValueHead, FixedKLController)
from trlx_b200.pipeline.offline_pipeline import PromptPipeline, pad_rows
from trlx_b200.pipeline.ppo_pipeline import PPORolloutStorage, RolloutBlock
from trlx_b200.trainer import register_trainer
from trlx_b200.trainer.accelerate_base_trainer import AccelerateRLTrainer
from trlx_b200.utils import Clock, infinite_dataloader, logging
from trlx_b200.utils.modeling import RunningMoments, gather_dict, logprobs_of_labels

logger = logging.get_logger(__name__)


def _is_config_like(obj) -> bool:
    return isinstance(obj, dict) or (hasattr(obj, "model_type") and not isinstance(obj, str)) or hasattr(obj, "family")


def _bucket_prompt_width(width: int, fixed: Optional[int] = None) -> int:
    """Padded prompt width used for rollouts.  CUDA graphs (decode step, prefill, training step) are keyed by shape, so
    widths are quantised: a fixed ``prompt_bucket`` when configured, otherwise geometric buckets — multiples of
    ``max(8, 2^floor(log2 w) / 4)`` — which keep the padding under 25 % and the number of distinct shapes logarithmic
    in the longest prompt (8, 16, 24, 32, 40, ... 64, 80, 96, 112, 128, 160, ...)."""
    width = max(int(width), 1)
    step = int(fixed) if fixed else max(8, (1 << (width.bit_length() - 1)) // 4)
    return -(-width // step) * step


class _SwappedReference:
    """Callable stand-in for a separate reference model (``trainer_kwargs.offload_reference``): forwards run through the
    *policy's* modules while the reference weights — kept as pinned host copies by
    :class:`~trlx_b200.models.modeling_nemo_ppo.RefLMHeads` — are swapped in, and the policy weights are restored afterwards.
    Costs two host↔device copies of the trainable parameters per call, saves one full model of device memory."""

    def __init__(self, model):
        from trlx_b200.models.modeling_nemo_ppo import RefLMHeads

        self.model = model
        self.swap = RefLMHeads(model.base_model, None, build_reference_model=True)

    @torch.no_grad()
    def __call__(self, *args, **kwargs):
        was_training = self.model.training
        self.model.eval()
        try:
            with self.swap.reference():
                return self.model(*args, **kwargs)
        finally:
            self.model.train(was_training)

    # the trainer treats ``ref_model`` like a module in a few places (device moves, eval, checkpoint filters)
    def eval(self):
        return self

    def requires_grad_(self, *_):
        return self

    def to(self, *_, **__):
        return self

    def parameters(self):
        return iter(())


@register_trainer
class AcceleratePPOTrainer(AccelerateRLTrainer):
    """PPO on the B200 runtime."""

    def __init__(self, config: TRLConfig, **kwargs):
        self.generate_kwargs = None
        super().__init__(config, **kwargs)
        if config.train.rollout_logging_dir is not None:
            self.log_rollouts = True
            self.setup_rollout_logging(config)
        else:
            self.log_rollouts = False

        self.store = PPORolloutStorage(self.tokenizer.pad_token_id, self.tokenizer.padding_side)
        self.store.clear_history()

        # a separate full reference model is only needed when neither a frozen branch nor an adapter provides one
        self.ref_model = None
        if getattr(self.model, "frozen_head", None) is None and not self.model.peft_type and \
                config.train.trainer_kwargs.get("offload_reference", False):
            # full fine-tuning without a second model in HBM: the reference policy lives in pinned host memory and is swapped
            # into the policy's own modules for the reference forward (reference: RefLMHeads, modeling_nemo_ppo.py:167-312)
            self.ref_model = _SwappedReference(self.model)
        elif getattr(self.model, "frozen_head", None) is None and not self.model.peft_type:
            self.ref_model = self.get_arch(self.config).to(self.runtime.device)
            if self.runtime.cuda and self.runtime.dtype != torch.float32:
                self.ref_model = self.ref_model.to(self.runtime.dtype)
            if hasattr(self, "_shard_like_policy"):  # model-parallel trainers: same TP / PP layout before copying the weights
                self._shard_like_policy(self.ref_model)
            self.ref_model.load_state_dict({k: v for k, v in self.model.raw_state_dict().items() if v.numel()}, strict=False)
            self.ref_model.eval().requires_grad_(False)

        if config.method.target is not None:
            self.kl_ctl = AdaptiveKLController(config.method.init_kl_coef, config.method.target, config.method.horizon)
        else:
            self.kl_ctl = FixedKLController(config.method.init_kl_coef)

        base_gen = dict(do_sample=True, use_cache=True, eos_token_id=self.tokenizer.eos_token_id,
                        pad_token_id=self.tokenizer.pad_token_id)
        self.generate_kwargs = {**base_gen, **config.method.gen_kwargs}
        if self.generate_sweep_kwarg is not None:
            self.generate_kwargs.pop(self.generate_sweep_kwarg[0], None)
        self.generate_experience_kwargs = ({**base_gen, **config.method.gen_experience_kwargs}
                                           if config.method.gen_experience_kwargs is not None else None)

        self.running_moments = RunningMoments()
        self.ref_mean = config.method.ref_mean
        self.ref_std = config.method.ref_std
        self.mean_kl = 0.0
        self.rank0_reward = bool(config.train.trainer_kwargs.get("rank0_reward", False))
        self.cache_trunk = bool(config.train.trainer_kwargs.get("cache_trunk", True))
        self._engine = None
        self._engine_failed = False
        self._graphed_steps = {}

    # ---- model ------------------------------------------------------------------------------------------------------------
    def get_arch(self, config: TRLConfig):
        model_class = (AutoModelForSeq2SeqLMWithHydraValueHead if config.model.model_arch_type == "seq2seq"
                       else AutoModelForCausalLMWithHydraValueHead)
        from_fn = model_class.from_config if _is_config_like(config.model.model_path) else model_class.from_pretrained
        return from_fn(config.model.model_path, num_layers_unfrozen=config.model.num_layers_unfrozen,
                       num_value_layers_unfrozen=config.method.num_value_layers_unfrozen,
                       peft_config=config.model.peft_config, **config.model.model_extra_configs)

    def setup_model(self):
        model = super().setup_model()
        if hasattr(model, "freeze_trunk_parameters"):
            n = model.freeze_trunk_parameters()
            if n:
                logger.info(f"shared frozen trunk: {n} parameter tensors below the branch layer (e.g. learned position "
                            "embeddings) are frozen explicitly; TRLX_B200_SHARE_TRUNK=0 restores full-trunk gradients")
        return model

    # ---- loss -------------------------------------------------------------------------------------------------------------
    def loss(self, batch: PPORLBatch) -> Tuple[torch.Tensor, Dict[str, Any]]:
        dev = self.runtime.device
        pad = self.tokenizer.pad_token_id
        query, response = batch.query_tensors.to(dev), batch.response_tensors.to(dev)
        old_logprobs, old_values, old_rewards = batch.logprobs.to(dev), batch.values.to(dev), batch.rewards.to(dev)
        width_tensor = getattr(batch, "width_tensor", None)  # device-side effective width (CUDA-graph replay)
        width = getattr(batch, "width", None)
        trunk_in = getattr(batch, "trunk_hidden", None)
        if width_tensor is None and width is not None and width < old_rewards.shape[1]:
            # a static-shape (full block width) batch on the eager path: trim to the widest response of THIS batch, the width
            # the reference collate would have produced — GAE / whitening run over padded columns too, so the width is maths
            rtok = min(width + 1, response.shape[1])
            response = response[:, :rtok]
            old_logprobs, old_values, old_rewards = old_logprobs[:, :width], old_values[:, :width], old_rewards[:, :width]
            if trunk_in is not None:
                trunk_in = trunk_in[:, : query.shape[1] + rtok]
        response_length = old_rewards.shape[1]
        advantages, returns = self.config.method.get_advantages_and_returns(old_values, old_rewards, response_length,
                                                                            width_tensor=width_tensor)

        if self.config.model.model_arch_type == "seq2seq":
            attention_mask = query.ne(pad).long()
            dec_mask = response.ne(pad).long()
            dec_mask[:, 0] = 1
            out = self.model(input_ids=query, attention_mask=attention_mask, decoder_input_ids=response,
                             decoder_attention_mask=dec_mask, return_dict=True)
            logprobs = logprobs_of_labels(out.logits[:, :-1, :], response[:, 1:])
            start, end = 0, response_length
            logprobs, values_pred = logprobs[:, start:end], out.value[:, start:end]
            mask = response.ne(pad).long()[:, start + 1:end + 1]
        else:
            tokens = torch.cat((query, response), dim=1)
            attention_mask = tokens.not_equal(pad).long()
            position_ids = (attention_mask.cumsum(-1) - 1).clamp_min(0)
            start = query.shape[1] - 1
            end = start + response_length
            labels = tokens[:, 1:]
            trunk = trunk_in
            if hasattr(self.model, "score") and self.model.can_share_trunk():
                # heads only on the response rows; trunk activation reused when the store carries it
                inp, am, pos = tokens[:, :-1], attention_mask[:, :-1], position_ids[:, :-1]
                if trunk is not None:
                    trunk = trunk[:, : inp.shape[1]]
                logprobs, values_pred, _, _ = self.model.score(inp, am, pos, labels, trunk_hidden=trunk, with_ref=False,
                                                               rows=(start, end))
            else:
                out = self.model(tokens, attention_mask, return_dict=True, position_ids=position_ids)
                logprobs = logprobs_of_labels(out.logits[:, :-1, :], labels)[:, start:end]
                values_pred = out.value[:, :-1][:, start:end]
            mask = attention_mask[:, start + 1:end + 1]

        return self.config.method.loss(logprobs=logprobs.float(), values=values_pred.float(), old_logprobs=old_logprobs,
                                       old_values=old_values, advantages=advantages, returns=returns, mask=mask.float(),
                                       width_tensor=width_tensor)

    # ---- bookkeeping ------------------------------------------------------------------------------------------------------
    def setup_rollout_logging(self, config):
        assert os.path.isdir(config.train.rollout_logging_dir)
        self.run_id = f"run-{uuid.uuid4()}"
        self.rollout_logging_dir = os.path.join(config.train.rollout_logging_dir, self.run_id)
        os.mkdir(self.rollout_logging_dir)
        with open(os.path.join(self.rollout_logging_dir, "config.json"), "w") as fh:
            fh.write(json.dumps(config.to_dict(), indent=2, default=str))

    def post_epoch_callback(self):
        """Export (optional) and drop the old rollouts, then collect fresh ones."""
        if self.log_rollouts:
            self.store.export_history(location=self.rollout_logging_dir)
        self.store.clear_history()
        self.make_experience(self.config.method.num_rollouts, self.iter_count)

    def post_backward_callback(self):
        self.kl_ctl.update(self.mean_kl, n_steps=self.config.train.batch_size)

    def _graphs_enabled(self) -> bool:
        """Whole-step CUDA graphs: CUDA, fused optimizer, one micro-batch per step, decoder-only hydra with shared trunk."""
        rt = self.runtime
        if rt.tp_size > 1 or rt.pp_size > 1:
            # the fused tensor-parallel kernels hand host-side epochs to their flag waits and fork copy streams per call: a replayed
            # graph would wait for flags that never come (2-GPU NeMoPPOTrainer deadlocked this way); model-parallel steps run eagerly
            return False
        return bool(self.runtime.cuda and self.config.train.parallel.cuda_graphs and self.num_mb == 1
                    and getattr(self.opt, "graph_capturable", False) and self.config.model.model_arch_type != "seq2seq"
                    and hasattr(self.model, "can_share_trunk") and self.model.can_share_trunk()
                    and getattr(self, "zero3", None) is None  # gathered parameters move between pooled buffers
                    and not self.config.train.trainer_kwargs.get("no_train_graph", False)
                    and os.environ.get("TRLX_B200_TRAIN_GRAPH", "1") == "1")

    def create_train_dataloader(self):
        return self.store.create_loader(self.config.train.batch_size, shuffle=True, static_shapes=self._graphs_enabled())

    def train_step(self, minibatch):
        """One optimizer step; replayed from a captured CUDA graph when the shapes allow it."""
        batch = minibatch[0] if len(minibatch) == 1 else None
        if batch is None or not self._graphs_enabled() or not hasattr(batch, "width") or not batch.query_tensors.is_cuda:
            return super().train_step(minibatch)
        from trlx_b200.trainer.accelerate_base_trainer import EventTimer
        from trlx_b200.trainer.graphed_step import GraphedPPOStep

        key = GraphedPPOStep.shape_key(batch)
        step = self._graphed_steps.get(key)
        if step is None:
            if len(self._graphed_steps) >= 8:  # shapes keep changing → not worth capturing
                return super().train_step(minibatch)
            step = self._graphed_steps[key] = GraphedPPOStep(self, batch)
        self.mb_count += 1
        timer = EventTimer()
        stats = step.run(batch, batch.width)
        stats["time/step"] = timer.stop()
        stats["time/forward"] = 0.0  # not separable inside one graph; see time/step
        stats["time/backward"] = 0.0
        self.scheduler.step()
        self.iter_count += 1
        self._after_weights_changed()
        return stats

    def _after_weights_changed(self):
        super()._after_weights_changed()
        if getattr(self, "_engine", None) is not None:
            self._engine.mark_dirty()  # its γ-folded weight copies are rebuilt lazily before the next rollout

    def prepare_learning(self):
        self.eval_dataloader = self.eval_pipeline.create_loader(self.config.method.chunk_size)
        self.make_experience(self.config.method.num_rollouts)
        self.train_dataloader = self.create_train_dataloader()
        self.n_inner_epochs = self.config.method.ppo_epochs
        self.total_steps = self.config.train.epochs * self.n_inner_epochs * len(self.train_dataloader)
        self.total_steps = min(self.total_steps, self.config.train.total_steps)

    def add_prompt_pipeline(self, pipeline: PromptPipeline):
        """Prompt source for ``make_experience`` (sharded across data-parallel ranks)."""
        sampler = None
        if self.runtime.distributed:
            from torch.utils.data.distributed import DistributedSampler

            sampler = DistributedSampler(pipeline, num_replicas=self.runtime.dp_size, rank=self.runtime.dp_rank,
                                         shuffle=True, seed=self.config.train.seed)
        loader = pipeline.create_loader(self.config.method.chunk_size, shuffle=True, sampler=sampler)
        self.prompt_iterator = infinite_dataloader(loader, sampler)

    def _extra_state(self):
        return {"kl_ctl": self.kl_ctl.value, "running_moments": self.running_moments.state_dict(),
                "ref_mean": self.ref_mean, "ref_std": self.ref_std, "mean_kl": self.mean_kl}

    def _load_extra_state(self, st):
        if "kl_ctl" in st:
            self.kl_ctl.value = st["kl_ctl"]
            self.running_moments.load_state_dict(st["running_moments"])
            self.ref_mean, self.ref_std, self.mean_kl = st["ref_mean"], st["ref_std"], st["mean_kl"]

    # ---- rollouts ---------------------------------------------------------------------------------------------------------
    def _get_engine(self):
        """The CUDA rollout engine, when the model / sampling options allow it."""
        if self._engine is not None or self._engine_failed:
            return self._engine
        try:
            from trlx_b200.engine.rollout import RolloutEngine

            gen = self.generate_experience_kwargs or self.generate_kwargs
            why = RolloutEngine.why_not(self.model, gen, self.config, self.stop_sequences)
            if getattr(self, "zero3", None) is not None:
                why = "ZeRO-3 partitions the parameters (the engine needs resident, address-stable weights)"
            if why is not None:
                logger.warning(f"rollouts use the PyTorch sampler + scoring pass, not the CUDA rollout engine: {why}")
            if why is None:
                gen_engine = dict(gen, _rollout_dtype=self.config.train.parallel.rollout_dtype)
                self._engine = RolloutEngine(self.model, self.tokenizer.pad_token_id, self.tokenizer.eos_token_id, gen_engine,
                                             cache_trunk=self.cache_trunk, seed=self.config.train.seed + self.runtime.rank)
            else:
                self._engine_failed = True
        except ops.ExtensionMissing:
            raise
        except ImportError as err:
            logger.warning(f"rollout engine unavailable: {err}")
            self._engine_failed = True
        return self._engine

    def _score_with_reward_fn(self, str_samples, str_prompts, str_outputs, metadata, device) -> torch.Tensor:
        """``reward_fn`` → ``[B, S]`` tensor (S = 1 for scalar rewards), ``-inf`` padded for dense rewards."""
        rewards = self.reward_fn(samples=str_samples, prompts=str_prompts, outputs=str_outputs, tokenizer=self.tokenizer,
                                 **metadata)
        rows = [torch.as_tensor(r, dtype=torch.float32).reshape(-1) for r in rewards]
        if not rows:
            return torch.zeros(0, 1, device=device)
        width = max(len(r) for r in rows)
        out = torch.full((len(rows), width), float("-inf"))
        for i, r in enumerate(rows):
            out[i, : len(r)] = r
        return out.to(device)

    def _collect_scores(self, prompt_tensors, samples, metadata, device, stats) -> torch.Tensor:
        t0 = time()
        if not self.rank0_reward or not self.runtime.distributed:
            str_samples, str_prompts, str_outputs = self.decode(prompt_tensors, samples, append_eos_token=True)
            scores = self._score_with_reward_fn(str_samples, str_prompts, str_outputs, metadata, device)
            stats["time/rollout_score"] = time() - t0
            return scores
        # reference-compatible funnel: gather → rank 0 scores everything → scatter
        rt = self.runtime
        pad = self.tokenizer.pad_token_id
        p, s = rt.pad_across_processes([prompt_tensors, samples], dim=1, pad_index=pad, pad_first=False)
        sizes = torch.full((len(p),), prompt_tensors.shape[1], device=device)
        gp, gs, gsz = rt.gather(p), rt.gather(s), rt.gather(sizes)
        gmeta = gather_dict(metadata)
        if rt.is_main_process:
            a, b, c = self.decode(gp, gs, gsz, append_eos_token=True)
            all_scores = self._score_with_reward_fn(a, b, c, gmeta, device)
            payload = [all_scores.cpu()]
        else:
            payload = [None]
        torch.distributed.broadcast_object_list(payload, src=0)
        n = len(prompt_tensors)
        stats["time/rollout_score"] = time() - t0
        return payload[0][rt.rank * n:(rt.rank + 1) * n].to(device)

    def _rollout_torch(self, batch, device, samples=None):
        """PyTorch path: sample (unless ``samples`` were already generated by the engine), re-tokenise the decoded / trimmed
        outputs as the reference does, then ONE shared-trunk scoring pass."""
        pad, eos = self.tokenizer.pad_token_id, self.tokenizer.eos_token_id
        if samples is None:
            samples = self.generate(batch["input_ids"], batch["attention_mask"])
        prompt_tensors = batch["input_ids"].to(device)
        seq2seq = self.config.model.model_arch_type == "seq2seq"
        str_samples, str_prompts, str_outputs = self.decode(prompt_tensors, samples, append_eos_token=True)
        outputs = self.tokenizer(str_outputs).input_ids
        if seq2seq:
            outputs = [[pad] + list(o) for o in outputs]
        sample_outputs = pad_rows([torch.tensor(o, dtype=torch.long) for o in outputs], pad, "right", min_len=1).to(device)
        with torch.no_grad():
            if seq2seq:
                attention_mask = batch["attention_mask"].to(device)
                dec_mask = sample_outputs.not_equal(pad).long()
                dec_mask[:, 0] = 1
                out = self.model(input_ids=prompt_tensors, attention_mask=attention_mask, decoder_input_ids=sample_outputs,
                                 decoder_attention_mask=dec_mask, return_dict=True)
                values = out.value
                if getattr(self.model, "frozen_head", None) is not None or self.model.peft_type:
                    ref_logits = self.model.forward_hydra(input_ids=prompt_tensors, attention_mask=attention_mask,
                                                          decoder_input_ids=sample_outputs,
                                                          decoder_attention_mask=dec_mask, return_dict=True).logits
                else:
                    ref_logits = self.ref_model(input_ids=prompt_tensors, attention_mask=attention_mask,
                                                decoder_input_ids=sample_outputs, decoder_attention_mask=dec_mask,
                                                return_dict=True).logits
                logprobs = logprobs_of_labels(out.logits[:, :-1, :], sample_outputs[:, 1:])
                ref_logprobs = logprobs_of_labels(ref_logits[:, :-1, :], sample_outputs[:, 1:])
                mask = sample_outputs.not_equal(pad).long()
                start, trunk = 0, None
            else:
                all_tokens = torch.cat((prompt_tensors, sample_outputs), dim=1)
                mask = all_tokens.not_equal(pad).long()
                position_ids = (mask.cumsum(-1) - 1).clamp_min(0)
                labels = torch.cat([all_tokens[:, 1:], all_tokens.new_full((len(all_tokens), 1), -1)], 1)
                trunk = None
                if hasattr(self.model, "score") and (self.model.can_share_trunk() or self.model.peft_type):
                    logprobs, values, ref_logprobs, trunk = self.model.score(all_tokens, mask, position_ids, labels)
                    logprobs, ref_logprobs = logprobs[:, :-1], ref_logprobs[:, :-1]
                else:
                    logits, *_, values = self.model(all_tokens, attention_mask=mask, position_ids=position_ids)
                    ref_logits = self.ref_model(all_tokens, attention_mask=mask, position_ids=position_ids,
                                                return_dict=True).logits
                    logprobs = logprobs_of_labels(logits[:, :-1, :], all_tokens[:, 1:])
                    ref_logprobs = logprobs_of_labels(ref_logits[:, :-1, :], all_tokens[:, 1:])
                start = prompt_tensors.shape[1] - 1
        return dict(samples=samples, prompt_tensors=prompt_tensors, sample_outputs=sample_outputs, logprobs=logprobs.float(),
                    ref_logprobs=ref_logprobs.float(), values=values.float()[:, :-1], mask=mask, start=start,
                    trunk=trunk if self.cache_trunk else None,
                    strings=(str_samples, str_prompts, str_outputs))

    def make_experience(self, num_rollouts: int = 1024, iter_count: int = 0):  # noqa: C901
        """Collect ``num_rollouts`` rollouts per rank into the store."""
        logger.info("Collecting rollouts")
        rt = self.runtime
        device = rt.device
        method = self.config.method
        pad = self.tokenizer.pad_token_id
        tbar = logging.tqdm(total=num_rollouts, disable=not rt.is_main_process, desc=f"[rollout 0 / {num_rollouts}]",
                            position=logging.get_verbosity() >= logging.WARNING, leave=logging.get_verbosity() < logging.WARNING)
        clock = Clock()
        collected = 0
        accumulated: List[Dict[str, Any]] = []
        engine = self._get_engine() if rt.cuda else None

        while collected < num_rollouts:
            stats: Dict[str, Any] = {}
            batch = next(self.prompt_iterator)
            self._last_prompt_width = int(batch["input_ids"].shape[1])
            metadata = {k: v for k, v in batch.items() if k not in ("input_ids", "attention_mask")}
            t_gen = time()
            if engine is not None:
                ids, am = batch["input_ids"], batch["attention_mask"]
                width = _bucket_prompt_width(ids.shape[1], self.config.train.trainer_kwargs.get("prompt_bucket"))
                if rt.distributed and rt.dp_size > 1:
                    # every data-parallel rank uses the same padded prompt width, so their rollout blocks — and therefore
                    # the CUDA-graph shape keys of the training step, whose capture runs cross-GPU barriers — line up
                    wt = torch.tensor([width], device=device, dtype=torch.int64)
                    rt.all_reduce(wt, "max", group=rt.dp_group)
                    width = int(wt.item())
                if width != ids.shape[1] and self.tokenizer.padding_side == "left":
                    # bucket the (left-padded) prompt width so decode / train-step CUDA graphs see few distinct shapes
                    ids = F.pad(ids, (width - ids.shape[1], 0), value=pad)
                    am = F.pad(am, (width - am.shape[1], 0), value=0)
                with rt.nvtx("rollout/engine"):
                    ro = engine.rollout(ids, am)
                if self.stop_sequences:
                    # generation ran on the engine; trimming at a stop sequence changes the text, so re-tokenise and re-score
                    # exactly like the reference (``accelerate_ppo_trainer.py:304-345``) instead of reusing decode-time scores
                    with rt.nvtx("rollout/rescore"):
                        ro = self._rollout_torch(dict(batch, input_ids=ids, attention_mask=am), device, samples=ro["samples"])
            else:
                with rt.nvtx("rollout/torch"):
                    ro = self._rollout_torch(batch, device)
            stats["time/rollout_generate"] = time() - t_gen

            prompt_tensors, sample_outputs = ro["prompt_tensors"], ro["sample_outputs"]
            if "strings" in ro and not (self.rank0_reward and rt.distributed):
                t0 = time()
                with rt.nvtx("rollout/reward_fn"):
                    scores = self._score_with_reward_fn(*ro["strings"], metadata, device)
                stats["time/rollout_score"] = time() - t0
            elif "samples_host" in ro and not (self.rank0_reward and rt.distributed):
                # the engine already sent the tokens to pinned host memory, ahead of its scoring kernels: wait for that copy
                # only, then detokenise + score on the CPU while the GPU finishes the reference / value passes
                host_tokens, host_ready = ro["samples_host"]
                host_ready.synchronize()
                with rt.nvtx("rollout/reward_fn"):
                    scores = self._collect_scores(host_tokens[:, :prompt_tensors.shape[1]], host_tokens, metadata, device, stats)
            else:
                with rt.nvtx("rollout/reward_fn"):
                    scores = self._collect_scores(prompt_tensors, ro["samples"], metadata, device, stats)
            scores_mask = scores != float("-inf")
            scores = torch.where(scores_mask, scores, torch.zeros_like(scores))

            if method.cliprange_reward:
                scores = torch.clip(scores, -method.cliprange_reward, method.cliprange_reward)
            summed = (scores * scores_mask).sum(dim=1)
            if self.ref_mean is None:
                self.ref_mean, self.ref_std = float(summed.mean()), float(summed.std()) if len(summed) > 1 else 1.0
            batch_mean, batch_std = self.running_moments.update(summed)
            stats["rollout_scores/mean"] = float(batch_mean)
            stats["rollout_scores/std"] = float(batch_std)
            stats["rollout_scores/running_mean"] = float(self.running_moments.mean)
            stats["rollout_scores/running_std"] = float(self.running_moments.std)
            if method.scale_reward == "running":
                scores = scores / self.running_moments.std
            elif method.scale_reward == "ref":
                scores = scores / self.ref_std

            # ---- the reference's per-sample slicing (``:455-504``): KL-penalty rewards, score placement, k3 statistics
            logprobs, ref_logprobs, values, mask, start = ro["logprobs"], ro["ref_logprobs"], ro["values"], ro["mask"], ro["start"]
            R = logprobs.shape[1] - start
            fused = (rt.cuda and scores.shape[1] == 1 and ops.enabled_for(logprobs) and hasattr(ops.C, "rollout_rewards")
                     and logprobs.dtype == torch.float32)
            if fused:  # one launch (csrc/rl_ops.cu: rollout_rewards_kernel)
                rewards, lp_s, v_s, slice_len, kl_sum = ops.C.rollout_rewards(
                    logprobs.contiguous(), ref_logprobs.float().contiguous(), values.float().contiguous(), mask.contiguous(),
                    scores[:, 0].float().contiguous(), start, float(self.kl_ctl.value))
                slice_len = slice_len.long()
                mean_kl = (kl_sum[0] / logprobs.shape[0]).float()
                mean_kl_per_token = (kl_sum[0] / logprobs.numel()).float()
            else:
                log_ratio = (logprobs - ref_logprobs) * mask[:, :-1]
                kl = log_ratio.exp() - 1 - log_ratio
                mean_kl_per_token = kl.mean()
                mean_kl = kl.sum(1).mean()
                slice_len = (mask[:, start:].sum(1) + 1).clamp(max=R)
                cols = torch.arange(R, device=device).unsqueeze(0)
                valid = cols < slice_len.unsqueeze(1)
                lp_s = torch.where(valid, logprobs[:, start:], torch.zeros_like(logprobs[:, start:]))
                v_s = torch.where(valid, values[:, start:], torch.zeros_like(values[:, start:]))
                rewards = torch.where(valid, -self.kl_ctl.value * log_ratio[:, start:], torch.zeros_like(lp_s))
                if scores.shape[1] == 1:
                    rewards.scatter_add_(1, (slice_len - 1).clamp_min(0).unsqueeze(1), scores[:, :1].to(rewards.dtype))
                else:
                    k = min(scores.shape[1], R)
                    dense = torch.zeros_like(rewards)
                    dense[:, :k] = (scores * scores_mask)[:, :k]
                    rewards = rewards + torch.where(valid, dense, torch.zeros_like(dense))

            q_lens = prompt_tensors.ne(pad).sum(1)
            block = RolloutBlock(queries=prompt_tensors, responses=sample_outputs, logprobs=lp_s, values=v_s, rewards=rewards,
                                 query_lens=q_lens, response_lens=slice_len,
                                 host_response_lens=slice_len.tolist(), host_query_lens=q_lens.tolist(),
                                 trunk_hidden=ro.get("trunk"))
            self.store.push_block(block)
            collected += len(prompt_tensors)

            if rt.distributed:
                rt.all_reduce(mean_kl, "mean")
            stats["time/rollout_time"] = clock.tick()
            stats["policy/sqrt_kl"] = torch.sqrt(mean_kl.clamp_min(0)).item()
            stats["policy/kl_per_token"] = torch.sqrt(mean_kl_per_token.clamp_min(0)).item()
            accumulated.append(stats)
            tbar.set_description(f"[rollout {min(collected, num_rollouts)} / {num_rollouts}]")
            tbar.update(min(len(prompt_tensors), num_rollouts))
        tbar.close()

        stats = {k: sum(xs[k] for xs in accumulated) / len(accumulated) for k in accumulated[-1]}
        stats["kl_ctl_value"] = self.kl_ctl.value
        self.mean_kl = stats["policy/sqrt_kl"] ** 2
        rt.log(stats, step=iter_count)

    def save_pretrained(self, directory: Optional[str] = None, **kwargs):
        """Export the wrapped model (``base_model.*``, ``v_head.*``, ``frozen_head.*`` keys) + tokenizer."""
        super().save_pretrained(directory, **kwargs)
