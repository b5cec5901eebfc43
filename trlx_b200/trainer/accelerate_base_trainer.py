"""Shared training loop for PPO / ILQL / SFT / RFT.

Public surface follows ``trlx/trainer/accelerate_base_trainer.py`` (class name kept as ``AccelerateRLTrainer`` so
configs naming ``Accelerate*Trainer`` keep working): ctor bookkeeping ``:46-146``, ``setup_model/optimizer/scheduler``
``:148-201``, ``decode`` ``:203-254``, ``generate``/``generate_eval`` ``:256-282``, ``save_pretrained`` ``:284-307``,
``save``/``load`` ``:309-333``, ``evaluate`` ``:339-500``, ``_accumulate`` ``:502-516``, ``learn`` ``:518-652``.

What is different underneath: there is no Accelerate/DeepSpeed.  :class:`trlx_b200.parallel.runtime.Runtime` owns the
process groups and trackers; the optimizer is the flat, partitioned :class:`~trlx_b200.parallel.optim.FusedAdamW`
whose ``step()`` *is* the gradient synchronisation (fused reduce-scatter + update + all-gather), so gradient
accumulation needs no ``no_sync`` dance; statistics stay on the device and are fetched with ONE transfer per
optimizer step; phase timings are device-timed.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
from abc import abstractmethod
from contextlib import contextmanager
from time import time
from typing import Any, Dict, List, Optional, Tuple

import torch

from trlx_b200.data.configs import TRLConfig
from trlx_b200.parallel.runtime import Runtime
from trlx_b200.pipeline import MiniBatchIterator
from trlx_b200.trainer import BaseRLTrainer, register_trainer
from trlx_b200.utils import (filter_non_scalars, get_distributed_config, get_git_tag, get_optimizer_class,
                             get_scheduler_class, logging, significant, resolve_output_dir)
from trlx_b200.utils.modeling import flatten_dict, freeze_bottom_causal_layers, freeze_bottom_seq2seq_layers
from trlx_b200.utils.tokenizer import load_tokenizer

logger = logging.get_logger(__name__)


class EventTimer:
    """Device-timed interval whose value is read lazily (after the step's single synchronising transfer)."""

    def __init__(self):
        self.start, self.end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.start.record()

    def stop(self):
        self.end.record()
        return self

    def __float__(self):
        self.end.synchronize()
        return self.start.elapsed_time(self.end) / 1e3


class PendingStats:
    """Statistics of an optimizer step on their way to the host.  Construction enqueues ONE device→host transfer of all 0-dim
    device tensors (packed, into pinned memory, followed by an event); :meth:`get` waits for that event only.  The training loop
    reads the statistics of step ``i`` after it has launched step ``i + 1``, so the GPU never idles on the host's bookkeeping
    (the values are copied in stream order, i.e. before a replayed CUDA graph overwrites its static output buffers)."""

    def __init__(self, stats: Dict[str, Any]):
        self.stats = dict(stats)
        self.keys = [k for k, v in stats.items() if isinstance(v, torch.Tensor) and v.numel() == 1]
        self.host, self.event, self.vals = None, None, None
        if self.keys:
            dev = [stats[k].detach().float().reshape(()) for k in self.keys]
            if dev[0].is_cuda:
                packed = torch.stack(dev)
                self.host = torch.empty(packed.shape, dtype=packed.dtype, pin_memory=True)
                self.host.copy_(packed, non_blocking=True)
                self.event = torch.cuda.Event()
                self.event.record()
            else:
                self.vals = [float(v) for v in dev]

    def get(self) -> Dict[str, Any]:
        if self.event is not None:
            self.event.synchronize()
            self.vals, self.event = self.host.tolist(), None
        out = self.stats
        if self.keys:
            out.update(zip(self.keys, self.vals))
        for k, v in list(out.items()):
            if isinstance(v, EventTimer):
                out[k] = float(v)
        return out


def _materialise(stats: Dict[str, Any]) -> Dict[str, Any]:
    """Turn 0-dim device tensors into python floats with a single device→host transfer."""
    return PendingStats(stats).get()


def _parse_max_time(value) -> Optional[float]:
    """Seconds from a number or a ``"DD:HH:MM:SS"`` string (``None`` → no limit)."""
    if value is None or value == "":
        return None
    if isinstance(value, (int, float)):
        return float(value)
    parts = [int(p) for p in str(value).split(":")]
    while len(parts) < 4:
        parts.insert(0, 0)
    d, h, m, sec = parts[-4:]
    return float(((d * 24 + h) * 60 + m) * 60 + sec)


@register_trainer
class AccelerateRLTrainer(BaseRLTrainer):
    """Abstract trainer on the B200 runtime."""

    #: ``train.trainer_kwargs`` keys that configure this framework's trainers (read from the config where they are used);
    #: everything else in ``trainer_kwargs`` is a constructor argument, exactly as in the reference (``trlx/trlx.py:92-98``)
    FRAMEWORK_KWARGS = ("prompt_bucket", "rank0_reward", "cache_trunk", "zero_stage", "max_time", "megatron_cfg",
                        "pretrained_model", "offload_reference", "no_train_graph", "max_nonfinite_steps", "lag_stats")

    def __init__(self, config: TRLConfig, **kwargs):
        for key in self.FRAMEWORK_KWARGS:
            kwargs.pop(key, None)
        super().__init__(config, **kwargs)
        self.max_length = config.train.seq_length
        if config.train.minibatch_size:
            assert config.train.batch_size % config.train.minibatch_size == 0, "Minibatch size must divide batch size"
            self.mb_size = config.train.minibatch_size
        else:
            self.mb_size = config.train.batch_size
        self.num_mb = config.train.batch_size // self.mb_size
        self.mb_count = 0
        self.gradient_accumulation_steps = 1

        # relative output locations are resolved once (TRLX_B200_OUT / never into the source checkout)
        config.train.checkpoint_dir = resolve_output_dir(config.train.checkpoint_dir)
        config.train.logging_dir = resolve_output_dir(config.train.logging_dir or "logs")
        if getattr(config.train, "rollout_logging_dir", None):
            config.train.rollout_logging_dir = resolve_output_dir(config.train.rollout_logging_dir)
        self.runtime = Runtime(config.train.parallel)
        self.accelerator = self.runtime  # attribute name kept for user code written against the reference
        self.runtime.barrier()

        self.tokenizer = load_tokenizer(config.tokenizer.tokenizer_path, **config.tokenizer.tokenizer_extra_configs)
        self.tokenizer.padding_side = config.tokenizer.padding_side
        self.tokenizer.truncation_side = config.tokenizer.truncation_side
        self.tokenizer.sep_token = "<sep>"
        if self.tokenizer.pad_token is None:
            self.tokenizer.pad_token = "<|padding|>"

        self._fit_random_init_vocab()
        self.model = self.setup_model()
        self.opt = self.setup_optimizer()
        self.scheduler = self.setup_scheduler()

        script_name = os.path.basename(sys.argv[0]).rsplit(".", 1)[0]
        mp = config.model.model_path
        model_name = mp.split("/")[-1] if isinstance(mp, str) else (mp.get("model_type", "model") if isinstance(mp, dict)
                                                                     else type(mp).__name__)
        num_gpus = "1gpu" if self.runtime.num_processes == 1 else f"{self.runtime.num_processes}gpus"
        run_name = config.train.run_name or "/".join([script_name, model_name, num_gpus]) + f":{get_git_tag()[0]}"
        self.run_name = run_name
        if self.runtime.is_main_process:
            cfg_dict = self.config.to_dict()
            cfg_dict["distributed"] = get_distributed_config(self.runtime)
            if cfg_dict["model"].get("peft_config") is not None:
                pc = cfg_dict["model"]["peft_config"]
                cfg_dict["model"]["peft_config"] = pc if isinstance(pc, dict) else str(pc)
            if not isinstance(cfg_dict["model"]["model_path"], str):
                cfg_dict["model"]["model_path"] = str(cfg_dict["model"]["model_path"])
            init_kwargs = {}
            if config.train.tracker == "wandb":
                init_kwargs = dict(entity=config.train.entity_name, group=config.train.group_name,
                                   tags=(config.train.tags or []) + ["/".join(get_git_tag())],
                                   mode="disabled" if os.environ.get("debug", False) else "online")
            self.runtime.init_tracker(config.train.tracker, config.train.project_name, flatten_dict(cfg_dict), run_name,
                                      logging_dir=config.train.logging_dir, **init_kwargs)

        self.nth_evaluation = 0
        self.iter_count = 0
        self.generate_sweep_kwarg = None
        self.generate_kwargs = dict(getattr(self, "generate_kwargs", None) or config.method.gen_kwargs)
        for k, v in config.method.gen_kwargs.items():
            if isinstance(v, list):
                if self.generate_sweep_kwarg is not None:
                    logger.info(f"Only a single sweep is allowed, {k} is going to be set to {v[0]}")
                    self.generate_kwargs[k] = v[0]
                else:
                    self.generate_sweep_kwarg = (k, v)

    # ---- setup --------------------------------------------------------------------------------------------------------
    def _fit_random_init_vocab(self) -> None:
        """A model described by a *config mapping* is initialised from scratch, so nothing ties its vocabulary size to a
        checkpoint: grow it to the tokenizer's when the tokenizer has more ids (offline runs pair synthetic tokenizers with
        small architecture presets; an id past the embedding table would otherwise be an index error deep inside the model)."""
        path = self.config.model.model_path
        if not isinstance(path, dict) or "vocab_size" not in path:
            return
        try:
            n_tok = len(self.tokenizer)
        except TypeError:
            n_tok = int(getattr(self.tokenizer, "vocab_size", 0) or 0)
        if n_tok > int(path["vocab_size"]):
            logger.warning(f"random-init model: vocab_size {path['vocab_size']} -> {n_tok} to cover the tokenizer")
            self.config.model.model_path = dict(path, vocab_size=n_tok)

    def setup_model(self):
        logger.info(f"Initializing model: {self.config.model.model_path}")
        model = self.get_arch(self.config)
        if self.config.model.peft_config is None:
            if self.config.model.model_arch_type == "seq2seq":
                freeze_bottom_seq2seq_layers(model.base_model, self.config.model.num_layers_unfrozen)
            else:
                freeze_bottom_causal_layers(model.base_model, self.config.model.num_layers_unfrozen)
        elif self.config.model.num_layers_unfrozen >= 0:
            logger.warning("The argument num_layers_unfrozen is ignored when using peft, to prevent unexpected behaviour."
                           "For Lora, use the `LoraConfig` argument `modules_to_save` instead.")
        model = model.to(self.runtime.device)
        if self.runtime.cuda and self.runtime.dtype != torch.float32:
            model = model.to(self.runtime.dtype)
        self._sync_initial_weights(model)
        if getattr(self.config.train.parallel, "activation_checkpointing", False):
            for m in model.modules():
                if hasattr(m, "gradient_checkpointing_enable") and m is not model:
                    m.gradient_checkpointing_enable()
        model.eval()
        return model

    def _sync_initial_weights(self, model):
        """Every rank starts from rank 0's weights (randomly initialised heads / from-config models are seeded per data-
        parallel rank).  The reference gets this from DDP's constructor broadcast inside ``accelerator.prepare``
        (``accelerate_base_trainer.py:71``); here it also guarantees identical full weights before TP/PP sharding."""
        if not self.runtime.distributed:
            return
        import torch.distributed as dist

        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                if t.numel():
                    dist.broadcast(t.data, src=0)

    def setup_optimizer(self):
        optimizer_class = get_optimizer_class(self.config.optimizer.name)
        kwargs = dict(self.config.optimizer.kwargs)
        from trlx_b200.parallel.optim import FusedAdamW

        if issubclass(optimizer_class, FusedAdamW):
            kwargs.setdefault("grad_clip", self.config.train.parallel.grad_clip)
            kwargs.setdefault("process_group", self.runtime.dp_group)
            kwargs.setdefault("zero_stage", self.config.train.parallel.zero_stage)
            kwargs.setdefault("bucket_mb", self.config.train.parallel.bucket_mb)
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.zero3 = None
        rt = self.runtime
        if (int(self.config.train.parallel.zero_stage) >= 3 and rt.tp_size == 1 and rt.pp_size == 1
                and (rt.dp_size > 1 or os.environ.get("TRLX_B200_ZERO3_FORCE") == "1")):
            # ZeRO-3: parameters, gradients and optimizer state partitioned over the data-parallel ranks (parallel/zero3.py)
            from trlx_b200.parallel.zero3 import Zero3ParamSharder, default_units

            self.zero3 = Zero3ParamSharder(self.model, default_units(self.model), rt.dp_group)
            params = self.zero3.shard_parameters()
            if issubclass(optimizer_class, FusedAdamW):
                kwargs["local_only"] = True
        opt = optimizer_class(self._optimizer_param_groups(params, optimizer_class), **kwargs)
        if hasattr(opt, "prepare"):
            opt.prepare()
        if issubclass(optimizer_class, FusedAdamW) and self.zero3 is None:
            from trlx_b200.ops.functional import mark_inplace_wgrad

            mark_inplace_wgrad(self.model)
        if "8bit" in optimizer_class.__name__:
            for module in self.model.modules():  # keep embedding state in 32 bits (reference: :183-191)
                if isinstance(module, torch.nn.Embedding):
                    module.weight._optim_32bit = True
        return opt

    def _optimizer_param_groups(self, params, optimizer_class):
        """Hook for trainers whose parameters do not all reduce over the same ranks (tensor-parallel trainers)."""
        return params

    def setup_scheduler(self):
        scheduler_class = get_scheduler_class(self.config.scheduler.name)
        return scheduler_class(self.opt, **self.config.scheduler.kwargs)

    # ---- text <-> tokens ------------------------------------------------------------------------------------------------
    def decode(self, prompts, samples, prompt_sizes=None, append_eos_token: bool = False) -> Tuple[List[str], List[str], List[str]]:
        """Token tensors → ``(samples, prompts, outputs)`` strings; trims at ``stop_sequences`` and restores a
        trailing EOS when the generation ended by itself (or was trimmed)."""
        if prompt_sizes is None:
            prompt_sizes = [prompts.shape[1]] * len(prompts)  # left-padded prompts
        if isinstance(prompts, torch.Tensor) and prompts.is_cuda:
            prompts = prompts.cpu()
        if isinstance(samples, torch.Tensor) and samples.is_cuda:
            samples = samples.cpu()
        if isinstance(prompt_sizes, torch.Tensor):
            prompt_sizes = prompt_sizes.tolist()
        seq2seq = self.config.model.model_arch_type == "seq2seq"
        tok = self.tokenizer
        str_samples, str_prompts, str_outputs = [], [], []
        # two batched detokeniser calls (the Rust side parallelises over rows) instead of 2 x B python-level decodes
        p_rows = prompts.tolist() if isinstance(prompts, torch.Tensor) else [list(map(int, p)) for p in prompts]
        s_rows = samples.tolist() if isinstance(samples, torch.Tensor) else [list(map(int, s)) for s in samples]
        sizes = [int(n) for n in prompt_sizes]
        dec_prompts = tok.batch_decode([row[:n] for row, n in zip(p_rows, sizes)], skip_special_tokens=True)
        dec_outputs = tok.batch_decode([row[(0 if seq2seq else n):] for row, n in zip(s_rows, sizes)], skip_special_tokens=True)
        for str_prompt, str_output, sample in zip(dec_prompts, dec_outputs, s_rows):
            trimmed = False
            for stop in self.stop_sequences or []:
                ix = str_output.find(stop)
                if ix >= 0:
                    str_output = str_output[:ix].rstrip()
                    trimmed = True
            last = int(sample[-1]) if len(sample) else None
            if append_eos_token and (trimmed or last == tok.eos_token_id or last == tok.pad_token_id):
                str_output += tok.eos_token
            str_prompts.append(str_prompt)
            str_outputs.append(str_output)
            str_samples.append(str_prompt + (tok.sep_token if seq2seq else "") + str_output)
        return str_samples, str_prompts, str_outputs

    def generate(self, input_ids, attention_mask=None, **kwargs):
        """Sample with the experience-generation kwargs (falls back to ``gen_kwargs``)."""
        base = getattr(self, "generate_experience_kwargs", None) or self.generate_kwargs
        return self._generate(input_ids, attention_mask, dict(base, **kwargs))

    def generate_eval(self, input_ids, attention_mask=None, **kwargs):
        """Sample with the evaluation ``gen_kwargs``."""
        return self._generate(input_ids, attention_mask, dict(self.generate_kwargs, **kwargs))

    def _generate(self, input_ids, attention_mask, kwargs):
        input_ids = input_ids.to(self.runtime.device)
        if attention_mask is not None:
            attention_mask = attention_mask.to(self.runtime.device)
        with torch.no_grad():
            return self.model.generate(input_ids=input_ids, attention_mask=attention_mask, **kwargs)

    # ---- persistence ------------------------------------------------------------------------------------------------------
    def save_pretrained(self, directory: Optional[str] = None, **kwargs):
        """HF-layout export (``config.json`` + weights + tokenizer) of the wrapped model."""
        if directory is None:
            directory = os.path.join(self.config.train.checkpoint_dir, "hf_model")
        directory = resolve_output_dir(directory)
        self.runtime.barrier()
        with self._full_params():
            if self.runtime.is_main_process:
                self.model.save_pretrained(directory, **kwargs)
        if self.runtime.is_main_process:
            try:
                self.tokenizer.save_pretrained(directory)
            except Exception as err:  # pragma: no cover
                logger.warning(f"could not save tokenizer: {err}")
        self.runtime.barrier()

    def _full_params(self, writeback: bool = False):
        """All parameters materialised on every rank (no-op unless ZeRO-3 partitions them); collective: every data-parallel
        rank must enter."""
        z = getattr(self, "zero3", None)
        return z.summon_full_params(writeback=writeback) if z is not None else contextlib.nullcontext()

    def _extra_state(self) -> Dict[str, Any]:
        return {}

    def _load_extra_state(self, state: Dict[str, Any]) -> None:
        pass

    def save(self, directory: Optional[str] = None, **kwargs):
        """Full training state: raw model tensors, optimizer shard of every rank, scheduler, RNG, counters."""
        directory = resolve_output_dir(directory or self.config.train.checkpoint_dir)
        os.makedirs(directory, exist_ok=True)
        rank = self.runtime.rank
        with self._full_params():
            if self.runtime.dp_rank == 0:  # one writer per model-parallel rank: its tensor / pipeline shard of the weights
                torch.save({k: v.detach().cpu() for k, v in self.model.raw_state_dict().items()},
                           os.path.join(directory, self._model_state_name()))
        if self.runtime.is_main_process:
            with open(os.path.join(directory, "state.json"), "w") as fh:
                json.dump({"iter_count": self.iter_count, "nth_evaluation": self.nth_evaluation,
                           "world_size": self.runtime.world_size}, fh)
        torch.save({"optimizer": self.opt.state_dict(), "scheduler": self.scheduler.state_dict(),
                    "rng": {"torch": torch.get_rng_state(),
                            "cuda": torch.cuda.get_rng_state() if self.runtime.cuda else None},
                    "extra": self._extra_state()},
                   os.path.join(directory, f"trainer_state_rank{rank}.pt"))
        if self.model.peft_type and self.runtime.is_main_process:
            self.model.save_pretrained(directory)
        self.runtime.barrier()

    def _model_state_name(self) -> str:
        """``model_state.pt`` for unsharded models; one file per (tensor, pipeline) rank otherwise."""
        rt = self.runtime
        if rt.tp_size == 1 and rt.pp_size == 1:
            return "model_state.pt"
        return f"model_state_mp_{rt.tp_rank:02d}" + (f"_{rt.pp_rank:03d}" if rt.pp_size > 1 else "") + ".pt"

    def load(self, directory: Optional[str] = None, **kwargs):
        """Restore what :meth:`save` wrote (or, for a plain ``hf_model`` export, just the weights)."""
        directory = resolve_output_dir(directory or self.config.train.checkpoint_dir, for_read=True)
        path = os.path.join(directory, self._model_state_name())
        if os.path.exists(path):
            sd = torch.load(path, map_location="cpu", weights_only=True)
            with self._full_params(writeback=True), torch.no_grad():
                own = self.model.raw_state_dict()
                for k, v in sd.items():
                    if k in own:
                        if own[k].shape != v.shape:
                            raise ValueError(f"checkpoint tensor {k} has shape {tuple(v.shape)}, the model expects "
                                             f"{tuple(own[k].shape)}: was it written with a different parallel layout?")
                        own[k].copy_(v.to(own[k].dtype))
        st_path = os.path.join(directory, f"trainer_state_rank{self.runtime.rank}.pt")
        if hasattr(self.opt, "resync_master"):
            # the weights were just written behind the optimizer's back: its fp32 master copy must follow (restored
            # below when the checkpoint carries optimizer state, which then overwrites it with the exact fp32 values)
            self.opt.resync_master()
        if os.path.exists(st_path):
            st = torch.load(st_path, map_location="cpu", weights_only=False)
            self.opt.load_state_dict(st["optimizer"])
            self.scheduler.load_state_dict(st["scheduler"])
            torch.set_rng_state(st["rng"]["torch"])
            if self.runtime.cuda and st["rng"]["cuda"] is not None:
                torch.cuda.set_rng_state(st["rng"]["cuda"])
            self._load_extra_state(st.get("extra", {}))
        js = os.path.join(directory, "state.json")
        if os.path.exists(js):
            with open(js) as fh:
                state = json.load(fh)
            self.iter_count = state.get("iter_count", 0)
            self.nth_evaluation = state.get("nth_evaluation", 0)
        self._after_weights_changed()
        self.runtime.barrier()

    def release_device_state(self):
        """Drop captured CUDA graphs / engine state (they hold NVLink symmetric-memory references; destroying them while the
        process group is being torn down at interpreter exit has deadlocked multi-rank runs)."""
        if getattr(self, "_graphed_steps", None):
            self._graphed_steps.clear()
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng._state = None
        if self.runtime.cuda:
            torch.cuda.synchronize()

    def _pre_optimizer_step(self):
        """Hook between the last backward and ``opt.step()`` (tensor-parallel gradient fix-ups live here)."""

    def _after_weights_changed(self):
        """Hook: engines holding derived weight copies refresh here."""

    # ---- evaluation -------------------------------------------------------------------------------------------------------
    def add_eval_pipeline(self, eval_pipeline):
        self.eval_pipeline = eval_pipeline

    def evaluate(self):  # noqa: C901
        """Generate on ``eval_pipeline``, score with ``reward_fn`` / ``metric_fn`` (main process), tabulate."""
        logger.info("Evaluating model")
        sweep_arg, sweep_values = self.generate_sweep_kwarg if self.generate_sweep_kwarg is not None else (None, [None])
        stats: Dict[str, Any] = {}
        table: List[List[tuple]] = []
        columns: List[str] = []
        seq2seq = self.config.model.model_arch_type == "seq2seq"
        pad_id = self.tokenizer.pad_token_id
        total = len(self.eval_dataloader) * len(sweep_values)
        tbar = logging.tqdm(total=total, desc="[eval]", disable=not self.runtime.is_main_process, position=0, leave=True)

        for i_sweep, sweep_value in enumerate(sweep_values):
            suffix = f"@{sweep_arg}={sweep_value}" if sweep_value is not None else ""
            all_samples, all_prompts, all_sizes, all_meta = [], [], [], []
            t0 = time()
            # eval batches are dealt round-robin to the model replicas (all TP/PP ranks of a replica run the same batch)
            world, rank = self.runtime.dp_size, self.runtime.dp_rank
            for i_prompt, prompts in enumerate(self.eval_dataloader):
                tbar.set_description(f"[generation sweep {i_sweep + 1}/{len(sweep_values)} | eval batch {i_prompt + 1}/{len(self.eval_dataloader)}]")
                tbar.update()
                if i_prompt % world != rank:
                    continue
                metadata = {k: v for k, v in prompts.items() if k not in ("input_ids", "attention_mask")}
                extra = {sweep_arg: sweep_value} if self.generate_sweep_kwarg else {}
                samples = self.generate_eval(prompts["input_ids"], prompts["attention_mask"], **extra)
                if seq2seq:
                    samples = samples[:, 1:].contiguous()
                all_samples.extend(samples.tolist())
                all_prompts.extend(prompts["input_ids"].tolist())
                all_sizes.extend([prompts["input_ids"].shape[1]] * len(prompts["input_ids"]))
                all_meta.append(metadata)
            if self.runtime.distributed:  # one pickled gather per sweep value instead of collectives per batch
                if not self.runtime.is_replica_leader:  # model-parallel peers hold duplicates of their leader's samples
                    all_samples, all_prompts, all_sizes, all_meta = [], [], [], []
                shards = self.runtime.gather_objects((all_samples, all_prompts, all_sizes, all_meta))
                all_samples = sum((sh[0] for sh in shards), [])
                all_prompts = sum((sh[1] for sh in shards), [])
                all_sizes = sum((sh[2] for sh in shards), [])
                all_meta = sum((sh[3] for sh in shards), [])
            stats["time/generate"] = time() - t0

            if self.runtime.is_main_process:
                str_samples, str_prompts, str_outputs = self.decode(all_prompts, all_samples, all_sizes)
                columns = ["prompt", "output"]
                columns_data = [str_prompts, str_outputs]
                metadata = {}
                for m in all_meta:
                    for k, v in m.items():
                        metadata.setdefault(k, []).extend(v)
                if self.reward_fn:
                    logger.info("Computing rewards")
                    rewards = self.reward_fn(samples=str_samples, prompts=str_prompts, outputs=str_outputs,
                                             tokenizer=self.tokenizer, **metadata)
                    if len(rewards) and isinstance(rewards[0], torch.Tensor):
                        rewards = torch.tensor([float(r.sum()) for r in rewards], dtype=torch.float64)
                    elif len(rewards) and isinstance(rewards[0], list):
                        rewards = torch.tensor([sum(r) for r in rewards], dtype=torch.float64)
                    else:
                        rewards = torch.as_tensor(rewards, dtype=torch.float64)
                    stats[f"reward/mean{suffix}"] = rewards.mean().item()
                    columns.append("reward")
                    columns_data.append(rewards.tolist())
                if self.metric_fn:
                    logger.info("Computing metrics")
                    t1 = time()
                    metrics = self.metric_fn(samples=str_samples, prompts=str_prompts, outputs=str_outputs, **metadata)
                    stats["time/metric"] = time() - t1
                    for k, xs in metrics.items():
                        stats[f"metrics/{k}{suffix}"] = torch.as_tensor(xs, dtype=torch.float64).mean(-1).item()
                        if isinstance(xs, float):
                            continue
                        columns.append(k)
                        columns_data.append(xs if isinstance(xs, list) else torch.as_tensor(xs).tolist())
                if self.generate_sweep_kwarg:
                    columns.insert(0, sweep_arg)
                    columns_data.insert(0, [sweep_value] * len(str_samples))
                table.append(list(zip(*columns_data)))
        tbar.close()

        logger.info("Summarizing evaluation")
        if self.runtime.is_main_process and table:
            rows = sum(list(map(list, zip(*table))), [])
            title = f"Evaluation #{self.nth_evaluation}"
            for k, x in stats.items():
                if k.startswith("reward") or k.startswith("metrics"):
                    title += f" {k}: {significant(x)}"
            try:
                from rich.console import Console
                from rich.table import Table

                rich_table = Table(*columns, title=title, show_lines=True)
                for ix in range(min(max(min(3, len(rows)), len(sweep_values)), len(rows))):
                    rich_table.add_row(*[str(significant(x)) for x in rows[ix]])
                Console().print(rich_table)
            except Exception:  # pragma: no cover - rich missing
                logger.info(title)
            if self.runtime._tracker_kind == "wandb":
                import wandb

                stats["samples"] = wandb.Table(columns, rows)
        self.nth_evaluation += 1
        return stats

    # ---- optimisation loop ------------------------------------------------------------------------------------------------
    @contextmanager
    def _accumulate(self):
        """Micro-batch bookkeeping.  Gradient synchronisation happens inside ``opt.step()`` (fused reduce-scatter), so
        unlike the reference (``:502-516``) no ``no_sync`` context is required while accumulating."""
        self.mb_count += 1
        assert self.mb_count // self.num_mb <= self.config.train.total_steps, "Beyond total steps, something is wrong"
        with contextlib.nullcontext():
            yield

    def _checkpoint_name(self) -> str:
        return f"checkpoint_{self.iter_count:0{len(str(self.total_steps))}d}"

    def _save_checkpoint(self, directory: str):
        if self.config.train.save_optimizer:
            logger.info(f"Saving intermediate optimizer & model checkpoint into {directory}")
            self.save(directory)
        pretrained = os.path.join(directory, "hf_model")
        logger.info(f"Saving pretrained model into {pretrained}")
        self.save_pretrained(pretrained)

    def train_step(self, minibatch) -> Dict[str, Any]:
        """One optimizer step over the micro-batches of ``minibatch`` → (device-resident) stats."""
        times: Dict[str, float] = {}
        stats_accum = []
        fwd = bwd = 0.0
        for mb_index, microbatch in enumerate(minibatch):
            with self._accumulate():
                with self.runtime.phase("forward", times):
                    loss, stats = self.loss(microbatch)
                fwd += times["forward"]
                if mb_index == len(minibatch) - 1:
                    self._arm_grad_overlap()
                with self.runtime.phase("backward", times):
                    self.model.train()
                    loss.backward()
                    self.model.eval()
                bwd += times["backward"]
                stats_accum.append(stats)
        n = len(stats_accum)
        stats = {k: sum(s[k] for s in stats_accum) / self.num_mb for k in stats_accum[0]}
        self._pre_optimizer_step()
        if self.zero3 is not None:
            self.zero3.reduce_pending()
        self.opt.step()
        self.opt.zero_grad()
        if self.zero3 is not None:
            self.zero3.finish_step()
        self.scheduler.step()
        self.iter_count += 1
        self._after_weights_changed()
        stats["time/forward"] = fwd / self.num_mb
        stats["time/backward"] = bwd / self.num_mb
        return stats

    def _arm_grad_overlap(self):
        """Before the backward that completes an optimizer step: let the fused data-parallel optimizer reduce / update /
        gather every gradient bucket as soon as it is final, on a side stream, while the rest of the backward runs.  Not
        with model parallelism: its gradient fix-ups (``_pre_optimizer_step``) run after the backward."""
        rt = self.runtime
        if getattr(self.opt, "can_overlap", False) and rt.tp_size == 1 and rt.pp_size == 1:
            self.opt.host_prepare()
            self.opt.arm_overlap()

    _LOSS_KEYS = ("loss", "losses/loss", "losses/total_loss")

    def _check_finite(self, stats: Dict[str, Any], step: Optional[int] = None) -> bool:
        """Divergence watchdog (the reference has none): a non-finite loss is logged, its step never overwrites a
        checkpoint, and after ``trainer_kwargs.max_nonfinite_steps`` (default 8) consecutive ones training stops with an error
        instead of burning the remaining budget on NaN weights.  Reads the already-materialised statistics: no extra sync."""
        import math

        bad = [k for k in self._LOSS_KEYS if isinstance(stats.get(k), (int, float)) and not math.isfinite(stats[k])]
        if not bad:
            self._nonfinite_streak = 0
            return True
        self._nonfinite_streak = getattr(self, "_nonfinite_streak", 0) + 1
        limit = int((self.config.train.trainer_kwargs or {}).get("max_nonfinite_steps", 8))
        logger.warning(f"step {self.iter_count if step is None else step}: non-finite {bad[0]} "
                       f"({self._nonfinite_streak} in a row; abort at {limit})")
        if self._nonfinite_streak >= limit:
            raise FloatingPointError(
                f"training diverged: {bad[0]} has been non-finite for {self._nonfinite_streak} consecutive steps "
                f"(last healthy checkpoint is under {self.config.train.checkpoint_dir}); lower the learning rate or enable "
                "`train.parallel.grad_clip`")
        return False

    def learn(self):  # noqa: C901
        """Train from ``self.store``; checkpoint / evaluate on their intervals; returns the last eval results."""
        logger.info("Starting training")
        self.prepare_learning()
        self.iter_count = getattr(self, "_resumed_iter", 0) or self.iter_count
        self.nth_evaluation = 0
        results = self.evaluate()
        self.runtime.log(filter_non_scalars(results) if self.runtime._tracker_kind != "wandb" else results, step=self.iter_count)

        tbar = logging.tqdm(initial=self.iter_count, total=self.total_steps, disable=not self.runtime.is_main_process,
                            position=0, leave=True)
        best_reward = -float("inf")
        time_limit = _parse_max_time((self.config.train.trainer_kwargs or {}).get("max_time", getattr(self, "_max_time", None)))
        t_start = time()
        # Statistics are read one optimizer step late (`PendingStats`): the transfer of step i is queued right behind it, the host
        # blocks on it only after step i + 1 has been launched.  Steps on which something depends on the numbers right away
        # (checkpoint, evaluation, last step, wall-clock budget) are read synchronously.  `trainer_kwargs.lag_stats=False` restores
        # the strictly synchronous loop.
        lag_ok = bool((self.config.train.trainer_kwargs or {}).get("lag_stats", True)) and time_limit is None
        lagging: List[Any] = []  # at most one (PendingStats, step index)

        def flush():
            while lagging:
                handle, step = lagging.pop(0)
                st = handle.get()
                self._check_finite(st, step)
                self.runtime.log(st, step=step)

        for _ in range(self.config.train.epochs):
            for _ in range(self.n_inner_epochs):
                train_dataloader = self.create_train_dataloader()
                for minibatch in MiniBatchIterator(train_dataloader, self.mb_size, self.num_mb):
                    with self.runtime.nvtx("train_step"):
                        stats = self.train_step(minibatch)
                    for gi, lr in enumerate(self.scheduler.get_last_lr()):
                        stats[f"learning_rate_group_{gi}"] = lr
                    handle = PendingStats(stats)
                    flush()  # the previous step's numbers: this step is already running on the device
                    due = (self.iter_count % self.config.train.checkpoint_interval == 0 or self.iter_count >= self.total_steps
                           or self.iter_count % self.config.train.eval_interval == 0)
                    if lag_ok and not due:
                        lagging.append((handle, self.iter_count))
                        tbar.update()
                        continue
                    stats = handle.get()
                    healthy = self._check_finite(stats)
                    if healthy and (self.iter_count % self.config.train.checkpoint_interval == 0
                                    or self.iter_count >= self.total_steps):
                        self._save_checkpoint(os.path.join(self.config.train.checkpoint_dir, self._checkpoint_name()))
                    if self.iter_count % self.config.train.eval_interval == 0 or self.iter_count >= self.total_steps:
                        results = self.evaluate()
                        stats.update(results)
                        if self.config.train.save_best:
                            if stats.get("reward/mean", -float("inf")) > best_reward:
                                best_reward, do_save = stats.get("reward/mean"), True
                            elif stats.get("metrics/reward", -float("inf")) > best_reward:
                                best_reward, do_save = stats.get("metrics/reward"), True
                            else:
                                do_save = False
                            flag = torch.tensor(int(do_save), device=self.runtime.device)
                            self.runtime.all_reduce(flag, "max")
                            if bool(flag.item()):
                                self._save_checkpoint(os.path.join(self.config.train.checkpoint_dir, "best_checkpoint"))
                    desc = " | ".join(f"{k}: {v:.2f}" for k, v in stats.items() if k.startswith("loss"))
                    tbar.set_description(f"[{desc}]")
                    tbar.update()
                    self.runtime.log(stats, step=self.iter_count)
                    if self.iter_count >= self.total_steps:
                        tbar.close()
                        return results
                    if time_limit is not None:
                        # wall-clock budget (the reference's NeMo path stops through PTL's StatelessTimer,
                        # ``trlx/trainer/nemo_ilql_trainer.py:72-77``): all ranks agree, checkpoint, stop cleanly
                        up = torch.tensor(int(time() - t_start >= time_limit), device=self.runtime.device)
                        self.runtime.all_reduce(up, "max")
                        if bool(up.item()):
                            logger.info(f"max_time reached after {self.iter_count} steps: saving a checkpoint and stopping")
                            self._save_checkpoint(os.path.join(self.config.train.checkpoint_dir, self._checkpoint_name()))
                            tbar.close()
                            return results
                self.post_backward_callback()
            flush()  # before the next rollouts log under a later step index
            self.post_epoch_callback()
        flush()
        tbar.close()
        return results

    # ---- hooks ------------------------------------------------------------------------------------------------------------
    @abstractmethod
    def create_train_dataloader(self):
        """New dataloader (fresh shuffle) for one inner epoch."""

    @abstractmethod
    def get_arch(self, config: TRLConfig):
        """Build the wrapped model for this method."""

    @abstractmethod
    def loss(self, batch) -> Tuple[torch.Tensor, Dict]:
        """Loss and statistics of one micro-batch."""

    @abstractmethod
    def prepare_learning(self):
        """Set ``eval_dataloader``, ``total_steps``, ``n_inner_epochs`` … before training."""

    def post_backward_callback(self):
        """After each pass over the store."""

    def post_epoch_callback(self):
        """After each outer epoch."""
