"""ILQL trainer (offline RL from samples + rewards).

Parity: ``trlx/trainer/accelerate_ilql_trainer.py`` — module-level ``make_experience`` ``:30-100``, trainer
``get_arch`` ``:121-136``, target-Q sync ``:138-140``, ``loss`` ``:142-160``, ``prepare_learning`` ``:165-177``, seq2seq
experience ``:179-244``.  On CUDA the loss runs in *gathered form* (:meth:`...modeling_ilql.AutoModelForCausalLMWithILQLHeads.loss_parts`),
so the five ``[B, A, V]`` tensors of the reference are never materialised.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Union, cast

import numpy as np
import torch

from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.ilql_types import ILQLBatch, ILQLSeq2SeqBatch
from trlx_b200.models.modeling_ilql import (AutoModelForCausalLMWithILQLHeads, AutoModelForSeq2SeqLMWithILQLHeads,
                                            ILQLConfig)
from trlx_b200.pipeline.offline_pipeline import (ILQLRolloutStorage, ILQLSeq2SeqRolloutStorage, tokenize_dialogue)
from trlx_b200.trainer import register_trainer
from trlx_b200.trainer.accelerate_base_trainer import AccelerateRLTrainer
from trlx_b200.trainer.accelerate_ppo_trainer import _is_config_like
from trlx_b200.utils import logging, to_device

logger = logging.get_logger(__name__)


def _print_table(title: str, columns: Sequence[str], row: Sequence[str]) -> None:
    try:
        from rich.console import Console
        from rich.table import Table

        table = Table(*columns, title=title, show_lines=True)
        table.add_row(*row)
        Console().print(table)
    except Exception:  # pragma: no cover
        logger.info(f"{title}: " + " | ".join(f"{c}={r}" for c, r in zip(columns, row)))


def _standardise(rewards) -> torch.Tensor:
    returns = torch.tensor(rewards, dtype=torch.float64)
    returns = returns - returns.mean()
    std = returns.std()
    if not torch.isnan(std):
        returns = returns / (std + torch.finfo(returns.dtype).eps)
    return returns


def _log_lengths(prompt_lengths, output_lengths, sample_lengths):
    row = [f"{l.mean():.2f} ∈ [{min(l)}, {max(l)}]" for l in (prompt_lengths, output_lengths, sample_lengths)]
    _print_table("Experience String Stats (mean ∈ [min, max])", ["Prompt Length", "Output Length", "Sample Length"], row)


def make_experience(samples, rewards, tokenizer=None, max_length: int = 2048, verbose: bool = True) -> ILQLRolloutStorage:
    """Tokenise dialogues and lay out ILQL indices.

    For every output token at position ``i`` the *action index* is ``i − 1`` (the state from which it is produced);
    ``states_ixs`` are the action indices plus the final position; ``dones`` is 1 everywhere except the terminal
    state; the (dataset-standardised) return is placed on the last action.
    """
    if verbose:
        logger.info("Collecting rollouts")
    if tokenizer is not None:
        samples = [tokenize_dialogue(s, tokenizer, max_length) for s in samples]
    all_input_ids, all_actions, all_states, all_dones = [], [], [], []
    for sample in samples:
        flat: List[int] = []
        actions: List[int] = []
        for msg in sample:
            if msg.is_output:
                actions.extend(range(len(flat) - 1, len(flat) + len(msg.tokens) - 1))
            flat.extend(msg.tokens)
        states = actions + [len(flat) - 1]
        all_input_ids.append(torch.tensor(flat, dtype=torch.long))
        all_actions.append(torch.tensor(actions, dtype=torch.long))
        all_states.append(torch.tensor(states, dtype=torch.long))
        all_dones.append(torch.tensor([1] * (len(states) - 1) + [0], dtype=torch.long))

    main = os.environ.get("RANK", "0") == "0"
    if tokenizer is not None and main and verbose and len(all_input_ids):
        split = int(all_states[0][1]) if len(all_states[0]) > 1 else len(all_input_ids[0])
        _print_table("Sample Example", ["Prompt", "Response", "Reward"],
                     [tokenizer.decode(all_input_ids[0][:split]), tokenizer.decode(all_input_ids[0][split:]), str(rewards[0])])
    if main and verbose and len(all_input_ids):
        sample_lengths = np.array([len(x) for x in all_input_ids])
        output_lengths = np.array([len(x) for x in all_actions])
        _log_lengths(sample_lengths - output_lengths, output_lengths, sample_lengths)

    returns = _standardise(rewards)
    rs = [torch.zeros(len(a)) for a in all_actions]
    for r, ret in zip(rs, returns):
        r[-1] = ret
    attention_mask = [torch.ones(len(x), dtype=torch.long) for x in all_input_ids]
    return ILQLRolloutStorage(all_input_ids, attention_mask, rs, all_states, all_actions, all_dones)


@register_trainer
class AccelerateILQLTrainer(AccelerateRLTrainer):
    def __init__(self, config: TRLConfig, **kwargs):
        super().__init__(config, **kwargs)
        if not isinstance(config.method, ILQLConfig):
            raise ValueError("config.method must be ILQLConfig")
        self.ilql: ILQLConfig = cast(ILQLConfig, config.method)
        self.generate_kwargs = dict(
            {k: v for k, v in config.method.gen_kwargs.items()
             if not (self.generate_sweep_kwarg and k == self.generate_sweep_kwarg[0])},
            max_length=self.max_length, logit_mask=self.logit_mask,
            eos_token_id=self.tokenizer.eos_token_id if self.tokenizer else 0,
            pad_token_id=self.tokenizer.pad_token_id if self.tokenizer else 0,
        )
        self.generate_experience_kwargs = None

    def _generate(self, input_ids, attention_mask, kwargs):
        """ILQL sampling on the CUDA engine (``engine/ilql.py``) when the model allows it, else the PyTorch loop of the model."""
        eng = self._ilql_engine()
        if eng is None:
            return super()._generate(input_ids, attention_mask, kwargs)
        return eng.generate(input_ids, attention_mask, **kwargs)

    def _ilql_engine(self):
        if getattr(self, "_ilql_engine_obj", None) is not None or getattr(self, "_ilql_engine_failed", False):
            return getattr(self, "_ilql_engine_obj", None)
        self._ilql_engine_obj = None
        try:
            from trlx_b200 import ops
            from trlx_b200.engine.ilql import ILQLDecodeEngine

            why = ("no CUDA device" if not self.runtime.cuda else
                   "seq2seq models generate through the PyTorch loop" if self.config.model.model_arch_type == "seq2seq" else
                   "ZeRO-3 partitions the parameters" if getattr(self, "zero3", None) is not None else
                   ILQLDecodeEngine.why_not(self.model))
            if why is None:
                self._ilql_engine_obj = ILQLDecodeEngine(self.model, self.tokenizer.pad_token_id, self.tokenizer.eos_token_id,
                                                         seed=self.config.train.seed + self.runtime.rank)
            else:
                if self.runtime.cuda:
                    logger.warning(f"ILQL generation uses the PyTorch sampling loop, not the CUDA engine: {why}")
                self._ilql_engine_failed = True
        except ImportError as err:
            logger.warning(f"ILQL decode engine unavailable: {err}")
            self._ilql_engine_failed = True
        return self._ilql_engine_obj

    def _after_weights_changed(self):
        super()._after_weights_changed()
        eng = getattr(self, "_ilql_engine_obj", None)
        if eng is not None:
            eng.mark_dirty()

    def get_arch(self, config):
        cls = (AutoModelForSeq2SeqLMWithILQLHeads if config.model.model_arch_type == "seq2seq"
               else AutoModelForCausalLMWithILQLHeads)
        from_fn = cls.from_config if _is_config_like(config.model.model_path) else cls.from_pretrained
        return from_fn(config.model.model_path, two_qs=config.method.two_qs, alpha=config.method.alpha,
                       peft_config=config.model.peft_config, **config.model.model_extra_configs)

    def post_backward_callback(self):
        if self.iter_count % self.config.method.steps_for_target_q_sync == 0:
            self.model.sync_target_q_heads()

    def loss(self, batch: Union[ILQLBatch, ILQLSeq2SeqBatch]):
        batch = to_device(batch, self.runtime.device)
        if self.config.model.model_arch_type == "seq2seq":
            logits, qs, target_qs, vs, _, _ = self.model(input_ids=batch.input_ids, attention_mask=batch.attention_mask,
                                                         actions_ixs=batch.actions_ixs, states_ixs=batch.states_ixs,
                                                         decoder_input_ids=batch.decoder_input_ids)
            return self.ilql.loss((logits, (qs, target_qs, vs)), batch)
        if batch.input_ids.is_cuda and hasattr(self.model, "loss_parts"):
            return self.ilql.loss(self.model.loss_parts(batch), batch)
        logits, qs, target_qs, vs, _ = self.model(input_ids=batch.input_ids, attention_mask=batch.attention_mask,
                                                  actions_ixs=batch.actions_ixs, states_ixs=batch.states_ixs)
        return self.ilql.loss((logits, (qs, target_qs, vs)), batch)

    def _sharded_loader(self):
        sampler = None
        if self.runtime.distributed:
            from torch.utils.data.distributed import DistributedSampler

            sampler = DistributedSampler(self.store, num_replicas=self.runtime.dp_size, rank=self.runtime.dp_rank,
                                         shuffle=True, seed=self.config.train.seed, drop_last=True)
            sampler.set_epoch(getattr(self, "_loader_epoch", 0))
            self._loader_epoch = getattr(self, "_loader_epoch", 0) + 1
        return self.store.create_loader(self.config.train.batch_size, sampler=sampler)

    def create_train_dataloader(self):
        return self._sharded_loader()

    def prepare_learning(self):
        self.train_dataloader = self.create_train_dataloader()
        self.eval_dataloader = self.eval_pipeline.create_loader(self.config.train.batch_size)
        self.n_inner_epochs = 1
        self.total_steps = min(self.config.train.epochs * len(self.train_dataloader), self.config.train.total_steps)

    def make_experience_seq2seq(self, samples, rewards, max_length: int = 2048):
        logger.info("Collecting rollouts")
        if self.tokenizer:
            samples = [tokenize_dialogue(s, self.tokenizer, max_length) for s in samples]
            keep = [i for i, s in enumerate(samples) if len(s) >= 2 and len(s[1].tokens) > 1]
            if len(keep) != len(samples):  # truncation left no room for an output (the reference would raise IndexError)
                logger.warning(f"dropping {len(samples) - len(keep)} samples whose output was truncated away "
                               f"(max_length={max_length})")
                samples, rewards = [samples[i] for i in keep], [rewards[i] for i in keep]
        all_input_ids, all_output_ids, all_actions, all_states, all_dones = [], [], [], [], []
        for sample in samples:
            all_input_ids.append(torch.tensor(sample[0].tokens, dtype=torch.long))
            all_output_ids.append(torch.tensor(sample[1].tokens, dtype=torch.long))
            length = 0
            actions: List[int] = []
            for phrase in sample:
                if phrase.is_output:
                    length = len(phrase.tokens)
                    actions.extend(range(0, length - 1))
            states = actions + [length - 1]
            all_actions.append(torch.tensor(actions, dtype=torch.long))
            all_states.append(torch.tensor(states, dtype=torch.long))
            all_dones.append(torch.tensor([1] * (len(states) - 1) + [0], dtype=torch.long))
        if self.tokenizer and self.runtime.is_main_process and len(samples):
            _print_table("Sample Example", ["Prompt", "Response", "Reward"],
                         [self.tokenizer.decode(all_input_ids[0]), self.tokenizer.decode(all_output_ids[0]), str(rewards[0])])
            out_l = np.array([len(x) for x in all_output_ids])
            in_l = np.array([len(x) for x in all_input_ids])
            _log_lengths(in_l, out_l, in_l + out_l)
        returns = torch.tensor(rewards, dtype=torch.float64)
        returns = (returns - returns.mean()) / (returns.std() + torch.finfo(returns.dtype).eps)
        rs = [torch.zeros(len(a)) for a in all_actions]
        for r, ret in zip(rs, returns):
            r[-1] = ret
        attention_mask = [torch.ones(len(x), dtype=torch.long) for x in all_input_ids]
        self.store = ILQLSeq2SeqRolloutStorage(all_input_ids, attention_mask, all_output_ids, rs, all_states, all_actions,
                                               all_dones)

    def make_experience(self, samples, rewards, max_length: int = 2048):
        if self.config.model.model_arch_type == "seq2seq":
            return self.make_experience_seq2seq(samples, rewards, max_length)
        self.store = make_experience(samples, rewards, self.tokenizer, max_length=max_length, verbose=True)
