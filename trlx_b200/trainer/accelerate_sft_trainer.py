"""Supervised fine-tuning trainer.

Parity: ``trlx/trainer/accelerate_sft_trainer.py`` — ``SFTConfig`` ``:16-26``, plain causal LM (+ optional adapter)
``:40-61``, masked cross-entropy ``:63-73``, ``make_experience`` ``:92-97``.  The loss goes through the fused LM-head
kernel on CUDA: ``CE = −log p(label)`` is read off the GEMM epilogue's online logsumexp, so the ``[B,T,V]`` logits of
the reference are not materialised.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from trlx_b200 import ops
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.method_configs import MethodConfig, register_method
from trlx_b200.models.modeling_base import PreTrainedModelWrapper, base_lm, export_base_state_dict
from trlx_b200.pipeline.offline_pipeline import DialogStore, PromptPipeline, tokenize_dialogue
from trlx_b200.trainer import register_trainer
from trlx_b200.trainer.accelerate_base_trainer import AccelerateRLTrainer
from trlx_b200.trainer.accelerate_ppo_trainer import _is_config_like


@dataclass
@register_method
class SFTConfig(MethodConfig):
    """:param gen_kwargs: generation kwargs used during evaluation"""

    gen_kwargs: dict


class CausalLMWrapper(PreTrainedModelWrapper):
    """Head-less wrapper: the plain LM (+ adapter) behind the common save/load/generate interface."""

    _supported_args = ["peft_config"]
    arch_type = "causal"

    def __init__(self, base_model, peft_config=None):
        super().__init__(base_model, peft_config=peft_config)

    def forward(self, input_ids=None, attention_mask=None, labels=None, position_ids=None, **kw):
        """With ``labels``: returns an object whose ``.loss`` is the mean next-token CE over labels != -100."""
        lm = base_lm(self.base_model)
        if labels is None or self.peft_type in ("PROMPT_TUNING", "PREFIX_TUNING"):
            return self.base_model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                   labels=labels, **kw)
        out = self.base_model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                              compute_logits=False, **kw)
        h = out.last_hidden_state[:, :-1]
        tgt = labels[:, 1:]
        lp, _ = ops.fused_logprob(h, lm.lm_head.weight, lm.lm_head.bias, torch.where(tgt == -100, -1, tgt))
        n = (tgt != -100).sum().clamp_min(1)
        out.loss = -(lp.float() * (tgt != -100)).sum() / n
        return out

    def generate(self, *args, **kwargs):
        from trlx_b200.models.generation import generate

        return generate(self.base_model, *args, **kwargs)

    def state_dict(self, *args, heads_only: bool = False, **kwargs):
        if heads_only:
            return {}
        # a plain LM is exported with bare HF keys so that any HF-style loader can read it
        return export_base_state_dict(self.base_model, prefix="")


@register_trainer
class AccelerateSFTTrainer(AccelerateRLTrainer):
    def __init__(self, config: TRLConfig, **kwargs):
        super().__init__(config, **kwargs)
        self.generate_kwargs = dict(
            {k: v for k, v in config.method.gen_kwargs.items()
             if not (self.generate_sweep_kwarg and k == self.generate_sweep_kwarg[0])},
            eos_token_id=self.tokenizer.eos_token_id, pad_token_id=self.tokenizer.pad_token_id)
        self.generate_experience_kwargs = None

    def get_arch(self, config):
        from_fn = CausalLMWrapper.from_config if _is_config_like(config.model.model_path) else CausalLMWrapper.from_pretrained
        return from_fn(config.model.model_path, peft_config=config.model.peft_config, **config.model.model_extra_configs)

    def loss(self, batch):
        dev = self.runtime.device
        input_ids, attention_mask = batch["input_ids"].to(dev), batch["attention_mask"].to(dev)
        labels = (batch["labels"] if "labels" in batch else batch["input_ids"]).to(dev).clone()
        labels[~attention_mask.bool()] = -100
        loss = self.model(input_ids=input_ids, attention_mask=attention_mask, labels=labels).loss
        return loss, {"loss": loss.detach()}

    def create_train_dataloader(self):
        sampler = None
        if self.runtime.distributed:
            from torch.utils.data.distributed import DistributedSampler

            sampler = DistributedSampler(self.store, num_replicas=self.runtime.dp_size, rank=self.runtime.dp_rank,
                                         shuffle=False, drop_last=True)
        if sampler is not None:
            try:
                return self.store.create_loader(self.config.train.batch_size, sampler=sampler)
            except TypeError:
                pass
        return self.store.create_loader(self.config.train.batch_size)

    def prepare_learning(self):
        self.train_dataloader = self.create_train_dataloader()
        self.eval_dataloader = self.eval_pipeline.create_loader(self.config.train.batch_size)
        self.n_inner_epochs = 1
        self.total_steps = min(self.config.train.epochs * len(self.train_dataloader), self.config.train.total_steps)

    def make_experience(self, samples, seq_length):
        if isinstance(samples[0], str):
            self.store = PromptPipeline(samples, seq_length, self.tokenizer)
        else:
            dialogs = [tokenize_dialogue(d, self.tokenizer, seq_length) for d in samples]
            self.store = DialogStore(dialogs, self.tokenizer)
