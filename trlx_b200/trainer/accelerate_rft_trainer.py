"""Rejection fine-tuning trainer (generate N per prompt → keep the best by per-prompt quantile → SFT on them).

Parity: ``trlx/trainer/accelerate_rft_trainer.py`` — ``RFTConfig`` ``:18-44``, grow step ``:117-150`` (sample, gather
across ranks, score on the main process, broadcast), improve step ``:152-172`` (threshold schedule between
``start_percentile`` and ``end_percentile``, de-duplication).
"""
from __future__ import annotations

import itertools
from collections import defaultdict
from dataclasses import dataclass

import numpy as np

from trlx_b200.data.method_configs import MethodConfig, register_method
from trlx_b200.pipeline.offline_pipeline import PromptPipeline
from trlx_b200.trainer import register_trainer
from trlx_b200.trainer.accelerate_sft_trainer import AccelerateSFTTrainer
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


@dataclass
@register_method
class RFTConfig(MethodConfig):
    """
    :param gen_kwargs: generation kwargs
    :param start_percentile: per-prompt score quantile used as the first acceptance threshold
    :param end_percentile: quantile reached at the last improve step of a growth cycle
    :param n_improve_steps: improve steps (threshold raises) between two generation (grow) steps
    :param n_generations_per_prompt: samples drawn per prompt in a grow step
    """

    gen_kwargs: dict
    start_percentile: float = 0.7
    end_percentile: float = 0.95
    n_improve_steps: int = 4
    n_generations_per_prompt: int = 32


@register_trainer
class AccelerateRFTTrainer(AccelerateSFTTrainer):
    def loss(self, batch):
        dev = self.runtime.device
        input_ids, attention_mask = batch["input_ids"].to(dev), batch["attention_mask"].to(dev)
        loss = self.model(input_ids=input_ids, attention_mask=attention_mask, labels=input_ids.clone()).loss
        return loss, {"loss": loss.detach()}

    def prepare_learning(self):
        self.epoch_count = 0
        self.iter_count = 0
        self.n_inner_epochs = 1
        # the number of selected samples varies per improve step → total steps come straight from the config
        self.total_steps = self.config.train.total_steps
        self.generations_per_prompt = defaultdict(list)
        self.eval_dataloader = self.eval_pipeline.create_loader(self.config.train.batch_size)
        self.make_experience()

    def add_prompt_pipeline(self, pipeline: PromptPipeline):
        sampler = None
        if self.runtime.distributed:
            from torch.utils.data.distributed import DistributedSampler

            sampler = DistributedSampler(pipeline, num_replicas=self.runtime.dp_size, rank=self.runtime.dp_rank, shuffle=False)
        self.prompt_dataloader = pipeline.create_loader(self.config.train.batch_size, sampler=sampler)

    def post_epoch_callback(self):
        self.make_experience()
        self.epoch_count += 1

    def make_experience(self):  # noqa: C901
        method = self.config.method
        rt = self.runtime
        if self.epoch_count % method.n_improve_steps == 0:
            generations = []
            for batch in logging.tqdm(self.prompt_dataloader, desc="Generating", disable=not rt.is_main_process):
                for _ in range(method.n_generations_per_prompt):
                    samples = self.generate(batch["input_ids"], batch["attention_mask"])
                    _, str_prompts, str_outputs = self.decode(batch["input_ids"], samples, append_eos_token=True)
                    generations.extend({"prompt": p, "output": o} for p, o in zip(str_prompts, str_outputs))
            if rt.distributed:
                generations = list(itertools.chain(*rt.gather_objects(generations)))
            if rt.is_main_process:
                scores = self.reward_fn(samples=[g["prompt"] + g["output"] for g in generations],
                                        prompts=[g["prompt"] for g in generations],
                                        outputs=[g["output"] for g in generations])
                scores = [float(s) for s in scores]
            else:
                scores = None
            scores = rt.broadcast_object(scores, src=0)
            for g, s in zip(generations, scores):
                self.generations_per_prompt[g["prompt"]].append({"output": g["output"], "score": s})

        scores = [[x["score"] for x in self.generations_per_prompt[p]] for p in self.generations_per_prompt]
        delta = (method.end_percentile - method.start_percentile) / method.n_improve_steps
        percentile = method.start_percentile + delta * (self.epoch_count % method.n_improve_steps)
        thresholds = np.array([np.quantile(np.array(s), percentile) for s in scores])
        # quantised rewards: never keep the minimum, never drop the maximum
        thresholds = np.clip(thresholds, thresholds.min() + 1e-3, thresholds.max() - 1e-3)
        selected = []
        for prompt, threshold in zip(self.generations_per_prompt, thresholds):
            for x in self.generations_per_prompt[prompt]:
                if x["score"] >= threshold:
                    selected.append((prompt, x["output"]))
        selected = sorted(set(selected))
        flat_scores = np.hstack(scores) if scores else np.zeros(1)
        rt.log({"scores_mean": float(np.mean(flat_scores)), "scores_max": float(np.max(flat_scores)),
                "thresholds_mean": float(np.mean(thresholds)), "len_samples_selected": len(selected)}, step=self.iter_count)
        if len(selected):
            # (prompt, output) PAIRS, as in the reference (``accelerate_rft_trainer.py:196-198``): the tokenizer encodes the two
            # segments separately, so no BPE merge forms across the prompt / output boundary
            self.store = PromptPipeline([[p, o] for p, o in selected], max_prompt_length=2048, tokenizer=self.tokenizer,
                                        add_special_tokens=True)
