"""Rejection fine-tuning on sentiment reward (reference: examples/rft_sentiments.py)."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import GPT2_SMALL, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.trainer.accelerate_rft_trainer import RFTConfig

default_config = TRLConfig(
    train=TrainConfig(seq_length=1024, epochs=100, total_steps=1000, batch_size=32, checkpoint_interval=10000, eval_interval=100,
                      pipeline="PromptPipeline", trainer="AccelerateRFTTrainer"),
    model=ModelConfig(model_path=offline_model("lvwerra/gpt2-imdb", GPT2_SMALL), num_layers_unfrozen=-1),
    tokenizer=TokenizerConfig(tokenizer_path="gpt2", truncation_side="right"),
    optimizer=OptimizerConfig(name="adamw", kwargs=dict(lr=3e-5, betas=(0.9, 0.95), eps=1.0e-8, weight_decay=1.0e-6)),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs=dict(T_max=1e12, eta_min=3e-5)),
    method=RFTConfig(name="RFTConfig", n_generations_per_prompt=4, start_percentile=0.9, end_percentile=0.95, n_improve_steps=1,
                     gen_kwargs=dict(max_new_tokens=40, top_k=0, top_p=1.0, temperature=1.0, do_sample=True)),
)


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    device = int(os.environ.get("LOCAL_RANK", 0)) if torch.cuda.is_available() else -1
    sentiment_fn = sentiment_scorer(device)

    def reward_fn(samples: List[str], **kwargs) -> List[float]:
        return [s["POSITIVE"] for s in sentiment_fn(samples)]

    texts, _ = load_imdb(512)
    prompts = [" ".join(review.split()[:4]) for review in texts]
    return trlx.train(reward_fn=reward_fn, prompts=prompts, eval_prompts=["I don't know much about Hungarian underground"] * 256,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
