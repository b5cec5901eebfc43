"""Rejection fine-tuning on sentiment reward (reference: examples/rft_sentiments.py)."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import GPT2_SMALL, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_sft_config
from trlx_b200.trainer.accelerate_rft_trainer import RFTConfig

default_config = default_sft_config().evolve(   # same optimiser recipe as SFT at the PPO learning rate, RFT method on top
    train=dict(trainer="AccelerateRFTTrainer", batch_size=32),
    model=dict(model_path=offline_model("lvwerra/gpt2-imdb", GPT2_SMALL)),
    optimizer=dict(kwargs=dict(lr=3e-5)),
    scheduler=dict(kwargs=dict(eta_min=3e-5)),
)
default_config.method = RFTConfig(name="RFTConfig", n_generations_per_prompt=4, n_improve_steps=1, start_percentile=0.9,
                                  end_percentile=0.95, gen_kwargs=dict(default_config.method.gen_kwargs, temperature=1.0))


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
