"""PPO: steer a GPT-2 towards positive movie reviews (reference: examples/ppo_sentiments.py; BASELINE config B2)."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import GPT2_SMALL, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_ppo_config


def main(hparams={}):
    config = TRLConfig.update(default_ppo_config().to_dict(), hparams)
    if isinstance(config.model.model_path, str):
        config.model.model_path = offline_model(config.model.model_path, GPT2_SMALL)
    device = int(os.environ.get("LOCAL_RANK", 0)) if torch.cuda.is_available() else -1
    sentiment_fn = sentiment_scorer(device)

    def reward_fn(samples: List[str], **kwargs) -> List[float]:
        return [s["POSITIVE"] for s in sentiment_fn(samples)]

    texts, _ = load_imdb()
    prompts = [" ".join(review.split()[:4]) for review in texts]  # a few words off each review
    return trlx.train(reward_fn=reward_fn, prompts=prompts, eval_prompts=["I don't know much about Hungarian underground"] * 256,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
