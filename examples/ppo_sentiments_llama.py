"""PPO on a Llama-family policy, top-2 layers trainable (reference: examples/ppo_sentiments_llama.py; BASELINE config B4)."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import LLAMA_TINY, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_ppo_config


def llama_config():
    """The default PPO recipe on a LLaMA-2-7B policy: lower learning rate, adaptive KL (target 6), 400 steps."""
    return default_ppo_config().evolve(
        train=dict(total_steps=400, save_best=False),
        model=dict(model_path=offline_model("NousResearch/Llama-2-7b-hf", LLAMA_TINY)),
        tokenizer=dict(tokenizer_path="NousResearch/Llama-2-7b-hf"),
        optimizer=dict(kwargs=dict(lr=1e-5)),
        scheduler=dict(kwargs=dict(T_max=10000, eta_min=1.0e-5)),
        method=dict(target=6),
    )


def main(hparams={}):
    config = TRLConfig.update(llama_config().to_dict(), hparams)
    device = int(os.environ.get("LOCAL_RANK", 0)) if torch.cuda.is_available() else -1
    sentiment_fn = sentiment_scorer(device)

    def reward_fn(samples: List[str], **kwargs) -> List[float]:
        return [s["POSITIVE"] for s in sentiment_fn(samples)]

    texts, _ = load_imdb()
    prompts = [" ".join(review.split()[:4]) for review in texts]
    return trlx.train(reward_fn=reward_fn, prompts=prompts, eval_prompts=["I don't know much about Hungarian underground"] * 64,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
