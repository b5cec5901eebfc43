"""PPO with *dense* (per-token) rewards: negative first half, positive second half (reference: examples/ppo_dense_sentiments.py)."""
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

    def dense_reward_fn(samples: List[str], prompts: List[str], outputs: List[str], tokenizer, **kwargs) -> List[List[float]]:
        halves = [s.split(".") for s in samples]
        first = [".".join(h[: len(h) // 2]) for h in halves]
        second = [".".join(h[len(h) // 2:]) for h in halves]
        neg_first = [s["NEGATIVE"] for s in sentiment_fn(first)]
        pos_second = [s["POSITIVE"] for s in sentiment_fn(second)]
        tok_scores = []
        for response, a, b in zip(outputs, neg_first, pos_second):
            n = max(len(tokenizer(response).input_ids), 1)
            row = [0.0] * n
            row[n // 2] = a   # reward for the negative opening lands mid-response …
            row[-1] = b       # … and for the positive ending on the last token
            tok_scores.append(row)
        return tok_scores

    texts, _ = load_imdb()
    prompts = [" ".join(review.split()[:4]) for review in texts]
    return trlx.train(reward_fn=dense_reward_fn, prompts=prompts,
                      eval_prompts=["I don't know much about Hungarian underground"] * 256, config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
