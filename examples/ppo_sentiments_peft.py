"""PPO with a LoRA adapter instead of unfrozen layers (reference: examples/ppo_sentiments_peft.py)."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import GPT2_SMALL, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_ppo_config
from trlx_b200.models.peft import LoraConfig, TaskType


def main(hparams={}):
    config = TRLConfig.update(default_ppo_config().to_dict(), hparams)
    if isinstance(config.model.model_path, str):
        config.model.model_path = offline_model(config.model.model_path, GPT2_SMALL)
    device = int(os.environ.get("LOCAL_RANK", 0)) if torch.cuda.is_available() else -1
    sentiment_fn = sentiment_scorer(device)
    config.model.peft_config = LoraConfig(r=8, task_type=TaskType.CAUSAL_LM, lora_alpha=32, lora_dropout=0.1)

    def reward_fn(samples: List[str], **kwargs) -> List[float]:
        return [s["POSITIVE"] for s in sentiment_fn(samples)]

    texts, _ = load_imdb()
    prompts = [" ".join(review.split()[:4]) for review in texts]
    return trlx.train(reward_fn=reward_fn, prompts=prompts, eval_prompts=["I don't know much about Hungarian underground"] * 256,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
