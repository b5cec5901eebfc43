"""Supervised fine-tuning on the chosen HH responses (reference: examples/hh/sft_hh.py)."""
import json
import os
import sys

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, offline_model, synthetic_dialogues
from examples.hh.ppo_hh import apply_preset
from examples.hh.reward import create_reward_fn
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_sft_config

default_config = default_sft_config().evolve(
    train=dict(total_steps=10000, batch_size=4, eval_interval=1000, checkpoint_dir="checkpoints/sft_hh"),
    model=dict(model_path="EleutherAI/gpt-j-6B"),
    tokenizer=dict(tokenizer_path="EleutherAI/gpt-j-6B", truncation_side="left"),
    optimizer=dict(kwargs=dict(lr=1e-6)),
    scheduler=dict(kwargs=dict(T_max=100000000, eta_min=1e-6)),
    method=dict(gen_kwargs=dict(max_new_tokens=128, top_k=20)),
)
apply_preset(default_config, os.environ.get("CONFIG_NAME"), "sft_hh")


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    if isinstance(config.model.model_path, str):
        config.model.model_path = offline_model(config.model.model_path, GPT2_TINY)
    data = synthetic_dialogues(1024)
    reward_fn = create_reward_fn(os.environ.get("REWARD_CHECKPOINT"), delta_reward=False)
    return trlx.train(config=config, samples=[x["prompt"] + x["chosen"] for x in data[:-64]],
                      eval_prompts=[x["prompt"] for x in data[-64:]][:280],
                      metric_fn=lambda **kwargs: {"reward": reward_fn(**kwargs)},
                      stop_sequences=["Human:", "human:", "Assistant:", "assistant:"])


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
