"""ILQL on Helpful/Harmless preference pairs: chosen = +1, rejected = −1 (reference: examples/hh/ilql_hh.py)."""
import json
import os
import sys
from itertools import islice

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, offline_model, synthetic_dialogues
from examples.hh.ppo_hh import apply_preset
from examples.hh.reward import create_reward_fn
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_ilql_config

default_config = default_ilql_config().evolve(
    train=dict(seq_length=1024, batch_size=4, total_steps=20000, checkpoint_interval=10000, eval_interval=1000,
               checkpoint_dir="checkpoints/ilql_hh"),
    model=dict(model_path="EleutherAI/gpt-j-6B"),
    tokenizer=dict(tokenizer_path="EleutherAI/gpt-j-6B", truncation_side="left"),
    optimizer=dict(kwargs=dict(lr=1e-6)),
    scheduler=dict(kwargs=dict(T_max=1000000000, eta_min=1e-6)),
    method=dict(tau=0.6, alpha=0.0001, steps_for_target_q_sync=1, gen_kwargs=dict(max_new_tokens=128, beta=[1, 4])),
)
_ILQL_MODELS = {"125M": "EleutherAI/pythia-125m-deduped", "1B": "EleutherAI/pythia-1.4b-deduped", "6B": "EleutherAI/pythia-6.9b-deduped",
                "20B": "EleutherAI/gpt-neox-20b"}
_name = os.environ.get("CONFIG_NAME")
if _name in _ILQL_MODELS:
    apply_preset(default_config, _name, "ilql_hh")
    default_config.model.model_path = _ILQL_MODELS[_name]
    default_config.train.batch_size = {"125M": 16, "1B": 8, "6B": 4, "20B": 1}[_name]
    default_config.train.total_steps = 3000 if _name == "20B" else 20000


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    if isinstance(config.model.model_path, str):
        config.model.model_path = offline_model(config.model.model_path, GPT2_TINY)
    data = synthetic_dialogues(1024)
    train, test = data[:-64], data[-64:]
    prompts_outputs = sum(([[x["prompt"], x["chosen"]], [x["prompt"], x["rejected"]]] for x in train), [])
    rewards = sum(([1, -1] for _ in train), [])
    eval_prompts = [{"prompt": x["prompt"], "original_output": x["chosen"]} for x in islice(test, 280)]
    reward_fn = create_reward_fn(os.environ.get("REWARD_CHECKPOINT"))
    return trlx.train(samples=prompts_outputs, rewards=rewards, config=config, eval_prompts=eval_prompts,
                      metric_fn=lambda **kwargs: {"reward": reward_fn(**kwargs)},
                      stop_sequences=["Human:", "human:", "Assistant:", "assistant:"])


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
