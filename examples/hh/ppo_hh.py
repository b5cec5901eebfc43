"""PPO on Helpful/Harmless dialogues with a separate reward model (reference: examples/hh/ppo_hh.py).

`CONFIG_NAME=125M|1B|6B|20B` picks the reference's size presets; the reward model runs in-process on the last GPU, or behind
`REWARD_HOST=host:port` (see `examples/hh/reward.py`)."""
import json
import os
import sys
from itertools import islice

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, offline_model, synthetic_dialogues
from examples.hh.reward import create_reward_fn
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ppo import PPOConfig

default_config = TRLConfig(
    train=TrainConfig(seq_length=1024, epochs=10000, total_steps=10000, batch_size=4, checkpoint_interval=10000, eval_interval=500,
                      pipeline="PromptPipeline", trainer="AcceleratePPOTrainer", checkpoint_dir="checkpoints/ppo_hh"),
    model=ModelConfig(model_path="EleutherAI/gpt-j-6B", num_layers_unfrozen=2),
    tokenizer=TokenizerConfig(tokenizer_path="EleutherAI/gpt-j-6B", truncation_side="left"),
    optimizer=OptimizerConfig(name="adamw", kwargs=dict(lr=8e-6, betas=(0.9, 0.95), eps=1.0e-8, weight_decay=1.0e-6)),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs=dict(T_max=10000, eta_min=8e-6)),
    method=PPOConfig(name="PPOConfig", num_rollouts=64, chunk_size=16, ppo_epochs=4, init_kl_coef=0.05, target=6, horizon=10000,
                     gamma=1, lam=0.95, cliprange=0.2, cliprange_value=0.2, vf_coef=1, scale_reward="running", ref_mean=None,
                     ref_std=None, cliprange_reward=10, gen_kwargs=dict(max_new_tokens=128, top_k=0, top_p=1.0, do_sample=True)),
)

PRESETS = {
    "125M": dict(batch_size=32, total_steps=1500, model="Dahoas/pythia-125M-static-sft", num_rollouts=128),
    "1B": dict(batch_size=8, total_steps=2500, lr=6e-6, model="Dahoas/pythia-1B-static-sft", chunk_size=16),
    "6B": dict(batch_size=4, seq_length=512, total_steps=6000, model="Dahoas/pythia-6B-static-sft", chunk_size=16),
    "20B": dict(batch_size=1, seq_length=512, total_steps=8000, lr=1e-6, model="EleutherAI/gpt-neox-20b", num_rollouts=16,
                chunk_size=4, ppo_epochs=2),
}


def apply_preset(config: TRLConfig, name, prefix: str) -> None:
    p = PRESETS.get(name or "")
    if not p:
        return
    config.train.checkpoint_dir = f"checkpoints/{prefix}_{name}"
    config.model.model_path = p["model"]
    config.tokenizer.tokenizer_path = "EleutherAI/gpt-neox-20b"
    for k in ("batch_size", "seq_length", "total_steps"):
        if k in p:
            setattr(config.train, k, p[k])
    if "lr" in p:
        config.optimizer.kwargs["lr"] = config.scheduler.kwargs["eta_min"] = p["lr"]
    for k in ("num_rollouts", "chunk_size", "ppo_epochs"):
        if k in p and hasattr(config.method, k):
            setattr(config.method, k, p[k])


apply_preset(default_config, os.environ.get("CONFIG_NAME"), "ppo_hh")


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    if isinstance(config.model.model_path, str):
        config.model.model_path = offline_model(config.model.model_path, GPT2_TINY)
    data = synthetic_dialogues(1024)
    prompts = [{"prompt": x["prompt"], "original_output": x["chosen"]} for x in data[:-64]]
    eval_prompts = [{"prompt": x["prompt"], "original_output": x["chosen"]} for x in islice(data[-64:], 280)]
    return trlx.train(prompts=prompts, eval_prompts=eval_prompts, reward_fn=create_reward_fn(os.environ.get("REWARD_CHECKPOINT")),
                      config=config, stop_sequences=["Human:", "human:", "Assistant:", "assistant:"])


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
