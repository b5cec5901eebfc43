"""Supervised fine-tuning on the chosen HH responses (reference: examples/hh/sft_hh.py)."""
import json
import os
import sys

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, offline_model, synthetic_dialogues
from examples.hh.ppo_hh import apply_preset
from examples.hh.reward import create_reward_fn
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.trainer.accelerate_sft_trainer import SFTConfig

default_config = TRLConfig(
    train=TrainConfig(seq_length=1024, epochs=100, total_steps=10000, batch_size=4, checkpoint_interval=10000, eval_interval=1000,
                      pipeline="PromptPipeline", trainer="AccelerateSFTTrainer", checkpoint_dir="checkpoints/sft_hh"),
    model=ModelConfig(model_path="EleutherAI/gpt-j-6B", num_layers_unfrozen=-1),
    tokenizer=TokenizerConfig(tokenizer_path="EleutherAI/gpt-j-6B", truncation_side="left"),
    optimizer=OptimizerConfig(name="adamw", kwargs=dict(lr=1e-6, betas=(0.9, 0.95), eps=1.0e-8, weight_decay=1.0e-6)),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs=dict(T_max=100000000, eta_min=1e-6)),
    method=SFTConfig(name="sftconfig", gen_kwargs=dict(max_new_tokens=128, top_k=20, top_p=1.0, do_sample=True)),
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
