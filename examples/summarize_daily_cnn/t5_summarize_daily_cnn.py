"""PPO summarisation with FLAN-T5, reward = overlap with the reference summary (reference:
examples/summarize_daily_cnn/t5_summarize_daily_cnn.py uses METEOR on CNN/DailyMail).  Prompts carry their reference summary as
metadata (`original_summaries`), which the trainer forwards to `reward_fn`."""
import json
import sys
from typing import List

import trlx_b200 as trlx
from examples._offline import T5_TINY, offline_model, overlap_f1, synthetic_summaries
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ppo import PPOConfig

config = TRLConfig(
    train=TrainConfig(seq_length=612, epochs=100, total_steps=100000, batch_size=12, checkpoint_interval=10000, eval_interval=500,
                      pipeline="PromptPipeline", trainer="AcceleratePPOTrainer"),
    model=ModelConfig(model_path=offline_model("google/flan-t5-large", T5_TINY), model_arch_type="seq2seq", num_layers_unfrozen=2),
    tokenizer=TokenizerConfig(tokenizer_path="google/flan-t5-large", truncation_side="right"),
    optimizer=OptimizerConfig(name="adamw", kwargs={"lr": 1.0e-5, "betas": [0.9, 0.999], "eps": 1.0e-8, "weight_decay": 1.0e-6}),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs={"T_max": 10000, "eta_min": 1.0e-6}),
    method=PPOConfig(name="PPOConfig", num_rollouts=512, chunk_size=12, ppo_epochs=4, init_kl_coef=0.05, target=6, horizon=10000,
                     gamma=0.99, lam=0.95, cliprange=0.2, cliprange_value=0.2, vf_coef=1.0, scale_reward=None, ref_mean=None,
                     ref_std=None, cliprange_reward=10, gen_kwargs={"max_new_tokens": 100},
                     gen_experience_kwargs={"max_new_tokens": 100, "do_sample": True, "temperature": 1.0, "top_k": 50, "top_p": 0.95}),
)


def reward_fn(samples: List[str], prompts: List[str], outputs: List[str], original_summaries: List[str], **kwargs):
    return [overlap_f1(output.strip(), ref) for ref, output in zip(original_summaries, outputs)]


def main(hparams={}):
    cfg = TRLConfig.update(config, hparams)
    data = synthetic_summaries(2048)
    to_prompt = lambda d: {"prompt": "Summarize: " + d["prompt"].replace("\nTL;DR:", ""), "original_summaries": d["label"].strip()}  # noqa: E731
    return trlx.train(reward_fn=reward_fn, prompts=[to_prompt(d) for d in data[:-128]], eval_prompts=[to_prompt(d) for d in data[-128:]],
                      config=cfg)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
