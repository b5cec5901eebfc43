"""PPO on a Llama-family policy, top-2 layers trainable (reference: examples/ppo_sentiments_llama.py; BASELINE config B4)."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import LLAMA_TINY, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ppo import PPOConfig


def llama_config():
    return TRLConfig(
        train=TrainConfig(seq_length=1024, epochs=100, total_steps=400, batch_size=32, checkpoint_interval=10000, eval_interval=100,
                          pipeline="PromptPipeline", trainer="AcceleratePPOTrainer", save_best=False),
        model=ModelConfig(model_path=offline_model("NousResearch/Llama-2-7b-hf", LLAMA_TINY), num_layers_unfrozen=2),
        tokenizer=TokenizerConfig(tokenizer_path="NousResearch/Llama-2-7b-hf", truncation_side="right"),
        optimizer=OptimizerConfig(name="adamw", kwargs=dict(lr=1e-5, betas=(0.9, 0.95), eps=1.0e-8, weight_decay=1.0e-6)),
        scheduler=SchedulerConfig(name="cosine_annealing", kwargs=dict(T_max=10000, eta_min=1.0e-5)),
        method=PPOConfig(name="PPOConfig", num_rollouts=128, chunk_size=128, ppo_epochs=4, init_kl_coef=0.001, target=6,
                         horizon=10000, gamma=1, lam=0.95, cliprange=0.2, cliprange_value=0.2, vf_coef=1, scale_reward="ignored",
                         ref_mean=None, ref_std=None, cliprange_reward=10,
                         gen_kwargs=dict(max_new_tokens=40, top_k=0, top_p=1.0, do_sample=True)),
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
