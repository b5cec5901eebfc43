"""PPO with an encoder-decoder policy (reference: examples/ppo_sentiments_t5.py): continue a review's first words."""
import json
import sys
from typing import Dict, List

import numpy as np

import trlx_b200 as trlx
from examples._offline import T5_TINY, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ppo import PPOConfig

default_config = TRLConfig(
    train=TrainConfig(seq_length=128, epochs=100, total_steps=100000, batch_size=12, checkpoint_interval=10000, eval_interval=100,
                      pipeline="PromptPipeline", trainer="AcceleratePPOTrainer", save_best=False),
    model=ModelConfig(model_path=offline_model("lvwerra/t5-imdb", T5_TINY), num_layers_unfrozen=-1, model_arch_type="seq2seq"),
    tokenizer=TokenizerConfig(tokenizer_path="lvwerra/t5-imdb", padding_side="right", truncation_side="right"),
    optimizer=OptimizerConfig(name="adamw", kwargs={"lr": 5.0e-5, "betas": [0.9, 0.999], "eps": 1.0e-8, "weight_decay": 1.0e-6}),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs={"T_max": 100000, "eta_min": 5.0e-5}),
    method=PPOConfig(name="PPOConfig", num_rollouts=128, chunk_size=12, ppo_epochs=4, init_kl_coef=0.05, target=6, horizon=10000,
                     gamma=0.99, lam=0.95, cliprange=0.2, cliprange_value=0.2, vf_coef=1, scale_reward=None, ref_mean=None,
                     ref_std=None, cliprange_reward=10,
                     gen_kwargs={"max_new_tokens": 50, "do_sample": True, "top_k": 0, "top_p": 1, "eos_token_id": -1}),
)


def review_prefixes(texts: List[str], lo: int = 2, hi: int = 8, seed: int = 2023) -> List[str]:
    """First `k ~ U[lo, hi)` words of every review longer than 200 characters (the reference samples token counts)."""
    rng = np.random.default_rng(seed)
    return [" ".join(t.split()[: int(rng.integers(lo, hi))]) for t in texts if len(t) > 200 or len(texts) < 4096]


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    sentiment_fn = sentiment_scorer()

    def reward_fn(samples: List[str], **kwargs) -> List[float]:
        return [s["POSITIVE"] for s in sentiment_fn(samples)]

    def metric_fn(samples: List[str], **kwargs) -> Dict[str, List[float]]:
        return {"sentiments": reward_fn(samples)}

    texts, _ = load_imdb()
    prompts = review_prefixes(texts)
    return trlx.train(reward_fn=reward_fn, metric_fn=metric_fn, prompts=prompts[:-64], eval_prompts=prompts[-64:], config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
