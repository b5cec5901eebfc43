"""Default configurations (parity: ``trlx/data/default_configs.py:17-148``).  Same hyper-parameters as the
reference; model / tokenizer names resolve offline (random-init presets, synthetic tokenizers) when no local
checkpoint exists.  The Megatron-style defaults return plain :class:`TRLConfig` objects with a tensor-parallel
layout instead of OmegaConf trees."""
from __future__ import annotations

from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ilql import ILQLConfig
from trlx_b200.models.modeling_ppo import PPOConfig
from trlx_b200.trainer.accelerate_sft_trainer import SFTConfig


# Every default shares the AdamW / cosine recipe and differs in a handful of numbers; the tables below are those numbers
# (values: ``trlx/data/default_configs.py:17-121``).
_ADAMW = dict(betas=(0.9, 0.95), eps=1.0e-8, weight_decay=1.0e-6)
_SAMPLING = dict(max_new_tokens=40, top_k=0, top_p=1.0, do_sample=True)


def _base(trainer: str, lr: float, model: str, unfrozen: int, method, **train) -> TRLConfig:
    train.setdefault("epochs", 100)
    train.setdefault("eval_interval", 100)
    return TRLConfig(
        method=method,
        model=ModelConfig(model_path=model, num_layers_unfrozen=unfrozen),
        tokenizer=TokenizerConfig(tokenizer_path="gpt2", truncation_side="right"),
        optimizer=OptimizerConfig(name="adamw", kwargs=dict(lr=lr, **_ADAMW)),
        scheduler=SchedulerConfig(name="cosine_annealing", kwargs=dict(T_max=1e12, eta_min=lr)),  # constant lr in practice
        train=TrainConfig(pipeline="PromptPipeline", trainer=trainer, **train),
    )


def default_ppo_config() -> TRLConfig:
    """GPT-2 (IMDB-tuned) with two unfrozen blocks, 128 rollouts of 40 tokens, 4 PPO epochs (the benchmark configuration)."""
    ppo = PPOConfig(name="PPOConfig", ppo_epochs=4, num_rollouts=128, chunk_size=128, gamma=1, lam=0.95, init_kl_coef=0.001,
                    target=None, horizon=10000, cliprange=0.2, cliprange_value=0.2, cliprange_reward=10, vf_coef=1,
                    scale_reward="ignored", ref_mean=None, ref_std=None, gen_kwargs=dict(_SAMPLING))
    return _base("AcceleratePPOTrainer", 3e-5, "lvwerra/gpt2-imdb", 2, ppo, seq_length=1024, batch_size=32, total_steps=10000,
                 checkpoint_interval=10000)


def default_ilql_config() -> TRLConfig:
    """Offline ILQL on GPT-2: two Q heads, expectile 0.7, top-20 advantage-shifted sampling."""
    ilql = ILQLConfig(name="ilqlconfig", two_qs=True, tau=0.7, gamma=0.99, alpha=0.001, beta=0, cql_scale=0.1, awac_scale=1,
                      steps_for_target_q_sync=5, gen_kwargs=dict(max_new_tokens=56, top_k=20, beta=1, temperature=1.0))
    return _base("AccelerateILQLTrainer", 5.0e-5, "gpt2", -1, ilql, seq_length=64, batch_size=128, total_steps=1000,
                 checkpoint_interval=1000)


def default_sft_config() -> TRLConfig:
    """Supervised fine-tuning of GPT-2, every layer trained."""
    return _base("AccelerateSFTTrainer", 1.0e-4, "gpt2", -1, SFTConfig(name="sftconfig", gen_kwargs=dict(_SAMPLING)),
                 seq_length=1024, batch_size=8, total_steps=1000, checkpoint_interval=10000)


def _megatron(model: str, tp: int, pp: int = 1, sp: bool = True) -> TRLConfig:
    cfg = default_ppo_config()
    return cfg.evolve(model=dict(model_path=model),
                      train=dict(trainer="NeMoPPOTrainer",
                                 parallel=dict(tensor_parallel=tp, pipeline_parallel=pp, sequence_parallel=sp)))


def default_nemo_20b_config() -> TRLConfig:
    """GPT-NeoX-20B-shaped model, TP=4, sequence parallel (``configs/nemo_configs/megatron_20b.yaml:51-62,82``)."""
    return _megatron("gpt-neox-20b", tp=4)


def default_nemo_2b_config() -> TRLConfig:
    return _megatron(dict(model_type="gpt_neox", vocab_size=50432, hidden_size=2048, num_hidden_layers=24,
                          num_attention_heads=16, intermediate_size=8192, max_position_embeddings=2048), tp=1)


def default_nemo_1_3b_config() -> TRLConfig:
    return _megatron(dict(model_type="gpt_neox", vocab_size=50432, hidden_size=2048, num_hidden_layers=24,
                          num_attention_heads=16, intermediate_size=8192, max_position_embeddings=2048), tp=1)
