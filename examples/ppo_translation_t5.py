"""PPO for translation with an encoder-decoder policy (reference: examples/ppo_translation_t5.py).

The reference rewards COMET and reports BLEU/chrF; offline, unigram-F1 against the reference translation stands in for all
three.  Experience is collected with beam search (`gen_experience_kwargs.num_beams = 4`), evaluation samples greedily."""
import json
import sys
from typing import Dict, List

import trlx_b200 as trlx
from examples._offline import T5_TINY, offline_model, overlap_f1, synthetic_translation
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ppo import PPOConfig
from trlx_b200.utils.tokenizer import load_tokenizer

default_config = TRLConfig(
    train=TrainConfig(seq_length=612, epochs=100, total_steps=100000, batch_size=12, checkpoint_interval=10000, eval_interval=200,
                      pipeline="PromptPipeline", trainer="AcceleratePPOTrainer", tracker="wandb"),
    model=ModelConfig(model_path=offline_model("t5-large", T5_TINY), model_arch_type="seq2seq", num_layers_unfrozen=-1),
    tokenizer=TokenizerConfig(tokenizer_path="t5-large", padding_side="right", truncation_side="right"),
    optimizer=OptimizerConfig(name="adamw", kwargs={"lr": 2.0e-6, "betas": [0.9, 0.999], "eps": 1.0e-8, "weight_decay": 1.0e-6}),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs={"T_max": 10000, "eta_min": 1.0e-6}),
    method=PPOConfig(name="PPOConfig", num_rollouts=256, chunk_size=12, ppo_epochs=4, init_kl_coef=0.05, target=6, horizon=10000,
                     gamma=0.99, lam=0.95, cliprange=0.2, cliprange_value=0.2, vf_coef=1.0, scale_reward=None, ref_mean=None,
                     ref_std=None, cliprange_reward=10, gen_kwargs={"max_new_tokens": 100},
                     gen_experience_kwargs={"max_new_tokens": 100, "do_sample": False, "num_beams": 4, "temperature": 1.0}),
)
PREFIX = "translate English to German: "


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    pairs = synthetic_translation(2048)
    train, valid = pairs[:-128], pairs[-128:]
    tokenizer = load_tokenizer(config.tokenizer.tokenizer_path)
    tokenizer.truncation_side = "right"
    max_length = config.train.seq_length - config.method.gen_kwargs["max_new_tokens"]
    translation_map = {}
    for p in pairs:  # key = the prompt exactly as the trainer will decode it after truncation
        ids = tokenizer(PREFIX + p["en"], truncation=True, max_length=max_length, add_special_tokens=False)["input_ids"]
        translation_map[tokenizer.decode(ids, skip_special_tokens=True).strip()] = {"src": p["en"], "tgt": p["de"]}

    def reward_fn(samples: List[str], prompts: List[str], outputs: List[str], **kwargs) -> List[float]:
        refs = [translation_map.get(prompt.strip(), {"tgt": ""})["tgt"] for prompt in prompts]
        return [overlap_f1(o.strip(), r) for o, r in zip(outputs, refs)]

    def metric_fn(samples: List[str], prompts: List[str], outputs: List[str], **kwargs) -> Dict[str, float]:
        scores = reward_fn(samples, prompts, outputs)
        exact = [float(o.strip() == translation_map.get(p.strip(), {"tgt": None})["tgt"]) for o, p in zip(outputs, prompts)]
        return {"overlap_f1": sum(scores) / max(len(scores), 1), "exact_match": sum(exact) / max(len(exact), 1)}

    return trlx.train(reward_fn=reward_fn, metric_fn=metric_fn, prompts=[PREFIX + p["en"] for p in train],
                      eval_prompts=[PREFIX + p["en"] for p in valid], config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
