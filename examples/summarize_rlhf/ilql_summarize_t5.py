"""ILQL on TL;DR preference comparisons with FLAN-T5: chosen = +1, rejected = −1
(reference: examples/summarize_rlhf/ilql_summarize_t5.py)."""
import json
import sys

import trlx_b200 as trlx
from examples._offline import T5_TINY, offline_model, overlap_f1
from examples.summarize_rlhf.reward_model.train_reward_model import make_pairs
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ilql import ILQLConfig

default_config = TRLConfig(
    train=TrainConfig(seq_length=550, batch_size=8, epochs=100, total_steps=5000, checkpoint_interval=10000, eval_interval=1000,
                      pipeline="PromptPipeline", trainer="AccelerateILQLTrainer", checkpoint_dir="ilql_summarize_t5"),
    model=ModelConfig(model_path=offline_model("pvduy/flant5-xl_openai_tldr_sft", T5_TINY), num_layers_unfrozen=-1,
                      model_arch_type="seq2seq"),
    tokenizer=TokenizerConfig(tokenizer_path="pvduy/flant5-xl_openai_tldr_sft", truncation_side="left"),
    optimizer=OptimizerConfig(name="adamw", kwargs=dict(lr=1e-6, betas=(0.9, 0.95), eps=1.0e-8, weight_decay=1.0e-6)),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs=dict(T_max=5000, eta_min=1e-6)),
    method=ILQLConfig(name="ilqlconfig", tau=0.6, gamma=0.99, cql_scale=0.1, awac_scale=1, alpha=0.0001, beta=0,
                      steps_for_target_q_sync=1, two_qs=True, gen_kwargs=dict(max_new_tokens=50, top_k=50, beta=[1, 2, 3], temperature=1.0)),
)


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    pairs = make_pairs(2048)
    train, test = pairs[:-128], pairs[-128:]
    samples = sum(([[p["prompt"], p["chosen"]], [p["prompt"], p["rejected"]]] for p in train), [])
    rewards = sum(([1, -1] for _ in train), [])
    refs = {p["prompt"].strip(): p["chosen"] for p in pairs}

    def metric_fn(samples, prompts, outputs, **kw):
        return {"overlap_f1": [overlap_f1(o, refs.get(p.strip(), "")) for p, o in zip(prompts, outputs)]}

    return trlx.train(samples=samples, rewards=rewards, eval_prompts=[p["prompt"] for p in test], metric_fn=metric_fn, config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
