"""ILQL with an encoder-decoder model (reference: examples/ilql_sentiments_t5.py): (prefix → continuation, label) triples."""
import json
import sys
from typing import Dict, List

import trlx_b200 as trlx
from examples._offline import T5_TINY, load_imdb, offline_model, sentiment_scorer
from examples.ppo_sentiments_t5 import review_prefixes
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ilql import ILQLConfig

default_config = TRLConfig(
    train=TrainConfig(seq_length=128, epochs=100, total_steps=1000, batch_size=32, checkpoint_interval=1000, eval_interval=100,
                      pipeline="PromptPipeline", trainer="AccelerateILQLTrainer", save_best=False),
    model=ModelConfig(model_path=offline_model("lvwerra/t5-imdb", T5_TINY), num_layers_unfrozen=-1, model_arch_type="seq2seq"),
    tokenizer=TokenizerConfig(tokenizer_path="lvwerra/t5-imdb", padding_side="right", truncation_side="right"),
    optimizer=OptimizerConfig(name="adamw", kwargs={"lr": 5.0e-5, "betas": [0.9, 0.999], "eps": 1.0e-8, "weight_decay": 1.0e-6}),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs={"T_max": 100000, "eta_min": 5.0e-5}),
    method=ILQLConfig(name="ILQLConfig", tau=0.7, gamma=0.99, cql_scale=0.1, awac_scale=1, alpha=0.001, beta=0,
                      steps_for_target_q_sync=5, two_qs=True, gen_kwargs=dict(max_new_tokens=56, top_k=20, beta=4, temperature=1.0)),
)


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    sentiment_fn = sentiment_scorer()

    def metric_fn(samples: List[str], **kwargs) -> Dict[str, List[float]]:
        return dict(sentiments=[s["POSITIVE"] for s in sentiment_fn(samples)])

    texts, labels = load_imdb()
    prefixes = review_prefixes(texts)
    samples = [[p, t[len(p):].strip()] for p, t in zip(prefixes, texts)]  # (prompt, continuation)
    return trlx.train(samples=samples[:-64], rewards=labels[: len(samples) - 64], eval_prompts=prefixes[-64:], metric_fn=metric_fn,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
