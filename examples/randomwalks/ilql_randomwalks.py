"""ILQL on random walks with an evaluation sweep over ``beta`` (reference: examples/randomwalks/ilql_randomwalks.py)."""
import json
import sys

import trlx_b200 as trlx
from examples.randomwalks import generate_random_walks
from examples.randomwalks.randomwalks import MODEL, TOKENIZER
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_ilql_config

default_config = default_ilql_config().evolve(
    train=dict(seq_length=11, batch_size=100, total_steps=1000, epochs=100, tracker=None, eval_interval=100),
    model=dict(model_path=MODEL), tokenizer=dict(tokenizer_path=TOKENIZER),
    optimizer=dict(kwargs=dict(lr=2e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-6)),
    method=dict(tau=0.8, gamma=0.99, cql_scale=0.1, awac_scale=1, alpha=0.1, beta=0, steps_for_target_q_sync=5, two_qs=True,
                gen_kwargs=dict(max_new_tokens=9, top_k=10, beta=[0, 1, 100], temperature=1.0)),
)


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    metric_fn, eval_prompts, walks, logit_mask = generate_random_walks(seed=config.train.seed)
    rewards = metric_fn(walks)["optimality"]
    # split each walk into (prompt = start node, output = rest of the path)
    samples = [[w[:1], w[1:]] for w in walks]
    # True = forbidden transition (no edge between the last node and the candidate next node)
    config.train.trainer_kwargs = dict(config.train.trainer_kwargs, logit_mask=~logit_mask)
    return trlx.train(samples=samples, rewards=rewards, eval_prompts=eval_prompts,
                      metric_fn=lambda samples, **kw: metric_fn(samples), config=config, stop_sequences=["|"])


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
