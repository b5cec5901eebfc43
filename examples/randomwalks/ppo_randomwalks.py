"""PPO on random walks (BASELINE.json config 1: CPU plumbing run; reference: examples/randomwalks/ppo_randomwalks.py)."""
import json
import sys

import trlx_b200 as trlx
from examples.randomwalks import online_task
from examples.randomwalks.randomwalks import MODEL, TOKENIZER
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_ppo_config

# the library's PPO defaults with the handful of values this 10-token toy task changes
default_config = default_ppo_config().evolve(
    train=dict(seq_length=10, epochs=20, batch_size=100, eval_interval=20, tracker=None),
    model=dict(model_path=MODEL, num_layers_unfrozen=-1),
    tokenizer=dict(tokenizer_path=TOKENIZER),
    optimizer=dict(kwargs=dict(lr=3.0e-4)),
    scheduler=dict(kwargs=dict(T_max=10000, eta_min=3.0e-4)),
    method=dict(init_kl_coef=0, vf_coef=1.2, cliprange_reward=1, gen_kwargs=dict(max_new_tokens=9)),
)


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    return trlx.train(config=config, **online_task(config.train.seed))


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
