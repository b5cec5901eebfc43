"""Rejection fine-tuning on random walks (reference: examples/randomwalks/rft_randomwalks.py)."""
import json
import sys

import trlx_b200 as trlx
from examples.randomwalks import online_task
from examples.randomwalks.randomwalks import MODEL, TOKENIZER
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_sft_config
from trlx_b200.trainer.accelerate_rft_trainer import RFTConfig

default_config = default_sft_config().evolve(
    train=dict(seq_length=10, batch_size=100, total_steps=200, epochs=100, tracker=None, eval_interval=20, trainer="AccelerateRFTTrainer"),
    model=dict(model_path=MODEL), tokenizer=dict(tokenizer_path=TOKENIZER),
    optimizer=dict(kwargs=dict(lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-6)),
)
default_config.method = RFTConfig(name="RFTConfig", n_generations_per_prompt=100, start_percentile=0.9, end_percentile=0.95,
                                  n_improve_steps=1, gen_kwargs=dict(max_new_tokens=9, top_k=0, top_p=1.0, temperature=1.0, do_sample=True))


def main(hparams={}):
    config = TRLConfig.update(default_config, hparams)
    return trlx.train(config=config, **online_task(config.train.seed))


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
