"""SFT on positive reviews with the model-parallel trainer (reference: examples/nemo_sft_sentiments.py)."""
import json
import sys
from typing import Dict, List

import trlx_b200 as trlx
from examples._offline import load_imdb, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_sft_config

default_config = default_sft_config()


def main(hparams={}):
    config = default_config.evolve(
        train=dict(seq_length=1024, batch_size=256, total_steps=1000, eval_interval=100, trainer="NeMoSFTTrainer",
                   trainer_kwargs=dict(pretrained_model=None, megatron_cfg="sft_megatron_20b.yaml")),
        optimizer=dict(name="distributed_fused_adam", kwargs=dict(lr=2e-5, weight_decay=1e-06, eps=1.0e-8, betas=(0.9, 0.95))),
        scheduler=dict(name="CosineAnnealing", kwargs=dict(warmup_steps=0, constant_steps=1e12, min_lr=1e-5)),
    )
    config = TRLConfig.update(config, hparams)
    texts, labels = load_imdb()
    positive = [t for t, l in zip(texts, labels) if l == 1]
    sentiment_fn = sentiment_scorer()

    def metric_fn(samples: List[str], **kwargs) -> Dict[str, List[float]]:
        return {"sentiments": [s["POSITIVE"] for s in sentiment_fn(samples)]}

    return trlx.train(samples=positive, eval_prompts=["I don't know much about Hungarian underground"] * 64, metric_fn=metric_fn,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
