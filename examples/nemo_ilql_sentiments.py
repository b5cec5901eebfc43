"""ILQL sentiments on the model-parallel trainer (reference: examples/nemo_ilql_sentiments.py; 20B recipe, TP=4 + SP)."""
import json
import sys
from typing import Dict, List

import trlx_b200 as trlx
from examples._offline import load_imdb, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_ilql_config

default_config = default_ilql_config()


def main(hparams={}):
    config = default_config.evolve(
        train=dict(seq_length=1024, batch_size=512, total_steps=200, trainer="NeMoILQLTrainer",
                   trainer_kwargs=dict(pretrained_model=None, megatron_cfg="megatron_20b.yaml")),
        method=dict(gen_kwargs=dict(beta=2.0, temperature=0.9)),
    )
    config = TRLConfig.update(config, hparams)
    sentiment_fn = sentiment_scorer()

    def metric_fn(samples: List[str], **kwargs) -> Dict[str, List[float]]:
        return {"sentiments": [s["POSITIVE"] for s in sentiment_fn(samples)]}

    texts, labels = load_imdb()
    return trlx.train(samples=texts, rewards=labels, eval_prompts=["I don't know much about Hungarian underground"] * 128,
                      metric_fn=metric_fn, config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
