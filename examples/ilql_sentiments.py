"""ILQL on labelled reviews (reference: examples/ilql_sentiments.py)."""
import json
import sys
from typing import Dict, List

import trlx_b200 as trlx
from examples._offline import GPT2_SMALL, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_ilql_config


def main(hparams={}):
    config = TRLConfig.update(default_ilql_config().to_dict(), hparams)
    if isinstance(config.model.model_path, str):
        config.model.model_path = offline_model(config.model.model_path, GPT2_SMALL)
    sentiment_fn = sentiment_scorer()

    def metric_fn(samples: List[str], **kwargs) -> Dict[str, List[float]]:
        return {"sentiments": [s["POSITIVE"] for s in sentiment_fn(samples)]}

    texts, labels = load_imdb()
    return trlx.train(samples=texts, rewards=labels, eval_prompts=["I don't know much about Hungarian underground"] * 256,
                      metric_fn=metric_fn, config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
