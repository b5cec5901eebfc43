"""Supervised fine-tuning on positive reviews (reference: examples/sft_sentiments.py)."""
import json
import sys
from typing import Dict, List

import trlx_b200 as trlx
from examples._offline import GPT2_SMALL, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_sft_config


def main(hparams={}):
    config = TRLConfig.update(default_sft_config().to_dict(), hparams)
    if isinstance(config.model.model_path, str):
        config.model.model_path = offline_model(config.model.model_path, GPT2_SMALL)
    texts, labels = load_imdb()
    positive = [t for t, l in zip(texts, labels) if l == 1]
    sentiment_fn = sentiment_scorer()

    def metric_fn(samples: List[str], **kwargs) -> Dict[str, List[float]]:
        return {"sentiments": [s["POSITIVE"] for s in sentiment_fn(samples)]}

    trainer = trlx.train(samples=positive, eval_prompts=["I don't know much about Hungarian underground"] * 64,
                         metric_fn=metric_fn, config=config)
    trainer.save_pretrained("reviews-sft")
    return trainer


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
