"""Instruction tuning on Alpaca-format records (reference: examples/alpaca/sft_alpaca.py); evaluation asks the tuned model to
rewrite negative reviews positively and scores the sentiment of what it wrote."""
import json
import os
from argparse import ArgumentParser
from typing import Dict, List

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_sft_config


# The (public) Stanford Alpaca prompt format: a fixed preamble, then "### Instruction" / optional "### Input" / "### Response".
_PREAMBLE = {True: "Below is an instruction that describes a task, paired with an input that provides further context. "
                   "Write a response that appropriately completes the request.",
             False: "Below is an instruction that describes a task. Write a response that appropriately completes the request."}


def preprocess(instruction: str, input: str, output: str):
    """``[prompt, output]`` pair for the SFT trainer (the loss covers the output only)."""
    sections = [_PREAMBLE[bool(input)], f"### Instruction:\n{instruction}"]
    if input:
        sections.append(f"### Input:\n{input}")
    sections.append("### Response:\n")
    return ["\n\n".join(sections), output]


def load_alpaca(path_or_name: str):
    if os.path.isfile(path_or_name):  # a local alpaca_data.json
        with open(path_or_name) as fh:
            return json.load(fh)
    tasks = [("Give three synonyms of the input word.", "happy", "joyful, cheerful, content"),
             ("Translate the input to upper case.", "hello world", "HELLO WORLD"),
             ("Name the capital of France.", "", "Paris"),
             ("Rewrite the input into a positive review.", "The movie was dull.", "The movie was delightful.")]
    return [dict(instruction=i, input=x, output=o) for i, x, o in tasks] * 64


def main(hparams={}, model_name="EleutherAI/gpt-j-6B", dataset="tatsu-lab/alpaca"):
    config = default_sft_config().evolve(
        train=dict(total_steps=2400, batch_size=4, seq_length=1024),
        model=dict(model_path=offline_model(model_name, GPT2_TINY)), tokenizer=dict(tokenizer_path=model_name),
        optimizer=dict(kwargs=dict(lr=2e-5)), scheduler=dict(kwargs=dict(eta_min=2e-5)), method=dict(gen_kwargs=dict(max_new_tokens=256)))
    config = TRLConfig.update(config.to_dict(), hparams)
    alpaca = [preprocess(x["instruction"], x["input"], x["output"]) for x in load_alpaca(dataset)]
    sentiment_fn = sentiment_scorer()

    def metric_fn(samples: List[str], prompts: List[str], outputs: List[str], **kwargs) -> Dict[str, List[float]]:
        return {"sentiments": [s["POSITIVE"] for s in sentiment_fn(outputs)]}

    texts, labels = load_imdb()
    bad_reviews = [t for t, l in zip(texts, labels) if l == 0][:256]
    zs_rewrite = [preprocess("Rewrite the input into a positive review.", x[:1024], "")[0] for x in bad_reviews]
    trainer = trlx.train(samples=alpaca, eval_prompts=zs_rewrite, metric_fn=metric_fn, config=config)
    slug = f"{model_name.split('/')[-1]}-{dataset.split('/')[-1]}"
    trainer.save_pretrained(f"{slug}-sft")
    return trainer


if __name__ == "__main__":
    parser = ArgumentParser()
    parser.add_argument("override_hparams", type=str, default="{}", nargs="?")
    parser.add_argument("--model_name", type=str, default="EleutherAI/gpt-j-6B")
    parser.add_argument("--dataset", type=str, default="tatsu-lab/alpaca")
    args = parser.parse_args()
    main(json.loads(args.override_hparams), args.model_name, args.dataset)
