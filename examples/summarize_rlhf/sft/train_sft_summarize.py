"""Stage 1 of the summarisation recipe: supervised fine-tuning on (post, reference summary) pairs
(reference: examples/summarize_rlhf/sft/train_gptj_summarize.py — HF Trainer there; here the framework's own SFT trainer)."""
import json
import sys

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, offline_model, overlap_f1, synthetic_summaries
from trlx_b200.data.default_configs import TRLConfig, default_sft_config


def main(hparams={}):
    config = default_sft_config().evolve(
        train=dict(seq_length=550, batch_size=16, total_steps=5000, eval_interval=500, checkpoint_dir="gptj-supervised-summarize-checkpoint"),
        model=dict(model_path=offline_model("EleutherAI/gpt-j-6B", GPT2_TINY)), tokenizer=dict(tokenizer_path="EleutherAI/gpt-j-6B"),
        optimizer=dict(kwargs=dict(lr=1e-5)), scheduler=dict(kwargs=dict(eta_min=1e-5)), method=dict(gen_kwargs=dict(max_new_tokens=50)))
    config = TRLConfig.update(config.to_dict(), hparams)
    data = synthetic_summaries(2048)
    refs = {d["prompt"].strip(): d["label"] for d in data}

    def metric_fn(samples, prompts, outputs, **kw):
        return {"overlap_f1": [overlap_f1(o, refs.get(p.strip(), "")) for p, o in zip(prompts, outputs)]}

    return trlx.train(samples=[[d["prompt"], d["label"]] for d in data[:-128]], eval_prompts=[d["prompt"] for d in data[-128:]],
                      metric_fn=metric_fn, config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
