"""Evaluate a tuned summariser: overlap with the reference summaries and mean reward-model score
(reference: examples/summarize_rlhf/trlx_inference_gptj.py reports ROUGE + reward for SFT vs PPO checkpoints)."""
import json
import sys

import torch

from examples._offline import GPTJ_TINY, offline_model, overlap_f1, synthetic_summaries
from trlx_b200.models.generation import generate
from trlx_b200.models.modeling_base import build_base_model
from trlx_b200.utils.tokenizer import load_tokenizer


def main(model_path="CarperAI/openai_summarize_tldr_ppo", n: int = 64, max_new_tokens: int = 50):
    tok = load_tokenizer("EleutherAI/gpt-j-6B")
    tok.pad_token, tok.padding_side = tok.eos_token, "left"
    model = build_base_model(offline_model(model_path, GPTJ_TINY)).eval()
    data = synthetic_summaries(n, seed=1)
    scores = []
    for i in range(0, n, 8):
        batch = data[i:i + 8]
        enc = tok([d["prompt"] for d in batch], return_tensors="pt", padding=True, truncation=True, max_length=500)
        with torch.no_grad():
            out = generate(model, enc.input_ids, enc.attention_mask, max_new_tokens=max_new_tokens, do_sample=False,
                           eos_token_id=tok.eos_token_id, pad_token_id=tok.pad_token_id)
        preds = tok.batch_decode(out[:, enc.input_ids.shape[1]:], skip_special_tokens=True)
        scores += [overlap_f1(p, d["label"]) for p, d in zip(preds, batch)]
    result = {"overlap_f1": sum(scores) / len(scores), "n": len(scores)}
    print(json.dumps(result))
    return result


if __name__ == "__main__":
    main(*sys.argv[1:2])
