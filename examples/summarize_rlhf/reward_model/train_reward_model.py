"""Train the pairwise reward model on (post, chosen summary, rejected summary) triples
(reference: examples/summarize_rlhf/reward_model/train_reward_model_gptj.py — HF Trainer + DeepSpeed there; here a plain loop
on the framework's fused, sharded AdamW, one process per GPU under torch.distributed.run)."""
import json
import os
import sys

import torch

from examples._offline import GPTJ_TINY, offline_model, synthetic_summaries
from examples.summarize_rlhf.reward_model.reward_model import GPTRewardModel
from trlx_b200.parallel.optim import FusedAdamW
from trlx_b200.parallel.runtime import Runtime
from trlx_b200.utils.tokenizer import load_tokenizer


def make_pairs(n: int = 2048, seed: int = 0):
    """chosen = the reference summary, rejected = a shuffled / truncated corruption of it."""
    import random

    rng = random.Random(seed)
    out = []
    for d in synthetic_summaries(n, seed):
        words = d["label"].split()
        bad = words[: max(len(words) // 3, 1)] if rng.random() < 0.5 else rng.sample(words, len(words))
        out.append(dict(prompt=d["prompt"], chosen=d["label"], rejected=" " + " ".join(bad)))
    return out


def encode(tokenizer, pairs, max_length: int):
    def enc(texts):
        return tokenizer([t + tokenizer.eos_token for t in texts], truncation=True, max_length=max_length, padding="max_length",
                         return_tensors="pt")
    c = enc([p["prompt"] + p["chosen"] for p in pairs])
    r = enc([p["prompt"] + p["rejected"] for p in pairs])
    keep = (c.input_ids != r.input_ids).any(1)
    return c.input_ids[keep], c.attention_mask[keep], r.input_ids[keep], r.attention_mask[keep]


def main(hparams={}):
    cfg = dict(model="EleutherAI/gpt-j-6B", tokenizer="EleutherAI/gpt-j-6B", lr=1e-5, epochs=1, batch_size=4, max_length=550,
               out="rm_checkpoint", steps=None, layers_frozen=0.7)
    cfg.update(hparams)
    rt = Runtime()
    tok = load_tokenizer(cfg["tokenizer"])
    tok.pad_token = tok.eos_token
    tok.padding_side = "right"
    model = GPTRewardModel(offline_model(cfg["model"], GPTJ_TINY), tok.pad_token_id).to(rt.device)
    if rt.cuda:
        model = model.to(torch.bfloat16)
    blocks = list(model.transformer.transformer.h)
    for blk in blocks[: int(cfg["layers_frozen"] * len(blocks))]:  # the reference freezes the bottom 70 % of the blocks
        blk.requires_grad_(False)
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=cfg["lr"], process_group=rt.dp_group)
    pairs = make_pairs()
    ci, cm, ri, rm = encode(tok, pairs[:-128], cfg["max_length"])
    vi = encode(tok, pairs[-128:], cfg["max_length"])
    step, bs = 0, cfg["batch_size"]
    for _ in range(cfg["epochs"]):
        perm = torch.randperm(len(ci))[rt.dp_rank::rt.dp_size]
        for i in range(0, len(perm) - bs + 1, bs):
            idx = perm[i:i + bs]
            ids = torch.cat([ci[idx], ri[idx]]).to(rt.device)
            mask = torch.cat([cm[idx], rm[idx]]).to(rt.device)
            out = model(ids, mask)
            out["loss"].backward()
            opt.step()
            opt.zero_grad()
            step += 1
            if step % 10 == 0 and rt.is_main_process:
                acc = (out["chosen_end_scores"] > out["rejected_end_scores"]).float().mean().item()
                print(f"step {step} loss {out['loss'].item():.4f} acc {acc:.2f}", flush=True)
            if cfg["steps"] and step >= cfg["steps"]:
                break
        if cfg["steps"] and step >= cfg["steps"]:
            break
    with torch.no_grad():
        ids = torch.cat([vi[0], vi[2]]).to(rt.device)
        mask = torch.cat([vi[1], vi[3]]).to(rt.device)
        out = model(ids, mask)
        acc = (out["chosen_end_scores"] > out["rejected_end_scores"]).float().mean().item()
    if rt.is_main_process:
        os.makedirs(cfg["out"], exist_ok=True)
        torch.save(model.state_dict(), os.path.join(cfg["out"], "reward_model.pt"))
        print(json.dumps({"eval_pairwise_accuracy": acc, "steps": step}))
    return acc


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
