"""Pairwise accuracy of a trained reward model on held-out comparisons
(reference: examples/summarize_rlhf/reward_model/gptj_reward_test.py).

    python -m examples.summarize_rlhf.reward_model.gptj_reward_test [rm_checkpoint/pytorch_model.bin]

Offline the comparisons are the synthetic (reference summary vs corrupted summary) pairs of ``train_reward_model.make_pairs``."""
import os
import random
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from examples._offline import GPTJ_TINY, _try_load_dataset, offline_model
from examples.summarize_rlhf.reward_model.reward_model import GPTRewardModel
from examples.summarize_rlhf.reward_model.train_reward_model import make_pairs
from trlx_b200.utils.tokenizer import load_tokenizer


def set_seed(seed_val: int = 42) -> None:
    random.seed(seed_val)
    np.random.seed(seed_val)
    torch.manual_seed(seed_val)


def create_comparison_dataset(path: str = "CarperAI/openai_summarize_comparisons", split: str = "test"):
    """``[{chosen, rejected}]`` with the prompt prepended; pairs whose two sides are (nearly) identical are dropped."""
    ds = _try_load_dataset(path, split)
    rows = list(ds) if ds is not None else make_pairs(256, seed=7)
    pairs = []
    for r in rows:
        chosen, rejected = r["prompt"] + "\n" + r["chosen"], r["prompt"] + "\n" + r["rejected"]
        if chosen != rejected and len(r["chosen"].split()) >= 2 and len(r["rejected"].split()) >= 2:
            pairs.append(dict(chosen=chosen, rejected=rejected))
    return pairs


class PairwiseDataset(Dataset):
    def __init__(self, pairs, tokenizer, max_length: int):
        self.items = []
        for p in pairs:
            c = tokenizer(p["chosen"] + tokenizer.eos_token, truncation=True, max_length=max_length, padding="max_length", return_tensors="pt")
            r = tokenizer(p["rejected"] + tokenizer.eos_token, truncation=True, max_length=max_length, padding="max_length", return_tensors="pt")
            if not torch.equal(c["input_ids"], r["input_ids"]):
                self.items.append((c["input_ids"][0], c["attention_mask"][0], r["input_ids"][0], r["attention_mask"][0]))

    def __len__(self) -> int:
        return len(self.items)

    def __getitem__(self, idx: int):
        return self.items[idx]


class DataCollatorReward:
    """Chosen sequences stacked on top of the rejected ones — the layout ``GPTRewardModel.forward`` expects."""

    def __call__(self, data):
        return {"input_ids": torch.cat([torch.stack([d[0] for d in data]), torch.stack([d[2] for d in data])]),
                "attention_mask": torch.cat([torch.stack([d[1] for d in data]), torch.stack([d[3] for d in data])]),
                "labels": torch.tensor([0] * len(data) + [1] * len(data))}


def main(checkpoint: str = "rm_checkpoint/pytorch_model.bin", max_length: int = 550, batch_size: int = 6) -> float:
    set_seed()
    tok = load_tokenizer("EleutherAI/gpt-j-6B")
    tok.pad_token = tok.eos_token
    tok.padding_side = "right"
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model = GPTRewardModel(offline_model("EleutherAI/gpt-j-6B", GPTJ_TINY), tok.pad_token_id)
    if os.path.exists(checkpoint):
        model.load_state_dict(torch.load(checkpoint, map_location="cpu"), strict=False)
    else:
        print(f"[gptj_reward_test] no checkpoint at {checkpoint}: scoring with an untrained head (accuracy ≈ chance)")
    model = model.to(device).eval()
    if device.type == "cuda":
        model = model.to(torch.bfloat16)
    loader = DataLoader(PairwiseDataset(create_comparison_dataset(), tok, max_length), batch_size=batch_size,
                        collate_fn=DataCollatorReward())
    correct = total = 0
    with torch.no_grad():
        for batch in loader:
            out = model(batch["input_ids"].to(device), batch["attention_mask"].to(device))
            correct += int((out["chosen_end_scores"] > out["rejected_end_scores"]).sum())
            total += out["chosen_end_scores"].numel()
    acc = correct / max(total, 1)
    print(f"Total accuracy: {acc:.4f} ({correct}/{total})")
    return acc


if __name__ == "__main__":
    main(*sys.argv[1:2])
