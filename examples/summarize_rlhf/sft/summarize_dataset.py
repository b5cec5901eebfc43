"""Datasets of the summarisation recipe (reference: examples/summarize_rlhf/sft/summarize_dataset.py): TL;DR posts for SFT,
pairwise comparisons for the reward model, and a helper that renders the JSONL dumps.  Hub datasets are used when they are on
disk; otherwise the synthetic summaries of ``examples/_offline.py`` stand in so every stage runs offline."""
import json
from typing import List

import torch
from torch.utils.data import Dataset

from examples._offline import _try_load_dataset, synthetic_summaries


def render_post(d: dict, with_summary: bool = True) -> str:
    head = f"SUBREDDIT: r/{d['subreddit']}\nTITLE: {d['title']}\nPOST: {d['post']}\nTL;DR: "
    return head + d["summary"] if with_summary else head


def get_dataset_from_jsonl(jsonl_file: str, return_summary: bool = True):
    """Posts of an OpenAI-format TL;DR dump: with the summary appended, or ``(prompts, summaries)`` when ``return_summary``
    is false."""
    with open(jsonl_file) as fh:
        rows = [json.loads(line) for line in fh if line.strip()]
    if return_summary:
        return [render_post(r) for r in rows]
    return [render_post(r, False) for r in rows], [r["summary"] for r in rows]


def _rows(path: str, split: str, n_offline: int = 2048) -> List[dict]:
    ds = _try_load_dataset(path, split)
    return list(ds) if ds is not None else synthetic_summaries(n_offline, seed=0 if "train" in split else 1)


class TLDRDataset(Dataset):
    """``prompt + label`` strings tokenised to a fixed length; ``labels = input_ids`` (causal-LM fine-tuning)."""

    def __init__(self, train_path: str, tokenizer, split: str, max_length: int = 550):
        self.posts = [r["prompt"] + r["label"] for r in _rows(train_path, split)]
        if "valid" in split:
            self.posts = self.posts[:2000]
        self.tokenizer, self.max_length = tokenizer, max_length

    def __len__(self) -> int:
        return len(self.posts)

    def __getitem__(self, idx: int):
        enc = self.tokenizer(self.posts[idx], truncation=True, max_length=self.max_length, padding="max_length")
        ids = torch.tensor(enc["input_ids"])
        return {"input_ids": ids, "attention_mask": torch.tensor(enc["attention_mask"]), "labels": ids}


class ComparisonDataset(Dataset):
    """Pairs of summaries of the same post from an OpenAI-format comparison dump; ``labels`` is the index (0/1) of the one the
    annotator preferred."""

    def __init__(self, comparison_path: str, tokenizer, max_length: int = 550):
        with open(comparison_path) as fh:
            rows = [json.loads(line) for line in fh if line.strip()]
        self.tokenizer, self.max_length = tokenizer, max_length
        self.first, self.second, self.labels = [], [], []
        for r in rows:
            info = r["info"]
            head = f"SUBREDDIT: r/{info['subreddit']}\nTITLE: {info['title']}\nPOST: {info['post']}\nTL;DR: "
            a, b = (head + s["text"] for s in r["summaries"][:2])
            self.first.append(a)
            self.second.append(b)
            self.labels.append(int(r["choice"]))

    def __len__(self) -> int:
        return len(self.labels)

    def __getitem__(self, idx: int):
        pair = self.tokenizer([self.first[idx], self.second[idx]], truncation=True, max_length=self.max_length,
                              padding="max_length", return_tensors="pt")
        return {"input_ids": pair["input_ids"], "attention_mask": pair["attention_mask"], "labels": torch.tensor(self.labels[idx])}


class AllSummDataset(Dataset):
    """Generic ``text`` + ``summary`` columns rendered as ``Summarize: … TL;DR: …`` (CNN/DailyMail-style corpora)."""

    def __init__(self, train_path: str, tokenizer, split: str, max_length: int = 1024):
        rows = _rows(train_path, split)
        self.texts = [f"Summarize: {r.get('text', r.get('prompt', ''))}. TL;DR: {r.get('summary', r.get('label', ''))}" for r in rows]
        if "valid" in split:
            self.texts = self.texts[:2000]
        self.tokenizer, self.max_length = tokenizer, max_length

    def __len__(self) -> int:
        return len(self.texts)

    def __getitem__(self, idx: int):
        enc = self.tokenizer(self.texts[idx], truncation=True, max_length=self.max_length, padding="max_length")
        ids = torch.tensor(enc["input_ids"])
        return {"input_ids": ids, "attention_mask": torch.tensor(enc["attention_mask"]), "labels": ids}
