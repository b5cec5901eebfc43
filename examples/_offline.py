"""Data / reward sources for the examples that work with or without network access.

The reference examples pull datasets (`imdb`, `Anthropic/hh-rlhf`, `CarperAI/openai_summarize_tldr`, …) and reward models
(`lvwerra/distilbert-imdb`, …) from the Hugging Face hub.  Each loader here first tries the real thing (when `datasets` /
the model is locally available) and otherwise falls back to a small *synthetic* stand-in with the same schema, so every
example runs end to end in an air-gapped container: that is also what `bench.py` and the tests exercise.
"""
from __future__ import annotations

import os
import random
from typing import Callable, Dict, List, Sequence, Tuple

POSITIVE = ("great wonderful brilliant moving funny delightful superb touching charming masterful excellent beautiful "
            "loved enjoyed amazing perfect").split()
NEGATIVE = ("awful boring terrible dull clumsy tedious bad painful weak lazy dreadful forgettable hated wasted poor "
            "disappointing").split()
NEUTRAL = ("the a this film movie story plot actor scene director camera script ending character was is and but with of "
           "in it I we they really quite rather very somewhat").split()


def _try_load_dataset(name: str, split: str):
    if os.environ.get("TRLX_B200_OFFLINE", "1") == "1":
        return None
    try:  # pragma: no cover - needs network
        from datasets import load_dataset

        return load_dataset(name, split=split)
    except Exception:
        return None


def synthetic_reviews(n: int = 2048, seed: int = 0) -> Tuple[List[str], List[int]]:
    """IMDB-shaped `(text, label)` pairs: the label decides which sentiment lexicon dominates."""
    rng = random.Random(seed)
    texts, labels = [], []
    for _ in range(n):
        label = rng.randint(0, 1)
        lex = POSITIVE if label else NEGATIVE
        words = []
        for _ in range(rng.randint(12, 40)):
            r = rng.random()
            words.append(rng.choice(lex) if r < 0.3 else (rng.choice(NEGATIVE if label else POSITIVE) if r < 0.34
                                                           else rng.choice(NEUTRAL)))
        texts.append(" ".join(words).capitalize() + ".")
        labels.append(label)
    return texts, labels


def load_imdb(n: int = 2048, seed: int = 0) -> Tuple[List[str], List[int]]:
    ds = _try_load_dataset("imdb", "train+test")
    if ds is not None:  # pragma: no cover
        return list(ds["text"]), list(ds["label"])
    return synthetic_reviews(n, seed)


def lexicon_sentiment(samples: Sequence[str]) -> List[float]:
    """P(positive) in [0, 1] from word counts — the offline stand-in for `lvwerra/distilbert-imdb`."""
    out = []
    for s in samples:
        words = [w.strip(".,!?").lower() for w in s.split()]
        p = sum(w in POSITIVE for w in words)
        n = sum(w in NEGATIVE for w in words)
        out.append((p + 0.5) / (p + n + 1.0))
    return out


def sentiment_scorer(device: int = -1) -> Callable[[Sequence[str]], List[Dict[str, float]]]:
    """Returns `fn(samples) -> [{"POSITIVE": p, "NEGATIVE": 1-p}]` (HF pipeline when available, lexicon otherwise)."""
    if os.environ.get("TRLX_B200_OFFLINE", "1") != "1":  # pragma: no cover - needs the hub
        try:
            from transformers import pipeline

            pipe = pipeline("sentiment-analysis", "lvwerra/distilbert-imdb", top_k=2, truncation=True, batch_size=256, device=device)
            return lambda samples: [{d["label"]: d["score"] for d in row} for row in pipe(list(samples))]
        except Exception:
            pass
    return lambda samples: [{"POSITIVE": p, "NEGATIVE": 1.0 - p} for p in lexicon_sentiment(samples)]


def synthetic_dialogues(n: int = 512, seed: int = 0) -> List[Dict[str, str]]:
    """HH-shaped records: `prompt`, `chosen`, `rejected` (helpful = on-topic and polite)."""
    rng = random.Random(seed)
    topics = ["bread", "a bicycle", "python", "the moon", "a garden", "chess", "tea", "a resume"]
    out = []
    for _ in range(n):
        t = rng.choice(topics)
        prompt = f"\n\nHuman: How do I learn about {t}?\n\nAssistant:"
        chosen = f" Happy to help. Start with the basics of {t}, practise a little every day, and ask questions as you go."
        rejected = rng.choice([" No.", " I do not care about that.", f" {t}? Figure it out yourself."])
        out.append(dict(prompt=prompt, chosen=chosen, rejected=rejected))
    return out


def synthetic_summaries(n: int = 512, seed: int = 0) -> List[Dict[str, str]]:
    """TL;DR-shaped records: `prompt` (post + 'TL;DR:'), `label` (reference summary)."""
    rng = random.Random(seed)
    subjects = ["my roommate", "my manager", "my landlord", "my sister", "a coworker"]
    events = ["keeps borrowing my things", "never answers messages", "wants to move out", "forgot my birthday",
              "asked me for money"]
    out = []
    for _ in range(n):
        s, e = rng.choice(subjects), rng.choice(events)
        filler = " ".join(rng.choice(NEUTRAL) for _ in range(rng.randint(20, 60)))
        out.append(dict(prompt=f"SUBREDDIT: r/advice\nPOST: {s.capitalize()} {e}. {filler}.\nTL;DR:", label=f" {s} {e}, what should I do?"))
    return out


def synthetic_translation(n: int = 512, seed: int = 0) -> List[Dict[str, str]]:
    """Toy en→'de' pairs (word-by-word cipher) for the seq2seq translation example."""
    rng = random.Random(seed)
    vocab = {"the": "der", "cat": "katze", "dog": "hund", "sees": "sieht", "a": "ein", "house": "haus", "small": "klein",
             "big": "gross", "likes": "mag", "bird": "vogel"}
    en = list(vocab)
    out = []
    for _ in range(n):
        words = [rng.choice(en) for _ in range(rng.randint(3, 8))]
        out.append(dict(en=" ".join(words), de=" ".join(vocab[w] for w in words)))
    return out


def overlap_f1(candidate: str, reference: str) -> float:
    """Unigram F1 — offline stand-in for ROUGE/COMET-style metrics."""
    c, r = candidate.lower().split(), reference.lower().split()
    if not c or not r:
        return 0.0
    common = sum(min(c.count(w), r.count(w)) for w in set(c))
    if common == 0:
        return 0.0
    p, rec = common / len(c), common / len(r)
    return 2 * p * rec / (p + rec)


def offline_model(name: str, fallback: dict) -> object:
    """`name` when checkpoints can be fetched, otherwise a small random-init config of the same family."""
    return name if os.environ.get("TRLX_B200_OFFLINE", "1") != "1" else fallback


GPT2_SMALL = dict(model_type="gpt2", vocab_size=50257, n_embd=768, n_layer=12, n_head=12, n_positions=1024)
GPT2_TINY = dict(model_type="gpt2", vocab_size=50257, n_embd=128, n_layer=4, n_head=4, n_positions=1024)
GPTJ_TINY = dict(model_type="gptj", vocab_size=50400, n_embd=128, n_layer=4, n_head=4, n_positions=2048, rotary_dim=16)
T5_TINY = dict(model_type="t5", vocab_size=32128, d_model=128, d_kv=32, d_ff=256, num_layers=2, num_heads=4,
               decoder_start_token_id=0, pad_token_id=0, eos_token_id=1)
LLAMA_TINY = dict(model_type="llama", vocab_size=32000, hidden_size=256, num_hidden_layers=4, num_attention_heads=8,
                  num_key_value_heads=4, intermediate_size=688, max_position_embeddings=2048)
