"""Tokenizer loading that works with no network and no tokenizer files on disk.

B200 boxes have no egress and an empty HF cache, so besides local
directories this module can *synthesise* tokenizers in memory with the ``tokenizers``
library and wrap them in ``PreTrainedTokenizerFast`` — which gives the whole HF call
surface the trainers use (``__call__`` with truncation/padding, ``pad``, ``decode``,
``batch_decode``, special tokens, ``save_pretrained``):

``toy://bytes``
    byte-level, 256 symbols + ``<|endoftext|>`` (id 256).
``toy://bpe?vocab=50257``
    byte-level BPE with deterministic synthetic merges, sized to a requested vocabulary
    (default 50257 = GPT-2's, EOS = last id) — for benchmarks with GPT-2/Llama-shaped heads.
``toy://chars?alphabet=abcdefghijklmnopqrstu``
    one token per listed character (+EOS) — the randomwalks task.

Unknown hub names (e.g. ``"gpt2"``) that cannot be resolved locally fall back to
``toy://bpe`` sized from the name when known, with a warning.
"""
from __future__ import annotations

import functools
import os
from typing import Dict, List, Tuple
from urllib.parse import parse_qs, urlparse

from trlx_b200.utils import logging

logger = logging.get_logger(__name__)

EOS = "<|endoftext|>"

_KNOWN_VOCABS = {"gpt2": 50257, "gptj": 50400, "gpt-j": 50400, "neox": 50432, "llama": 32000, "opt": 50272,
                 "bloom": 250880, "t5": 32100, "pythia": 50304}


@functools.lru_cache(maxsize=1)
def _byte_alphabet() -> Dict[int, str]:
    """GPT-2's reversible byte → printable-unicode map."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def _synthetic_merges(n_merges: int) -> Tuple[Dict[str, int], List[Tuple[str, str]]]:
    b2u = _byte_alphabet()
    vocab = {b2u[b]: b for b in range(256)}
    letters = [chr(c) for c in range(ord("a"), ord("z") + 1)]
    frontier = letters + [b2u[ord(" ")]]
    merges: List[Tuple[str, str]] = []
    while len(merges) < n_merges:
        nxt = []
        for left in frontier:
            for right in letters:
                tok = left + right
                if tok in vocab:
                    continue
                vocab[tok] = len(vocab)
                merges.append((left, right))
                nxt.append(tok)
                if len(merges) == n_merges:
                    return vocab, merges
        frontier = nxt
    return vocab, merges


def _wrap(tok, model_max_length: int = 1 << 20):
    from transformers import PreTrainedTokenizerFast

    return PreTrainedTokenizerFast(
        tokenizer_object=tok, bos_token=EOS, eos_token=EOS, unk_token=EOS, model_max_length=model_max_length
    )


def build_toy_tokenizer(spec: str):
    """Build one of the ``toy://`` tokenizers described in the module docstring."""
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers

    url = urlparse(spec)
    kind = url.netloc or url.path.strip("/")
    query = {k: v[0] for k, v in parse_qs(url.query).items()}

    if kind == "chars":
        alphabet = query.get("alphabet", "abcdefghijklmnopqrstuvwxyz")
        vocab = {ch: i for i, ch in enumerate(alphabet)}
        vocab[EOS] = len(vocab)
        tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token=EOS))
        tok.pre_tokenizer = pre_tokenizers.Split("", "isolated")
        tok.decoder = decoders.Fuse()
        return _wrap(tok)

    if kind in ("bytes", "bpe"):
        if kind == "bytes":
            vocab, merges = _synthetic_merges(0)
        else:
            size = int(query.get("vocab", 50257))
            vocab, merges = _synthetic_merges(max(size - 257, 0))
        tok = Tokenizer(models.BPE(vocab=vocab, merges=merges))
        tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)
        tok.decoder = decoders.ByteLevel()
        return _wrap(tok)

    raise ValueError(f"unknown toy tokenizer spec: {spec}")


def load_tokenizer(path, **extra):
    """Resolve ``TokenizerConfig.tokenizer_path`` to a tokenizer object.

    Accepts an already-built tokenizer, a ``toy://`` spec, a local directory, or a hub name
    (resolved offline; falls back to a synthetic BPE of the matching vocabulary size).
    """
    if not isinstance(path, str):
        return path  # user handed us a tokenizer object
    if path.startswith("toy://"):
        return build_toy_tokenizer(path)
    from transformers import AutoTokenizer

    try:
        return AutoTokenizer.from_pretrained(path, **extra)
    except Exception as err:  # no files / no network
        if os.path.isdir(path):
            raise
        size = next((v for k, v in _KNOWN_VOCABS.items() if k in path.lower()), 50257)
        logger.warning(
            f"tokenizer '{path}' is not available offline ({type(err).__name__}); "
            f"using synthetic toy://bpe?vocab={size} instead"
        )
        return build_toy_tokenizer(f"toy://bpe?vocab={size}")
