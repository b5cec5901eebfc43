"""Prompt pipeline, dialogue tokenisation and offline (SFT / ILQL) stores.

Parity: ``trlx/pipeline/offline_pipeline.py`` — ``DialogMessage`` ``:22-35``, ``tokenize_dialogue``
``:38-87``, ``DialogStore`` ``:90-115``, ``PromptPipeline`` ``:118-188``, ILQL stores ``:191-289``.

Layout choice: ragged per-sample tensors are padded with one vectorised helper
(:func:`pad_rows`) that supports left or right padding directly, instead of the reference's
flip → ``pad_sequence`` → flip idiom; the ILQL stores are one generic columnar class.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Type, Union

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader

from trlx_b200.data.ilql_types import ILQLBatch, ILQLElement, ILQLSeq2SeqBatch, ILQLSeq2SeqElement
from trlx_b200.pipeline import BasePipeline, BaseRolloutStore, register_datapipeline


def pad_rows(rows: Sequence[torch.Tensor], pad_value, side: str = "right", dtype=None,
             min_len: int = 0) -> torch.Tensor:
    """Stack 1-D tensors of different lengths into ``[N, max_len]`` padded on ``side``."""
    width = max(max((int(r.shape[0]) for r in rows), default=0), min_len)
    dtype = dtype or (rows[0].dtype if len(rows) else torch.long)
    out = torch.full((len(rows), width), pad_value, dtype=dtype)
    for i, r in enumerate(rows):
        n = int(r.shape[0])
        if n == 0:
            continue
        if side == "left":
            out[i, width - n:] = r
        else:
            out[i, :n] = r
    return out


@dataclass
class DialogMessage:
    """One turn of a dialogue: ``is_output`` marks model turns (trained on); ``tokens`` are ids."""

    is_output: bool
    tokens: Tuple[int, ...]


def tokenize_dialogue(dialogue: Union[str, Iterable[str]], tokenizer, max_length: int = 2048) -> List[DialogMessage]:
    """Tokenise ``(prompt_1, output_1, prompt_2, output_2, …)`` into :class:`DialogMessage` s.

    * a bare string is treated as ``(bos, string)``;
    * EOS is appended to the last phrase if missing;
    * the total is truncated to ``max_length`` tokens from ``tokenizer.truncation_side``;
    * messages truncated to nothing are dropped; if the result starts with an output, a BOS
      prompt is inserted (dropping one more token if the budget was exactly exhausted).
    """
    if isinstance(dialogue, str):
        phrases = [tokenizer.bos_token or tokenizer.eos_token, dialogue]
    else:
        phrases = list(dialogue)
        if len(phrases) % 2 != 0:
            raise ValueError("Dialogue must have an even number of phrases, alternating prompt and output")
    if not phrases[-1].endswith(tokenizer.eos_token):
        phrases[-1] = phrases[-1] + tokenizer.eos_token

    ids = [tuple(tokenizer(p, add_special_tokens=False).input_ids) for p in phrases]
    from_left = tokenizer.truncation_side == "left"  # truncate the *start* of the dialogue

    # Walk messages starting from the side that is kept; hand out the remaining budget.
    order = range(len(ids) - 1, -1, -1) if from_left else range(len(ids))
    budget = max_length
    kept: Dict[int, Tuple[int, ...]] = {}
    for i in order:
        take = min(len(ids[i]), max(budget, 0))
        if take:
            kept[i] = ids[i][len(ids[i]) - take:] if from_left else ids[i][:take]
        budget -= len(ids[i])

    out = [DialogMessage(is_output=(i % 2 == 1), tokens=kept[i]) for i in sorted(kept)]
    if out and out[0].is_output:
        if sum(len(m.tokens) for m in out) == max_length:
            if from_left:
                out[0].tokens = out[0].tokens[1:]
            else:
                out[-1].tokens = out[-1].tokens[:-1]
            out = [m for m in out if len(m.tokens) > 0]  # the trimmed message may have had a single token
        if out and out[0].is_output:
            out.insert(0, DialogMessage(False, (tokenizer.bos_token_id,)))
    return out


class DialogStore(BaseRolloutStore):
    """SFT store: ``labels`` equal the ids on output tokens and ``-100`` on prompt tokens."""

    def __init__(self, dialogs: List[List[DialogMessage]], tokenizer):
        super().__init__()
        self.tokenizer = tokenizer
        self.history = []
        for d in dialogs:
            flat = [t for m in d for t in m.tokens]
            lab = [t if m.is_output else -100 for m in d for t in m.tokens]
            self.history.append(
                dict(
                    input_ids=torch.tensor(flat, dtype=torch.long),
                    attention_mask=torch.ones(len(flat), dtype=torch.bool),
                    labels=torch.tensor(lab, dtype=torch.long),
                )
            )

    def create_loader(self, batch_size: int, shuffle: bool = False) -> DataLoader:
        from transformers.tokenization_utils_base import BatchEncoding

        side = getattr(self.tokenizer, "padding_side", "right")
        pad_id = self.tokenizer.pad_token_id

        def collate(elems: Iterable[dict]) -> BatchEncoding:
            elems = list(elems)
            return BatchEncoding(
                dict(
                    input_ids=pad_rows([e["input_ids"] for e in elems], pad_id, side),
                    attention_mask=pad_rows([e["attention_mask"].long() for e in elems], 0, side),
                    # like the reference, labels are padded with the pad id (the loss masks on attention)
                    labels=pad_rows([e["labels"] for e in elems], pad_id, side),
                )
            )

        return DataLoader(self, batch_size=batch_size, collate_fn=collate, shuffle=shuffle)


@register_datapipeline
class PromptPipeline(BasePipeline):
    """Prompts for rollouts / evaluation.

    :param prompts: list of strings, or list of dicts with a ``"prompt"`` key whose other keys are
        carried along and forwarded to ``reward_fn`` / ``metric_fn`` as keyword arguments
    :param max_prompt_length: prompts are truncated to this many tokens (tokenizer's truncation side)
    :param tokenizer: HF-style tokenizer
    :param add_special_tokens: forwarded to the tokenizer
    """

    def __init__(self, prompts: Union[List[Dict[str, Any]], List[str]], max_prompt_length: int, tokenizer,
                 add_special_tokens: bool = False):
        super().__init__()
        if len(prompts) and isinstance(prompts[0], dict):
            metadata = [{k: v for k, v in p.items() if k != "prompt"} for p in prompts]
            texts = [p["prompt"] for p in prompts]
        else:
            metadata = [{} for _ in prompts]
            texts = list(prompts)
        enc = tokenizer(texts, truncation=True, padding=False, max_length=max_prompt_length,
                        add_special_tokens=add_special_tokens)
        self.tokenizer = tokenizer
        self.prompts = [
            {"input_ids": ids, "attention_mask": mask, **meta}
            for ids, mask, meta in zip(enc["input_ids"], enc["attention_mask"], metadata)
        ]

    def __getitem__(self, ix: int):
        return self.prompts[ix]

    def __len__(self) -> int:
        return len(self.prompts)

    def create_loader(self, batch_size: int, shuffle: bool = False, sampler=None, drop_last: bool = False) -> DataLoader:
        from transformers.tokenization_utils_base import BatchEncoding

        tok = self.tokenizer

        def collate(xs):
            side = getattr(tok, "padding_side", "left")
            ids = [torch.tensor(x["input_ids"], dtype=torch.long) for x in xs]
            out = BatchEncoding(
                dict(
                    input_ids=pad_rows(ids, tok.pad_token_id, side),
                    attention_mask=pad_rows([torch.ones_like(i) for i in ids], 0, side),
                )
            )
            for key in xs[0]:
                if key not in ("input_ids", "attention_mask"):
                    out[key] = [x[key] for x in xs]
            return out

        def collate_pinned(xs):
            out = collate(xs)
            if torch.cuda.is_available():  # prompts are staged in pinned host memory → async H2D in the trainers
                for k in ("input_ids", "attention_mask"):
                    out[k] = out[k].pin_memory()
            return out

        return DataLoader(self, batch_size=batch_size, collate_fn=collate_pinned, shuffle=shuffle if sampler is None else False,
                          sampler=sampler, num_workers=0, drop_last=drop_last)


# ---- ILQL ------------------------------------------------------------------------------------------
class _ColumnarStore(BaseRolloutStore):
    """Ragged columnar store: one python list of tensors per field of ``element_cls``."""

    element_cls: Type = None
    batch_cls: Type = None
    _float_cols = ("rewards",)

    def __init__(self, *columns):
        super().__init__()
        names = [f for f in self.element_cls.__dataclass_fields__]
        if len(columns) != len(names):
            raise TypeError(f"{type(self).__name__} expects columns {names}")
        for n, c in zip(names, columns):
            setattr(self, n, c)
        self._names = names

    def __getitem__(self, ix: int):
        return self.element_cls(*(getattr(self, n)[ix] for n in self._names))

    def __len__(self) -> int:
        return len(getattr(self, self._names[0]))

    @classmethod
    def collate(cls, elems):
        elems = list(elems)
        cols = []
        for n in cls.element_cls.__dataclass_fields__:
            rows = [getattr(e, n) for e in elems]
            cols.append(pad_rows(rows, 0.0 if n in cls._float_cols else 0, "right"))
        return cls.batch_cls(*cols)

    def create_loader(self, batch_size: int, shuffle: bool = True, sampler=None, drop_last: Optional[bool] = None):
        if drop_last is None:
            drop_last = dist.is_available() and dist.is_initialized()
        return DataLoader(self, batch_size=batch_size, shuffle=shuffle if sampler is None else False, sampler=sampler,
                          collate_fn=type(self).collate, drop_last=drop_last)


class ILQLRolloutStorage(_ColumnarStore):
    """``(input_ids, attention_mask, rewards, states_ixs, actions_ixs, dones)`` per sample."""

    element_cls = ILQLElement
    batch_cls = ILQLBatch

    def __init__(self, input_ids, attention_mask, rewards, states_ixs, actions_ixs, dones):
        super().__init__(input_ids, attention_mask, rewards, states_ixs, actions_ixs, dones)


class ILQLSeq2SeqRolloutStorage(_ColumnarStore):
    """Seq2seq variant with an extra ``decoder_input_ids`` column."""

    element_cls = ILQLSeq2SeqElement
    batch_cls = ILQLSeq2SeqBatch

    def __init__(self, input_ids, attention_mask, decoder_input_ids, rewards, states_ixs, actions_ixs, dones):
        super().__init__(input_ids, attention_mask, decoder_input_ids, rewards, states_ixs, actions_ixs, dones)


def ilql_collate_fn(elems: Iterable[ILQLElement]) -> ILQLBatch:
    return ILQLRolloutStorage.collate(elems)


def ilql_seq2seq_collate_fn(elems: Iterable[ILQLSeq2SeqElement]) -> ILQLSeq2SeqBatch:
    return ILQLSeq2SeqRolloutStorage.collate(elems)
