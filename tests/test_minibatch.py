from dataclasses import dataclass

import torch
from torch.utils.data import DataLoader, Dataset

from trlx_b200.pipeline import MiniBatchIterator
from trlx_b200.pipeline.offline_pipeline import (ILQLRolloutStorage, ILQLSeq2SeqRolloutStorage, PromptPipeline)
from trlx_b200.utils.tokenizer import build_toy_tokenizer


@dataclass
class Pair:
    a: torch.Tensor
    b: torch.Tensor


class Pairs(Dataset):
    def __init__(self, n):
        self.a = torch.arange(n)
        self.b = torch.arange(n) * 10

    def __len__(self):
        return len(self.a)

    def __getitem__(self, i):
        return self.a[i], self.b[i]

    @staticmethod
    def collate(xs):
        return Pair(torch.stack([x[0] for x in xs]), torch.stack([x[1] for x in xs]))


def _loader(n, bs):
    return DataLoader(Pairs(n), batch_size=bs, collate_fn=Pairs.collate)


def test_even_split_preserves_order_and_type():
    out = list(MiniBatchIterator(_loader(16, 8), mb_size=4, num_mb=2))
    assert len(out) == 2 and all(len(mbs) == 2 for mbs in out)
    flat = torch.cat([mb.a for mbs in out for mb in mbs])
    assert flat.tolist() == list(range(16)) and isinstance(out[0][0], Pair)
    assert out[1][1].b.tolist() == [120, 130, 140, 150]


def test_short_tail_batches():
    out = list(MiniBatchIterator(_loader(10, 8), mb_size=4, num_mb=2))
    assert [len(m) for m in out] == [2, 1]          # last batch has 2 rows → one short micro-batch
    assert out[1][0].a.tolist() == [8, 9]
    out = list(MiniBatchIterator(_loader(13, 8), mb_size=4, num_mb=2))
    assert [mb.a.numel() for mb in out[1]] == [4, 1]


def test_single_microbatch_and_small_dataset():
    out = list(MiniBatchIterator(_loader(3, 8), mb_size=8, num_mb=1))
    assert len(out) == 1 and out[0][0].a.tolist() == [0, 1, 2]
    out = list(MiniBatchIterator(_loader(3, 8), mb_size=2, num_mb=4))
    assert [mb.a.tolist() for mb in out[0]] == [[0, 1], [2]]


def test_with_prompt_pipeline_and_ilql_stores():
    tok = build_toy_tokenizer("toy://bytes")
    tok.pad_token = "<|padding|>"
    pipe = PromptPipeline(["a", "bb", "ccc", "dddd"] * 2, 8, tok)
    for mbs in MiniBatchIterator(pipe.create_loader(4), mb_size=2, num_mb=2):
        assert len(mbs) == 2 and all(mb["input_ids"].shape[0] == 2 for mb in mbs)
        assert set(mbs[0].keys()) == {"input_ids", "attention_mask"}

    n = 6
    cols = [[torch.arange(3 + i) for i in range(n)], [torch.ones(3 + i, dtype=torch.long) for i in range(n)],
            [torch.zeros(2 + i) for i in range(n)], [torch.arange(3 + i) for i in range(n)],
            [torch.arange(2 + i) for i in range(n)], [torch.ones(3 + i, dtype=torch.long) for i in range(n)]]
    store = ILQLRolloutStorage(*cols)
    batches = list(MiniBatchIterator(store.create_loader(4, shuffle=False, drop_last=False), mb_size=2, num_mb=2))
    assert sum(len(mb.input_ids) for mbs in batches for mb in mbs) == n
    assert batches[0][0].rewards.dtype == torch.float32 and batches[0][0].input_ids.shape[1] == 6

    s2s = ILQLSeq2SeqRolloutStorage(cols[0], cols[1], [torch.arange(2 + i) for i in range(n)], *cols[2:])
    b = next(iter(s2s.create_loader(3, shuffle=False, drop_last=False)))
    assert b.decoder_input_ids.shape[0] == 3 and len(b) == 3
