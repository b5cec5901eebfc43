"""PPO rollout storage.

Parity: ``trlx/pipeline/ppo_pipeline.py`` — ``ppo_collate_fn`` ``:14-50`` (queries padded on the
tokenizer's padding side, everything else right-padded) and ``PPORolloutStorage`` ``:53-104``
(``push`` / ``clear_history`` / ``export_history`` / ``create_loader``).

B200 design: besides the reference's python list of per-sample CPU tensors, the
store accepts whole *blocks* of rollouts as dense, already padded, **device-resident** tensors
(:class:`RolloutBlock`).  ``create_loader`` then returns a :class:`DeviceBatchLoader` that
builds every minibatch with one ``index_select`` per field on the GPU — no ``.cpu()``, no
per-sample slicing, no re-padding and no host→device copy per minibatch
(reference: ``accelerate_ppo_trainer.py:462-502`` + ``:138-142``).  ``history`` / ``__getitem__``
still expose ``PPORLElement`` views for user code and for ``export_history``.
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass
from functools import partial
from typing import Iterable, Iterator, List, Optional

import torch
from torch.utils.data import DataLoader

from trlx_b200.data.ppo_types import PPORLBatch, PPORLElement
from trlx_b200.pipeline import BaseRolloutStore
from trlx_b200.pipeline.offline_pipeline import pad_rows


def ppo_collate_fn(padding_side: str, pad_token_id: int, elems: Iterable[PPORLElement]) -> PPORLBatch:
    elems = list(elems)
    q_side = "left" if padding_side == "left" else "right"
    return PPORLBatch(
        pad_rows([e.query_tensor for e in elems], pad_token_id, q_side),
        pad_rows([e.response_tensor for e in elems], pad_token_id, "right"),
        pad_rows([e.logprobs for e in elems], 0.0, "right"),
        pad_rows([e.values for e in elems], 0.0, "right"),
        pad_rows([e.rewards for e in elems], 0.0, "right"),
    )


@dataclass
class RolloutBlock:
    """A dense chunk of rollouts living on one device.

    ``queries`` ``[N,Q]`` padded on ``query_side``; ``responses`` / ``logprobs`` / ``values`` /
    ``rewards`` ``[N,R]`` right-padded; ``query_lens`` / ``response_lens`` ``[N]`` (int64, also
    mirrored on the host in ``host_response_lens`` so batch widths can be chosen without a sync).
    """

    queries: torch.Tensor
    responses: torch.Tensor
    logprobs: torch.Tensor
    values: torch.Tensor
    rewards: torch.Tensor
    query_lens: torch.Tensor
    response_lens: torch.Tensor
    host_response_lens: Optional[List[int]] = None
    host_query_lens: Optional[List[int]] = None
    trunk_hidden: Optional[torch.Tensor] = None  # [N, Q+R, H] frozen-trunk activation at the branch point (optional)

    def __len__(self) -> int:
        return int(self.queries.shape[0])


@dataclass
class PPORLBatchCached(PPORLBatch):
    """:class:`PPORLBatch` + the cached frozen-trunk activation ``[B, Q+R, H]`` aligned with ``cat(query, response)``."""

    trunk_hidden: Optional[torch.Tensor] = None


class DeviceBatchLoader:
    """Iterates minibatches of a list of :class:`RolloutBlock` entirely on the device.

    Every batch is trimmed to the longest response (and longest query) *in that batch* — the same
    widths the reference collate produces, which matters because GAE / whitening are taken over
    padded positions too.  ``static_shapes=True`` keeps the full block width instead
    (for CUDA-graph replay); the fused GAE kernel then receives the batch width as a scalar.
    """

    def __init__(self, blocks: List[RolloutBlock], batch_size: int, shuffle: bool, pad_token_id: int,
                 query_side: str = "left", static_shapes: bool = False, generator: Optional[torch.Generator] = None):
        self.batch_size, self.shuffle, self.static_shapes = batch_size, shuffle, static_shapes
        self.generator = generator
        dev = blocks[0].queries.device
        Q = max(b.queries.shape[1] for b in blocks)
        R = max(b.responses.shape[1] for b in blocks)

        def widen(t, width, value, left=False):
            if t.shape[1] == width:
                return t
            pad = t.new_full((t.shape[0], width - t.shape[1]), value)
            return torch.cat([pad, t] if left else [t, pad], dim=1)

        left = query_side == "left"
        self.queries = torch.cat([widen(b.queries, Q, pad_token_id, left) for b in blocks])
        self.responses = torch.cat([widen(b.responses, R, pad_token_id) for b in blocks])
        self.logprobs = torch.cat([widen(b.logprobs, R, 0.0) for b in blocks])
        self.values = torch.cat([widen(b.values, R, 0.0) for b in blocks])
        self.rewards = torch.cat([widen(b.rewards, R, 0.0) for b in blocks])
        self.trunk = None
        if all(b.trunk_hidden is not None for b in blocks):
            parts = []
            for b in blocks:
                t = b.trunk_hidden
                qb, rb = b.queries.shape[1], b.responses.shape[1]
                t = t[:, : qb + rb]
                if t.shape[1] < qb + rb:  # engines cache T-1 positions (the last token is never an input)
                    t = torch.cat([t, t.new_zeros(t.shape[0], qb + rb - t.shape[1], t.shape[2])], 1)
                tq, tr = t[:, :qb], t[:, qb:]
                if qb < Q:
                    zq = t.new_zeros(t.shape[0], Q - qb, t.shape[2])
                    tq = torch.cat([zq, tq] if left else [tq, zq], 1)
                if rb < R:
                    tr = torch.cat([tr, t.new_zeros(t.shape[0], R - rb, t.shape[2])], 1)
                parts.append(torch.cat([tq, tr], 1))
            self.trunk = torch.cat(parts)
        self.Q, self.R = Q, R
        self.response_lens = torch.cat([b.response_lens for b in blocks])
        self.query_lens = torch.cat([b.query_lens for b in blocks])
        self.host_rlens = sum((b.host_response_lens or b.response_lens.tolist() for b in blocks), [])
        self.host_qlens = sum((b.host_query_lens or b.query_lens.tolist() for b in blocks), [])
        self.left, self.device, self.n = left, dev, int(self.queries.shape[0])

    def __len__(self) -> int:
        return (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[PPORLBatch]:
        order = torch.randperm(self.n, generator=self.generator) if self.shuffle else torch.arange(self.n)
        for lo in range(0, self.n, self.batch_size):
            idx_host = order[lo: lo + self.batch_size]
            idx = idx_host.to(self.device, non_blocking=True)
            q, r = self.queries.index_select(0, idx), self.responses.index_select(0, idx)
            lp, v, rw = (t.index_select(0, idx) for t in (self.logprobs, self.values, self.rewards))
            th = self.trunk.index_select(0, idx) if self.trunk is not None else None
            qmax, rmax = self.Q, self.R
            ids = idx_host.tolist()
            width = min(max(max(self.host_rlens[i] for i in ids), 1), self.R)  # widest scored response of this batch
            if not self.static_shapes:
                rmax = width
                qmax = max(max(self.host_qlens[i] for i in ids), 1)
                q = q[:, q.shape[1] - qmax:] if self.left else q[:, :qmax]
                # responses keep one token more than the longest scored slice: position i is scored against token i+1
                rtok = min(rmax + 1, self.R)
                r, lp, v, rw = r[:, :rtok], lp[:, :rmax], v[:, :rmax], rw[:, :rmax]
                if th is not None:
                    th = th[:, self.Q - qmax: self.Q + rtok] if self.left else torch.cat([th[:, :qmax], th[:, self.Q: self.Q + rtok]], 1)
            batch = PPORLBatchCached(q, r, lp, v, rw, trunk_hidden=th) if th is not None else PPORLBatch(q, r, lp, v, rw)
            batch.width = width  # python int: in static-shape mode the kernels take it through a device scalar
            yield batch


class PPORolloutStorage(BaseRolloutStore):
    """Rollout storage for PPO (element list and/or device-resident blocks)."""

    def __init__(self, pad_token_id: int, padding_side: str):
        super().__init__()
        self.pad_token_id = pad_token_id
        self.padding_side = padding_side
        self._elements: List[PPORLElement] = []
        self._blocks: List[RolloutBlock] = []
        self._placeholder = True  # the reference starts with ``history = [None]``

    # -- element view ---------------------------------------------------------------------------
    def _block_elements(self) -> List[PPORLElement]:
        out = []
        for b in self._blocks:
            qlen = b.host_query_lens or b.query_lens.tolist()
            rlen = b.host_response_lens or b.response_lens.tolist()
            Q = b.queries.shape[1]
            for i in range(len(b)):
                ql, rl = int(qlen[i]), int(rlen[i])
                q = b.queries[i, Q - ql:] if self.padding_side == "left" else b.queries[i, :ql]
                out.append(PPORLElement(q, b.responses[i, :rl], b.logprobs[i, :rl], b.values[i, :rl], b.rewards[i, :rl]))
        return out

    @property
    def history(self) -> List[PPORLElement]:
        if self._placeholder and not self._elements and not self._blocks:
            return [None]
        return self._elements + self._block_elements()

    @history.setter
    def history(self, value):
        if value is None:
            return
        self._elements = [v for v in value if v is not None]
        self._blocks = []
        self._placeholder = False

    def push(self, exps: Iterable[PPORLElement]):
        self._elements += list(exps)

    def push_block(self, block: RolloutBlock):
        self._blocks.append(block)

    def clear_history(self):
        self._elements, self._blocks, self._placeholder = [], [], False

    def export_history(self, location: str, only_text: bool = True):
        """Dump rollouts to ``location/epoch-<time>.json`` (Algorithm-Distillation export)."""
        assert os.path.exists(location)
        keep = ("query_tensor", "response_tensor") if only_text else None
        rows = []
        for exp in self.history:
            d = {k: v.cpu().tolist() for k, v in exp.asdict().items()}
            rows.append({k: v for k, v in d.items() if keep is None or k in keep})
        with open(os.path.join(location, f"epoch-{time.time()}.json"), "w") as fh:
            fh.write(json.dumps(rows, indent=2))

    def __getitem__(self, index: int) -> PPORLElement:
        return self.history[index]

    def __len__(self) -> int:
        n = len(self._elements) + sum(len(b) for b in self._blocks)
        return n if (n or not self._placeholder) else 1

    def create_loader(self, batch_size: int, shuffle: bool, static_shapes: bool = False, generator=None):
        if self._blocks and not self._elements:
            return DeviceBatchLoader(self._blocks, batch_size, shuffle, self.pad_token_id, self.padding_side,
                                     static_shapes=static_shapes, generator=generator)
        return DataLoader(self, batch_size, shuffle=shuffle, generator=generator,
                          collate_fn=partial(ppo_collate_fn, self.padding_side, self.pad_token_id))
