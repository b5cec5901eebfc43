"""Whole-train-step CUDA graph for PPO (forward → fused loss → backward → fused optimizer → zero-grad).

The PPO inner loop of the reference (``trlx/trainer/accelerate_base_trainer.py:563-588``) is launch-bound on small
models: ~300 tiny kernels per optimizer step driven from Python.  Here the step is captured ONCE per shape and
replayed: minibatches come from the device-resident rollout store with *static* shapes, the only per-batch scalar that
changes the maths (the widest response of the batch, which sets the GAE / whitening / mean denominators) is passed
through a device int that the kernels read, and the optimizer's host half (lr schedule, bias corrections) is staged
into pinned memory that the captured H2D copy picks up at replay time.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from trlx_b200.pipeline.ppo_pipeline import PPORLBatchCached
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


class GraphedPPOStep:
    def __init__(self, trainer, example: PPORLBatchCached):
        self.trainer = trainer
        dev = trainer.runtime.device
        self.key = self.shape_key(example)
        self.q = torch.empty_like(example.query_tensors)
        self.r = torch.empty_like(example.response_tensors)
        self.lp = torch.empty_like(example.logprobs)
        self.v = torch.empty_like(example.values)
        self.rw = torch.empty_like(example.rewards)
        th = getattr(example, "trunk_hidden", None)
        self.th = torch.empty_like(th) if th is not None else None
        self.width = torch.zeros(1, dtype=torch.int32, device=dev)
        self.width_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.batch = PPORLBatchCached(self.q, self.r, self.lp, self.v, self.rw, trunk_hidden=self.th)
        self.batch.width_tensor = self.width
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.stats: Dict[str, torch.Tensor] = {}
        self.launches = 0

    @staticmethod
    def shape_key(batch) -> Tuple:
        th = getattr(batch, "trunk_hidden", None)
        return (tuple(batch.query_tensors.shape), tuple(batch.response_tensors.shape), tuple(batch.rewards.shape),
                None if th is None else tuple(th.shape))

    def _load(self, batch, width: int):
        self.q.copy_(batch.query_tensors, non_blocking=True)
        self.r.copy_(batch.response_tensors, non_blocking=True)
        self.lp.copy_(batch.logprobs, non_blocking=True)
        self.v.copy_(batch.values, non_blocking=True)
        self.rw.copy_(batch.rewards, non_blocking=True)
        if self.th is not None:
            self.th.copy_(batch.trunk_hidden, non_blocking=True)
        self.width_host[0] = int(width)
        self.width.copy_(self.width_host, non_blocking=True)

    def _body(self):
        tr = self.trainer
        loss, stats = tr.loss(self.batch)
        rt = tr.runtime
        if getattr(tr.opt, "can_overlap", False) and rt.tp_size == 1 and rt.pp_size == 1:
            tr.opt.arm_overlap()  # bucket kernels fork onto a side stream inside the graph; device_step() joins them
        tr.model.train()
        loss.backward()
        tr.model.eval()
        tr._pre_optimizer_step()
        tr.opt.device_step()
        tr.opt.zero_grad()
        return stats

    def capture(self, batch, width: int):
        """Warm up on a side stream (undoing its optimizer effects), then capture."""
        from trlx_b200 import ops

        tr = self.trainer
        self._load(batch, width)
        snap = tr.opt.snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                tr.opt.host_prepare()
                before = ops.launch_count()
                self._body()
                self.launches = ops.launch_count() - before
        torch.cuda.current_stream().wait_stream(side)
        tr.opt.restore(snap)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.stats = self._body()
        tr.opt.restore(snap)  # capture does not execute, but keep state pristine regardless

    def run(self, batch, width: int) -> Dict[str, Any]:
        from trlx_b200 import ops

        if self.graph is None:
            self.capture(batch, width)
        self._load(batch, width)
        # the optimizer step IS the replay: host half (lr, bias corrections → pinned staging), then the captured device half
        self.trainer.opt.step(graph=self.graph)
        ops.add_launches(self.launches)
        return dict(self.stats)
