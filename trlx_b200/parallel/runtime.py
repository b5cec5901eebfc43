"""Process/runtime layer — what HF Accelerate (+DeepSpeed) provides to the reference trainers
(``trlx/trainer/accelerate_base_trainer.py:56-64``: process-group init, device placement, gather/broadcast helpers,
trackers, ``save_state`` / ``load_state``), rebuilt directly on ``torch.distributed``:

* one process per GPU, backend ``nccl`` on CUDA (``gloo`` on CPU so the same code runs in CPU tests);
* DP × TP × PP process groups from :class:`~trlx_b200.data.configs.ParallelConfig`;
* object / tensor collectives used by rollouts and evaluation;
* trackers: ``wandb`` (if importable), ``tensorboard``, ``jsonl`` or none;
* device-timed phases (CUDA events) for ``time/*`` stats and NVTX ranges for profilers.
"""
from __future__ import annotations

import contextlib
import datetime
import json
import os
import time
from typing import Any, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


class Runtime:
    def __init__(self, parallel=None, seed: Optional[int] = None):
        from trlx_b200.data.configs import ParallelConfig

        self.parallel = parallel or ParallelConfig()
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.cuda = torch.cuda.is_available()
        if self.cuda:
            torch.cuda.set_device(self.local_rank % torch.cuda.device_count())
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device("cpu")
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            backend = "nccl" if self.cuda else "gloo"
            kwargs = dict(backend=backend, rank=self.rank, world_size=self.world_size,
                          timeout=datetime.timedelta(minutes=30))
            if self.cuda:
                kwargs["device_id"] = self.device
            dist.init_process_group(**kwargs)
        self.distributed = dist.is_available() and dist.is_initialized() and self.world_size > 1
        self.dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[self.parallel.precision]
        if not self.cuda:
            self.dtype = torch.float32
        self._build_groups()
        self.tracker = None
        self._tracker_kind = None
        self._jsonl = None

    # ---- topology ---------------------------------------------------------------------------------------------------
    def _build_groups(self):
        tp, pp = int(self.parallel.tensor_parallel), int(self.parallel.pipeline_parallel)
        if self.world_size == 1 and tp * pp > 1:
            # a single process cannot be mis-sharded: run the recipe unsharded (debugging / smoke runs of the NeMo-style examples)
            import warnings

            warnings.warn(f"tensor_parallel={tp} x pipeline_parallel={pp} requested but only one process is running: "
                          "continuing without model parallelism")
            tp = pp = 1
            self.parallel.tensor_parallel = self.parallel.pipeline_parallel = 1
            self.parallel.sequence_parallel = False
        if self.world_size % (tp * pp) != 0:
            raise ValueError(f"world_size {self.world_size} is not divisible by tp*pp = {tp * pp}")
        self.tp_size, self.pp_size = tp, pp
        self.dp_size = self.world_size // (tp * pp)
        # rank = ((pp_rank * dp) + dp_rank) * tp + tp_rank  — TP innermost so a TP group sits on adjacent GPUs
        self.tp_rank = self.rank % tp
        self.dp_rank = (self.rank // tp) % self.dp_size
        self.pp_rank = self.rank // (tp * self.dp_size)
        self.tp_group = self.dp_group = self.pp_group = self.stage_group = None
        from trlx_b200.utils.modeling import set_statistics_group

        set_statistics_group(None)
        if not self.distributed or (tp == 1 and pp == 1):
            self.dp_group = None  # WORLD
            self._publish_state()
            return
        for p in range(pp):
            for d in range(self.dp_size):
                ranks = [((p * self.dp_size) + d) * tp + t for t in range(tp)]
                g = dist.new_group(ranks)
                if self.rank in ranks:
                    self.tp_group = g
        for p in range(pp):
            for t in range(tp):
                ranks = [((p * self.dp_size) + d) * tp + t for d in range(self.dp_size)]
                g = dist.new_group(ranks)
                if self.rank in ranks:
                    self.dp_group = g
        for d in range(self.dp_size):
            for t in range(tp):
                ranks = [((p * self.dp_size) + d) * tp + t for p in range(pp)]
                g = dist.new_group(ranks)
                if self.rank in ranks:
                    self.pp_group = g
        # every rank of one pipeline stage (tensor x data parallel): the replica set of parameters that sequence parallelism
        # leaves replicated inside the TP group (block norms) — their gradients are summed over TP and averaged over DP in ONE
        # pass of the fused optimizer kernel over this group
        self.stage_group = None
        for p in range(pp):
            ranks = [((p * self.dp_size) + d) * tp + t for d in range(self.dp_size) for t in range(tp)]
            g = dist.new_group(ranks)
            if self.rank in ranks:
                self.stage_group = g
        set_statistics_group(self.dp_group)  # model-parallel peers hold identical data: statistics span DP only
        self._publish_state()

    def _publish_state(self):
        """Make the layout visible to group-less building blocks (``parallel/state.py``)."""
        from trlx_b200.parallel.state import set_model_parallel

        set_model_parallel(tp_group=self.tp_group, tp_rank=self.tp_rank, tp_size=self.tp_size, pp_group=self.pp_group,
                           pp_rank=self.pp_rank, pp_size=self.pp_size, dp_group=self.dp_group, dp_rank=self.dp_rank,
                           dp_size=self.dp_size)

    @property
    def is_replica_leader(self) -> bool:
        """One rank per model replica (first tensor-parallel rank of the first pipeline stage)."""
        return self.tp_rank == 0 and self.pp_rank == 0

    @property
    def is_main_process(self) -> bool:
        return self.rank == 0

    @property
    def num_processes(self) -> int:
        return self.world_size

    def describe(self) -> Dict[str, Any]:
        return dict(mixed_precision=self.parallel.precision, num_gpus=self.world_size, dp=self.dp_size, tp=self.tp_size,
                    pp=self.pp_size, zero_stage=self.parallel.zero_stage, gradient_clipping=self.parallel.grad_clip)

    # ---- collectives ------------------------------------------------------------------------------------------------
    def barrier(self):
        if self.distributed:
            dist.barrier()

    wait_for_everyone = barrier

    def all_reduce(self, t: torch.Tensor, op: str = "sum", group=None) -> torch.Tensor:
        if self.distributed:
            dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[
                "sum" if op == "mean" else op], group=group)
            if op == "mean":
                t /= dist.get_world_size(group)
        return t

    def gather(self, t: torch.Tensor) -> torch.Tensor:
        """Concatenate ``t`` from all ranks along dim 0 (same shapes required)."""
        if not self.distributed:
            return t
        t = t.contiguous()
        out = [torch.empty_like(t) for _ in range(self.world_size)]
        dist.all_gather(out, t)
        return torch.cat(out, 0)

    def pad_across_processes(self, tensors: Sequence[torch.Tensor], dim: int = 1, pad_index: int = 0,
                             pad_first: bool = False) -> List[torch.Tensor]:
        """Right- (or left-) pad each tensor along ``dim`` to the max size over ranks."""
        if not self.distributed:
            return list(tensors)
        out = []
        for t in tensors:
            if t.dim() <= dim:
                out.append(t)
                continue
            size = torch.tensor([t.shape[dim]], device=t.device)
            dist.all_reduce(size, op=dist.ReduceOp.MAX)
            width = int(size.item())
            if width == t.shape[dim]:
                out.append(t)
                continue
            shape = list(t.shape)
            shape[dim] = width - t.shape[dim]
            pad = t.new_full(shape, pad_index)
            out.append(torch.cat([pad, t] if pad_first else [t, pad], dim))
        return out

    def gather_objects(self, obj: Any) -> List[Any]:
        if not self.distributed:
            return [obj]
        out = [None] * self.world_size
        dist.all_gather_object(out, obj)
        return out

    def broadcast_object(self, obj: Any, src: int = 0) -> Any:
        if not self.distributed:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    # ---- trackers ---------------------------------------------------------------------------------------------------
    def init_tracker(self, kind: Optional[str], project: str, config: Dict[str, Any], run_name: str,
                     logging_dir: Optional[str] = None, **init_kwargs) -> None:
        if not self.is_main_process or kind is None:
            return
        kind = str(kind).lower()
        if kind == "wandb":
            try:
                import wandb

                self.tracker = wandb.init(project=project, config=config, name=run_name, **init_kwargs)
                self._tracker_kind = "wandb"
                return
            except Exception as err:
                logger.warning(f"wandb tracker unavailable ({type(err).__name__}: {err}); falling back to jsonl")
                kind = "jsonl"
        if kind == "tensorboard":
            try:
                from torch.utils.tensorboard import SummaryWriter

                self.tracker = SummaryWriter(log_dir=os.path.join(logging_dir or "logs", run_name.replace("/", "_")))
                self._tracker_kind = "tensorboard"
                flat = {k: (v if isinstance(v, (int, float, str, bool)) else str(v)) for k, v in config.items()}
                self.tracker.add_text("config", json.dumps(flat, default=str))
                return
            except Exception as err:
                logger.warning(f"tensorboard tracker unavailable ({type(err).__name__}); falling back to jsonl")
                kind = "jsonl"
        if kind == "jsonl":
            d = logging_dir or "logs"
            os.makedirs(d, exist_ok=True)
            self._jsonl = open(os.path.join(d, run_name.replace("/", "_").replace(":", "_") + ".jsonl"), "a")
            self._tracker_kind = "jsonl"
            return
        raise ValueError(f"Only supported trackers are `wandb`, `tensorboard` and `jsonl`. Got: `{kind}`. "
                         "Set `tracker` to `None` to disable tracking.")

    def log(self, stats: Dict[str, Any], step: Optional[int] = None) -> None:
        if not self.is_main_process or self._tracker_kind is None:
            return
        if self._tracker_kind == "wandb":
            self.tracker.log(stats, step=step)
            return
        scalars = {}
        for k, v in stats.items():
            try:
                scalars[k] = float(v)
            except (TypeError, ValueError):
                continue
        if self._tracker_kind == "tensorboard":
            for k, v in scalars.items():
                self.tracker.add_scalar(k, v, global_step=step)
        elif self._tracker_kind == "jsonl":
            self._jsonl.write(json.dumps({"step": step, **scalars}) + "\n")
            self._jsonl.flush()

    def end_training(self):
        if self._tracker_kind == "wandb":
            self.tracker.finish()
        elif self._tracker_kind == "tensorboard":
            self.tracker.close()
        elif self._jsonl is not None:
            self._jsonl.close()
        self._tracker_kind = None

    # ---- timing / profiling -----------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def nvtx(self, name: str):
        """NVTX range only (no events, no synchronisation): phase markers for timelines (``ncu --nvtx``, CUPTI traces)."""
        if self.cuda:
            torch.cuda.nvtx.range_push(name)
            try:
                yield
            finally:
                torch.cuda.nvtx.range_pop()
        else:
            yield

    @contextlib.contextmanager
    def phase(self, name: str, sink: Optional[Dict[str, float]] = None):
        """NVTX range + device-timed duration (seconds) stored into ``sink[name]``."""
        if self.cuda:
            torch.cuda.nvtx.range_push(name)
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            try:
                yield
            finally:
                end.record()
                torch.cuda.nvtx.range_pop()
                if sink is not None:
                    end.synchronize()
                    sink[name] = start.elapsed_time(end) / 1e3
        else:
            t0 = time.time()
            try:
                yield
            finally:
                if sink is not None:
                    sink[name] = time.time() - t0
