"""Thin Accelerator over torch.distributed / DDP with bf16 autocast (stand-in for HF Accelerate, see ../README.md)."""
import contextlib
import os
import types

import torch
import torch.distributed as dist

from . import state  # noqa: F401


class _AutocastModule(torch.nn.Module):
    """What Accelerate does for mixed_precision=bf16: wrap forward in autocast (outputs kept in fp32)."""

    def __init__(self, module, dtype):
        super().__init__()
        self.module = module
        self._dtype = dtype

    def forward(self, *a, **k):
        with torch.autocast("cuda", dtype=self._dtype):
            return self.module(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)


class _EngineLikeDDP(torch.nn.parallel.DistributedDataParallel):
    """DDP whose missing attributes resolve on the wrapped module.  The reference reads ``model.peft_type`` / ``model.frozen_head``
    and calls ``model.generate`` / ``model.forward_hydra`` on whatever ``accelerator.prepare`` returned
    (``trlx/trainer/accelerate_ppo_trainer.py:70-80``); its multi-GPU default is DeepSpeed (``configs/accelerate/zero2-bf16.yaml``),
    whose engine forwards attribute access exactly like this — plain ``torch`` DDP does not, and the trainer would not construct."""

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)


class Accelerator:
    def __init__(self, log_with=None, project_dir=None, **kw):
        self.rank = int(os.environ.get("RANK", 0))
        self.num_processes = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_rank)
            self.device = torch.device("cuda", self.local_rank)
        else:
            self.device = torch.device("cpu")
        if self.num_processes > 1 and not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
        self.mixed_precision = os.environ.get("ACCELERATE_MIXED_PRECISION", "bf16" if torch.cuda.is_available() else "no")
        self.state = types.SimpleNamespace(deepspeed_plugin=None)
        self.gradient_state = state.GradientState()
        self.gradient_accumulation_steps = 1
        self._pre_hooks = []

    is_main_process = property(lambda self: self.rank == 0)
    is_local_main_process = property(lambda self: self.local_rank == 0)

    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o = o.to(self.device)
                if self.num_processes > 1 and any(p.requires_grad for p in o.parameters()):
                    o = _EngineLikeDDP(
                        o, device_ids=[self.local_rank] if self.device.type == "cuda" else None, find_unused_parameters=True)
                if self.mixed_precision == "bf16" and self.device.type == "cuda":
                    inner = o
                    fwd = inner.forward

                    def autocast_forward(*a, __fwd=fwd, **k):
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            return __fwd(*a, **k)

                    inner.forward = autocast_forward
            elif isinstance(o, torch.utils.data.DataLoader):
                o = self.prepare_data_loader(o)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def prepare_data_loader(self, loader):
        if self.num_processes == 1:
            return loader
        from torch.utils.data import DataLoader
        from torch.utils.data.distributed import DistributedSampler

        shuffle = isinstance(getattr(loader, "sampler", None), torch.utils.data.RandomSampler)
        sampler = DistributedSampler(loader.dataset, num_replicas=self.num_processes, rank=self.rank, shuffle=shuffle)
        return DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=sampler, collate_fn=loader.collate_fn,
                          drop_last=loader.drop_last)

    def unwrap_model(self, model):
        while hasattr(model, "module") and isinstance(model, torch.nn.parallel.DistributedDataParallel):
            model = model.module
        return model

    def backward(self, loss):
        loss.backward()

    @contextlib.contextmanager
    def no_sync(self, model):
        ctx = model.no_sync() if isinstance(model, torch.nn.parallel.DistributedDataParallel) else contextlib.nullcontext()
        with ctx:
            yield

    @contextlib.contextmanager
    def main_process_first(self):
        yield

    @contextlib.contextmanager
    def accumulate(self, model):
        yield

    def init_trackers(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    def wait_for_everyone(self):
        if dist.is_initialized():
            dist.barrier()

    def gather(self, t):
        """All-gather along dim 0; lists / tuples / dicts of tensors are gathered leaf by leaf (Accelerate's behaviour)."""
        if self.num_processes == 1:
            return t
        if isinstance(t, (list, tuple)):
            return type(t)(self.gather(x) for x in t)
        if isinstance(t, dict):
            return {k: self.gather(v) for k, v in t.items()}
        if not isinstance(t, torch.Tensor):
            return t
        if t.dim() == 0:
            t = t[None]
        out = [torch.empty_like(t) for _ in range(self.num_processes)]
        dist.all_gather(out, t.contiguous())
        return torch.cat(out, 0)

    gather_for_metrics = gather

    def pad_across_processes(self, tensors, dim=0, pad_index=0, pad_first=False):
        single = not isinstance(tensors, (list, tuple))
        ts = [tensors] if single else list(tensors)
        if self.num_processes > 1:
            res = []
            for t in ts:
                if t.dim() <= dim:
                    res.append(t)
                    continue
                size = torch.tensor([t.shape[dim]], device=t.device)
                dist.all_reduce(size, op=dist.ReduceOp.MAX)
                w = int(size.item())
                if w > t.shape[dim]:
                    shape = list(t.shape)
                    shape[dim] = w - t.shape[dim]
                    pad = t.new_full(shape, pad_index)
                    t = torch.cat([pad, t] if pad_first else [t, pad], dim)
                res.append(t)
            ts = res
        return ts[0] if single else ts

    def get_state_dict(self, model, unwrap=True):
        return self.unwrap_model(model).state_dict()

    def save(self, obj, path):
        if self.is_main_process:
            torch.save(obj, path)

    def save_state(self, output_dir=None, **k):
        os.makedirs(output_dir, exist_ok=True)

    def load_state(self, input_dir=None, **k):
        pass

    def register_load_state_pre_hook(self, hook):
        self._pre_hooks.append(hook)
        return types.SimpleNamespace(remove=lambda: None)

    def free_memory(self):
        pass
