#!/usr/bin/env python
"""Headline benchmark: PPO samples/s on the reference's default PPO recipe (BASELINE.json config 2).

``ppo_sentiments`` shape from ``trlx/data/default_configs.py:17-59`` (reference): GPT-2 124M (L12 H768 V50257),
``num_layers_unfrozen=2``, batch 32/rank, ``num_rollouts=128``, ``chunk_size=128``, ``ppo_epochs=4``,
``max_new_tokens=40``, AdamW lr 3e-5; prompts = 4-word review openers (synthetic), ``reward_fn = len``; random-init
weights and a synthetic 50257-entry BPE tokenizer (no network).

One *step* = one full PPO iteration exactly as ``trainer.learn()`` executes it per epoch:
``make_experience(128 rollouts per rank)`` (generate → score → KL rewards → store) followed by ``ppo_epochs`` passes of
4 minibatch updates each (16 optimizer steps incl. gradient sync) and the KL-controller update.
``value`` = 128·N·K / device time (CUDA events, max over ranks);  ``e2e`` = the same K iterations timed by wall clock
around the public trainer calls, which each iteration copy the prompt batch host→device from pinned memory and read the
sampled tokens and the loss statistics back device→host.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 8 --steps 5 --warmup 3
    python bench.py --impl reference      # the unmodified reference from baseline/_ref via baseline/run_reference.py
    python bench.py --impl eager          # this framework with every custom kernel disabled (PyTorch eager + NCCL)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--model", default="gpt2")
    return ap.parse_args()


BASELINE_PUBLISHED = None  # the reference publishes no throughput number


def reference_arm(args):
    """Run the UNMODIFIED reference (``baseline/_ref``) through its own public API — ``trlx.train()`` → stock
    ``AcceleratePPOTrainer.make_experience`` / ``learn`` — on the same recipe (see ``baseline/run_reference.py``).  Runs in
    this process' place (``exec``) so none of this framework's modules, kernels or engine are loaded on that path."""
    runner = os.path.join(ROOT, "baseline", "run_reference.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "trlx")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is not installed (pip install --no-deps "
                          "--target baseline/_ref <copy of /root/reference>)"}))
        return 0
    os.execv(sys.executable, [sys.executable, runner, "--gpus", str(args.gpus), "--steps", str(args.steps),
                              "--warmup", str(args.warmup)])


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML in a background thread
    (first sample immediately, then every 50 ms, last sample at stop) — `nvidia-smi -lms` needs most of a second to start on an
    8-GPU box and missed short runs entirely; it remains the fallback when the NVML bindings are unavailable."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None
        self.thread, self.stop_flag, self.sm, self.mx, self.reasons = None, None, [], [], set()

    def _nvml_sample(self, pynvml, handle):
        try:
            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)))
            try:
                bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
            except Exception:
                bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
            for name, bit in self.REASON_BITS.items():
                if bits & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        try:
            import threading

            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.stop_flag = threading.Event()
            self._nvml_sample(pynvml, handle)

            def loop():
                while not self.stop_flag.wait(0.05):
                    self._nvml_sample(pynvml, handle)
                self._nvml_sample(pynvml, handle)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def build_trainer(args, tmp):
    import torch

    import trlx_b200  # noqa: F401
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.pipeline.offline_pipeline import PromptPipeline
    from trlx_b200.utils import set_seed
    from trlx_b200.utils.loading import get_trainer

    cfg = default_ppo_config()
    cfg.model.model_path = args.model  # preset → random-init GPT-2 124M architecture (no hub access)
    cfg.tokenizer.tokenizer_path = "toy://bpe?vocab=50257"
    cfg.train.tracker = None
    cfg.train.checkpoint_dir = tmp
    cfg.train.checkpoint_interval = 10 ** 9
    cfg.train.eval_interval = 10 ** 9
    set_seed(cfg.train.seed)
    trainer = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=lambda samples, **kw: [float(len(s)) for s in samples],
                                             metric_fn=None, stop_sequences=[])
    # synthetic "first four words of a review" prompts (examples/ppo_sentiments.py:44-45 of the reference)
    import random

    rng = random.Random(1234)
    words = ["the", "movie", "was", "really", "quite", "film", "i", "thought", "this", "plot", "acting", "felt", "very",
             "good", "bad", "long", "story", "an", "great", "boring", "after", "watching", "director", "scenes"]
    prompts = [" ".join(rng.choice(words) for _ in range(4)) for _ in range(4096)]
    max_prompt_length = cfg.train.seq_length - cfg.method.gen_kwargs["max_new_tokens"]
    trainer.add_prompt_pipeline(PromptPipeline(prompts, max_prompt_length, trainer.tokenizer))
    trainer.add_eval_pipeline(PromptPipeline(prompts[:8], max_prompt_length, trainer.tokenizer))
    trainer.n_inner_epochs = cfg.method.ppo_epochs
    trainer.total_steps = 10 ** 9
    trainer.config.train.total_steps = 10 ** 9
    return trainer, cfg


def _optimizer_note(trainer) -> str:
    """Which gradient-reduction path ran (for the record next to a multi-GPU number)."""
    fg = next((g for g in (getattr(trainer.opt, "_flat", None) or []) if g is not None), None)
    if fg is None:
        return "torch fallback"
    if fg.world == 1:
        return "fused AdamW (single rank)"
    return (f"fused RS+AdamW+AG over NVLink, {len(fg.buckets)} buckets, "
            f"{'overlapped with backward' if getattr(trainer.opt, 'can_overlap', False) else 'after backward'}, "
            f"{'NVLS multimem' if getattr(fg, 'mc_grad', 0) else 'P2P ld/st'}")


def ppo_iteration(trainer):
    """Exactly the per-epoch body of ``AccelerateRLTrainer.learn`` for PPO (without eval / checkpoint IO)."""
    from trlx_b200.pipeline import MiniBatchIterator
    from trlx_b200.trainer.accelerate_base_trainer import PendingStats

    trainer.store.clear_history()
    trainer.make_experience(trainer.config.method.num_rollouts, trainer.iter_count)
    last, pending = None, None
    for _ in range(trainer.n_inner_epochs):
        loader = trainer.create_train_dataloader()
        for minibatch in MiniBatchIterator(loader, trainer.mb_size, trainer.num_mb):
            stats = trainer.train_step(minibatch)
            handle = PendingStats(stats)  # device→host read of the step's loss / statistics, queued behind the step ...
            if pending is not None:
                last = pending.get()      # ... and consumed one step late, as in `learn()`
            pending = handle
        trainer.post_backward_callback()
    if pending is not None:
        last = pending.get()              # every step's statistics reach the host inside the iteration
    return last


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    if args.impl == "eager":
        os.environ["TRLX_B200_DISABLE_KERNELS"] = "1"
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from trlx_b200 import ops
    from trlx_b200.utils import logging as tlog

    tlog.set_verbosity(tlog.ERROR)
    tlog.disable_progress_bar()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run", file=sys.stderr)
    tmp = tempfile.mkdtemp(prefix="trlx_b200_bench_")
    trainer, cfg = build_trainer(args, tmp)
    dev = trainer.runtime.device
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ppo_iteration(trainer)
    barrier()

    # ---- device-timed run ------------------------------------------------------------------------------------------
    sampler = ClockSampler(trainer.runtime.local_rank)
    sampler.start()
    ops.reset_launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations
        ppo_iteration(trainer)
    end.record()
    barrier()
    launches = ops.launch_count()
    clocks = sampler.stop()
    ms = torch.tensor([start.elapsed_time(end)], device=dev, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps

    # ---- end-to-end (wall clock around the public trainer calls; H2D prompt copies + D2H reads included) -----------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        ppo_iteration(trainer)
    barrier()
    wall = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    e2e_s_per_step = wall.item() / args.steps

    m = cfg.method
    samples_per_step = m.num_rollouts * world
    chunk, q = m.chunk_size, getattr(trainer, "_last_prompt_width", 8)
    r = m.gen_kwargs["max_new_tokens"]
    chunks_per_step = (m.num_rollouts + chunk - 1) // chunk
    h2d = chunks_per_step * chunk * q * 8 * 2  # int64 input_ids + attention_mask from pinned host memory
    opt_steps = m.ppo_epochs * ((m.num_rollouts + cfg.train.batch_size - 1) // cfg.train.batch_size)
    d2h = chunks_per_step * chunk * (q + r) * 8 + opt_steps * 24 * 4 + chunks_per_step * 4 * 8
    value = samples_per_step / (ms_per_step / 1e3)
    out = {
        "metric": "ppo_samples_per_sec", "value": round(value, 2), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.impl == "ours" else "bf16 (eager)", "data": "synthetic",
        "impl": args.impl,
        "config": {"model": "gpt2-124M (random init, L12 H768 V50257)", "global_batch": cfg.train.batch_size * world,
                   "seq_len": cfg.train.seq_length, "parallelism": f"dp{world}", "num_rollouts_per_gpu": m.num_rollouts,
                   "chunk_size": m.chunk_size, "ppo_epochs": m.ppo_epochs, "max_new_tokens": r, "num_layers_unfrozen": 2,
                   "optimizer_steps_per_step": opt_steps, "l2": "flushed (256 MiB write) before every timed iteration",
                   "learn_tokens_per_sec": round(samples_per_step * m.ppo_epochs * (q + r) / (ms_per_step / 1e3), 1),
                   "optimizer": _optimizer_note(trainer)},
        "clocks": clocks,
        "e2e": {"value": round(samples_per_step / e2e_s_per_step, 2), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
    }
    if os.environ.get("BENCH_BREAKDOWN") and rank == 0 and world == 1:  # rank-0-only re-run: would desynchronise collectives
        # coarse per-phase wall-clock (synchronised) for one more iteration — diagnostic only, not part of the metric
        import collections

        acc = collections.OrderedDict()

        def timed(name, fn, *a, **k):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t) * 1e3
            return r

        from trlx_b200.pipeline import MiniBatchIterator

        eng = trainer._engine
        if eng is not None:
            orig_prefill, orig_rollout = eng._prefill, eng.rollout
            eng._prefill = lambda *a, **k: timed("engine.prefill", orig_prefill, *a, **k)
            eng.rollout = lambda *a, **k: timed("engine.rollout(total)", orig_rollout, *a, **k)
        orig_decode = trainer.decode
        trainer.decode = lambda *a, **k: timed("decode(strings)", orig_decode, *a, **k)
        trainer.store.clear_history()
        timed("make_experience(total)", trainer.make_experience, cfg.method.num_rollouts, 0)
        for _ in range(trainer.n_inner_epochs):
            loader = timed("create_loader", trainer.create_train_dataloader)
            for minibatch in MiniBatchIterator(loader, trainer.mb_size, trainer.num_mb):
                timed("train_step(16x)", trainer.train_step, minibatch)
        acc["train_graphs_captured"] = len(getattr(trainer, "_graphed_steps", {}) or {})
        print("BREAKDOWN_MS " + json.dumps({k: round(v, 2) for k, v in acc.items()}), file=sys.stderr)
    if os.environ.get("BENCH_PROFILE") and rank == 0:
        # per-kernel device time of the two phases (torch.profiler/CUPTI; diagnostic only — never a bench number)
        from torch.profiler import ProfilerActivity, profile

        from trlx_b200.pipeline import MiniBatchIterator

        if not os.environ.get("BENCH_PROFILE_GRAPH"):
            os.environ["TRLX_B200_TRAIN_GRAPH"] = "0"  # the eager step also shows which autograd node launched a kernel
            trainer._graphed_steps = {}
        os.makedirs("gpurun_out", exist_ok=True)
        for phase in ("train", "rollout"):
            trainer.store.clear_history()
            if phase == "train":
                trainer.make_experience(cfg.method.num_rollouts, 0)
                loader = trainer.create_train_dataloader()
                batches = list(MiniBatchIterator(loader, trainer.mb_size, trainer.num_mb))
                trainer.train_step(batches[0])
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                if phase == "train":
                    for mb in batches[1:5]:
                        trainer.train_step(mb)
                else:
                    trainer.make_experience(cfg.method.num_rollouts, 0)
                torch.cuda.synchronize()
            table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70)
            suffix = "_graph" if os.environ.get("BENCH_PROFILE_GRAPH") else ""
            with open(f"gpurun_out/profile_{phase}{suffix}.txt", "w") as fh:
                fh.write(table)
            if os.environ.get("BENCH_PROFILE_TRACE"):
                prof.export_chrome_trace(f"gpurun_out/trace_{phase}{suffix}.json")
    if rank == 0:
        print(json.dumps(out), flush=True)
    # Teardown order matters with captured CUDA graphs that reference NVLink symmetric memory: drop the graphs first, drain
    # the device, then leave the process group.  The interpreter is exited directly afterwards — the measurement is done and
    # NCCL/symmetric-memory destructors racing at interpreter shutdown have hung multi-rank runs (run20).
    try:
        getattr(trainer, "_graphed_steps", {}).clear()
        if getattr(trainer, "_engine", None) is not None:
            trainer._engine._state = None
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
    finally:
        sys.stdout.flush()
        sys.stderr.flush()
        if dist.is_initialized() and world > 1:
            os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
