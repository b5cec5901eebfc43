#!/usr/bin/env python
"""Reference arm of ``bench.py``: the UNMODIFIED reference (``baseline/_ref``, ``pip install --no-deps --target``) driven
through its own public API — ``trlx.train(reward_fn=…, prompts=…, config=default_ppo_config())`` → stock
``AcceleratePPOTrainer.make_experience`` / ``learn`` — on the same recipe, synthetic prompts and random-init GPT-2 124M as
``bench.py`` times for this framework.  Nothing from ``trlx_b200`` is on the timed path: the only uses of this repo are the
*data preparation* (a synthetic 50257-entry BPE tokenizer saved as a plain HF tokenizer directory) and ``bench.ClockSampler``.

Third-party packages that are not installable offline are replaced by the stand-ins in ``baseline/shims`` (see its README);
``transformers_compat`` adapts three transformers-5 API changes without touching the reference's code.

One "step" = one epoch of the reference's ``learn()``: ``ppo_epochs × (num_rollouts / batch_size)`` optimizer steps on the
current store, then ``post_epoch_callback`` → ``make_experience(num_rollouts)``.  Timing is taken by a hook at the end of
``post_epoch_callback`` (instrumentation only): CUDA events + barrier for the device-timed run, wall clock for a second run
of the same length (its inputs come from the host dataloader every chunk and its statistics are read back with ``.item()``).
"""
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def prepare_assets(tmp, tiny=False):
    """Random-init GPT-2 checkpoint + tokenizer directory (no network): data preparation, outside the timed path."""
    import torch
    import transformers

    sys.path.insert(0, ROOT)
    from trlx_b200.utils.tokenizer import build_toy_tokenizer

    vocab = 1024 if tiny else 50257
    tok = build_toy_tokenizer(f"toy://bpe?vocab={vocab}")
    tok_dir = os.path.join(tmp, "tokenizer")
    tok.save_pretrained(tok_dir)
    cfg = transformers.GPT2Config(vocab_size=vocab, bos_token_id=vocab - 1, eos_token_id=vocab - 1)
    if tiny:
        cfg = transformers.GPT2Config(vocab_size=vocab, n_embd=64, n_layer=4, n_head=2, n_positions=128,
                                      bos_token_id=vocab - 1, eos_token_id=vocab - 1)
    torch.manual_seed(1000)
    model = transformers.GPT2LMHeadModel(cfg)
    model_dir = os.path.join(tmp, "model")
    os.makedirs(model_dir, exist_ok=True)
    cfg.architectures = ["GPT2LMHeadModel"]  # what the hub checkpoint's config.json carries
    cfg.save_pretrained(model_dir)
    sd = {k: v for k, v in model.state_dict().items() if k != "lm_head.weight"}  # tied to wte
    torch.save(sd, os.path.join(model_dir, "pytorch_model.bin"))
    sys.path.remove(ROOT)
    return model_dir, tok_dir


def main(args):
    import torch
    import torch.distributed as dist
    import transformers  # noqa: F401  imported BEFORE the shims are importable: it must see "accelerate not installed"
    import transformers.modeling_utils  # noqa: F401
    import transformers.generation  # noqa: F401
    from transformers import AutoModelForCausalLM, AutoTokenizer  # noqa: F401

    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, os.path.join(HERE, "_ref"))

    import transformers_compat  # noqa: F401  (before the reference imports transformers' model classes)

    sys.path.insert(0, ROOT)
    from bench import ClockSampler
    sys.path.remove(ROOT)

    import trlx  # the reference
    assert os.path.realpath(trlx.__file__).startswith(os.path.realpath(os.path.join(HERE, "_ref"))), trlx.__file__
    from trlx.data.default_configs import default_ppo_config
    from trlx.trainer.accelerate_ppo_trainer import AcceleratePPOTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()

    # Safety net: a reference run that stalls (an untested collective pattern on a new topology, a wedged rendezvous) must not
    # hang the harness that compares the two arms.  Every rank arms the same timer; when it fires, rank 0 reports the arm as
    # unavailable — the contract bench.py documents — and all ranks leave.
    import threading

    limit = float(os.environ.get("REF_TIMEOUT", "900"))

    def give_up():
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"no result after {limit:.0f} s on {world} rank(s)"}), flush=True)
        os._exit(0)

    watchdog = threading.Timer(limit, give_up)
    watchdog.daemon = True
    watchdog.start()
    if cuda:
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl" if cuda else "gloo")
    tiny = bool(os.environ.get("REF_TINY"))
    tmp = tempfile.mkdtemp(prefix="trlx_ref_bench_")
    if rank == 0:
        model_dir, tok_dir = prepare_assets(tmp, tiny)
        paths = [model_dir, tok_dir]
    else:
        paths = [None, None]
    if world > 1:
        dist.broadcast_object_list(paths, src=0)
    model_dir, tok_dir = paths

    cfg = default_ppo_config()
    cfg.model.model_path = model_dir
    cfg.tokenizer.tokenizer_path = tok_dir
    cfg.train.tracker = None
    cfg.train.checkpoint_dir = os.path.join(tmp, "ckpts")
    cfg.train.checkpoint_interval = 10 ** 9
    cfg.train.eval_interval = 10 ** 9
    cfg.train.total_steps = 10 ** 9
    W, K = args.warmup, args.steps
    cfg.train.epochs = W + 2 * K + 1  # learn() returns (after a final save + eval) before the last epoch's callback
    if tiny:
        cfg.train.seq_length = 64
        cfg.train.batch_size = 8
        cfg.method.num_rollouts = 16
        cfg.method.chunk_size = 16
        cfg.method.gen_kwargs["max_new_tokens"] = 8

    import random

    rng = random.Random(1234)
    words = ["the", "movie", "was", "really", "quite", "film", "i", "thought", "this", "plot", "acting", "felt", "very",
             "good", "bad", "long", "story", "an", "great", "boring", "after", "watching", "director", "scenes"]
    prompts = [" ".join(rng.choice(words) for _ in range(4)) for _ in range(4096)]

    dev = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
    flush = torch.empty(256 * 1024 * 1024 if cuda else 1, dtype=torch.uint8, device=dev)
    state = {"epoch": 0, "ms": None, "wall": None, "clocks": None, "sampler": None, "t0": None}
    if cuda:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if dist.is_initialized():
            dist.barrier()
        if cuda:
            torch.cuda.synchronize()

    orig_cb = AcceleratePPOTrainer.post_epoch_callback

    def hooked(self):
        orig_cb(self)  # store.clear_history(); make_experience(num_rollouts, iter_count)
        state["epoch"] += 1
        e = state["epoch"]
        if e == W:
            state["sampler"] = ClockSampler(local_rank)
            state["sampler"].start()
            barrier()
            state["t0"] = time.perf_counter()
            if cuda:
                ev0.record()
        elif e == W + K:
            if cuda:
                ev1.record()
            barrier()
            state["ms"] = ev0.elapsed_time(ev1) if cuda else (time.perf_counter() - state["t0"]) * 1e3
            state["clocks"] = state["sampler"].stop()
            barrier()
            state["t0"] = time.perf_counter()
        elif e == W + 2 * K:
            barrier()
            state["wall"] = time.perf_counter() - state["t0"]
        if cuda:
            flush.fill_(1)  # L2 flush between iterations, as in the other arm

    AcceleratePPOTrainer.post_epoch_callback = hooked
    if W == 0:
        raise SystemExit("reference arm needs --warmup >= 1")

    trainer = trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) for s in samples], prompts=prompts,
                         eval_prompts=prompts[:8], config=cfg)

    ms = torch.tensor([state["ms"]], dtype=torch.float64, device=dev)
    wall = torch.tensor([state["wall"]], dtype=torch.float64, device=dev)
    if dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / K
    m = cfg.method
    samples = m.num_rollouts * world
    r = m.gen_kwargs["max_new_tokens"]
    q = 8
    chunks = (m.num_rollouts + m.chunk_size - 1) // m.chunk_size
    opt_steps = m.ppo_epochs * ((m.num_rollouts + cfg.train.batch_size - 1) // cfg.train.batch_size)
    out = {
        "impl": "reference", "metric": "ppo_samples_per_sec", "value": round(samples / (ms_per_step / 1e3), 2),
        "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_per_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 autocast (fp32 master weights, Accelerate mixed_precision=bf16)", "data": "synthetic",
        "config": {"model": "gpt2-124M (random init, L12 H768 V50257)", "global_batch": cfg.train.batch_size * world,
                   "seq_len": cfg.train.seq_length, "parallelism": f"dp{world}", "num_rollouts_per_gpu": m.num_rollouts,
                   "chunk_size": m.chunk_size, "ppo_epochs": m.ppo_epochs, "max_new_tokens": r,
                   "num_layers_unfrozen": cfg.model.num_layers_unfrozen, "optimizer_steps_per_step": opt_steps,
                   "l2": "flushed (256 MiB write) before every timed iteration",
                   "reference": f"trlx {getattr(trlx, '__version__', '0.7.0')} from baseline/_ref, transformers "
                                f"{__import__('transformers').__version__} via baseline/shims/transformers_compat"},
        "clocks": state["clocks"],
        "e2e": {"value": round(samples / (wall.item() / K), 2), "unit": "samples/s",
                "h2d_bytes_per_step": int(chunks * m.chunk_size * q * 8 * 2 + opt_steps * cfg.train.batch_size * (q + 4 * r) * 4),
                "d2h_bytes_per_step": int(chunks * m.chunk_size * (q + r) * 8 * 3 + opt_steps * 24 * 4)},
        "gpu_launches": 0,
    }
    if rank == 0:
        print(json.dumps(out), flush=True)
    sys.stdout.flush()
    if world > 1:
        barrier()
        os._exit(0)
    return 0


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    sys.exit(main(ap.parse_args()))
