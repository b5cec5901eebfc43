"""Throughput of the other configurations BASELINE.json names, through the public trainer API (the headline benchmark is
``bench.py``; this harness is for the secondary rows):

    python scripts/bench_configs.py --config ilql_gptj      [--gpus N via torch.distributed.run] [--steps K --warmup W] [--tiny]
    python scripts/bench_configs.py --config llama_lora_fp8
    python scripts/bench_configs.py --config neox20b_tp4    (8 ranks: tensor-parallel 4 x data-parallel 2)
    python scripts/bench_configs.py --config randomwalks    (CPU plumbing run)

Random-init weights of the named architecture, synthetic data of the named shape; ``--tiny`` swaps in a 2-layer model of the same
family so the whole path can be exercised on a CPU.  Timing: CUDA events on the launching stream (wall clock on CPU), barrier +
synchronize on both sides, max over ranks; one JSON line from rank 0.
"""
import argparse
import json
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GPTJ_6B = dict(model_type="gptj", vocab_size=50400, n_embd=4096, n_layer=28, n_head=16, n_positions=2048, rotary_dim=64)
LLAMA2_7B = dict(model_type="llama", vocab_size=32000, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                 num_key_value_heads=32, intermediate_size=11008, max_position_embeddings=4096)
NEOX_20B = dict(model_type="gpt_neox", vocab_size=50432, hidden_size=6144, num_hidden_layers=44, num_attention_heads=64,
                intermediate_size=24576, max_position_embeddings=2048, rotary_pct=0.25)
GPT2_SMALL = dict(model_type="gpt2", vocab_size=50257, n_embd=768, n_layer=12, n_head=12, n_positions=1024)


def tiny(arch: dict) -> dict:
    out = dict(arch)
    for key, val in (("n_embd", 64), ("hidden_size", 64), ("n_layer", 2), ("num_hidden_layers", 2), ("n_head", 4),
                     ("num_attention_heads", 4), ("num_key_value_heads", 4), ("intermediate_size", 128), ("rotary_dim", 8),
                     ("vocab_size", 512)):
        if key in out:
            out[key] = val
    return out


def words(n: int, k: int, seed: int = 0):
    rng = random.Random(seed)
    vocab = ["the", "movie", "was", "really", "quite", "film", "thought", "this", "plot", "acting", "felt", "very", "good", "bad"]
    return [" ".join(rng.choice(vocab) for _ in range(k)) for _ in range(n)]


def build(name: str, small: bool, tmp: str):
    """→ ``(trainer, one_step() -> n_samples, description)``"""
    from trlx_b200.data.default_configs import default_ilql_config, default_ppo_config
    from trlx_b200.pipeline import MiniBatchIterator
    from trlx_b200.pipeline.offline_pipeline import PromptPipeline
    from trlx_b200.utils import set_seed
    from trlx_b200.utils.loading import get_trainer

    common = dict(tracker=None, checkpoint_dir=tmp, checkpoint_interval=10 ** 9, eval_interval=10 ** 9, total_steps=10 ** 9)
    tok = "toy://bpe?vocab=512" if small else None

    def online(cfg, prompts):
        set_seed(cfg.train.seed, cfg.train.parallel)  # as trlx.train(): model-parallel peers share one RNG stream
        trainer = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=lambda samples, **kw: [float(len(s)) for s in samples],
                                                 **cfg.train.trainer_kwargs)
        budget = cfg.train.seq_length - cfg.method.gen_kwargs["max_new_tokens"]
        trainer.add_prompt_pipeline(PromptPipeline(prompts, budget, trainer.tokenizer))
        trainer.add_eval_pipeline(PromptPipeline(prompts[:4], budget, trainer.tokenizer))
        trainer.n_inner_epochs = cfg.method.ppo_epochs

        def step():
            trainer.store.clear_history()
            trainer.make_experience(cfg.method.num_rollouts, trainer.iter_count)
            for _ in range(trainer.n_inner_epochs):
                for mb in MiniBatchIterator(trainer.create_train_dataloader(), trainer.mb_size, trainer.num_mb):
                    trainer.train_step(mb)
            return cfg.method.num_rollouts
        return trainer, step

    if name == "randomwalks":
        sys.path.insert(0, ROOT)
        from examples.randomwalks import online_task
        from examples.randomwalks.ppo_randomwalks import default_config

        cfg = default_config.evolve(train=common)
        task = online_task(cfg.train.seed)
        trainer, step = online(cfg, task["prompts"])
        trainer.reward_fn = task["reward_fn"]
        return trainer, step, "PPO on random walks (GPT-2-small-class toy model)"
    if name == "llama_lora_fp8":
        arch = tiny(LLAMA2_7B) if small else LLAMA2_7B
        seq, new = (64, 16) if small else (2048, 256)
        cfg = default_ppo_config().evolve(
            train=dict(seq_length=seq, batch_size=4 if small else 16, parallel=dict(rollout_dtype="fp8"), **common),
            model=dict(model_path=arch, num_layers_unfrozen=2,
                       peft_config=dict(peft_type="LORA", task_type="CAUSAL_LM", r=8, lora_alpha=32, lora_dropout=0.0)),
            tokenizer=dict(tokenizer_path=tok or "toy://bpe?vocab=32000"),
            method=dict(num_rollouts=8 if small else 64, chunk_size=4 if small else 32,
                        gen_kwargs=dict(max_new_tokens=new, _rollout_dtype="fp8")))
        trainer, step = online(cfg, words(256, 8))
        return trainer, step, f"PPO, Llama-2-7B-shaped{' (tiny)' if small else ''}, LoRA r=8, fp8 rollout, seq {seq}"
    if name == "neox20b_tp4":
        arch = tiny(NEOX_20B) if small else NEOX_20B
        world = int(os.environ.get("WORLD_SIZE", "1"))
        pp = int(os.environ.get("BENCH_PP", "1"))  # optional pipeline stages on top (world = tp x pp x dp)
        per_stage = max(world // pp, 1)
        tp = 4 if per_stage % 4 == 0 else (2 if per_stage % 2 == 0 else 1)
        tp = int(os.environ.get("BENCH_TP", tp))  # e.g. BENCH_TP=2 on 4 ranks: TP = 2 x DP = 2
        cfg = default_ppo_config().evolve(
            train=dict(seq_length=64 if small else 1024, batch_size=4 if small else 8, trainer="NeMoPPOTrainer",
                       parallel=dict(tensor_parallel=tp, pipeline_parallel=pp, sequence_parallel=tp > 1), **common),
            model=dict(model_path=arch, num_layers_unfrozen=2), tokenizer=dict(tokenizer_path=tok or "toy://bpe?vocab=50432"),
            method=dict(num_rollouts=8 if small else 32, chunk_size=4 if small else 16,
                        gen_kwargs=dict(max_new_tokens=16 if small else 64)))
        trainer, step = online(cfg, words(256, 8))
        return trainer, step, (f"PPO, GPT-NeoX-20B-shaped{' (tiny)' if small else ''}, TP={tp} x PP={pp} x "
                               f"DP={max(world // (tp * pp), 1)}")
    if name in ("neox20b_ilql_tp4", "neox20b_sft_tp4"):
        from trlx_b200.data.default_configs import default_sft_config

        arch = tiny(NEOX_20B) if small else NEOX_20B
        world = int(os.environ.get("WORLD_SIZE", "1"))
        pp = int(os.environ.get("BENCH_PP", "1"))
        per_stage = max(world // pp, 1)
        tp = 4 if per_stage % 4 == 0 else (2 if per_stage % 2 == 0 else 1)
        ilql = name == "neox20b_ilql_tp4"
        base = default_ilql_config() if ilql else default_sft_config()
        cfg = base.evolve(
            train=dict(seq_length=64, batch_size=8 if small else 32, trainer="NeMoILQLTrainer" if ilql else "NeMoSFTTrainer",
                       parallel=dict(tensor_parallel=tp, pipeline_parallel=pp, sequence_parallel=tp > 1), **common),
            model=dict(model_path=arch), tokenizer=dict(tokenizer_path=tok or "toy://bpe?vocab=50432"))
        set_seed(cfg.train.seed, cfg.train.parallel)
        trainer = get_trainer(cfg.train.trainer)(config=cfg, **cfg.train.trainer_kwargs)
        texts = words(512, 12)
        if ilql:
            trainer.make_experience([[t[: len(t) // 2], t[len(t) // 2:]] for t in texts], [float(len(t) % 7) for t in texts],
                                    cfg.train.seq_length)
        else:
            trainer.make_experience(texts, cfg.train.seq_length)
        trainer.add_eval_pipeline(PromptPipeline(texts[:4], 32, trainer.tokenizer))
        trainer.prepare_learning()
        state = {"it": None}

        def step():
            if state["it"] is None:
                state["it"] = iter(MiniBatchIterator(trainer.create_train_dataloader(), trainer.mb_size, trainer.num_mb))
            try:
                mb = next(state["it"])
            except StopIteration:
                state["it"] = iter(MiniBatchIterator(trainer.create_train_dataloader(), trainer.mb_size, trainer.num_mb))
                mb = next(state["it"])
            trainer.train_step(mb)
            return cfg.train.batch_size
        kind = "ILQL" if ilql else "SFT"
        return trainer, step, f"{kind}, GPT-NeoX-20B-shaped{' (tiny)' if small else ''}, TP={tp} x PP={pp} x DP={max(world // (tp * pp), 1)}"
    if name == "ilql_gptj":
        arch = tiny(GPTJ_6B) if small else GPTJ_6B
        cfg = default_ilql_config().evolve(
            train=dict(seq_length=64, batch_size=8 if small else 32, trainer="AccelerateILQLTrainer",
                       parallel=dict(zero_stage=2), **common),
            model=dict(model_path=arch), tokenizer=dict(tokenizer_path=tok or "toy://bpe?vocab=50400"))
        set_seed(cfg.train.seed, cfg.train.parallel)
        trainer = get_trainer(cfg.train.trainer)(config=cfg, **cfg.train.trainer_kwargs)
        texts = words(512, 12)
        trainer.make_experience([[t[: len(t) // 2], t[len(t) // 2:]] for t in texts], [float(len(t) % 7) for t in texts], cfg.train.seq_length)
        trainer.add_eval_pipeline(PromptPipeline(texts[:4], 32, trainer.tokenizer))
        trainer.prepare_learning()
        loader = {"it": None}

        def step():
            if loader["it"] is None:
                loader["it"] = iter(MiniBatchIterator(trainer.create_train_dataloader(), trainer.mb_size, trainer.num_mb))
            try:
                mb = next(loader["it"])
            except StopIteration:
                loader["it"] = iter(MiniBatchIterator(trainer.create_train_dataloader(), trainer.mb_size, trainer.num_mb))
                mb = next(loader["it"])
            trainer.train_step(mb)
            return cfg.train.batch_size
        return trainer, step, f"ILQL, GPT-J-6B-shaped{' (tiny)' if small else ''}, sharded optimizer state (zero_stage 2)"
    raise SystemExit(f"unknown --config {name}")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", required=True, choices=["randomwalks", "ilql_gptj", "llama_lora_fp8", "neox20b_tp4", "neox20b_ilql_tp4", "neox20b_sft_tp4"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tiny", action="store_true", help="2-layer model of the same family (CPU / plumbing runs)")
    args = ap.parse_args(argv)
    if os.environ.get("BENCH_HANG_DUMP"):  # debugging aid: dump every thread's stack if the run is still going after N seconds
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["BENCH_HANG_DUMP"]), exit=True)
    import torch
    import torch.distributed as dist

    from trlx_b200.utils import logging as tlog

    tlog.set_verbosity(tlog.ERROR)
    tlog.disable_progress_bar()
    tmp = tempfile.mkdtemp(prefix="trlx_b200_cfgbench_")
    trainer, step, what = build(args.config, args.tiny, tmp)
    cuda = trainer.runtime.cuda
    world = trainer.runtime.world_size

    def barrier():
        if dist.is_initialized():
            dist.barrier()
        if cuda:
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    samples = 0
    if cuda:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        samples += step()
    if cuda:
        b.record()
    barrier()
    seconds = a.elapsed_time(b) / 1e3 if cuda else time.perf_counter() - t0
    t = torch.tensor([seconds], dtype=torch.float64, device=trainer.runtime.device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dp = max(trainer.runtime.dp_size, 1)
    eng = getattr(trainer, "_engine", None) or getattr(trainer, "_ilql_engine_obj", None)
    engine = ("none (PyTorch sampler)" if eng is None else
              f"{type(eng).__name__}{' fp8' if getattr(eng, 'fp8', False) else ' bf16'}{' lora-merged' if getattr(eng, 'lora', False) else ''}"
              f"{' megakernel' if getattr(eng, 'mega', False) else ''}")
    if trainer.runtime.is_main_process:
        print(json.dumps({"config": args.config, "what": what, "rollout_engine": engine, "zero3": getattr(trainer, "zero3", None) is not None,
                          "metric": "samples_per_sec", "value": round(samples * dp / t.item(), 3),
                          "unit": "samples/s", "n_gpus": world if cuda else 0, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": round(t.item() / args.steps * 1e3, 3), "tiny": bool(args.tiny),
                          "dtype": "bf16" if cuda else "fp32 (cpu)", "data": "synthetic"}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
