"""Multi-process (gloo, CPU) tests of the distributed *logic*: groups, global statistics, gradient sync of the
optimizer front-end, tensor/sequence-parallel equivalence.  The NVLink kernels themselves are covered by GPU tests."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, out_dir, args):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        result = fn(rank, world, *args)
        torch.save(result, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def run_distributed(fn, world=2, args=()):
    out_dir = tempfile.mkdtemp()
    mp.spawn(_worker, args=(world, _free_port(), fn, out_dir, args), nprocs=world, join=True)
    return [torch.load(os.path.join(out_dir, f"r{r}.pt"), weights_only=False) for r in range(world)]


# ---- statistics -------------------------------------------------------------------------------------------------------------
def _stats_job(rank, world):
    from trlx_b200.utils.modeling import RunningMoments, get_global_statistics, whiten

    torch.manual_seed(rank)
    x = torch.randn(5 + 3 * rank, 4) * (1 + rank) + rank
    mean, var, n = get_global_statistics(x)
    rm = RunningMoments()
    rm.update(x)
    return dict(x=x, mean=mean, var=var, n=n, white=whiten(x), rm_mean=rm.mean, rm_std=rm.std)


def test_global_statistics_and_whiten():
    res = run_distributed(_stats_job, 2)
    allx = torch.cat([r["x"].reshape(-1) for r in res])
    for r in res:
        assert torch.isclose(r["mean"], allx.mean(), atol=1e-5) and torch.isclose(r["var"], allx.var(unbiased=False), rtol=1e-4)
        assert int(r["n"]) == allx.numel()
        torch.testing.assert_close(r["white"], (r["x"] - allx.mean()) * torch.rsqrt(allx.var(unbiased=False) + 1e-8), atol=1e-4, rtol=1e-4)
        assert abs(r["rm_mean"] - allx.mean().item()) < 1e-4 and abs(r["rm_std"] - allx.std().item()) < 1e-3


# ---- runtime groups -----------------------------------------------------------------------------------------------------------
def _groups_job(rank, world):
    from trlx_b200.data.configs import ParallelConfig
    from trlx_b200.parallel.runtime import Runtime

    rt = Runtime(ParallelConfig(tensor_parallel=2, pipeline_parallel=1))
    t = torch.tensor([float(rank)])
    dist.all_reduce(t, group=rt.tp_group)
    d = torch.tensor([float(rank)])
    dist.all_reduce(d, group=rt.dp_group)
    gathered = rt.gather_objects({"r": rank})
    return dict(tp_rank=rt.tp_rank, dp_rank=rt.dp_rank, tp_sum=t.item(), dp_sum=d.item(), n=len(gathered), dp=rt.dp_size)


def test_runtime_dp_tp_groups():
    res = run_distributed(_groups_job, 4)
    assert [r["tp_rank"] for r in res] == [0, 1, 0, 1] and [r["dp_rank"] for r in res] == [0, 0, 1, 1]
    assert [r["tp_sum"] for r in res] == [1.0, 1.0, 5.0, 5.0]  # TP groups {0,1} {2,3}
    assert [r["dp_sum"] for r in res] == [2.0, 4.0, 2.0, 4.0]  # DP groups {0,2} {1,3}
    assert all(r["n"] == 4 and r["dp"] == 2 for r in res)


# ---- optimizer gradient sync ----------------------------------------------------------------------------------------------------
def _optim_job(rank, world):
    from trlx_b200.parallel.optim import FusedAdamW

    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(6, 5))
    opt = FusedAdamW([w], lr=0.1, betas=(0.9, 0.95), weight_decay=0.0)
    torch.manual_seed(100 + rank)
    x = torch.randn(4, 5)
    for _ in range(2):
        opt.zero_grad()
        (w @ x.t()).pow(2).mean().backward()
        opt.step()
    return dict(w=w.detach().clone(), x=x)


def test_optimizer_averages_gradients_across_ranks():
    res = run_distributed(_optim_job, 2)
    torch.testing.assert_close(res[0]["w"], res[1]["w"])  # replicas stay in sync
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(6, 5))
    opt = torch.optim.AdamW([w], lr=0.1, betas=(0.9, 0.95), weight_decay=0.0)
    for _ in range(2):
        opt.zero_grad()
        sum((w @ r["x"].t()).pow(2).mean() for r in res).div(2).backward()  # mean over ranks == DDP semantics
        opt.step()
    torch.testing.assert_close(res[0]["w"], w.detach(), atol=1e-6, rtol=1e-5)


# ---- tensor / sequence parallel --------------------------------------------------------------------------------------------------
def _tp_job(rank, world, family, sequence_parallel):
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx_b200.parallel.tensor_parallel import apply_tensor_parallel

    cfgs = {
        "gpt2": dict(model_type="gpt2", vocab_size=48, n_embd=32, n_layer=3, n_head=4, n_positions=32),
        "llama": dict(model_type="llama", vocab_size=48, hidden_size=32, num_hidden_layers=3, num_attention_heads=4,
                      num_key_value_heads=2, intermediate_size=64, max_position_embeddings=32),
        "bloom": dict(model_type="bloom", vocab_size=48, hidden_size=32, n_layer=2, n_head=4),
    }
    torch.manual_seed(0)
    model = AutoModelForCausalLMWithHydraValueHead.from_config(cfgs[family], num_layers_unfrozen=1).eval()
    torch.manual_seed(1)
    ids = torch.randint(0, 48, (2, 8))
    mask = torch.ones(2, 8, dtype=torch.long)
    mask[0, :2] = 0
    pos = (mask.cumsum(-1) - 1).clamp_min(0)
    ref = model(ids, mask, position_ids=pos, return_dict=True)
    ref_hydra = model.forward_hydra(ids, mask, position_ids=pos, return_dict=True).logits
    ref.logits.float().pow(2).mean().add(ref.value.pow(2).mean()).backward()
    ref_grad = model.base_model.transformer.h[-1].mlp.down.weight.grad.clone()
    model.zero_grad()

    apply_tensor_parallel(model, None, rank, world, sequence_parallel=sequence_parallel)
    out = model(ids, mask, position_ids=pos, return_dict=True)
    hydra = model.forward_hydra(ids, mask, position_ids=pos, return_dict=True).logits
    out.logits.float().pow(2).mean().add(out.value.pow(2).mean()).backward()
    shard_grad = model.base_model.transformer.h[-1].mlp.down.weight.grad.clone()
    f = ref_grad.shape[1] // world
    return dict(logit_err=(out.logits - ref.logits).abs().max().item(), value_err=(out.value - ref.value).abs().max().item(),
                hydra_err=(hydra - ref_hydra).abs().max().item(),
                grad_err=(shard_grad - ref_grad[:, rank * f:(rank + 1) * f]).abs().max().item(),
                qkv_rows=model.base_model.transformer.h[0].attn.qkv.weight.shape[0])


@pytest.mark.parametrize("family,sp", [("gpt2", False), ("gpt2", True), ("llama", False), ("llama", True), ("bloom", False)])
def test_tensor_parallel_matches_single_rank(family, sp):
    res = run_distributed(_tp_job, 2, args=(family, sp))
    for r in res:
        assert r["logit_err"] < 1e-4 and r["value_err"] < 1e-4 and r["hydra_err"] < 1e-4 and r["grad_err"] < 1e-4, r


def test_tp_state_dict_resharding_roundtrip():
    from trlx_b200.nn.arch import spec_from_hf_config
    from trlx_b200.nn.transformer import CausalLM
    from trlx_b200.parallel.tensor_parallel import shard_state_dict, unshard_state_dicts

    spec = spec_from_hf_config(dict(model_type="llama", vocab_size=48, hidden_size=32, num_hidden_layers=2, num_attention_heads=4,
                                    num_key_value_heads=2, intermediate_size=64, max_position_embeddings=32))
    sd = CausalLM(spec).state_dict()
    shards = [shard_state_dict(spec, sd, r, 2) for r in range(2)]
    assert shards[0]["transformer.h.0.attn.qkv.weight"].shape[0] == (spec.q_size + 2 * spec.kv_size) // 2
    back = unshard_state_dicts(spec, shards)
    for k, v in sd.items():
        torch.testing.assert_close(back[k], v)


# ---- pipeline parallelism -------------------------------------------------------------------------------------------------
def _pp_job(rank, world, family, tied, virtual=1, M=5):
    import copy

    from trlx_b200.nn.arch import spec_from_hf_config
    from trlx_b200.nn.transformer import CausalLM
    from trlx_b200.parallel import pipeline_parallel as pp

    cfgs = {
        "gpt2": dict(model_type="gpt2", vocab_size=61, n_embd=32, n_layer=4, n_head=4, n_positions=64),
        "llama": dict(model_type="llama", vocab_size=61, hidden_size=32, num_hidden_layers=4, num_attention_heads=4,
                      num_key_value_heads=2, intermediate_size=64, max_position_embeddings=64),
    }
    torch.manual_seed(0)
    spec = spec_from_hf_config(dict(cfgs[family], tie_word_embeddings=tied))
    assert spec.tie_word_embeddings == tied
    full = CausalLM(spec).float()
    staged = copy.deepcopy(full)
    stage = pp.apply_pipeline_parallel(staged, None, rank, world, virtual)
    assert stage.virtual == virtual
    g = torch.Generator().manual_seed(1)
    mbs = []
    for i in range(M):
        T = 6 + (i % 3)  # the activation shape changes between micro-batches
        ids = torch.randint(0, 61, (2, T), generator=g)
        mask = torch.ones_like(ids)
        mask[0, :i % 2] = 0
        mbs.append(dict(input_ids=ids, attention_mask=mask, labels=ids))

    # oracle: the unpartitioned model
    ref_losses = []
    for mb in mbs:
        out = full(**mb)
        out.loss.backward()
        ref_losses.append(out.loss.detach())
    ref_grads = {n: p.grad.clone() for n, p in full.named_parameters() if p.grad is not None}

    # (1) differentiable relay, one micro-batch at a time
    relay_losses = []
    for mb in mbs:
        out = staged(**mb)
        out.loss.backward()
        relay_losses.append(out.loss.detach())
    pp.allreduce_tied_embedding_grads(stage)
    relay_grads = {n: p.grad.clone() for n, p in staged.named_parameters() if p.grad is not None and p.numel()}
    staged.zero_grad()

    # (2) 1F1B schedule
    def loss_fn(mb):
        out = staged(**mb)
        return out.loss, {"loss": out.loss.detach()}

    stats = pp.run_schedule(stage, mbs, loss_fn, torch.device("cpu"))
    pp.allreduce_tied_embedding_grads(stage)
    sched_grads = {n: p.grad.clone() for n, p in staged.named_parameters() if p.grad is not None and p.numel()}
    shared = pp.broadcast_stats(stage, {"loss": sum(s["loss"] for s in stats) / M} if stage.last else None, torch.device("cpu"))

    # (3) inference relay with a KV cache: every stage sees the same logits as the full model
    with torch.no_grad():
        ids = mbs[0]["input_ids"]
        a = staged(input_ids=ids[:, :4], use_cache=True)
        b = staged(input_ids=ids[:, 4:5], past_key_values=a.past_key_values, use_cache=True)
        fa = full(input_ids=ids[:, :5])
    return dict(ref_losses=ref_losses, relay_losses=relay_losses, ref_grads=ref_grads, relay_grads=relay_grads,
                sched_grads=sched_grads, mean_loss=shared["loss"], step_logits=b.logits[:, -1], full_logits=fa.logits[:, -1],
                owned=(stage.lo, stage.hi))


def test_interleaved_schedule_plan():
    """Every (P, V, M): each rank runs every operation exactly once, and a round's transfers are matched by construction."""
    from trlx_b200.parallel.pipeline_parallel import interleaved_rounds

    for P in (2, 3, 4, 8):
        for V in (1, 2, 3):
            for M in (1, 3, 4, 5, 8, 16):
                rounds = interleaved_rounds(P, V, M)
                for r in range(P):
                    mine = [a[r] for a in rounds if a[r] is not None]
                    assert len(mine) == len(set(mine)) == 2 * M * V
    # activations in flight on a rank never exceed its warm-up depth + 1 (what bounds the memory of the schedule)
    for P, V, M in ((4, 2, 8), (4, 1, 8), (2, 3, 6)):
        rounds = interleaved_rounds(P, V, M)
        for r in range(P):
            live = peak = 0
            for acts in rounds:
                if acts[r] is not None:
                    live += 1 if acts[r][0] == "F" else -1
                    peak = max(peak, live)
            assert live == 0 and peak <= min(2 * (P - r - 1) + (V - 1) * P, M * V) + 1, (P, V, M, r, peak)
    # the bubble shrinks with the number of chunks: P = 4, 16 micro-batches, measured in whole-stage units
    cost = {V: len(interleaved_rounds(4, V, 16)) / V for V in (1, 2, 4)}
    assert cost[4] < cost[2] < cost[1]


@pytest.mark.parametrize("world,family,tied,virtual,M", [(2, "gpt2", True, 1, 5), (3, "llama", False, 1, 5),
                                                         (2, "gpt2", True, 2, 4), (2, "llama", False, 2, 5)])
def test_pipeline_parallel_matches_single_rank(world, family, tied, virtual, M):
    res = run_distributed(_pp_job, world, (family, tied, virtual, M))
    ref = res[0]["ref_grads"]
    seen = set()
    for r in res:
        for a, b in zip(r["relay_losses"], r["ref_losses"]):  # every stage reports the true loss in relay mode
            torch.testing.assert_close(a, b, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(r["mean_loss"], sum(r["ref_losses"]) / len(r["ref_losses"]), atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(r["step_logits"], r["full_logits"], atol=1e-4, rtol=1e-4)
        for kind in ("relay_grads", "sched_grads"):
            for n, g in r[kind].items():
                torch.testing.assert_close(g, ref[n], atol=2e-5, rtol=1e-4, msg=lambda m: f"{kind} {n}: {m}")
                seen.add(n)
    assert seen == set(ref), f"parameters without a gradient on any stage: {set(ref) - seen}"


def _pp_trainer_job(rank, world, kind, tmp, virtual=1):
    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_ppo_config, default_sft_config

    gpt2 = dict(model_type="gpt2", vocab_size=257, n_embd=32, n_layer=4, n_head=2, n_positions=64, eos_token_id=256, bos_token_id=256)
    common = dict(seq_length=16, batch_size=4, minibatch_size=1, total_steps=3, epochs=4, checkpoint_interval=100, eval_interval=3,
                  tracker=None, checkpoint_dir=tmp, seed=3, parallel=dict(pipeline_parallel=world, virtual_pipeline_parallel=virtual))
    prompts = ["hello a", "what is", "b", "count the a", "zz", "dog dog", "a dog", "the"]
    if kind == "ppo":
        cfg = default_ppo_config().evolve(
            train=dict(common, trainer="NeMoPPOTrainer"), model=dict(model_path=gpt2, num_layers_unfrozen=-1),
            tokenizer=dict(tokenizer_path="toy://bytes"),
            method=dict(num_rollouts=4, chunk_size=4, ppo_epochs=1, gen_kwargs=dict(max_new_tokens=5, top_k=0, top_p=1.0, do_sample=True)))
        trainer = trlx.train(reward_fn=lambda samples, **kw: [float(s.count("a")) for s in samples], prompts=prompts,
                             eval_prompts=prompts[:2], config=cfg)
    else:
        cfg = default_sft_config().evolve(
            train=dict(common, trainer="NeMoSFTTrainer"), model=dict(model_path=gpt2),
            tokenizer=dict(tokenizer_path="toy://bytes"), method=dict(gen_kwargs=dict(max_new_tokens=4, do_sample=False)))
        trainer = trlx.train(samples=[[p, " yes a"] for p in prompts], eval_prompts=prompts[:2], config=cfg)
    lm = trainer.model.base_model
    owned = [i for i, blk in enumerate(lm.transformer.h) if sum(p.numel() for p in blk.parameters()) > 0]
    return dict(iters=trainer.iter_count, owned=owned)


@pytest.mark.parametrize("kind", ["sft", "ppo"])
def test_pipeline_parallel_trainers_run(kind, tmp_path):
    res = run_distributed(_pp_trainer_job, 2, (kind, str(tmp_path)))
    assert [r["iters"] for r in res] == [3, 3]
    assert res[0]["owned"] == [0, 1] and res[1]["owned"] == [2, 3]


def test_interleaved_pipeline_ppo_trainer_runs(tmp_path):
    """``virtual_pipeline_parallel=2``: every rank owns two non-adjacent chunks, the optimizer step runs the interleaved schedule."""
    res = run_distributed(_pp_trainer_job, 2, ("ppo", str(tmp_path), 2))
    assert [r["iters"] for r in res] == [3, 3]
    assert res[0]["owned"] == [0, 2] and res[1]["owned"] == [1, 3]


def _sp_odd_length_job(rank, world, family):
    """Sequence parallelism with a length the group cannot shard (7 tokens on 2 ranks) and with a KV-cache decode: the forward
    falls back to replicated activations, gradients (incl. the block norms after the SP gradient all-reduce) stay exact."""
    from trlx_b200.models.generation import generate
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx_b200.parallel.tensor_parallel import allreduce_sequence_parallel_grads, apply_tensor_parallel

    cfgs = {"gpt2": dict(model_type="gpt2", vocab_size=48, n_embd=32, n_layer=3, n_head=4, n_positions=32),
            "llama": dict(model_type="llama", vocab_size=48, hidden_size=32, num_hidden_layers=3, num_attention_heads=4,
                          num_key_value_heads=2, intermediate_size=64, max_position_embeddings=32)}
    torch.manual_seed(0)
    model = AutoModelForCausalLMWithHydraValueHead.from_config(cfgs[family], num_layers_unfrozen=-1).eval()
    torch.manual_seed(1)
    out = {}
    refs = {}
    for T in (7, 8):  # indivisible, then divisible (sharded) — the same module must handle both back to back
        ids = torch.randint(0, 48, (2, T))
        mask = torch.ones(2, T, dtype=torch.long)
        mask[0, :2] = 0
        ref = model(ids, mask, return_dict=True)
        ref.logits.float().pow(2).mean().add(ref.value.pow(2).mean()).backward()
        norm = model.base_model.transformer.h[1].norm1.weight
        refs[T] = (ids, mask, ref.logits.detach(), norm.grad.clone(), model.base_model.transformer.h[-1].mlp.down.weight.grad.clone(),
                   model.base_model.transformer.wte.weight.grad.clone())
        model.zero_grad()
    gen_ref = generate(model.base_model, refs[7][0], attention_mask=refs[7][1], max_new_tokens=4, do_sample=False, pad_token_id=0,
                       eos_token_id=47)
    apply_tensor_parallel(model, None, rank, world, sequence_parallel=True)
    for T, (ids, mask, logits, norm_grad, down_grad, wte_grad) in refs.items():
        res = model(ids, mask, return_dict=True)
        res.logits.float().pow(2).mean().add(res.value.pow(2).mean()).backward()
        allreduce_sequence_parallel_grads(model, None)
        f = down_grad.shape[1] // world
        out[T] = dict(logits=(res.logits - logits).abs().max().item(),
                      norm=(model.base_model.transformer.h[1].norm1.weight.grad - norm_grad).abs().max().item(),
                      down=(model.base_model.transformer.h[-1].mlp.down.weight.grad - down_grad[:, rank * f:(rank + 1) * f]).abs().max().item(),
                      wte=(model.base_model.transformer.wte.weight.grad - wte_grad).abs().max().item())
        model.zero_grad()
    gen = generate(model.base_model, refs[7][0], attention_mask=refs[7][1], max_new_tokens=4, do_sample=False, pad_token_id=0,
                   eos_token_id=47)
    out["gen_equal"] = bool(torch.equal(gen, gen_ref))
    return out


@pytest.mark.parametrize("family", ["gpt2", "llama"])
def test_sequence_parallel_handles_indivisible_lengths_and_cached_decoding(family):
    for r in run_distributed(_sp_odd_length_job, 2, args=(family,)):
        for T in (7, 8):
            assert r[T]["logits"] < 1e-4 and r[T]["norm"] < 1e-5 and r[T]["down"] < 1e-5 and r[T]["wte"] < 1e-5, r
        assert r["gen_equal"], r


# ---- whole trainers on model-parallel layouts (gloo, tiny models) through scripts/bench_configs.py ----------------------------------
def _run_config_bench(world: int, env: dict) -> dict:
    import json
    import subprocess
    import sys

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "scripts/bench_configs.py", "--config", "neox20b_tp4", "--tiny", "--steps", "1",
           "--warmup", "1"]
    full_env = dict(os.environ, OMP_NUM_THREADS="1", BENCH_HANG_DUMP="400", **env)
    proc = subprocess.run(cmd, env=full_env, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stderr[-3000:]
    return json.loads(lines[-1])


def test_ppo_trainer_with_tensor_and_sequence_parallelism_end_to_end():
    """Rollouts (KV-cache decoding), scoring with the hydra branch and PPO updates on TP = 2 with sequence parallelism, for
    whatever sequence lengths the prompts produce."""
    rec = _run_config_bench(2, {})
    assert "TP=2 x PP=1" in rec["what"] and rec["value"] > 0


def test_ppo_trainer_with_tensor_and_pipeline_parallelism_end_to_end():
    """TP = 2 x PP = 2: sharded separate reference model, 1F1B updates, generation relayed through the stages."""
    rec = _run_config_bench(4, {"BENCH_PP": "2"})
    assert "TP=2 x PP=2" in rec["what"] and rec["value"] > 0


@pytest.mark.parametrize("world,pp", [(2, 1), (4, 2)])
def test_model_parallel_trainer_save_and_resume(world, pp, tmp_path):
    """Every (tensor, pipeline) rank writes and reads back its own shard (``model_state_mp_XX[_YYY].pt``, ``mp_rank_XX[_YYY]/``);
    the script also takes one more optimizer step on the original and the resumed trainer and requires identical weights
    (optimizer shards, master weights and scheduler came back)."""
    import subprocess
    import sys

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "scripts/check_mp_resume.py"]
    proc = subprocess.run(cmd, env=dict(os.environ, OMP_NUM_THREADS="1", PP=str(pp), CKPT_DIR=str(tmp_path)), capture_output=True,
                          text=True, timeout=900)
    assert proc.returncode == 0, proc.stderr[-3000:]
    assert proc.stdout.count("resume_err 0.000e+00") == world, proc.stdout
    saved = sorted(os.listdir(tmp_path / "ck"))
    assert sum(n.startswith("model_state_mp_") for n in saved) == world  # dp = 1: one file per model-parallel rank


# ---- ZeRO-3 parameter partitioning -------------------------------------------------------------------------------------------
def _zero3_job(rank, world, stage, tmp):
    from trlx_b200.data.default_configs import default_sft_config
    from trlx_b200.pipeline import MiniBatchIterator
    from trlx_b200.pipeline.offline_pipeline import PromptPipeline
    from trlx_b200.utils import set_seed
    from trlx_b200.utils.loading import get_trainer

    arch = dict(model_type="gpt2", vocab_size=128, n_embd=32, n_layer=3, n_head=2, n_positions=64)
    cfg = default_sft_config().evolve(
        train=dict(seq_length=24, batch_size=4, tracker=None, checkpoint_dir=os.path.join(tmp, f"s{stage}"), checkpoint_interval=10 ** 9,
                   eval_interval=10 ** 9, total_steps=10 ** 9, parallel=dict(zero_stage=stage)),
        model=dict(model_path=arch), tokenizer=dict(tokenizer_path="toy://bytes"), optimizer=dict(name="adamw", kwargs=dict(lr=1e-2)))
    set_seed(cfg.train.seed, cfg.train.parallel)
    trainer = get_trainer(cfg.train.trainer)(config=cfg)
    texts = ["the movie was " + "very " * (i % 5) + "good" for i in range(32)]
    trainer.make_experience(texts, cfg.train.seq_length)
    trainer.add_eval_pipeline(PromptPipeline(texts[:2], 16, trainer.tokenizer))
    trainer.prepare_learning()
    it = iter(MiniBatchIterator(trainer.create_train_dataloader(), trainer.mb_size, trainer.num_mb))
    losses = [float(trainer.train_step(next(it))["loss"]) for _ in range(3)]
    z = getattr(trainer, "zero3", None)
    resident = sum(p.numel() for p in trainer.model.parameters())  # what the rank holds between steps
    with trainer._full_params():
        full = {k: v.detach().clone() for k, v in trainer.model.state_dict().items()}
    trainer.save(os.path.join(tmp, f"ckpt{stage}"))
    trainer.save_pretrained(os.path.join(tmp, f"hf{stage}"))
    # generation goes through the per-unit gather hooks
    ids = torch.tensor([[3, 4, 5]])
    out = trainer.generate(ids, torch.ones_like(ids), max_new_tokens=4, do_sample=False)
    if z is not None:
        z.release_all()
    return dict(losses=losses, full=full, resident=resident, sharded=z is not None, gen=out.cpu(),
                units=len(z.units) if z is not None else 0)


def test_zero3_parameter_partitioning_matches_replicated_training(tmp_path):
    """``zero_stage: 3`` (parameters + gradients + optimizer state partitioned, per-block gathers) reproduces the replicated
    run: same losses, same final weights, same greedy generation; between steps a rank holds no full parameter."""
    ref = run_distributed(_zero3_job, 2, args=(1, str(tmp_path)))
    got = run_distributed(_zero3_job, 2, args=(3, str(tmp_path)))
    assert all(r["sharded"] for r in got) and not any(r["sharded"] for r in ref) and got[0]["units"] >= 4
    assert got[0]["resident"] == 0 and ref[0]["resident"] > 0
    assert got[0]["losses"] == pytest.approx(ref[0]["losses"], rel=1e-4, abs=1e-5)
    for k, v in ref[0]["full"].items():
        torch.testing.assert_close(got[0]["full"][k], v, atol=1e-5, rtol=1e-4, msg=k)
        torch.testing.assert_close(got[1]["full"][k], v, atol=1e-5, rtol=1e-4, msg=k)
    assert torch.equal(got[0]["gen"], ref[0]["gen"])
    assert os.path.exists(tmp_path / "hf3" / "config.json")


# ---- tensor x pipeline parallelism with sequence parallelism ---------------------------------------------------------------------
def _tp_pp_sp_job(rank, world, virtual=1):
    """4 ranks = TP 2 x PP 2, sequence parallel on: stage boundaries carry each TP rank's sequence shard."""
    import copy

    from trlx_b200.nn.arch import spec_from_hf_config
    from trlx_b200.nn.transformer import CausalLM
    from trlx_b200.parallel import pipeline_parallel as pp
    from trlx_b200.parallel.tensor_parallel import allreduce_sequence_parallel_grads, apply_tensor_parallel

    tp_size, pp_size = 2, 2
    pp_rank, tp_rank = rank // tp_size, rank % tp_size
    tp_groups = [dist.new_group([s * tp_size + t for t in range(tp_size)]) for s in range(pp_size)]
    pp_groups = [dist.new_group([s * tp_size + t for s in range(pp_size)]) for t in range(tp_size)]
    tp_group, pp_group = tp_groups[pp_rank], pp_groups[tp_rank]

    torch.manual_seed(0)
    spec = spec_from_hf_config(dict(model_type="gpt2", vocab_size=61, n_embd=32, n_layer=4, n_head=4, n_positions=64,
                                    tie_word_embeddings=False))
    full = CausalLM(spec).float()
    staged = copy.deepcopy(full)
    tpc = apply_tensor_parallel(staged, tp_group, tp_rank, tp_size, sequence_parallel=True)
    stage = pp.apply_pipeline_parallel(staged, pp_group, pp_rank, pp_size, virtual)
    g = torch.Generator().manual_seed(1)
    mbs = []
    for i in range(4):
        T = 8 if i % 2 == 0 else 6  # even lengths: shardable over TP = 2
        ids = torch.randint(0, 61, (2, T), generator=g)
        mbs.append(dict(input_ids=ids, attention_mask=torch.ones_like(ids), labels=ids))
    ref_losses = []
    for mb in mbs:
        out = full(**mb)
        out.loss.backward()
        ref_losses.append(out.loss.detach())
    ref_ln = full.transformer.h[stage.lo].norm1.weight.grad.clone()
    ref_down = full.transformer.h[stage.lo].mlp.down.weight.grad.clone()

    boundary_shapes = []
    orig_exchange = pp._exchange

    def spy(st, send_next=None, send_prev=None, recv_prev_shape=None, recv_next_shape=None, **kw):
        for sh in (recv_prev_shape, recv_next_shape):
            if sh is not None:
                boundary_shapes.append(tuple(sh))
        return orig_exchange(st, send_next=send_next, send_prev=send_prev, recv_prev_shape=recv_prev_shape,
                             recv_next_shape=recv_next_shape, **kw)

    pp._exchange = spy

    def loss_fn(mb):
        out = staged(**mb)
        return out.loss, {"loss": out.loss.detach()}

    if virtual > 1:  # interleaved schedule: record the shapes of the buffers its grouped transfers receive into
        orig_empty = torch.empty

        def empty_spy(*a, **k):
            if a and isinstance(a[0], tuple) and len(a[0]) == 3:
                boundary_shapes.append(tuple(a[0]))
            return orig_empty(*a, **k)

        pp.torch.empty = empty_spy
    try:
        stats = pp.run_schedule(stage, mbs, loss_fn, torch.device("cpu"))
    finally:
        pp._exchange = orig_exchange
        if virtual > 1:
            pp.torch.empty = orig_empty
    allreduce_sequence_parallel_grads(staged, tp_group)
    shared = pp.broadcast_stats(stage, {"loss": sum(s["loss"] for s in stats) / len(mbs)} if stage.last else None, torch.device("cpu"))
    blk = staged.transformer.h[stage.lo]
    f = ref_down.shape[1] // tp_size
    return dict(mean_loss=shared["loss"], ref_mean=sum(ref_losses) / len(ref_losses),
                ln_err=(blk.norm1.weight.grad - ref_ln).abs().max().item(),
                down_err=(blk.mlp.down.weight.grad - ref_down[:, tp_rank * f:(tp_rank + 1) * f]).abs().max().item(),
                shapes=boundary_shapes, sp=tpc.sequence_parallel)


@pytest.mark.parametrize("virtual", [1, 2])
def test_tensor_pipeline_sequence_parallel_matches_single_rank(virtual):
    res = run_distributed(_tp_pp_sp_job, 4, (virtual,))
    for r in res:
        assert r["sp"]
        torch.testing.assert_close(r["mean_loss"], r["ref_mean"], atol=1e-5, rtol=1e-5)
        assert r["ln_err"] < 2e-5 and r["down_err"] < 2e-5, r
        # every boundary tensor is a sequence SHARD: T / tp positions, not T
        assert r["shapes"] and all(sh[1] in (4, 3) for sh in r["shapes"]), r["shapes"]
