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
