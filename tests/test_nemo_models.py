"""Model-parallel building blocks of the NeMo-style models (``trlx_b200/models/modeling_nemo_*.py``): single-process semantics
plus gloo world-2 equivalence against the dense computation, including gradients."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from test_distributed_cpu import run_distributed


# ---- single process ---------------------------------------------------------------------------------------------------------
def test_parallel_linear_is_dense_without_a_group():
    from trlx_b200.models.modeling_nemo_ppo import ParallelLinear, ValueHead, make_parallel_head

    torch.manual_seed(0)
    col, row = ParallelLinear(8, 16, dtype=torch.float32), ParallelLinear(16, 4, dtype=torch.float32)
    assert col.column and not row.column and col.weight.shape == (16, 8) and row.weight.shape == (4, 16)
    x = torch.randn(3, 5, 8)
    torch.testing.assert_close(col(x), F.linear(x, col.weight, col.bias))
    head = make_parallel_head(8, 1, dtype=torch.float32)
    assert head(x).shape == (3, 5, 1)
    assert ValueHead(8, dtype=torch.float32)(x).shape == (3, 5)
    assert ValueHead(8, dtype=torch.float32, seq_first=True)(x).shape == (5, 3)


def test_ref_lm_heads_swap_reference_weights():
    from trlx_b200.models.modeling_base import build_base_model
    from trlx_b200.models.modeling_nemo_ppo import RefLMHeads

    torch.manual_seed(0)
    lm = build_base_model(dict(model_type="gpt2", vocab_size=50, n_embd=16, n_layer=2, n_head=2, n_positions=32), "causal")
    heads = RefLMHeads(lm, nn.Linear(16, 1))
    ids = torch.randint(0, 50, (2, 7))
    ref_logits0 = lm(ids).logits.detach().clone()
    with torch.no_grad():  # "train" the policy
        for p in lm.parameters():
            p.add_(0.05 * torch.randn_like(p))
    policy_logits = lm(ids).logits.detach().clone()
    logits, values, ref_logits = heads(ids, run_policy_model=True, run_reference_model=True, run_value_head=True)
    torch.testing.assert_close(logits, policy_logits)
    torch.testing.assert_close(ref_logits, ref_logits0)
    assert values.shape == (2, 7, 1) and heads.reference_model_offloaded
    torch.testing.assert_close(lm(ids).logits, policy_logits)  # policy weights are back
    with heads.reference():
        torch.testing.assert_close(lm(ids).logits, ref_logits0)
    # loading a pretrained state dict re-snapshots the reference
    heads.load_state_dict(lm.state_dict())
    _, _, ref_logits2 = heads(ids, run_policy_model=False, run_reference_model=True)
    torch.testing.assert_close(ref_logits2, policy_logits)


def test_reshard_for_pipeline_parallelism_keeps_and_renumbers_stage_layers():
    from trlx_b200.models.modeling_nemo_ppo import reshard_for_pipeline_parallelism

    sd = {f"transformer.h.{i}.mlp.up.weight": i for i in range(6)}
    sd.update({"transformer.wte.weight": "emb", "transformer.ln_f.weight": "norm", "lm_head.weight": "head"})
    first, last = reshard_for_pipeline_parallelism(6, sd, 0, 3), reshard_for_pipeline_parallelism(6, sd, 2, 3)
    assert first == {"transformer.h.0.mlp.up.weight": 0, "transformer.h.1.mlp.up.weight": 1, "transformer.wte.weight": "emb",
                     "lm_head.weight": "head"}
    assert last["transformer.h.0.mlp.up.weight"] == 4 and last["transformer.h.1.mlp.up.weight"] == 5
    assert last["transformer.ln_f.weight"] == "norm" and "transformer.h.2.mlp.up.weight" not in last


def test_vocab_parallel_cross_entropy_single_rank_matches_dense():
    from trlx_b200.models.modeling_nemo_sft import vocab_parallel_cross_entropy

    torch.manual_seed(1)
    x = torch.randn(3, 5, 11, requires_grad=True)
    t = torch.randint(0, 11, (3, 5))
    loss = vocab_parallel_cross_entropy(x, t)
    ref = F.cross_entropy(x.transpose(1, 2), t, reduction="none")
    torch.testing.assert_close(loss, ref)
    loss.sum().backward()
    g = x.grad.clone()
    x.grad = None
    ref.sum().backward()
    torch.testing.assert_close(g, x.grad)


def test_parallel_ilql_heads_polyak_sync_and_shapes():
    from trlx_b200.models.modeling_ilql import ILQLConfig
    from trlx_b200.models.modeling_nemo_ilql import ParallelILQLHeads

    torch.manual_seed(0)
    cfg = ILQLConfig(name="ilqlconfig", tau=0.7, gamma=0.99, cql_scale=0.1, awac_scale=1.0, alpha=0.25, beta=0.0, steps_for_target_q_sync=1,
                     two_qs=True, gen_kwargs={})
    heads = ParallelILQLHeads(cfg, 8, 13, dtype=torch.float32)
    qs, tqs, vs = heads(torch.randn(2, 4, 8))
    assert len(qs) == 2 and qs[0].shape == (2, 4, 13) and tqs[1].shape == (2, 4, 13) and vs.shape == (2, 4, 1)
    assert all(not p.requires_grad for p in heads.target_q_heads.parameters())
    before = [p.clone() for p in heads.target_q_heads.parameters()]
    with torch.no_grad():
        for p in heads.q_heads.parameters():
            p.add_(1.0)
    heads.sync_target_q_heads()
    for b, p in zip(before, heads.target_q_heads.parameters()):
        torch.testing.assert_close(p, b + 0.25)  # alpha * (b + 1) + (1 - alpha) * b


def test_megatron_trainer_plan_and_cyclic_sequence(tmp_path):
    from trlx_b200.trainer.nemo_ilql_trainer import ShuffledCyclicSequence, megatron_trainer

    recipe = dict(trainer=dict(devices=4, num_nodes=2, precision=16, max_steps=123, max_time="00:00:10:30"),
                  exp_manager=dict(explicit_log_dir=str(tmp_path), resume_if_exists=True),
                  model=dict(seed=7, tensor_model_parallel_size=2, pipeline_model_parallel_size=1, sequence_parallel=True,
                             num_layers=2, hidden_size=16, num_attention_heads=2, ffn_hidden_size=64, encoder_seq_length=32,
                             max_position_embeddings=32, optim=dict(name="distributed_fused_adam", lr=1e-4)))
    (tmp_path / "run").mkdir()
    (tmp_path / "run" / "checkpoint_10").mkdir()
    plan = megatron_trainer(recipe)
    assert (plan.devices, plan.num_nodes, plan.precision, plan.max_steps, plan.seed) == (4, 2, "fp16", 123, 7)
    assert plan.max_time_seconds() == 630 and plan.distributed_optimizer and plan.grad_scaler["growth_interval"] == 1000
    assert plan.parallel["tensor_parallel"] == 2 and plan.parallel["sequence_parallel"] is True
    assert plan.resume_from_checkpoint.endswith("checkpoint_10")
    seq = ShuffledCyclicSequence(10, ["a", "b", "c"], seed=3)
    items = [seq[i] for i in range(len(seq))]
    assert len(items) == 10 and set(items) == {"a", "b", "c"} and items == [seq[i] for i in range(10)]


def test_learn_stops_at_max_time(tmp_path):
    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_sft_config

    cfg = default_sft_config().evolve(
        train=dict(total_steps=10 ** 6, epochs=10 ** 6, batch_size=2, seq_length=16, tracker=None, checkpoint_interval=10 ** 6,
                   eval_interval=10 ** 6, checkpoint_dir=str(tmp_path), trainer_kwargs=dict(max_time="00:00:00:01")),
        model=dict(model_path=dict(model_type="gpt2", vocab_size=257, n_embd=16, n_layer=1, n_head=2, n_positions=32,
                                   eos_token_id=256, bos_token_id=256)),
        tokenizer=dict(tokenizer_path="toy://bytes"), method=dict(gen_kwargs=dict(max_new_tokens=2)))
    trainer = trlx.train(samples=["ab", "cd", "ef", "gh"], eval_prompts=["a"], config=cfg)
    assert 0 < trainer.iter_count < 10 ** 6
    assert any(name.startswith("checkpoint_") for name in os.listdir(tmp_path))


# ---- two tensor-parallel ranks ----------------------------------------------------------------------------------------------
def _tp_state(rank, world):
    from trlx_b200.parallel.state import set_model_parallel

    return set_model_parallel(tp_group=dist.group.WORLD, tp_rank=rank, tp_size=world)


def _parallel_head_job(rank, world):
    from trlx_b200.models.modeling_nemo_ppo import make_parallel_head, shard_head_state_dict
    from trlx_b200.utils.modeling import make_head

    _tp_state(rank, world)
    out = {}
    for name, n_out in (("value", 1), ("vocab", 40)):
        torch.manual_seed(0)
        dense = make_head(8, n_out, torch.float32)
        head = make_parallel_head(8, n_out, dtype=torch.float32)
        head.load_state_dict(shard_head_state_dict(dense.state_dict(), head))
        torch.manual_seed(1)  # identical (replicated) activations on both ranks
        x = torch.randn(3, 5, 8, requires_grad=True)
        y = head(x)
        (y * torch.arange(1, n_out + 1, dtype=torch.float32)).sum().backward()
        gx = x.grad.clone()
        x.grad = None
        y_ref = dense(x)
        (y_ref * torch.arange(1, n_out + 1, dtype=torch.float32)).sum().backward()
        # local parameter gradients against the matching slices of the dense gradients
        want = shard_head_state_dict({k: p.grad for k, p in dense.named_parameters()}, head)
        got = {k: p.grad for k, p in head.state_dict(keep_vars=True).items()}
        out[name] = dict(y=(y - y_ref).abs().max().item(), gx=(gx - x.grad).abs().max().item(),
                         gp=max((got[k] - want[k]).abs().max().item() for k in want),
                         local_rows=head[0].weight.shape[0])
    return out


def test_parallel_heads_match_dense_heads_on_two_ranks():
    for res in run_distributed(_parallel_head_job, 2):
        for name in ("value", "vocab"):
            assert res[name]["y"] < 1e-5 and res[name]["gx"] < 1e-4 and res[name]["gp"] < 5e-4, res  # fp32 summation order
        assert res["value"]["local_rows"] == 8 and res["vocab"]["local_rows"] == 8  # 2n = 16 rows split over two ranks


def _vocab_ce_job(rank, world):
    from trlx_b200.models.modeling_nemo_sft import vocab_parallel_cross_entropy

    st = _tp_state(rank, world)
    torch.manual_seed(5)
    full = torch.randn(4, 6, 10)
    target = torch.randint(0, 10, (4, 6))
    local = full.chunk(world, dim=-1)[rank].clone().requires_grad_(True)
    loss = vocab_parallel_cross_entropy(local, target, st)
    loss.sum().backward()
    ref_in = full.clone().requires_grad_(True)
    ref = F.cross_entropy(ref_in.transpose(1, 2), target, reduction="none")
    ref.sum().backward()
    return dict(loss=(loss - ref).abs().max().item(),
                grad=(local.grad - ref_in.grad.chunk(world, dim=-1)[rank]).abs().max().item())


def test_vocab_parallel_cross_entropy_on_two_ranks():
    for res in run_distributed(_vocab_ce_job, 2):
        assert res["loss"] < 1e-5 and res["grad"] < 1e-6, res


def _sft_gpt_job(rank, world):
    from trlx_b200.models.modeling_base import build_base_model
    from trlx_b200.models.modeling_nemo_sft import SFTGPT

    _tp_state(rank, world)
    torch.manual_seed(0)
    lm = build_base_model(dict(model_type="gpt2", vocab_size=48, n_embd=16, n_layer=1, n_head=2, n_positions=32), "causal")
    ids = torch.randint(0, 48, (2, 9))
    mask = torch.ones_like(ids)
    dense_loss, _ = SFTGPT(language_model=lm, vocab_parallel=False)(ids, mask)
    par_loss, _ = SFTGPT(language_model=lm, vocab_parallel=True)(ids, mask)
    par_loss.backward()
    return dict(diff=abs(dense_loss.item() - par_loss.item()), has_grad=lm.lm_head.weight.grad is not None)


def test_sft_gpt_vocab_parallel_loss_matches_dense():
    for res in run_distributed(_sft_gpt_job, 2):
        assert res["diff"] < 1e-5 and res["has_grad"], res


_TINY = dict(model_type="gpt2", vocab_size=50, n_embd=16, n_layer=3, n_head=2, n_positions=32)


@pytest.mark.parametrize("unfrozen", [1, -1])
def test_ppogpt_single_process(unfrozen):
    """Hydra branch (``num_layers_unfrozen > 0``) or host-resident reference copies (full fine-tuning) as the reference policy."""
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.models.modeling_nemo_ppo import PPOGPT
    from trlx_b200.parallel.state import set_model_parallel

    set_model_parallel()
    torch.manual_seed(0)
    model = PPOGPT(default_ppo_config().evolve(model=dict(model_path=_TINY, num_layers_unfrozen=unfrozen)))
    ids = torch.randint(0, 50, (2, 6))
    out = model(ids)
    assert out.logits.shape == (2, 6, 50) and out.value.shape == (2, 6)
    assert (model.ref_heads is not None) == (unfrozen == -1)
    torch.testing.assert_close(model.reference_logits(ids), out.logits.detach())  # untrained: reference == policy
    with torch.no_grad():
        for p in model.model.base_model.parameters():
            if p.requires_grad:
                p.add_(0.05)
    assert (model.reference_logits(ids) - model(ids).logits).abs().max() > 1e-4  # the reference stayed behind
    assert model.generate(ids, max_new_tokens=3, do_sample=False, pad_token_id=0, eos_token_id=49).shape[1] <= 9


def test_ilqlgpt_shifted_logits():
    from trlx_b200.data.default_configs import default_ilql_config
    from trlx_b200.models.modeling_nemo_ilql import ILQLGPT
    from trlx_b200.parallel.state import set_model_parallel

    set_model_parallel()
    torch.manual_seed(0)
    cfg = default_ilql_config().evolve(model=dict(model_path=_TINY))
    model = ILQLGPT(cfg.method, cfg, dtype=torch.float32)
    ids = torch.randint(0, 50, (2, 6))
    logits, (qs, target_qs, vs) = model(ids)
    assert logits.shape == (2, 6, 50) and len(qs) == len(target_qs) == (2 if cfg.method.two_qs else 1) and vs.shape == (2, 6, 1)
    shifted = model.shifted_logits(ids, beta=2.0)
    tq = torch.minimum(*target_qs) if len(target_qs) == 2 else target_qs[0]
    want = F.log_softmax(logits[:, -1].float(), -1) + 2.0 * (tq[:, -1] - vs[:, -1])
    torch.testing.assert_close(shifted, want)


def _ppogpt_tp_job(rank, world):
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.models.modeling_nemo_ppo import PPOGPT
    from trlx_b200.parallel.state import set_model_parallel

    cfg = default_ppo_config().evolve(model=dict(model_path=_TINY, num_layers_unfrozen=1))
    ids = torch.arange(12).view(2, 6) % 50
    set_model_parallel()
    torch.manual_seed(0)
    dense = PPOGPT(cfg)
    d = dense(ids)
    _tp_state(rank, world)
    torch.manual_seed(0)
    sharded = PPOGPT(cfg)
    s = sharded(ids)
    return dict(logits=(d.logits - s.logits).abs().max().item(), value=(d.value - s.value).abs().max().item(),
                head_rows=sharded.value_head.v_head[0].weight.shape[0])


def test_ppogpt_tensor_parallel_matches_dense():
    for res in run_distributed(_ppogpt_tp_job, 2):
        assert res["logits"] < 1e-4 and res["value"] < 1e-4 and res["head_rows"] == 16, res


def test_megatron_batch_sampler_slices_global_batches_per_rank():
    from trlx_b200.models.megatron_api import MegatronBatchSampler

    r0 = list(MegatronBatchSampler(22, 0, 2, 8, 0, 2))
    r1 = list(MegatronBatchSampler(22, 0, 2, 8, 1, 2))
    assert r0 == [[0, 1, 2, 3], [8, 9, 10, 11]] and r1 == [[4, 5, 6, 7], [12, 13, 14, 15]]
    assert list(MegatronBatchSampler(22, 8, 2, 8, 1, 2)) == [[12, 13, 14, 15]]  # resume after 8 consumed samples
    assert len(MegatronBatchSampler(22, 0, 2, 8, 0, 2)) == 2
    with pytest.raises(ValueError):
        MegatronBatchSampler(22, 0, 3, 8, 0, 2)


def test_model_api_training_step_checkpoint_and_inference_mode(tmp_path):
    from trlx_b200.data.default_configs import default_ppo_config, default_sft_config
    from trlx_b200.data.ppo_types import PPORLBatch
    from trlx_b200.models.modeling_nemo_ppo import PPOGPT
    from trlx_b200.models.modeling_nemo_sft import SFTGPT
    from trlx_b200.parallel.state import set_model_parallel

    set_model_parallel()
    torch.manual_seed(0)
    cfg = default_ppo_config().evolve(model=dict(model_path=_TINY, num_layers_unfrozen=1), train=dict(batch_size=4, minibatch_size=2))
    model = PPOGPT(cfg)
    B, Q, R = 4, 5, 3
    batch = PPORLBatch(query_tensors=torch.randint(1, 50, (B, Q)), response_tensors=torch.randint(1, 50, (B, R)),
                       logprobs=torch.randn(B, R) * 0.1 - 2, values=torch.randn(B, R) * 0.1, rewards=torch.randn(B, R) * 0.1)
    opt, _ = model.configure_optimizers()
    before = [p.detach().clone() for p in model.parameters() if p.requires_grad]
    stats = model.training_step(batch, opt)
    assert "loss" in stats and any(not torch.equal(a, b) for a, b in zip(before, (p for p in model.parameters() if p.requires_grad)))
    lp, rlp, val = model.infer_logprobs_and_values(torch.randint(1, 50, (2, 7)))
    assert lp.shape == rlp.shape == val.shape == (2, 6)
    path = model.save_pretrained(str(tmp_path))
    assert path.endswith(os.path.join("mp_rank_00", "model_weights.ckpt")) and os.path.exists(path)
    clone = PPOGPT(cfg)
    clone.load_from_pretrained(str(tmp_path), strict=True)
    ids = torch.randint(1, 50, (2, 7))
    torch.testing.assert_close(clone(ids).logits, model(ids).logits)
    loader = model.build_data_loader(list(range(20)), lambda xs: xs)
    assert [len(b) for b in loader] == [4] * 5

    sft = SFTGPT(default_sft_config().evolve(model=dict(model_path=_TINY)), dtype=torch.float32)
    sft.activation_checkpointing_(True)
    with sft.inference_mode():
        assert not torch.is_grad_enabled() and not sft.training
        assert not any(getattr(m, "gradient_checkpointing", False) for m in sft.modules())
    assert sft.training and any(getattr(m, "gradient_checkpointing", False) for m in sft.modules())
    opt, _ = sft.configure_optimizers()
    l0 = sft.training_step(dict(input_ids=torch.arange(16).view(2, 8) % 50), opt)["loss"]
    l1 = sft.training_step(dict(input_ids=torch.arange(16).view(2, 8) % 50), opt)["loss"]
    assert l1 < l0
    assert sft.maybe_initalize_per_dp_rng(3).initial_seed() == 3


def test_ilqlgpt_training_step_and_generate():
    from trlx_b200.data.default_configs import default_ilql_config
    from trlx_b200.models.modeling_nemo_ilql import ILQLGPT
    from trlx_b200.parallel.state import set_model_parallel
    from trlx_b200.trainer.accelerate_ilql_trainer import make_experience
    from trlx_b200.utils.tokenizer import build_toy_tokenizer

    set_model_parallel()
    torch.manual_seed(0)
    cfg = default_ilql_config().evolve(model=dict(model_path=_TINY))
    model = ILQLGPT(cfg.method, cfg, dtype=torch.float32)
    tok = build_toy_tokenizer("toy://chars?alphabet=abcdefghij")
    store = make_experience([["ab", "cd"], ["ef", "gh"], ["ab", "ij"], ["cd", "ef"]], [1.0, 0.5, 0.2, 0.9], tok, max_length=16,
                            verbose=False)
    batch = next(iter(store.create_loader(4)))
    opt, _ = model.configure_optimizers()
    stats = model.training_step(batch, opt)
    assert stats["loss"] > 0 and "losses/loss_q" in stats
    out = model.generate(torch.randint(1, 10, (2, 3)), max_new_tokens=3)
    assert out.shape == (2, 6)
