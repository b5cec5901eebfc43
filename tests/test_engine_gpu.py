"""Rollout engine (CUDA-graph decode on the sm_100a kernels) vs the PyTorch sampler + scoring pass."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(family="gpt2"):
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx_b200.utils.modeling import freeze_bottom_causal_layers

    torch.manual_seed(0)
    if family == "gpt2":
        cfg = dict(model_type="gpt2", vocab_size=1000, n_embd=128, n_layer=4, n_head=2, n_positions=128,
                   eos_token_id=999, bos_token_id=999)
    else:
        cfg = dict(model_type="llama", vocab_size=1000, hidden_size=128, num_hidden_layers=4, num_attention_heads=2,
                   num_key_value_heads=1, intermediate_size=256, max_position_embeddings=128, eos_token_id=999, bos_token_id=1)
    m = AutoModelForCausalLMWithHydraValueHead.from_config(cfg, num_layers_unfrozen=2)
    freeze_bottom_causal_layers(m.base_model, 2)
    m = m.cuda().to(torch.bfloat16).eval()
    # make the policy branch differ from the reference branch so ref log-probs are a real test
    with torch.no_grad():
        for p in m.base_model.transformer.h[-1].parameters():
            p.add_(torch.randn_like(p) * 0.02)
    return m


@pytest.mark.parametrize("family", ["gpt2", "llama"])
@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("defer_ref", ["1", "0"])
def test_greedy_rollout_matches_torch_path(family, graph, defer_ref, monkeypatch):
    """``defer_ref``: reference log-probs from one batched pass over the cached trunk activations after the loop (default)
    vs the frozen branch inside every decode step."""
    from trlx_b200.engine.rollout import RolloutEngine
    from trlx_b200.models.generation import generate

    monkeypatch.setenv("TRLX_B200_DEFER_REF", defer_ref)

    m = _model(family)
    pad = eos = 999
    B, Q, R = 6, 9, 12
    gen = dict(max_new_tokens=R, do_sample=False, eos_token_id=eos, pad_token_id=pad, top_k=0, top_p=1.0)
    assert RolloutEngine.supports(m, gen)
    eng = RolloutEngine(m, pad, eos, gen, cache_trunk=True, seed=1, use_cuda_graph=graph)
    ids = torch.randint(1, 900, (B, Q), device="cuda")
    mask = torch.ones(B, Q, dtype=torch.long, device="cuda")
    for b, npad in enumerate([0, 3, 1, 5, 0, 2]):
        mask[b, :npad] = 0
        ids[b, :npad] = pad
    for _ in range(2):  # second call exercises graph replay / state reset
        ro = eng.rollout(ids, mask)
    ref = generate(m.base_model, ids, attention_mask=mask, **gen)
    Rg = ro["sample_outputs"].shape[1]
    # greedy decoding of a random bf16 model has near-ties, so token-for-token equality with the PyTorch sampler is only
    # required for the first step; optimality of every chosen token is checked against teacher-forced logits below
    assert (ro["samples"][:, Q] == ref[:, Q]).float().mean() >= 0.8, (ro["samples"][:, Q], ref[:, Q])

    all_tokens = ro["samples"]
    amask = ro["mask"]
    pos = (amask.cumsum(-1) - 1).clamp_min(0)
    labels = torch.cat([all_tokens[:, 1:], all_tokens.new_full((B, 1), -1)], 1)
    with torch.no_grad():
        lp, val, rlp, trunk = m.score(all_tokens, amask, pos, labels)
    start = Q - 1
    # rows that finished early have pad positions that the engine zero-fills
    resp_valid = torch.arange(Rg, device="cuda")[None] < (amask[:, Q:].sum(1, keepdim=True) + (1 if pad == eos else 0))
    sel = resp_valid
    torch.testing.assert_close(ro["logprobs"][:, start:][sel], lp[:, :-1][:, start:][sel].float(), atol=6e-2, rtol=5e-2)
    torch.testing.assert_close(ro["ref_logprobs"][:, start:][sel], rlp[:, :-1][:, start:][sel].float(), atol=6e-2, rtol=5e-2)
    torch.testing.assert_close(ro["values"][:, start:][sel], val[:, :-1][:, start:][sel].float(), atol=6e-2, rtol=5e-2)
    with torch.no_grad():
        logits = m(all_tokens, amask, position_ids=pos)[0].float()
    best = torch.log_softmax(logits, -1).max(-1).values[:, :-1][:, start:]
    assert ((best - ro["logprobs"][:, start:])[sel] < 0.1).all(), "engine picked a token that is not (near-)argmax"
    # prompt-position log-probs (used for the KL statistic) and the cached trunk activation
    pm = amask[:, :start].bool() & amask[:, 1:Q].bool()
    torch.testing.assert_close(ro["logprobs"][:, :start][pm], lp[:, :start][pm].float(), atol=6e-2, rtol=5e-2)
    torch.testing.assert_close(ro["ref_logprobs"][:, :start][pm], rlp[:, :start][pm].float(), atol=6e-2, rtol=5e-2)
    t_eng, t_ref = ro["trunk"], trunk[:, : ro["trunk"].shape[1]]
    tm = amask[:, : t_eng.shape[1]].bool() & torch.cat([torch.ones(B, Q, dtype=torch.bool, device="cuda"), resp_valid[:, 1:Rg]], 1)[:, : t_eng.shape[1]]
    torch.testing.assert_close(t_eng[tm].float(), t_ref[tm].float(), atol=8e-2, rtol=5e-2)


def test_sampling_rollout_statistics():
    from trlx_b200.engine.rollout import RolloutEngine

    m = _model("gpt2")
    pad = eos = 999
    gen = dict(max_new_tokens=8, do_sample=True, eos_token_id=eos, pad_token_id=pad, top_k=0, top_p=1.0)
    eng = RolloutEngine(m, pad, eos, gen, seed=3)
    ids = torch.randint(1, 900, (16, 5), device="cuda")
    a = eng.rollout(ids, torch.ones_like(ids))
    b = eng.rollout(ids, torch.ones_like(ids))
    assert not torch.equal(a["sample_outputs"], b["sample_outputs"])  # fresh noise per call
    assert (a["logprobs"][:, 4:] <= 0).all() and torch.isfinite(a["logprobs"]).all()
    assert a["trunk"].shape[1] == a["samples"].shape[1] - 1


def test_ppo_trainer_runs_on_engine(tmp_path):
    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_ppo_config

    cfg = default_ppo_config().evolve(
        train=dict(seq_length=32, batch_size=8, total_steps=4, epochs=2, checkpoint_interval=100, eval_interval=100,
                   tracker=None, checkpoint_dir=str(tmp_path), seed=3),
        model=dict(model_path=dict(model_type="gpt2", vocab_size=512, n_embd=128, n_layer=4, n_head=2, n_positions=64,
                                   eos_token_id=256, bos_token_id=256), num_layers_unfrozen=2),
        tokenizer=dict(tokenizer_path="toy://bytes"),
        method=dict(num_rollouts=16, chunk_size=8, ppo_epochs=2, gen_kwargs=dict(max_new_tokens=8, top_k=0, top_p=1.0, do_sample=True)),
    )
    trainer = trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) for s in samples],
                         prompts=["hello world", "the quick", "a", "brown fox jumps"] * 4, eval_prompts=["hi"] * 2, config=cfg)
    assert trainer._engine is not None, "the CUDA rollout engine must be the path that ran"
    assert trainer.iter_count == 4


@pytest.mark.parametrize("all_four", [False, True])
def test_fp8_rollout_stays_close_to_bf16(monkeypatch, all_four):
    """rollout_dtype = fp8: teacher-forced log-probs of the fp8 engine's own samples, re-scored by the bf16 torch model, stay
    within quantisation noise of what the engine reported — with the two norm → GEMM pairs in e4m3, and with all four GEMMs of
    a block in e4m3 (``TRLX_B200_FP8_ALL=1``; automatic from hidden size 2048)."""
    from trlx_b200.engine.rollout import RolloutEngine

    monkeypatch.setenv("TRLX_B200_FP8_ALL", "1" if all_four else "0")
    m = _model("gpt2")
    gen = dict(max_new_tokens=8, do_sample=False, eos_token_id=999, pad_token_id=999, top_k=0, top_p=1.0, _rollout_dtype="fp8")
    eng = RolloutEngine(m, 999, 999, gen, seed=0)
    assert eng.fp8 and eng.fp8_all == all_four
    torch.manual_seed(1)
    prompts = torch.randint(0, 990, (16, 7), device="cuda")
    ro = eng.rollout(prompts, torch.ones_like(prompts))
    tokens, mask = ro["samples"], ro["mask"]
    pos = (mask.cumsum(-1) - 1).clamp_min(0)
    with torch.no_grad():
        out = m(tokens, attention_mask=mask, position_ids=pos, return_dict=True)
    lp = torch.log_softmax(out.logits[:, :-1].float(), -1).gather(-1, tokens[:, 1:, None]).squeeze(-1)
    start = ro["start"]
    valid = mask[:, start + 1:].bool()
    diff = (lp[:, start:] - ro["logprobs"][:, start:])[valid].abs()
    assert diff.mean().item() < (0.2 if all_four else 0.15) and torch.isfinite(ro["values"]).all()
    assert (eng.fp8_w[0].out_w is not None) == all_four


def _mid_model(L=3, H=256, nh=4, F=1024):
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx_b200.utils.modeling import freeze_bottom_causal_layers

    torch.manual_seed(0)
    cfg = dict(model_type="gpt2", vocab_size=1000, n_embd=H, n_layer=L, n_head=nh, n_inner=F, n_positions=128,
               eos_token_id=999, bos_token_id=999)
    m = AutoModelForCausalLMWithHydraValueHead.from_config(cfg, num_layers_unfrozen=1)
    freeze_bottom_causal_layers(m.base_model, 1)
    m = m.cuda().to(torch.bfloat16).eval()
    with torch.no_grad():  # biases / norm parameters away from their (0, 1) initial values so every fused term is exercised
        for n, p in m.base_model.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    return m


@pytest.mark.parametrize("B", [20, 128])
def test_decode_megakernel_matches_kernel_per_op_path(B, monkeypatch):
    """csrc/decode_mega.cu (all blocks of a decode step in one cluster-resident launch) vs the kernel-per-op decode graph:
    same greedy tokens, log-probs, values and trunk activations."""
    from trlx_b200.engine.rollout import RolloutEngine

    m = _mid_model()
    pad = eos = 999
    Q, R = 7, 10
    gen = dict(max_new_tokens=R, do_sample=False, eos_token_id=eos, pad_token_id=pad, top_k=0, top_p=1.0)
    torch.manual_seed(5)
    ids = torch.randint(1, 900, (B, Q), device="cuda")
    mask = torch.ones(B, Q, dtype=torch.long, device="cuda")
    for b in range(B):
        npad = b % 4
        mask[b, :npad] = 0
        ids[b, :npad] = pad
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TRLX_B200_DECODE_MEGA", flag)
        eng = RolloutEngine(m, pad, eos, gen, cache_trunk=True, seed=1)
        assert eng.mega == (flag == "1")
        for _ in range(2):
            outs[flag] = eng.rollout(ids, mask)
    a, b_ = outs["1"], outs["0"]
    same = (a["sample_outputs"] == b_["sample_outputs"]).all(1)
    assert same.float().mean() >= 0.9, "greedy tokens diverge between the megakernel and the kernel-per-op path"
    start = Q - 1
    for key in ("logprobs", "values", "ref_logprobs"):
        torch.testing.assert_close(a[key][same], b_[key][same], atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(a["trunk"][same].float(), b_["trunk"][same].float(), atol=3e-2, rtol=3e-2)


def test_engine_logprobs_against_fp32_oracle():
    """Teacher-forced log-probs / values of the engine's own samples from an fp32 copy of the model (not the repo's bf16
    forward, which shares kernels with the engine)."""
    import copy

    from trlx_b200.engine.rollout import RolloutEngine

    m = _mid_model()
    pad = eos = 999
    B, Q, R = 32, 6, 12
    gen = dict(max_new_tokens=R, do_sample=True, eos_token_id=eos, pad_token_id=pad, top_k=0, top_p=1.0)
    eng = RolloutEngine(m, pad, eos, gen, cache_trunk=True, seed=2)
    ids = torch.randint(1, 900, (B, Q), device="cuda")
    ro = eng.rollout(ids, torch.ones_like(ids))
    m32 = copy.deepcopy(m).float()  # fp32 tensors never take the bf16 kernel paths: plain PyTorch maths
    tokens, amask = ro["samples"], ro["mask"]
    pos = (amask.cumsum(-1) - 1).clamp_min(0)
    with torch.no_grad():
        out = m32(tokens, attention_mask=amask, position_ids=pos, return_dict=True)
    lp = torch.log_softmax(out.logits[:, :-1].float(), -1).gather(-1, tokens[:, 1:, None]).squeeze(-1)
    val = out.value[:, :-1].float()
    start = ro["start"]
    valid = amask[:, start + 1:].bool()
    dlp = (lp[:, start:] - ro["logprobs"][:, start:])[valid].abs()
    dv = (val[:, start:] - ro["values"][:, start:])[valid].abs()
    assert dlp.max().item() < 2e-2 + 2e-2 * lp.abs().max().item(), dlp.max()
    assert dlp.mean().item() < 1e-2 and dv.mean().item() < 1e-2, (dlp.mean(), dv.mean())


def test_engine_top_k_top_p_sampling():
    """top-k / top-p run inside the engine (no fall-back to the PyTorch sampler): every sampled token lies in the filtered
    set of the teacher-forced distribution and the reported log-probs are the RAW policy log-probs."""
    from trlx_b200.engine.rollout import RolloutEngine

    m = _mid_model()
    pad = eos = 999
    B, Q, R = 24, 6, 10
    gen = dict(max_new_tokens=R, do_sample=True, eos_token_id=eos, pad_token_id=pad, top_k=5, top_p=0.9, temperature=0.8)
    assert RolloutEngine.why_not(m, gen) is None
    eng = RolloutEngine(m, pad, eos, gen, cache_trunk=True, seed=4)
    assert eng.filtered
    ids = torch.randint(1, 900, (B, Q), device="cuda")
    ro = eng.rollout(ids, torch.ones_like(ids))
    tokens, amask = ro["samples"], ro["mask"]
    pos = (amask.cumsum(-1) - 1).clamp_min(0)
    with torch.no_grad():
        logits = m(tokens, attention_mask=amask, position_ids=pos, return_dict=True).logits.float()
    start = ro["start"]
    lp_all = torch.log_softmax(logits[:, :-1], -1)
    nxt = tokens[:, 1:]
    valid = amask[:, 1:].bool()
    valid[:, :start] = False
    rank = (lp_all > lp_all.gather(-1, nxt[..., None])).sum(-1)  # how many tokens are strictly more likely
    assert (rank[valid] < 5 + 2).all(), "a token outside the top-k set (allowing bf16 near-ties) was sampled"
    got = ro["logprobs"][:, start:][valid[:, start:]]
    want = lp_all.gather(-1, nxt[..., None]).squeeze(-1)[:, start:][valid[:, start:]]
    torch.testing.assert_close(got, want, atol=6e-2, rtol=5e-2)


@pytest.mark.parametrize("fp8", [False, True])
def test_engine_serves_lora_models_with_merged_weights(fp8):
    """PEFT / LoRA policies run on the engine: rollouts read W + (alpha/r) B A (re-merged after weight changes), reference
    log-probs come from one adapter-free pass; checked against teacher-forced scoring with adapters on / off."""
    from trlx_b200.engine.rollout import RolloutEngine
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead

    torch.manual_seed(0)
    cfg = dict(model_type="gpt2", vocab_size=1000, n_embd=256, n_layer=3, n_head=4, n_positions=128, eos_token_id=999, bos_token_id=999)
    m = AutoModelForCausalLMWithHydraValueHead.from_config(
        cfg, peft_config=dict(peft_type="LORA", task_type="CAUSAL_LM", r=8, lora_alpha=32, lora_dropout=0.0))
    m = m.cuda().to(torch.bfloat16).eval()

    def perturb(scale):
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "lora_B" in n:
                    p.copy_(torch.randn_like(p) * scale)

    perturb(0.05)
    pad = eos = 999
    B, Q, R = 16, 6, 8
    gen = dict(max_new_tokens=R, do_sample=False, eos_token_id=eos, pad_token_id=pad, top_k=0, top_p=1.0)
    if fp8:
        gen["_rollout_dtype"] = "fp8"
    assert RolloutEngine.why_not(m, gen) is None
    eng = RolloutEngine(m, pad, eos, gen, seed=1)
    assert eng.lora and eng.fp8 == fp8
    ids = torch.randint(1, 900, (B, Q), device="cuda")
    tol = 0.2 if fp8 else 6e-2
    for round_ in range(2):
        ro = eng.rollout(ids, torch.ones_like(ids))
        tokens, amask = ro["samples"], ro["mask"]
        pos = (amask.cumsum(-1) - 1).clamp_min(0)
        with torch.no_grad():
            on = m(tokens, attention_mask=amask, position_ids=pos, return_dict=True)
            off = m(tokens, attention_mask=amask, position_ids=pos, return_dict=True, ignore_peft_adapter=True)
        nxt = tokens[:, 1:, None]
        lp_on = torch.log_softmax(on.logits[:, :-1].float(), -1).gather(-1, nxt).squeeze(-1)
        lp_off = torch.log_softmax(off.logits[:, :-1].float(), -1).gather(-1, nxt).squeeze(-1)
        start = ro["start"]
        valid = amask[:, 1:].bool()
        valid[:, :start] = False
        assert (lp_on - lp_off)[valid].abs().mean() > 0.02, "the adapters must change the policy for this test to mean anything"
        assert (ro["logprobs"] - lp_on)[valid].abs().mean() < tol / 3 and (ro["logprobs"] - lp_on)[valid].abs().max() < tol * 2
        torch.testing.assert_close(ro["ref_logprobs"][valid], lp_off[valid], atol=6e-2, rtol=5e-2)
        assert ro["trunk"] is None
        perturb(0.08)          # "optimizer step": new adapters ...
        eng.mark_dirty()       # ... must be re-merged before the next rollout


def test_ilql_decode_engine_matches_model_generate():
    """ILQL advantage-shifted decoding on the engine (paged KV, fused heads + sampling kernel) vs the model's PyTorch loop,
    greedy (temperature 0) so the two are comparable token by token."""
    from trlx_b200.engine.ilql import ILQLDecodeEngine
    from trlx_b200.models.modeling_ilql import AutoModelForCausalLMWithILQLHeads

    torch.manual_seed(0)
    cfg = dict(model_type="gpt2", vocab_size=600, n_embd=256, n_layer=3, n_head=4, n_positions=128, eos_token_id=599, bos_token_id=599)
    m = AutoModelForCausalLMWithILQLHeads.from_config(cfg, two_qs=True).cuda().to(torch.bfloat16).eval()
    with torch.no_grad():
        for p in m.ilql_heads.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    assert ILQLDecodeEngine.why_not(m) is None
    eng = ILQLDecodeEngine(m, 599, 599, seed=0)
    B, Q, R = 12, 5, 10
    ids = torch.randint(1, 590, (B, Q), device="cuda")
    mask = torch.ones_like(ids)
    mask[1, :2] = 0
    ids[1, :2] = 599
    kw = dict(beta=1.5, max_new_tokens=R, temperature=0.0, top_k=8, pad_token_id=599, eos_token_id=599)
    for _ in range(2):
        got = eng.generate(ids, mask, **kw)
    want = m.generate(ids, attention_mask=mask, **kw)
    n = min(got.shape[1], want.shape[1])
    assert torch.equal(got[:, :Q], ids)
    first_equal = (got[:, Q] == want[:, Q]).float().mean()
    assert first_equal >= 0.9, (got[:, Q], want[:, Q])
    rows_equal = (got[:, :n] == want[:, :n]).all(1).float().mean()
    assert rows_equal >= 0.6, rows_equal  # bf16 near-ties may flip a later argmax; whole-row agreement must still dominate
    # sampling mode runs and respects the vocabulary / EOS conventions
    s = eng.generate(ids, mask, beta=1.0, max_new_tokens=R, temperature=1.0, top_k=20, pad_token_id=599, eos_token_id=599)
    assert s.shape[0] == B and (s >= 0).all() and (s < 600).all()


def test_engine_serves_models_with_a_value_branch():
    """``num_value_layers_unfrozen > 0``: the value function has its own transformer branch.  The engine samples without it and
    scores every position in one batched pass over the cached trunk activations — same values as ``model.score``."""
    from trlx_b200.engine.rollout import RolloutEngine
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx_b200.utils.modeling import freeze_bottom_causal_layers

    torch.manual_seed(0)
    cfg = dict(model_type="gpt2", vocab_size=1000, n_embd=256, n_layer=4, n_head=4, n_positions=128, eos_token_id=999, bos_token_id=999)
    m = AutoModelForCausalLMWithHydraValueHead.from_config(cfg, num_layers_unfrozen=2, num_value_layers_unfrozen=1)
    freeze_bottom_causal_layers(m.base_model, 2)
    m = m.cuda().to(torch.bfloat16).eval()
    pad = eos = 999
    B, Q, R = 12, 6, 8
    gen = dict(max_new_tokens=R, do_sample=True, eos_token_id=eos, pad_token_id=pad, top_k=0, top_p=1.0)
    assert RolloutEngine.why_not(m, gen) is None
    eng = RolloutEngine(m, pad, eos, gen, seed=1)
    assert eng.value_branch
    ids = torch.randint(1, 900, (B, Q), device="cuda")
    ro = eng.rollout(ids, torch.ones_like(ids))
    tokens, amask = ro["samples"], ro["mask"]
    pos = (amask.cumsum(-1) - 1).clamp_min(0)
    labels = torch.cat([tokens[:, 1:], tokens.new_full((B, 1), -1)], 1)
    with torch.no_grad():
        lp, val, _, _ = m.score(tokens, amask, pos, labels, with_ref=False)
    start = ro["start"]
    valid = amask[:, start + 1:].bool()
    torch.testing.assert_close(ro["values"][:, start:][valid], val[:, :-1][:, start:][valid].float(), atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(ro["logprobs"][:, start:][valid], lp[:, :-1][:, start:][valid].float(), atol=6e-2, rtol=5e-2)
