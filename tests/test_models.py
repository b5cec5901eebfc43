import os

import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from trlx_b200.data.ilql_types import ILQLBatch
from trlx_b200.models.modeling_ilql import (AutoModelForCausalLMWithILQLHeads, AutoModelForSeq2SeqLMWithILQLHeads, ILQLConfig,
                                            ILQLHeads, batched_index_select, topk_mask)
from trlx_b200.models.modeling_ppo import (AutoModelForCausalLMWithHydraValueHead, AutoModelForCausalLMWithValueHead,
                                           AutoModelForSeq2SeqLMWithHydraValueHead, AutoModelForSeq2SeqLMWithValueHead,
                                           hf_get_branch_class)

CAUSAL = {
    "gpt2": dict(model_type="gpt2", vocab_size=64, n_embd=32, n_layer=3, n_head=2, n_positions=64),
    "gptj": dict(model_type="gptj", vocab_size=64, n_embd=32, n_layer=3, n_head=2, rotary_dim=8, n_positions=64),
    "gpt_neox": dict(model_type="gpt_neox", vocab_size=64, hidden_size=32, num_hidden_layers=3, num_attention_heads=2,
                     intermediate_size=64, max_position_embeddings=64),
    "llama": dict(model_type="llama", vocab_size=64, hidden_size=32, num_hidden_layers=3, num_attention_heads=4,
                  num_key_value_heads=2, intermediate_size=48, max_position_embeddings=64),
    "opt": dict(model_type="opt", vocab_size=64, hidden_size=32, num_hidden_layers=3, num_attention_heads=2, ffn_dim=64,
                max_position_embeddings=64, word_embed_proj_dim=32),
    # OPT-350m layout: narrower word embeddings (project_in / project_out), post-LN blocks, no final norm
    "opt350": dict(model_type="opt", vocab_size=64, hidden_size=32, num_hidden_layers=3, num_attention_heads=2, ffn_dim=64,
                   max_position_embeddings=64, word_embed_proj_dim=16, do_layer_norm_before=False),
    "bloom": dict(model_type="bloom", vocab_size=64, hidden_size=32, n_layer=3, n_head=2),
    "gpt_bigcode": dict(model_type="gpt_bigcode", vocab_size=64, n_embd=32, n_layer=3, n_head=2, n_positions=64),
    "gpt_neo": dict(model_type="gpt_neo", vocab_size=64, hidden_size=32, num_layers=4, num_heads=2, max_position_embeddings=64,
                    attention_types=[[["global", "local"], 2]], window_size=4),
}
T5 = dict(model_type="t5", vocab_size=64, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_decoder_layers=3, num_heads=4)


def _inputs(B=3, T=7):
    torch.manual_seed(1)
    ids = torch.randint(2, 60, (B, T))
    mask = torch.ones(B, T, dtype=torch.long)
    mask[0, :2] = 0
    return ids, mask, (mask.cumsum(-1) - 1).clamp_min(0)


@pytest.mark.parametrize("family", list(CAUSAL))
def test_matches_huggingface_reference_implementation(family):
    transformers = pytest.importorskip("transformers")
    from trlx_b200.models.modeling_base import build_base_model, import_base_state_dict

    torch.manual_seed(0)
    cfg = transformers.AutoConfig.for_model(**CAUSAL[family])
    hf = transformers.AutoModelForCausalLM.from_config(cfg).eval()
    ours = build_base_model(CAUSAL[family]).eval()
    import_base_state_dict(ours, hf.state_dict(), strict=True)
    ids, mask, pos = _inputs()
    kw = dict(input_ids=ids, attention_mask=mask)
    if family != "bloom":
        kw["position_ids"] = pos
    with torch.no_grad():
        a = hf(**kw).logits
        b = ours(input_ids=ids, attention_mask=mask, position_ids=pos).logits
    assert (a - b)[mask.bool()].abs().max() < 1e-4


@pytest.mark.parametrize("family", list(CAUSAL))
def test_value_head_wrapper_forward_generate_save_load(family, tmp_path):
    torch.manual_seed(0)
    model = AutoModelForCausalLMWithValueHead.from_config(CAUSAL[family]).eval()
    ids, mask, pos = _inputs()
    logits, *_, value = model(ids, mask, position_ids=pos)
    assert logits.shape == (3, 7, 64) and value.shape == (3, 7)
    out = model(ids, mask, position_ids=pos, return_dict=True)
    assert torch.equal(out.logits, logits) and len(out.hidden_states) == len(model.base_model.transformer.h) + 1
    gen = model.generate(ids, attention_mask=mask, max_new_tokens=4, do_sample=True, eos_token_id=63, pad_token_id=63)
    assert gen.shape[0] == 3 and 7 < gen.shape[1] <= 11 and torch.equal(gen[:, :7], ids)
    model.save_pretrained(str(tmp_path))
    assert {"config.json", "pytorch_model.bin"} <= set(os.listdir(tmp_path))
    sd = torch.load(tmp_path / "pytorch_model.bin")
    assert any(k.startswith("base_model.") for k in sd) and "v_head.0.weight" in sd
    again = AutoModelForCausalLMWithValueHead.from_pretrained(str(tmp_path)).eval()
    l2, *_, v2 = again(ids, mask, position_ids=pos)
    torch.testing.assert_close(l2, logits)
    torch.testing.assert_close(v2, value)
    for (k1, t1), (k2, t2) in zip(sorted(model.state_dict().items()), sorted(again.state_dict().items())):
        assert k1 == k2 and torch.equal(t1, t2)


@pytest.mark.parametrize("family", list(CAUSAL))
@pytest.mark.parametrize("k", [1, 2])
def test_hydra_equals_forward_at_init(family, k, tmp_path):
    torch.manual_seed(0)
    model = AutoModelForCausalLMWithHydraValueHead.from_config(CAUSAL[family], num_layers_unfrozen=k).eval()
    ids, mask, pos = _inputs()
    out = model(ids, mask, position_ids=pos, return_dict=True)
    hydra = model.forward_hydra(ids, mask, position_ids=pos, return_dict=True, output_hidden_states=True)
    torch.testing.assert_close(hydra.logits.float(), out.logits.float(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(hydra.hidden_states[-1], out.hidden_states[-1], atol=1e-5, rtol=1e-5)
    assert torch.equal(model.forward_hydra(ids, mask, position_ids=pos, return_dict=False).float(), hydra.logits.float())
    assert all(not p.requires_grad for p in model.frozen_head.parameters())
    assert isinstance(model.frozen_head, hf_get_branch_class(model.config))
    torch.testing.assert_close(model.frozen_head.lm_head.weight, model.base_model.lm_head.weight)
    # single-pass scoring agrees with the two separate forwards
    labels = torch.cat([ids[:, 1:], ids.new_full((3, 1), -1)], 1)
    lp, val, ref_lp, trunk = model.score(ids, mask, pos, labels)
    ref = torch.log_softmax(out.logits[:, :-1].float(), -1).gather(-1, ids[:, 1:, None]).squeeze(-1)
    torch.testing.assert_close(lp[:, :-1], ref, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(ref_lp, lp, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(val, out.value, atol=1e-5, rtol=1e-5)
    # checkpoint carries the branch; loading WITHOUT the argument rebuilds it from the key indices
    model.save_pretrained(str(tmp_path))
    again = AutoModelForCausalLMWithHydraValueHead.from_pretrained(str(tmp_path)).eval()
    assert again.num_layers_unfrozen == k and again.frozen_head is not None
    torch.testing.assert_close(again.forward_hydra(ids, mask, position_ids=pos, return_dict=False), hydra.logits)


def test_value_branch_with_own_layers():
    model = AutoModelForCausalLMWithHydraValueHead.from_config(CAUSAL["gpt2"], num_layers_unfrozen=1, num_value_layers_unfrozen=2).eval()
    ids, mask, pos = _inputs()
    out = model(ids, mask, position_ids=pos, return_dict=True)
    assert out.value.shape == (3, 7) and any(p.requires_grad for p in model.v_head.parameters())
    assert any(k.startswith("v_head.decoder_blocks.1.") for k in model.state_dict())


def test_unsupported_branch_architecture():
    with pytest.raises(ValueError):
        hf_get_branch_class(dict(model_type="xlnet"))


def test_seq2seq_wrappers(tmp_path):
    torch.manual_seed(0)
    model = AutoModelForSeq2SeqLMWithHydraValueHead.from_config(T5, num_layers_unfrozen=1).eval()
    ids, mask, _ = _inputs()
    dec = torch.randint(2, 60, (3, 5))
    out = model(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True)
    assert out.logits.shape == (3, 5, 64) and out.value.shape == (3, 5)
    hydra = model.forward_hydra(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True)
    torch.testing.assert_close(hydra.logits, out.logits, atol=1e-5, rtol=1e-5)
    gen = model.generate(ids, attention_mask=mask, max_new_tokens=4, do_sample=False)
    assert gen.shape[0] == 3 and gen[:, 0].eq(0).all()
    model.save_pretrained(str(tmp_path))
    again = AutoModelForSeq2SeqLMWithHydraValueHead.from_pretrained(str(tmp_path)).eval()
    torch.testing.assert_close(again(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True).logits, out.logits)
    plain = AutoModelForSeq2SeqLMWithValueHead.from_config(T5)
    assert plain(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True).value.shape == (3, 5)


# ---- ILQL ---------------------------------------------------------------------------------------------------------------
@settings(deadline=None, max_examples=30)
@given(st.integers(1, 4), st.integers(1, 9), st.integers(1, 5), st.integers(1, 6))
def test_batched_index_select_matches_loop(B, T, H, n):
    x = torch.randn(B, T, H)
    idx = torch.randint(0, T, (B, n))
    out = batched_index_select(x, idx, 1)
    for b in range(B):
        assert torch.equal(out[b], x[b, idx[b]])


def test_topk_mask():
    x = torch.tensor([[1.0, 5.0, 3.0, 2.0]])
    assert topk_mask(x, 2).tolist() == [[float("-inf"), 5.0, 3.0, float("-inf")]]
    assert torch.equal(topk_mask(x, 10), x)


@settings(deadline=None, max_examples=20)
@given(st.integers(1, 3), st.integers(2, 8), st.booleans(), st.floats(0.0, 1.0))
def test_ilql_heads_shapes_and_polyak(B, T, two_qs, alpha):
    heads = ILQLHeads(16, 11, two_qs, alpha, torch.float32)
    hs = torch.randn(B, T, 16)
    s_ix = torch.randint(0, T, (B, 4))
    a_ix = torch.randint(0, T, (B, 3))
    qs, tqs, vs = heads(hs, states_ixs=s_ix, actions_ixs=a_ix)
    assert len(qs) == len(tqs) == (2 if two_qs else 1) and qs[0].shape == (B, 3, 11) and vs.shape == (B, 4, 1)
    assert all(not p.requires_grad for p in heads.target_q_heads.parameters())
    with torch.no_grad():
        for p in heads.q_heads.parameters():
            p.add_(1.0)
    before = [p.clone() for p in heads.target_q_heads.parameters()]
    heads.sync_target_q_heads()
    for b, t, q in zip(before, heads.target_q_heads.parameters(), heads.q_heads.parameters()):
        torch.testing.assert_close(t, alpha * q + (1 - alpha) * b)


def _ilql_batch(B=3, T=8, A=4, V=64):
    torch.manual_seed(2)
    ids = torch.randint(1, V, (B, T))
    actions_ixs = torch.sort(torch.randint(0, T - 1, (B, A)), 1).values
    states_ixs = torch.cat([actions_ixs, torch.full((B, 1), T - 1)], 1)
    dones = torch.ones(B, A + 1, dtype=torch.long)
    dones[:, -1] = 0
    return ILQLBatch(ids, torch.ones(B, T, dtype=torch.long), torch.randn(B, A), states_ixs, actions_ixs, dones)


def test_ilql_loss_gathered_form_equals_materialised_form():
    cfg = ILQLConfig(name="ilqlconfig", tau=0.7, gamma=0.99, cql_scale=0.1, awac_scale=1, alpha=0.1, beta=0.5,
                     steps_for_target_q_sync=1, two_qs=True, gen_kwargs={})
    model = AutoModelForCausalLMWithILQLHeads.from_config(CAUSAL["gpt2"], two_qs=True, alpha=0.1)
    batch = _ilql_batch()
    logits, qs, tqs, vs, _ = model(batch.input_ids, batch.attention_mask, actions_ixs=batch.actions_ixs, states_ixs=batch.states_ixs)
    loss_full, stats_full = cfg.loss((logits, (qs, tqs, vs)), batch)
    loss_parts, stats_parts = cfg.loss(model.loss_parts(batch), batch)
    torch.testing.assert_close(loss_parts, loss_full, atol=1e-5, rtol=1e-5)
    assert set(stats_full) == set(stats_parts) and "losses/loss_awac" in stats_full and "qvalues/1/mean" in stats_full
    # manual recomputation of the Q term
    actions = batch.input_ids[:, 1:].gather(1, batch.actions_ixs)
    V = vs[..., 0]
    target = batch.rewards + 0.99 * (V[:, 1:] * batch.dones[:, 1:]).detach()
    lq = sum((((q.gather(-1, actions[..., None]).squeeze(-1) - target) * batch.dones[:, :-1]) ** 2).sum() / batch.dones[:, :-1].sum() for q in qs)
    torch.testing.assert_close(stats_full["losses/loss_q"], lq.detach(), atol=1e-5, rtol=1e-5)
    loss_parts.backward()
    assert model.ilql_heads.q_heads[0][0].weight.grad is not None and model.ilql_heads.target_q_heads[0][0].weight.grad is None


def test_ilql_wrappers_generate_and_save(tmp_path):
    model = AutoModelForCausalLMWithILQLHeads.from_config(CAUSAL["gpt2"], two_qs=True, alpha=0.5).eval()
    ids, mask, _ = _inputs()
    out = model.generate(ids, mask, max_new_tokens=5, beta=1.0, top_k=5, temperature=0.7, eos_token_id=63, pad_token_id=63)
    assert out.shape[0] == 3 and out.shape[1] <= 12
    greedy = model.generate(ids, mask, max_new_tokens=3, temperature=0.0, eos_token_id=63, pad_token_id=63)
    assert torch.equal(greedy, model.generate(ids, mask, max_new_tokens=3, temperature=0.0, eos_token_id=63, pad_token_id=63))
    logit_mask = torch.zeros(64, 64, dtype=torch.bool)
    logit_mask[:, 10:] = True  # only tokens < 10 are ever allowed
    masked = model.generate(ids, mask, max_new_tokens=4, logit_mask=logit_mask, eos_token_id=63, pad_token_id=63)
    assert (masked[:, 7:] < 10).logical_or(masked[:, 7:] == 63).all()
    model.save_pretrained(str(tmp_path))
    again = AutoModelForCausalLMWithILQLHeads.from_pretrained(str(tmp_path), two_qs=True).eval()
    for (k1, t1), (k2, t2) in zip(sorted(model.state_dict().items()), sorted(again.state_dict().items())):
        assert k1 == k2 and torch.equal(t1, t2)
    s2s = AutoModelForSeq2SeqLMWithILQLHeads.from_config(T5, two_qs=False).eval()
    gen = s2s.generate(ids, mask, max_new_tokens=3, eos_token_id=1, pad_token_id=0)
    assert gen.shape[0] == 3 and gen.shape[1] <= 4


def test_beam_search_exact_when_width_covers_vocab_and_never_worse_than_greedy():
    from trlx_b200.models.generation import generate
    from trlx_b200.nn.arch import spec_from_hf_config
    from trlx_b200.nn.transformer import CausalLM

    torch.manual_seed(0)
    V = 30
    spec = spec_from_hf_config(dict(model_type="gpt2", vocab_size=V, n_embd=32, n_layer=2, n_head=2, n_positions=64, eos_token_id=V - 1))
    m = CausalLM(spec).eval()
    ids = torch.randint(0, V - 2, (3, 5))
    mask = torch.ones_like(ids)
    mask[0, :2] = 0

    def seq_logprob(seq, rows=slice(None)):
        am = torch.cat([mask[rows], torch.ones(seq.shape[0], seq.shape[1] - 5, dtype=torch.long)], 1)
        lp = torch.log_softmax(m(input_ids=seq, attention_mask=am).logits[:, :-1].float(), -1)
        return lp.gather(-1, seq[:, 1:, None]).squeeze(-1)[:, 4:].sum(-1)

    with torch.no_grad():
        greedy = generate(m, ids, mask, max_new_tokens=5, do_sample=False, eos_token_id=None, pad_token_id=0)
        beam = generate(m, ids, mask, max_new_tokens=5, num_beams=4, eos_token_id=None, pad_token_id=0)
        assert beam.shape == greedy.shape and (seq_logprob(beam) >= seq_logprob(greedy) - 1e-4).all()
        # beam width = vocabulary size makes a 2-step search exhaustive
        wide = generate(m, ids[:1], mask[:1], max_new_tokens=2, num_beams=V, eos_token_id=None, pad_token_id=0)
        best = max(((seq_logprob(torch.cat([ids[:1], torch.tensor([[a, b]])], 1), slice(0, 1)).item(), a, b)
                    for a in range(V) for b in range(V)))
        assert wide[0, -2:].tolist() == [best[1], best[2]]


def test_beam_search_seq2seq_runs_and_pads_after_eos():
    from trlx_b200.models.generation import generate
    from trlx_b200.models.modeling_base import build_base_model

    torch.manual_seed(0)
    t5 = build_base_model(dict(model_type="t5", vocab_size=40, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4, eos_token_id=1,
                               pad_token_id=0, decoder_start_token_id=0), "seq2seq").eval()
    ids = torch.randint(2, 40, (2, 6))
    out = generate(t5, ids, torch.ones_like(ids), max_new_tokens=8, num_beams=3, eos_token_id=1, pad_token_id=0)
    assert out.shape[0] == 2 and out[:, 0].eq(0).all()
    for row in out.tolist():
        if 1 in row:
            assert all(t == 0 for t in row[row.index(1) + 1:])


def test_activation_checkpointing_matches_plain_backward():
    from trlx_b200.nn.arch import spec_from_hf_config
    from trlx_b200.nn.transformer import CausalLM

    torch.manual_seed(0)
    spec = spec_from_hf_config(dict(model_type="llama", vocab_size=40, hidden_size=32, num_hidden_layers=3, num_attention_heads=4,
                                    num_key_value_heads=2, intermediate_size=64, max_position_embeddings=32))
    m = CausalLM(spec)
    ids = torch.randint(0, 40, (2, 9))
    mask = torch.ones_like(ids)
    mask[0, :3] = 0
    m(input_ids=ids, attention_mask=mask, labels=ids).loss.backward()
    ref = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad()
    m.gradient_checkpointing_enable()
    out = m(input_ids=ids, attention_mask=mask, labels=ids)
    out.loss.backward()
    for n, p in m.named_parameters():
        torch.testing.assert_close(p.grad, ref[n], atol=1e-6, rtol=1e-5)


def test_push_to_hub_fails_cleanly_when_the_hub_is_disabled(monkeypatch):
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithValueHead

    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    model = AutoModelForCausalLMWithValueHead.from_config(dict(model_type="gpt2", vocab_size=20, n_embd=8, n_layer=1, n_head=2,
                                                                n_positions=16))
    with pytest.raises(RuntimeError, match="disabled"):
        model.push_to_hub("someone/some-model")


def test_ops_attention_cpu_path_matches_manual_softmax():
    from trlx_b200 import ops

    torch.manual_seed(0)
    q, k, v = (torch.randn(2, 3, 5, 8) for _ in range(3))
    causal = ops.attention(q, k, v, None, causal=True)
    s = q @ k.transpose(-1, -2) * 8 ** -0.5
    mask = torch.ones(5, 5, dtype=torch.bool).tril()
    ref = torch.softmax(s.masked_fill(~mask, float("-inf")), -1) @ v
    torch.testing.assert_close(causal, ref, atol=1e-5, rtol=1e-5)
    bias = torch.zeros(2, 1, 5, 5).masked_fill(~mask, torch.finfo(torch.float32).min)
    torch.testing.assert_close(ops.attention(q, k, v, bias, scale=8 ** -0.5), ref, atol=1e-5, rtol=1e-5)


def test_norm_module_cpu_matches_functional():
    from trlx_b200.nn.arch import spec_from_hf_config
    from trlx_b200.nn.transformer import Norm

    spec = spec_from_hf_config(dict(model_type="gpt2", vocab_size=20, n_embd=16, n_layer=1, n_head=2, n_positions=8))
    norm = Norm(spec)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
    x = torch.randn(3, 4, 16)
    torch.testing.assert_close(norm(x), torch.nn.functional.layer_norm(x, (16,), norm.weight, norm.bias, norm.eps))


def test_from_pretrained_accepts_base_model_checkpoints_and_never_stays_silently_random(tmp_path):
    """A GPT-2 directory saved from the *base* class (``wte.weight``, ``h.0.*`` — no ``transformer.`` prefix, no
    ``lm_head``) must load; a checkpoint that matches nothing must raise instead of leaving random weights."""
    import json

    import transformers

    cfg = transformers.GPT2Config(vocab_size=64, n_embd=32, n_layer=2, n_head=2, n_positions=32)
    torch.manual_seed(3)
    hf = transformers.GPT2LMHeadModel(cfg).eval()
    d = tmp_path / "base_only"
    d.mkdir()
    json.dump({**cfg.to_dict(), "architectures": ["GPT2Model"]}, open(d / "config.json", "w"))
    torch.save(hf.transformer.state_dict(), d / "pytorch_model.bin")  # keys: wte.weight, h.0.ln_1.weight, …
    model = AutoModelForCausalLMWithValueHead.from_pretrained(str(d)).eval()
    ids, mask, pos = _inputs(B=2, T=6)
    with torch.no_grad():
        want = hf(ids, attention_mask=mask, position_ids=pos).logits
        got = model(ids, attention_mask=mask, position_ids=pos, return_dict=True).logits
    assert torch.allclose(got[mask.bool()], want[mask.bool()], atol=1e-4)

    bad = tmp_path / "garbage"
    bad.mkdir()
    json.dump(cfg.to_dict(), open(bad / "config.json", "w"))
    torch.save({"something.else": torch.zeros(3)}, bad / "pytorch_model.bin")
    with pytest.raises(ValueError, match="randomly initialised"):
        AutoModelForCausalLMWithValueHead.from_pretrained(str(bad))


def test_sampling_filter_matches_hf_logits_warpers():
    """``top_k_top_p_filter`` (the PyTorch sampler, and the oracle of the CUDA sampling kernel) keeps exactly the tokens HF's
    ``TopKLogitsWarper`` → ``TopPLogitsWarper`` chain keeps — what the reference's ``generate`` kwargs mean."""
    warpers = pytest.importorskip("transformers.generation.logits_process")
    from trlx_b200.models.generation import top_k_top_p_filter

    torch.manual_seed(0)
    logits = torch.randn(5, 50) * 3
    logits[2, :7] = logits[2, 0]  # ties
    for t in (0.7, 1.0, 1.3):
        for k in (0, 1, 5, 50):
            for p in (1.0, 0.9, 0.5, 0.05):
                x = logits / t
                ref = x.clone()
                if k > 0:
                    ref = warpers.TopKLogitsWarper(k)(None, ref)
                if p < 1.0:
                    ref = warpers.TopPLogitsWarper(p)(None, ref)
                got = top_k_top_p_filter(x.clone(), k, p)
                assert torch.equal(torch.isinf(ref), torch.isinf(got)), (t, k, p)
                torch.testing.assert_close(got[~torch.isinf(got)], ref[~torch.isinf(ref)])


def test_beam_search_matches_hf_generate():
    """``generate(num_beams=…)`` returns the sequences HF's beam search returns for the same weights (left-padded prompts, several
    beam widths and length penalties, EOS reachable)."""
    transformers = pytest.importorskip("transformers")
    from trlx_b200.models.generation import generate
    from trlx_b200.models.modeling_base import build_base_model, import_base_state_dict

    cfgd = dict(model_type="gpt2", vocab_size=24, n_embd=32, n_layer=2, n_head=2, n_positions=64, eos_token_id=23, bos_token_id=23,
                pad_token_id=23)
    for seed in (0, 1):
        torch.manual_seed(seed)
        hf = transformers.AutoModelForCausalLM.from_config(transformers.AutoConfig.for_model(**cfgd)).eval()
        ours = build_base_model(cfgd).eval()
        import_base_state_dict(ours, hf.state_dict(), strict=True)
        ids = torch.randint(0, 22, (3, 5), generator=torch.Generator().manual_seed(seed + 10))
        mask = torch.ones_like(ids)
        mask[0, :2] = 0
        for nb, lp in ((2, 1.0), (4, 1.0), (4, 0.0), (3, 2.0)):
            with torch.no_grad():
                a = hf.generate(ids, attention_mask=mask, num_beams=nb, max_new_tokens=7, do_sample=False, length_penalty=lp,
                                early_stopping=False, pad_token_id=23, eos_token_id=23)
                b = generate(ours, ids, attention_mask=mask, num_beams=nb, max_new_tokens=7, do_sample=False, length_penalty=lp,
                             pad_token_id=23, eos_token_id=23)
            assert a.shape == b.shape and torch.equal(a, b), (seed, nb, lp, a, b)


def test_greedy_generation_controls_match_hf_generate():
    """Length controls and stopping of ``generate`` (``max_new_tokens`` / ``max_length`` / ``min_new_tokens`` / ``min_length`` /
    several EOS ids, left-padded prompts, finished rows padded) give HF's sequences for the same weights."""
    transformers = pytest.importorskip("transformers")
    from trlx_b200.models.generation import generate
    from trlx_b200.models.modeling_base import build_base_model, import_base_state_dict

    cfgd = dict(model_type="gpt2", vocab_size=24, n_embd=32, n_layer=2, n_head=2, n_positions=64, eos_token_id=23, bos_token_id=23,
                pad_token_id=23)
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        hf = transformers.AutoModelForCausalLM.from_config(transformers.AutoConfig.for_model(**cfgd)).eval()
        ours = build_base_model(cfgd).eval()
        import_base_state_dict(ours, hf.state_dict(), strict=True)
        ids = torch.randint(0, 22, (4, 6), generator=torch.Generator().manual_seed(seed + 3))
        mask = torch.ones_like(ids)
        mask[1, :3] = 0
        for kw in (dict(max_new_tokens=8), dict(max_length=11), dict(max_new_tokens=8, min_new_tokens=5),
                   dict(max_new_tokens=6, min_length=10), dict(max_new_tokens=8, eos_token_id=[23, 5])):
            kw = {"eos_token_id": 23, **kw}
            with torch.no_grad():
                a = hf.generate(ids, attention_mask=mask, do_sample=False, pad_token_id=23, **kw)
                b = generate(ours, ids, attention_mask=mask, do_sample=False, pad_token_id=23, **kw)
            assert a.shape == b.shape and torch.equal(a, b), (seed, kw, a, b)


@pytest.mark.parametrize("family", list(CAUSAL) + ["t5_tied", "t5_untied"])
def test_cached_greedy_decoding_matches_hf_for_every_family(family):
    """Incremental (KV-cache) greedy decoding from left-padded prompts produces HF's tokens for every decoder family — rotary
    offsets, ALiBi, local attention windows, multi-query caches, post-LN OPT — and for T5 (tied and untied embeddings)."""
    transformers = pytest.importorskip("transformers")
    from trlx_b200.models.generation import generate
    from trlx_b200.models.modeling_base import build_base_model, import_base_state_dict

    g = torch.Generator().manual_seed(3)
    ids = torch.randint(2, 60, (3, 6), generator=g)
    mask = torch.ones_like(ids)
    torch.manual_seed(0)
    if family.startswith("t5"):
        cfgd = dict(model_type="t5", vocab_size=64, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_decoder_layers=2, num_heads=4,
                    eos_token_id=1, pad_token_id=0, decoder_start_token_id=0, tie_word_embeddings=family == "t5_tied")
        hf = transformers.AutoModelForSeq2SeqLM.from_config(transformers.AutoConfig.for_model(**cfgd)).eval()
        ours = build_base_model(cfgd, "seq2seq").eval()
        import_base_state_dict(ours, hf.state_dict(), strict=False)
        mask[0, 4:] = 0  # encoder inputs are right-padded
        with torch.no_grad():
            a = hf.generate(ids, attention_mask=mask, do_sample=False, max_new_tokens=8)
            b = generate(ours, ids, attention_mask=mask, do_sample=False, max_new_tokens=8, pad_token_id=0, eos_token_id=1,
                         decoder_start_token_id=0)
    else:
        cfgd = dict(CAUSAL[family], eos_token_id=63, pad_token_id=63, bos_token_id=63)
        hf = transformers.AutoModelForCausalLM.from_config(transformers.AutoConfig.for_model(**cfgd)).eval()
        ours = build_base_model(cfgd).eval()
        import_base_state_dict(ours, hf.state_dict(), strict=True)
        mask[0, :2] = 0
        with torch.no_grad():
            a = hf.generate(ids, attention_mask=mask, do_sample=False, max_new_tokens=8, pad_token_id=63, eos_token_id=63)
            b = generate(ours, ids, attention_mask=mask, do_sample=False, max_new_tokens=8, pad_token_id=63, eos_token_id=63)
    assert a.shape == b.shape and torch.equal(a, b), (a, b)
