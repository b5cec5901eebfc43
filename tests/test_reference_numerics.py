"""Numerical parity with the UNMODIFIED reference (installed under ``baseline/_ref`` for the bench's reference arm): its PPO
advantage / loss maths, ILQL loss, log-prob gather, whitening and running moments are evaluated in a separate interpreter (the
repository's ``trlx`` alias package would shadow it here) on fixed random inputs, and compared with this framework's
implementations of the same entry points.  Skipped when the reference is not installed."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "trlx")), reason="baseline/_ref is not installed")

PRODUCER = textwrap.dedent("""
    import sys, torch
    import transformers  # before the stand-in `accelerate` is importable (its version probe runs at import time)
    sys.path.insert(0, {shims!r}); sys.path.insert(0, {ref!r})
    import transformers_compat  # noqa: F401
    import trlx
    assert {ref!r} in trlx.__file__, trlx.__file__
    from trlx.data.default_configs import default_ppo_config, default_ilql_config
    from trlx.data.ilql_types import ILQLBatch
    from trlx.utils.modeling import logprobs_of_labels, whiten, RunningMoments, get_tensor_stats
    inp = torch.load({inp!r})
    out = {{}}
    ppo = default_ppo_config().method
    adv, ret = ppo.get_advantages_and_returns(inp["values"], inp["rewards"], inp["values"].shape[1])
    loss, stats = ppo.loss(inp["logprobs"], inp["new_values"], inp["old_logprobs"], inp["values"], adv, ret, inp["mask"])
    out["ppo"] = dict(adv=adv, ret=ret, loss=loss.detach(), stats={{k: float(v) for k, v in stats.items()}})
    adv_raw, _ = ppo.get_advantages_and_returns(inp["values"], inp["rewards"], inp["values"].shape[1], use_whitening=False)
    out["adv_raw"] = adv_raw
    ilql = default_ilql_config().method
    b = inp["ilql"]
    batch = ILQLBatch(b["input_ids"], b["attention_mask"], b["rewards"], b["states_ixs"], b["actions_ixs"], b["dones"])
    loss, stats = ilql.loss((b["logits"], (b["qs"], b["target_qs"], b["vs"])), batch)
    out["ilql"] = dict(loss=loss.detach(), stats={{k: float(v) for k, v in stats.items()}})
    out["logprobs"] = logprobs_of_labels(inp["logits"], inp["labels"])
    out["whiten"] = whiten(inp["values"], shift_mean=True)
    out["whiten_noshift"] = whiten(inp["values"], shift_mean=False)
    rm = RunningMoments()
    out["moments"] = [tuple(float(x) for x in rm.update(chunk)) + (float(rm.mean), float(rm.std)) for chunk in inp["chunks"]]
    from trlx.models.modeling_ppo import AdaptiveKLController, FixedKLController
    a, f = AdaptiveKLController(0.05, 6.0, 1000), FixedKLController(0.2)
    trace = []
    for kl, n in ((8.0, 128), (1.0, 64), (40.0, 256), (-3.0, 32), (6.0, 128)):
        a.update(kl, n); f.update(kl, n)
        trace.append((float(a.value), float(f.value)))
    out["kl_ctl"] = trace
    from trlx.utils import significant, filter_non_scalars
    from trlx.utils.modeling import flatten_dict, get_tensor_stats
    from trlx.models.modeling_ilql import topk_mask, batched_index_select
    out["significant"] = [significant(x) for x in (0, 1234.5678, 0.000123456, -98.7654, 3, 1e-12, 12345678.9)]
    nested = dict(a=1.5, b=dict(c=2, d=dict(e=torch.tensor(3.0))), f="text", g=[1, 2])
    out["flat"] = {{k: (float(v) if isinstance(v, (int, float, torch.Tensor)) else repr(v)) for k, v in flatten_dict(nested).items()}}
    out["scalars"] = sorted(filter_non_scalars(dict(a=1, b=2.5, d=[1], e=torch.tensor(1.0), f=object())).keys())
    out["topk"] = topk_mask(inp["logits"][0], 5)
    out["bis"] = batched_index_select(inp["logits"], inp["labels"][:, :4] % inp["logits"].shape[1], 1)
    out["tstats"] = {{k: float(v) for k, v in get_tensor_stats(inp["values"], inp["mask"], inp["mask"].sum()).items()}}
    from trlx.utils.modeling import (freeze_bottom_causal_layers, hf_get_decoder_blocks, hf_get_hidden_size,
                                     hf_get_num_hidden_layers, hf_get_decoder_final_norm, hf_get_lm_head)
    fams = {{}}
    for name, kw in {families!r}.items():
        row = {{}}
        for k in (0, 1, 2, -1):
            torch.manual_seed(0)
            hf = transformers.AutoModelForCausalLM.from_config(transformers.AutoConfig.for_model(**kw))
            freeze_bottom_causal_layers(hf, k)
            row[k] = sum(p.numel() for p in hf.parameters() if p.requires_grad)
        row["blocks"] = len(hf_get_decoder_blocks(hf)); row["hidden"] = hf_get_hidden_size(hf.config)
        row["layers"] = hf_get_num_hidden_layers(hf.config)
        row["norm_numel"] = sum(p.numel() for p in hf_get_decoder_final_norm(hf).parameters())
        row["head_shape"] = tuple(hf_get_lm_head(hf).weight.shape)
        fams[name] = row
    out["families"] = fams
    from trlx.data.configs import TRLConfig
    from trlx.data.default_configs import default_sft_config
    cfgs = {{}}
    for path in {yamls!r}:
        cfgs[path] = TRLConfig.load_yaml(path).to_dict()
    cfgs["default_ppo"], cfgs["default_ilql"], cfgs["default_sft"] = (default_ppo_config().to_dict(), default_ilql_config().to_dict(),
                                                                     default_sft_config().to_dict())
    cfgs["updated"] = TRLConfig.update(default_ppo_config().to_dict(), {{"train.seq_length": 77, "method.gamma": 0.5,
                                                                        "optimizer.kwargs.lr": 1e-3}}).to_dict()
    cfgs["evolved"] = default_ilql_config().evolve(train=dict(batch_size=3), method=dict(tau=0.9, gen_kwargs=dict(beta=2))).to_dict()
    out["configs"] = cfgs
    torch.save(out, {outp!r})
""")


FAMILIES = {
    "gpt2": dict(model_type="gpt2", vocab_size=64, n_embd=32, n_layer=3, n_head=2, n_positions=64),
    "gptj": dict(model_type="gptj", vocab_size=64, n_embd=32, n_layer=3, n_head=2, rotary_dim=8, n_positions=64),
    "gpt_neox": dict(model_type="gpt_neox", vocab_size=64, hidden_size=32, num_hidden_layers=3, num_attention_heads=2,
                     intermediate_size=64, max_position_embeddings=64),
    "llama": dict(model_type="llama", vocab_size=64, hidden_size=32, num_hidden_layers=3, num_attention_heads=4,
                  num_key_value_heads=2, intermediate_size=48, max_position_embeddings=64),
    "opt": dict(model_type="opt", vocab_size=64, hidden_size=32, num_hidden_layers=3, num_attention_heads=2, ffn_dim=64,
                max_position_embeddings=64, word_embed_proj_dim=32),
    "bloom": dict(model_type="bloom", vocab_size=64, hidden_size=32, n_layer=3, n_head=2),
    "gpt_bigcode": dict(model_type="gpt_bigcode", vocab_size=64, n_embd=32, n_layer=3, n_head=2, n_positions=64),
    "gpt_neo": dict(model_type="gpt_neo", vocab_size=64, hidden_size=32, num_layers=4, num_heads=2, max_position_embeddings=64,
                    attention_types=[[["global", "local"], 2]], window_size=4),
}


def _yamls():
    """TRLConfig YAML files shipped with the reference checkout (when it is available next to the installed copy)."""
    root = os.environ.get("TRLX_REFERENCE", "/root/reference")
    cands = [os.path.join(root, "configs", "test_config.yml"),
             os.path.join(root, "examples", "experiments", "grounded_program_synthesis", "configs", "trlx_ppo_config.yml")]
    return [c for c in cands if os.path.exists(c)]


def _inputs():
    g = torch.Generator().manual_seed(0)
    B, R, V = 6, 9, 37
    mask = torch.ones(B, R)
    mask[0, 6:] = 0
    mask[3, 2:] = 0
    inp = dict(values=torch.randn(B, R, generator=g), rewards=torch.randn(B, R, generator=g) * 0.3 * mask,
               logprobs=-torch.rand(B, R, generator=g) * 3, old_logprobs=-torch.rand(B, R, generator=g) * 3,
               new_values=torch.randn(B, R, generator=g), mask=mask,
               logits=torch.randn(B, R, V, generator=g), labels=torch.randint(0, V, (B, R), generator=g),
               chunks=[torch.randn(11, generator=g) * 2 + 1, torch.randn(7, generator=g) - 3, torch.randn(16, generator=g)])
    # ILQL: 4 sequences of 8 tokens, 5 actions / 6 states each
    T, A = 8, 5
    ids = torch.randint(0, V, (4, T), generator=g)
    actions_ixs = torch.arange(2, 2 + A).repeat(4, 1)
    states_ixs = torch.arange(2, 3 + A).repeat(4, 1)
    dones = torch.ones(4, A + 1, dtype=torch.long)
    dones[:, -1] = 0
    dones[1, 3:] = 0
    inp["ilql"] = dict(input_ids=ids, attention_mask=torch.ones_like(ids), rewards=torch.randn(4, A, generator=g),
                       states_ixs=states_ixs, actions_ixs=actions_ixs, dones=dones, logits=torch.randn(4, T, V, generator=g),
                       qs=tuple(torch.randn(4, A, V, generator=g) for _ in range(2)),
                       target_qs=tuple(torch.randn(4, A, V, generator=g) for _ in range(2)), vs=torch.randn(4, A + 1, 1, generator=g))
    return inp


@pytest.fixture(scope="module")
def reference_outputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("refnum")
    inp, outp = str(d / "in.pt"), str(d / "out.pt")
    torch.save(_inputs(), inp)
    code = PRODUCER.format(shims=os.path.join(ROOT, "baseline", "shims"), ref=REF, inp=inp, outp=outp, yamls=_yamls(), families=FAMILIES)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    res = subprocess.run([sys.executable, "-c", code], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    return torch.load(outp, weights_only=False)


def test_ppo_advantages_and_loss_match_the_reference(reference_outputs):
    from trlx_b200.data.default_configs import default_ppo_config

    inp, ref = _inputs(), reference_outputs
    ppo = default_ppo_config().method
    adv, ret = ppo.get_advantages_and_returns(inp["values"], inp["rewards"], inp["values"].shape[1])
    torch.testing.assert_close(adv, ref["ppo"]["adv"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(ret, ref["ppo"]["ret"], atol=1e-5, rtol=1e-5)
    raw, _ = ppo.get_advantages_and_returns(inp["values"], inp["rewards"], inp["values"].shape[1], use_whitening=False)
    torch.testing.assert_close(raw, ref["adv_raw"], atol=1e-5, rtol=1e-5)
    loss, stats = ppo.loss(inp["logprobs"], inp["new_values"], inp["old_logprobs"], inp["values"], adv, ret, inp["mask"])
    torch.testing.assert_close(loss.detach(), ref["ppo"]["loss"], atol=1e-5, rtol=1e-5)
    mine = {k: float(v) for k, v in stats.items()}
    assert set(ref["ppo"]["stats"]) <= set(mine), set(ref["ppo"]["stats"]) - set(mine)
    for k, v in ref["ppo"]["stats"].items():
        assert abs(mine[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, mine[k], v)


def test_ilql_loss_matches_the_reference(reference_outputs):
    from trlx_b200.data.default_configs import default_ilql_config
    from trlx_b200.data.ilql_types import ILQLBatch

    b, ref = _inputs()["ilql"], reference_outputs["ilql"]
    batch = ILQLBatch(b["input_ids"], b["attention_mask"], b["rewards"], b["states_ixs"], b["actions_ixs"], b["dones"])
    loss, stats = default_ilql_config().method.loss((b["logits"], (b["qs"], b["target_qs"], b["vs"])), batch)
    torch.testing.assert_close(loss.detach(), ref["loss"], atol=1e-4, rtol=1e-5)
    mine = {k: float(v) for k, v in stats.items()}
    assert set(ref["stats"]) <= set(mine), set(ref["stats"]) - set(mine)
    for k, v in ref["stats"].items():
        assert abs(mine[k] - v) <= 2e-4 * max(1.0, abs(v)), (k, mine[k], v)


def test_statistics_helpers_match_the_reference(reference_outputs):
    from trlx_b200.utils.modeling import RunningMoments, logprobs_of_labels, whiten

    inp, ref = _inputs(), reference_outputs
    torch.testing.assert_close(logprobs_of_labels(inp["logits"], inp["labels"]), ref["logprobs"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(whiten(inp["values"], shift_mean=True), ref["whiten"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(whiten(inp["values"], shift_mean=False), ref["whiten_noshift"], atol=1e-5, rtol=1e-5)
    rm = RunningMoments()
    for chunk, want in zip(inp["chunks"], ref["moments"]):
        got = tuple(float(x) for x in rm.update(chunk)) + (float(rm.mean), float(rm.std))
        for a, b in zip(got, want):
            assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (got, want)


def test_kl_controllers_match_the_reference(reference_outputs):
    from trlx_b200.models.modeling_ppo import AdaptiveKLController, FixedKLController

    a, f = AdaptiveKLController(0.05, 6.0, 1000), FixedKLController(0.2)
    for (kl, n), (want_a, want_f) in zip(((8.0, 128), (1.0, 64), (40.0, 256), (-3.0, 32), (6.0, 128)), reference_outputs["kl_ctl"]):
        a.update(kl, n)
        f.update(kl, n)
        assert abs(float(a.value) - want_a) < 1e-9 and abs(float(f.value) - want_f) < 1e-12


def test_small_utilities_match_the_reference(reference_outputs):
    from trlx_b200.models.modeling_ilql import batched_index_select, topk_mask
    from trlx_b200.utils import filter_non_scalars, significant
    from trlx_b200.utils.modeling import flatten_dict, get_tensor_stats

    inp, ref = _inputs(), reference_outputs
    assert [significant(x) for x in (0, 1234.5678, 0.000123456, -98.7654, 3, 1e-12, 12345678.9)] == ref["significant"]
    nested = dict(a=1.5, b=dict(c=2, d=dict(e=torch.tensor(3.0))), f="text", g=[1, 2])
    flat = {k: (float(v) if isinstance(v, (int, float, torch.Tensor)) else repr(v)) for k, v in flatten_dict(nested).items()}
    assert flat == ref["flat"]
    assert sorted(filter_non_scalars(dict(a=1, b=2.5, d=[1], e=torch.tensor(1.0), f=object())).keys()) == ref["scalars"]
    torch.testing.assert_close(topk_mask(inp["logits"][0], 5), ref["topk"])
    torch.testing.assert_close(batched_index_select(inp["logits"], inp["labels"][:, :4] % inp["logits"].shape[1], 1), ref["bis"])
    mine = {k: float(v) for k, v in get_tensor_stats(inp["values"], inp["mask"], inp["mask"].sum()).items()}
    assert set(mine) == set(ref["tstats"])
    for k, v in ref["tstats"].items():
        assert abs(mine[k] - v) < 1e-5, (k, mine[k], v)


def test_config_trees_match_the_reference(reference_outputs):
    """Default configs, YAML loading, dotted ``update`` and ``evolve``: the resulting trees contain everything the reference's do,
    with equal values (this framework adds keys — ``train.parallel`` … — but never renames or drops one)."""
    from trlx_b200.data.configs import TRLConfig
    from trlx_b200.data.default_configs import default_ilql_config, default_ppo_config, default_sft_config

    def covers(ours, ref, path=""):
        for k, v in ref.items():
            assert k in ours, f"{path}{k} missing"
            if isinstance(v, dict) and isinstance(ours[k], dict):
                covers(ours[k], v, f"{path}{k}.")
            else:
                assert ours[k] == v or (isinstance(v, float) and abs(ours[k] - v) < 1e-12), f"{path}{k}: {ours[k]!r} != {v!r}"

    want = reference_outputs["configs"]
    mine = {p: TRLConfig.load_yaml(p).to_dict() for p in _yamls()}
    mine.update(default_ppo=default_ppo_config().to_dict(), default_ilql=default_ilql_config().to_dict(),
                default_sft=default_sft_config().to_dict(),
                updated=TRLConfig.update(default_ppo_config().to_dict(), {"train.seq_length": 77, "method.gamma": 0.5,
                                                                          "optimizer.kwargs.lr": 1e-3}).to_dict(),
                evolved=default_ilql_config().evolve(train=dict(batch_size=3), method=dict(tau=0.9, gen_kwargs=dict(beta=2))).to_dict())
    assert set(want) == set(mine)
    for name, ref_tree in want.items():
        covers(mine[name], ref_tree, f"[{os.path.basename(name)}] ")


def test_layer_freezing_and_getters_match_the_reference_for_every_family(reference_outputs):
    """``freeze_bottom_causal_layers(model, k)`` leaves the same number of trainable elements as the reference does on the HF model
    of the same configuration, for k in {0, 1, 2, -1} and all eight decoder families; block / hidden-size / final-norm / LM-head
    getters agree."""
    from trlx_b200.models.modeling_base import build_base_model
    from trlx_b200.utils.modeling import (freeze_bottom_causal_layers, hf_get_decoder_blocks, hf_get_decoder_final_norm,
                                          hf_get_hidden_size, hf_get_lm_head, hf_get_num_hidden_layers)

    for name, kw in FAMILIES.items():
        want = reference_outputs["families"][name]
        for k in (0, 1, 2, -1):
            model = build_base_model(kw)
            freeze_bottom_causal_layers(model, k)
            got = sum(p.numel() for p in model.parameters() if p.requires_grad)
            assert got == want[k], (name, k, got, want[k])
        assert len(hf_get_decoder_blocks(model)) == want["blocks"] and hf_get_hidden_size(model.config) == want["hidden"]
        assert hf_get_num_hidden_layers(model.config) == want["layers"]
        assert sum(p.numel() for p in hf_get_decoder_final_norm(model).parameters()) == want["norm_numel"]
        assert tuple(hf_get_lm_head(model).weight.shape) == want["head_shape"]
