import math
import os

import pytest
import torch

import trlx_b200.utils as utils
from trlx_b200.utils import modeling as mu


@pytest.mark.parametrize("name", [o.value for o in utils.OptimizerName])
def test_optimizer_factories(name):
    cls = utils.get_optimizer_class(name)
    p = torch.nn.Parameter(torch.randn(300, 20))
    kwargs = dict(lr=0.1) if name == "sgd" else dict(lr=0.1, betas=(0.9, 0.95))
    opt = cls([p], **kwargs)
    before = p.detach().clone()
    (p ** 2).sum().backward()
    opt.step()
    assert not torch.equal(before, p.detach())
    assert utils.get_optimizer_class(utils.OptimizerName(name)) is cls


def test_unknown_optimizer_and_scheduler():
    with pytest.raises(ValueError):
        utils.get_optimizer_class("nope")
    with pytest.raises(ValueError):
        utils.get_scheduler_class("nope")


@pytest.mark.parametrize("name", [s.value for s in utils.SchedulerName])
def test_scheduler_factories(name):
    cls = utils.get_scheduler_class(name)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    sched = cls(opt, T_max=10) if name == "cosine_annealing" else cls(opt)
    opt.step()
    sched.step()
    assert len(sched.get_last_lr()) == 1


def test_fused_adamw_cpu_matches_torch():
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(10, 10))
    b = torch.nn.Parameter(a.detach().clone())
    o1 = utils.get_optimizer_class("adamw")([a], lr=1e-2, betas=(0.9, 0.95), weight_decay=0.1)
    o2 = torch.optim.AdamW([b], lr=1e-2, betas=(0.9, 0.95), weight_decay=0.1)
    for _ in range(3):
        for p, o in ((a, o1), (b, o2)):
            o.zero_grad()
            (p ** 3).sum().backward()
            o.step()
    torch.testing.assert_close(a, b)


def test_8bit_adam_tracks_fp32_adam():
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(64, 128))
    b = torch.nn.Parameter(a.detach().clone())
    o1 = utils.get_optimizer_class("adamw_8bit_bnb")([a], lr=1e-2)
    o2 = torch.optim.AdamW([b], lr=1e-2)
    for _ in range(5):
        for p, o in ((a, o1), (b, o2)):
            o.zero_grad()
            ((p - 1) ** 2).sum().backward()
            o.step()
    assert (a - b).abs().max() < 1e-2 and (a - b).abs().mean() < 3e-3


@pytest.mark.parametrize("preset,n_layers,hidden", [("gpt2", 12, 768), ("gpt-j-6b", 28, 4096), ("gpt-neox-20b", 44, 6144),
                                                     ("opt-125m", 12, 768), ("llama-2-7b", 32, 4096)])
def test_getters_on_meta_models(preset, n_layers, hidden):
    from trlx_b200.nn.arch import resolve_config, spec_from_hf_config
    from trlx_b200.nn.transformer import CausalLM

    spec = spec_from_hf_config(resolve_config(preset))
    with torch.device("meta"):
        model = CausalLM(spec)
    assert mu.hf_get_hidden_size(model.config) == hidden
    assert mu.hf_get_num_hidden_layers(model.config) == n_layers
    assert len(mu.hf_get_decoder_blocks(model)) == n_layers
    assert mu.hf_get_decoder(model) is model.transformer
    assert mu.hf_get_decoder_final_norm(model) is model.transformer.ln_f
    assert mu.hf_get_lm_head(model) is model.lm_head


def test_getters_on_huggingface_models():
    transformers = pytest.importorskip("transformers")
    with torch.device("meta"):
        m = transformers.AutoModelForCausalLM.from_config(transformers.GPT2Config(n_layer=2, n_embd=32, n_head=2))
    assert len(mu.hf_get_decoder_blocks(m)) == 2 and mu.hf_get_hidden_size(m.config) == 32
    assert mu.hf_get_decoder_final_norm(m) is m.transformer.ln_f


def test_running_moments_match_closed_form():
    torch.manual_seed(0)
    rm = mu.RunningMoments()
    xs = []
    for _ in range(10):
        x = torch.randn(50) * 3 + 2
        xs.append(x)
        m, s = rm.update(x)
        assert torch.isclose(m, x.mean(), atol=1e-5) and torch.isclose(s, x.std(), atol=1e-4)
    allx = torch.cat(xs)
    assert math.isclose(rm.mean, allx.mean().item(), rel_tol=1e-4)
    assert math.isclose(rm.std, allx.std().item(), rel_tol=1e-4)


def test_whiten_and_global_statistics():
    x = torch.randn(7, 13) * 4 + 1
    w = mu.whiten(x)
    assert abs(w.mean().item()) < 1e-5 and abs(w.std().item() - 1) < 1e-3
    w2 = mu.whiten(x, shift_mean=False)
    assert torch.allclose(w2.mean(), x.mean(), atol=1e-4)
    mean, var, n = mu.get_global_statistics(x)
    assert torch.isclose(mean, x.mean()) and torch.isclose(var, x.var(unbiased=False), rtol=1e-4) and n == x.numel()


def test_logprobs_of_labels():
    logits = torch.randn(3, 5, 11)
    labels = torch.randint(0, 11, (3, 5))
    ref = torch.log_softmax(logits, -1).gather(-1, labels[..., None]).squeeze(-1)
    torch.testing.assert_close(mu.logprobs_of_labels(logits, labels), ref)


def test_misc_helpers():
    assert utils.significant(123456.0) == 123000 and utils.significant(0.00123456, 1) == 0.0012
    assert utils.significant("x") == "x" and utils.significant(0) == 0
    assert mu.flatten_dict({"a": {"b": 1, "c": {"d": 2}}, "e": 3}) == {"a/b": 1, "a/c/d": 2, "e": 3}
    assert utils.filter_non_scalars({"a": 1, "b": "x", "c": torch.tensor(2.0), "d": [1, 2]}) == {"a": 1.0, "c": 2.0}
    tree = {"x": [torch.ones(2), (torch.zeros(1),)], "y": 3}
    out = utils.tree_map(lambda v: v + 1, tree)
    assert out["y"] == 4 and torch.equal(out["x"][0], torch.full((2,), 2.0))
    c = utils.Clock()
    c.tick(10)
    assert c.get_stat(n_samp=10) >= 0
    it = utils.infinite_dataloader([1, 2])
    assert [next(it) for _ in range(5)] == [1, 2, 1, 2, 1]
    stats = mu.get_tensor_stats(torch.tensor([1.0, 2.0, 100.0]), torch.tensor([1.0, 1.0, 0.0]), 2)
    assert stats["max"] == 2 and stats["min"] == 1 and stats["mean"] == 1.5


def test_freezing_rules():
    from trlx_b200.nn.arch import spec_from_hf_config
    from trlx_b200.nn.transformer import CausalLM

    spec = spec_from_hf_config(dict(model_type="gpt2", vocab_size=50, n_embd=16, n_layer=4, n_head=2, n_positions=32))

    def trainable(k):
        m = CausalLM(spec)
        mu.freeze_bottom_causal_layers(m, k)
        return {n for n, p in m.named_parameters() if p.requires_grad}

    assert not any(".h." in n or "wte" in n for n in trainable(0)) and any("ln_f" in n for n in trainable(0))
    two = trainable(2)
    assert all(f".h.{i}." not in n for n in two for i in (0, 1)) and any(".h.3." in n for n in two) and not any("wte" in n for n in two)
    assert len(trainable(-1)) == len(list(CausalLM(spec).named_parameters()))


def test_logging_rank_filter():
    import io
    import logging as pylogging

    from trlx_b200.utils import logging

    stream = io.StringIO()
    handler = pylogging.StreamHandler(stream)
    logging.add_handler(handler)
    try:
        logger = logging.get_logger("trlx_b200.test")
        logging.set_verbosity(logging.INFO)
        logger.info("shown")
        logger.info("hidden", ranks=["3"])
        logger.info("everywhere", ranks=[])
    finally:
        logging.remove_handler(handler)
    out = stream.getvalue()
    assert "[RANK 0] shown" in out and "hidden" not in out and "everywhere" in out
    logging.disable_progress_bar()
    assert not logging.is_progress_bar_enabled() and list(logging.tqdm([1, 2])) == [1, 2]
    logging.enable_progress_bar()


@pytest.mark.parametrize("bucket_elems", [1 << 62, 200])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_flat_optimizer_partition_with_virtual_ranks(world, bucket_elems):
    """N-virtual-rank simulation of the sharded, bucketed optimizer layout (no process group): every rank owns 1/world of every
    gradient bucket, the owned slices tile the flat buffer exactly, stay 16-byte aligned, and `reduce-scatter → per-shard AdamW →
    all-gather` bucket by bucket reproduces the unsharded update."""
    import copy

    from trlx_b200.ops.reference import adamw_step
    from trlx_b200.parallel.optim import _ALIGN, _FlatGroup

    torch.manual_seed(world)
    shapes = [(7, 5), (33,), (16, 16), (1,), (129, 3)]
    base = [torch.nn.Parameter(torch.randn(*s).to(torch.bfloat16)) for s in shapes]
    groups = [_FlatGroup(copy.deepcopy(base), world, r, None, symmetric=False, bucket_elems=bucket_elems) for r in range(world)]
    g0 = groups[0]
    assert g0.numel % (_ALIGN * world) == 0 and all(o % _ALIGN == 0 for o in g0.offsets)
    assert len(g0.buckets) == (1 if bucket_elems > 10 ** 6 else 2)
    assert sorted(i for b in g0.buckets for i in b["params"]) == list(range(len(shapes)))
    covered = torch.zeros(g0.numel, dtype=torch.int32)
    for g in groups:
        assert g.numel == g0.numel and g.offsets == g0.offsets and g.layout() == g0.layout()
        assert g.shard == sum(b["shard"] for b in g.buckets) and g.master.numel() == g.shard
        for b in g.buckets:
            assert b["mine"] % _ALIGN == 0 and b["shard"] % _ALIGN == 0 and b["lo"] <= b["mine"] and b["mine"] + b["shard"] <= b["hi"]
            covered[b["mine"]:b["mine"] + b["shard"]] += 1
    assert bool((covered == 1).all())  # no gaps, no overlap
    for p, o in zip(g0.params, g0.offsets):  # parameters are views of the flat buffer, padding is zero
        assert p.data_ptr() == g0.flat_param.data_ptr() + 2 * o
    # every rank holds a different local gradient; the step uses their mean (data parallelism)
    grads = [torch.randn(g0.numel) for _ in range(world)]
    mean_grad = torch.stack(grads).mean(0)
    hp = dict(lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.01)
    full_w = g0.flat_param.float().clone()
    expect = adamw_step(full_w.clone(), mean_grad.clone(), torch.zeros_like(full_w), torch.zeros_like(full_w), 1, **hp)
    gathered = torch.empty_like(full_w)
    for g in groups:
        for b in g.buckets:
            sl = slice(b["mine"], b["mine"] + b["shard"])
            ms = slice(b["moff"], b["moff"] + b["shard"])
            reduced = torch.stack([gr[sl] for gr in grads]).mean(0)                      # reduce-scatter
            gathered[sl] = adamw_step(g.master[ms].clone(), reduced, g.exp_avg[ms], g.exp_avg_sq[ms], 1, **hp)  # all-gather
    torch.testing.assert_close(gathered, expect)


def test_relative_outputs_never_land_in_the_source_checkout(tmp_path, monkeypatch):
    """`resolve_output_dir`: TRLX_B200_OUT wins; inside the source checkout relative paths are redirected to a temp root (a 498 MB
    example checkpoint written into the repo once voided a whole round of measurements); elsewhere they stay cwd-relative."""
    import os

    import trlx_b200.utils as U

    monkeypatch.delenv("TRLX_B200_OUT", raising=False)
    assert U.resolve_output_dir("/abs/path") == "/abs/path" and U.resolve_output_dir(None) is None
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(U.__file__))))
    monkeypatch.chdir(repo)
    out = U.resolve_output_dir("ckpts")
    assert os.path.isabs(out) and not os.path.realpath(out).startswith(os.path.realpath(repo))
    monkeypatch.chdir(tmp_path)
    assert U.resolve_output_dir("ckpts") == "ckpts"
    monkeypatch.setenv("TRLX_B200_OUT", str(tmp_path / "o"))
    assert U.resolve_output_dir("ckpts") == str(tmp_path / "o" / "ckpts")
    (tmp_path / "here").mkdir()
    assert U.resolve_output_dir("here", for_read=True) == "here"  # an existing cwd-relative checkpoint is read where it is


def test_inplace_wgrad_marks_only_unshared_linear_weights():
    """``gradient_accumulation_fusion`` eligibility: trainable ``nn.Linear`` weights owned by exactly one module; tied embeddings /
    LM heads, frozen weights and non-Linear parameters keep the autograd accumulation.  Without the CUDA extension the sink is
    never used (the wgrad GEMM path is CUDA-only)."""
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx_b200.ops import functional as Fn

    cfg = dict(model_type="gpt2", vocab_size=64, n_embd=32, n_layer=2, n_head=2, n_positions=32, tie_word_embeddings=True)
    model = AutoModelForCausalLMWithHydraValueHead.from_config(cfg, num_layers_unfrozen=1)
    lm = model.base_model
    assert lm.lm_head.weight is lm.transformer.wte.weight
    for p in lm.transformer.h[0].parameters():
        p.requires_grad_(False)
    n = Fn.mark_inplace_wgrad(model)
    marked = {name for name, p in model.named_parameters() if getattr(p, "_b200_inplace_ok", False)}
    assert n == len(marked) and n > 0
    assert not any("wte" in k or "lm_head" in k or "wpe" in k or "norm" in k or k.endswith(".bias") for k in marked), marked
    assert not any(k.startswith("base_model.transformer.h.0.") for k in marked)  # frozen block
    assert any(k.startswith("base_model.transformer.h.1.") and k.endswith("weight") for k in marked)
    w = lm.transformer.h[1].mlp.up.weight
    x = torch.randn(4, 32)
    assert Fn._wgrad_sink(w, x, x) is None  # no sink registered / CPU tensors: autograd path


def test_every_environment_switch_is_documented():
    """docs/env.md lists every ``TRLX_B200_*`` variable the code reads."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    docs = os.path.join(root, "docs", "env.md")
    if not os.path.exists(docs):
        pytest.skip("docs/ not shipped with this snapshot")
    documented = set(re.findall(r"TRLX_B200_[A-Z0-9_]+", open(docs).read()))
    used = set()
    for base in ("trlx_b200", "baseline"):
        for d, _, files in os.walk(os.path.join(root, base)):
            if "_ref" in d or "__pycache__" in d or os.sep + "build" in d:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".cuh")):
                    used |= set(re.findall(r"TRLX_B200_[A-Z0-9_]+", open(os.path.join(d, f), errors="ignore").read()))
    for f in ("bench.py", "__graft_entry__.py"):
        used |= set(re.findall(r"TRLX_B200_[A-Z0-9_]+", open(os.path.join(root, f)).read()))
    assert not (used - documented), sorted(used - documented)
