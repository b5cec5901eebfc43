"""Example tasks against the reference checkout (``/root/reference`` or ``$TRLX_REFERENCE``; skipped elsewhere)."""
import importlib.util
import os

import pytest
import torch

REF = os.environ.get("TRLX_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "examples", "randomwalks", "randomwalks.py")), reason="no reference checkout")
def test_randomwalks_task_is_the_reference_task():
    """The toy graph task (BFS implementation here, networkx there): same evaluation prompts, same sampled walks, same logit mask,
    same ``lengths`` / ``optimality`` metrics — for the default parameters, another graph, and the ``|``-delimited variant."""
    pytest.importorskip("networkx")
    ref = _load(os.path.join(REF, "examples", "randomwalks", "randomwalks.py"), "_ref_randomwalks")
    ours = _load(os.path.join(ROOT, "examples", "randomwalks", "randomwalks.py"), "_our_randomwalks")
    for kw in (dict(), dict(seed=7, n_nodes=15, max_length=8, n_walks=200, p_edge=0.2), dict(gpt2_tokenizer=True, n_walks=100)):
        m_ref, eval_ref, walks_ref, mask_ref = ref.generate_random_walks(**kw)
        m_our, eval_our, walks_our, mask_our = ours.generate_random_walks(**kw)
        assert eval_ref == eval_our and walks_ref == walks_our
        assert torch.equal(torch.as_tensor(mask_ref), torch.as_tensor(mask_our))
        d = "|" if kw.get("gpt2_tokenizer") else ""
        samples = walks_ref[:60] + [d.join("bcdefghij"), d.join("ba"), d.join("bcb"), d.join("cdcdcdcdcd"), d.join("ecba")]
        a, b = m_ref(samples), m_our(samples)
        assert set(a) == set(b)
        for k in a:
            assert [float(x) for x in a[k]] == pytest.approx([float(x) for x in b[k]], abs=1e-9), k


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "examples", "summarize_rlhf", "reward_model", "reward_model.py")),
                    reason="no reference checkout")
def test_pairwise_reward_model_objective_is_the_reference_objective():
    """The summarisation reward model: for the same per-token rewards, the vectorised pairwise loss (divergence index, padding,
    end-of-sequence scores) equals the reference's per-pair Python loop — training batches and the inference form."""
    ref_mod = _load(os.path.join(REF, "examples", "summarize_rlhf", "reward_model", "reward_model.py"), "_ref_reward_model")
    from examples.summarize_rlhf.reward_model.reward_model import GPTRewardModel

    PAD, T, H = 0, 12, 8
    g = torch.Generator().manual_seed(0)
    prompt = torch.randint(1, 50, (5, 4), generator=g)
    lens_c, lens_r = [12, 9, 7, 12, 6], [10, 12, 7, 5, 11]
    chosen, rejected = torch.full((5, T), PAD), torch.full((5, T), PAD)
    for i in range(5):
        chosen[i, :4], rejected[i, :4] = prompt[i], prompt[i]
        chosen[i, 4:lens_c[i]] = torch.randint(1, 50, (lens_c[i] - 4,), generator=g)
        rejected[i, 4:lens_r[i]] = torch.randint(50, 99, (lens_r[i] - 4,), generator=g)
    hidden = torch.randn(10, T, H, generator=g)

    class Trunk(torch.nn.Module):
        def forward(self, input_ids, **kw):
            return (hidden[: input_ids.shape[0]],)

    ref = object.__new__(ref_mod.GPTRewardModel)
    torch.nn.Module.__init__(ref)
    ref.transformer, ref.v_head, ref.PAD_ID = Trunk(), torch.nn.Linear(H, 1, bias=False), PAD
    ours = object.__new__(GPTRewardModel)
    torch.nn.Module.__init__(ours)
    ours.v_head, ours.PAD_ID = ref.v_head, PAD
    ours.rewards = lambda input_ids, attention_mask=None: ref.v_head(hidden[: input_ids.shape[0]]).squeeze(-1)

    pair = torch.cat([chosen, rejected])
    a, b = ref(input_ids=pair), ours(pair)
    assert set(a) == set(b) == {"loss", "chosen_end_scores", "rejected_end_scores"}
    for k in a:
        torch.testing.assert_close(b[k], a[k], atol=1e-6, rtol=1e-5)
    same = torch.cat([chosen, chosen])  # inference form: identical halves -> scores at the last non-pad token
    a, b = ref(input_ids=same), ours(same)
    assert set(a) == set(b) == {"chosen_end_scores"}
    torch.testing.assert_close(b["chosen_end_scores"], a["chosen_end_scores"])


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "examples", "experiments", "grounded_program_synthesis", "lang.py")),
                    reason="no reference checkout")
def test_list_dsl_interpreter_agrees_with_the_reference_interpreter():
    """The grounded-program-synthesis DSL: the parser-based interpreter here returns what the reference's ``eval``-based one
    returns, on programs drawn by the reference's own sampler and on malformed / ill-typed ones (``"ERROR"``)."""
    import random

    ref = _load(os.path.join(REF, "examples", "experiments", "grounded_program_synthesis", "lang.py"), "_ref_lang")
    ours = _load(os.path.join(ROOT, "examples", "experiments", "grounded_program_synthesis", "lang.py"), "_our_lang")
    random.seed(0)
    data = ref.create_synthetic_dataset(600)
    ri, oi = ref.Interpreter(), ours.Interpreter()
    programs = [d["output"] for d in data] + [
        "take([1,2,3],2)", "take([1,2,3],5)", "drop([1,2,3],1)", "foo([1])", "add_n([1],[2])", "minimum([])", "maximum([3,1])",
        "div_n([4,5],0)", "div_n([4,-5],2)", "expand_copy([1,2])", "reverse(3)", "take([1,2,3],", "mul_n(reverse([1,2]),-3)",
        "sub_n([1,2],1) extra", "1", "[1,2]", "sort_asc(minimum([1,2]))"]
    assert len(programs) > 300
    for p in programs:
        assert ri(p) == oi(p), (p, ri(p), oi(p))
    assert all(oi(d["output"]) == d["io_out"] for d in data)
