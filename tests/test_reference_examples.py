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
