import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from trlx_b200.data.ppo_types import PPORLElement
from trlx_b200.pipeline.offline_pipeline import DialogStore, PromptPipeline, pad_rows, tokenize_dialogue
from trlx_b200.pipeline.ppo_pipeline import DeviceBatchLoader, PPORolloutStorage, RolloutBlock, ppo_collate_fn
from trlx_b200.utils.tokenizer import build_toy_tokenizer, load_tokenizer


@pytest.fixture(scope="module")
def tok():
    t = build_toy_tokenizer("toy://bytes")
    t.pad_token = "<|padding|>"
    return t


def test_toy_tokenizers_roundtrip():
    t = build_toy_tokenizer("toy://bpe?vocab=600")
    assert len(t) == 600 and t.eos_token_id == 599
    text = "the quick brown fox, über 9000!"
    assert t.decode(t(text)["input_ids"]) == text
    c = build_toy_tokenizer("toy://chars?alphabet=abcd")
    assert c("abca")["input_ids"] == [0, 1, 2, 0] and c.decode([3, 0]) == "da"
    fallback = load_tokenizer("gpt2")  # no network → synthetic BPE of GPT-2's size
    assert len(fallback) == 50257 and fallback.eos_token_id == 50256


def test_tokenize_dialogue_single_and_multi_turn(tok):
    out = tokenize_dialogue("abc", tok)
    assert [m.is_output for m in out] == [False, True]
    assert out[0].tokens == (tok.bos_token_id,) and out[1].tokens == (97, 98, 99, tok.eos_token_id)
    multi = tokenize_dialogue(["ab", "cd", "ef", "gh"], tok)
    assert [m.is_output for m in multi] == [False, True, False, True] and multi[-1].tokens[-1] == tok.eos_token_id
    with pytest.raises(ValueError):
        tokenize_dialogue(["a", "b", "c"], tok)


@settings(deadline=None, max_examples=60)
@given(st.lists(st.text(alphabet="abcdef ", min_size=1, max_size=12), min_size=2, max_size=6).filter(lambda x: len(x) % 2 == 0),
       st.integers(min_value=2, max_value=40), st.sampled_from(["left", "right"]))
def test_tokenize_dialogue_truncation_invariants(phrases, max_length, side):
    tok = build_toy_tokenizer("toy://bytes")
    tok.truncation_side = side
    full = tokenize_dialogue(list(phrases), tok, max_length=10 ** 6)
    out = tokenize_dialogue(list(phrases), tok, max_length=max_length)
    flat_full = [t for m in full for t in m.tokens]
    flat = [t for m in out for t in m.tokens]
    assert 0 < len(flat) <= max_length
    assert all(len(m.tokens) > 0 for m in out)
    assert not out[0].is_output  # always starts with a prompt (BOS inserted when needed)
    body = flat[1:] if (out[0].tokens == (tok.bos_token_id,) and flat_full[: len(flat)] != flat and flat_full[-len(flat):] != flat) else flat
    if side == "right":
        assert flat_full[: len(body)] == body or flat_full[: len(flat)] == flat
    else:
        assert flat_full[-len(body):] == body or flat_full[-len(flat):] == flat


def test_dialog_store_masks_prompt_tokens(tok):
    tok.padding_side = "right"
    dialogs = [tokenize_dialogue(["ab", "cde"], tok), tokenize_dialogue(["x", "y"], tok)]
    store = DialogStore(dialogs, tok)
    batch = next(iter(store.create_loader(2)))
    assert batch["input_ids"].shape == batch["labels"].shape == batch["attention_mask"].shape
    assert batch["labels"][0, :2].tolist() == [-100, -100] and batch["labels"][0, 2:6].tolist() == [99, 100, 101, tok.eos_token_id]
    assert batch["attention_mask"][1].sum() == 3


def test_prompt_pipeline_metadata_and_padding(tok):
    tok.padding_side = "left"
    prompts = [{"prompt": "hello world", "gold": 1}, {"prompt": "hi", "gold": 2}]
    pipe = PromptPipeline(prompts, max_prompt_length=5, tokenizer=tok)
    assert len(pipe) == 2 and len(pipe[0]["input_ids"]) == 5 and pipe[1]["gold"] == 2
    batch = next(iter(pipe.create_loader(2)))
    assert batch["input_ids"].shape == (2, 5) and batch["gold"] == [1, 2]
    assert batch["attention_mask"][1].tolist() == [0, 0, 0, 1, 1] and batch["input_ids"][1, 0] == tok.pad_token_id
    assert "prompt" in prompts[0]  # input dicts are not mutated (the reference pops the key)


def _elements(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        q, r = int(torch.randint(1, 6, (1,), generator=g)), int(torch.randint(1, 7, (1,), generator=g))
        out.append(PPORLElement(torch.randint(1, 50, (q,), generator=g), torch.randint(1, 50, (r,), generator=g),
                                torch.randn(r, generator=g), torch.randn(r, generator=g), torch.randn(r, generator=g)))
    return out


def test_ppo_collate_and_store(tmp_path):
    elems = _elements(5)
    batch = ppo_collate_fn("left", 0, elems)
    Q, R = max(len(e.query_tensor) for e in elems), max(len(e.response_tensor) for e in elems)
    assert batch.query_tensors.shape == (5, Q) and batch.rewards.shape == (5, R)
    for i, e in enumerate(elems):
        assert batch.query_tensors[i, Q - len(e.query_tensor):].tolist() == e.query_tensor.tolist()
        assert batch.response_tensors[i, : len(e.response_tensor)].tolist() == e.response_tensor.tolist()
        assert (batch.query_tensors[i, : Q - len(e.query_tensor)] == 0).all()
    store = PPORolloutStorage(0, "left")
    assert store.history == [None]
    store.clear_history()
    store.push(elems)
    assert len(store) == 5 and store[2] is elems[2]
    store.export_history(str(tmp_path))
    assert len(list(tmp_path.iterdir())) == 1
    seen = sum(len(b.query_tensors) for b in store.create_loader(2, shuffle=True))
    assert seen == 5


def test_device_block_loader_matches_collate():
    elems = _elements(7, seed=3)
    Q, R = 6, 7
    block = RolloutBlock(
        queries=pad_rows([e.query_tensor for e in elems], 0, "left", min_len=Q),
        responses=pad_rows([e.response_tensor for e in elems], 0, "right", min_len=R),
        logprobs=pad_rows([e.logprobs for e in elems], 0.0, "right", min_len=R),
        values=pad_rows([e.values for e in elems], 0.0, "right", min_len=R),
        rewards=pad_rows([e.rewards for e in elems], 0.0, "right", min_len=R),
        query_lens=torch.tensor([len(e.query_tensor) for e in elems]),
        response_lens=torch.tensor([len(e.response_tensor) for e in elems]),
        trunk_hidden=torch.randn(7, Q + R, 4),
    )
    store = PPORolloutStorage(0, "left")
    store.clear_history()
    store.push_block(block)
    assert len(store) == 7
    hist = store.history
    for h, e in zip(hist, elems):
        assert h.query_tensor.tolist() == e.query_tensor.tolist() and torch.equal(h.rewards, e.rewards)
    loader = store.create_loader(3, shuffle=False)
    assert isinstance(loader, DeviceBatchLoader) and len(loader) == 3
    for i, b in enumerate(loader):
        ref = ppo_collate_fn("left", 0, elems[3 * i: 3 * i + 3])
        assert torch.equal(b.query_tensors, ref.query_tensors)
        w = ref.response_tensors.shape[1]
        assert torch.equal(b.response_tensors[:, :w], ref.response_tensors) and b.response_tensors.shape[1] <= w + 1
        assert torch.equal(b.rewards, ref.rewards) and torch.equal(b.values, ref.values)
        assert b.trunk_hidden.shape[:2] == (len(ref.query_tensors), b.query_tensors.shape[1] + b.response_tensors.shape[1])


def test_reward_model_repository_packaging(tmp_path):
    """``examples/hh/to_triton.py``: the serving entry (weights + serving.json) the reward server is started from."""
    import json
    import os
    import sys

    sys.path.insert(0, os.getcwd())
    from examples.hh.to_triton import build_repository

    ckpt = tmp_path / "rm"
    ckpt.mkdir()
    (ckpt / "pytorch_model.bin").write_bytes(b"weights")
    (ckpt / "config.json").write_text("{}")
    (ckpt / "notes.tmp").write_text("ignored")
    version_dir = build_repository(str(ckpt), str(tmp_path / "store"), "gptj-rm-static", max_batch_size=8, port=8123)
    assert sorted(os.listdir(version_dir)) == ["config.json", "pytorch_model.bin"]
    spec = json.loads((tmp_path / "store" / "gptj-rm-static" / "serving.json").read_text())
    assert spec["max_batch_size"] == 8 and spec["port"] == 8123 and spec["files"] == ["config.json", "pytorch_model.bin"]
