import os
import sys
from unittest import mock

import pytest
import torch

from trlx_b200.models.modeling_ilql import AutoModelForCausalLMWithILQLHeads
from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead, AutoModelForSeq2SeqLMWithHydraValueHead
from trlx_b200.models.peft import get_peft_config
from trlx_b200.trainer.accelerate_sft_trainer import CausalLMWrapper

GPT2 = dict(model_type="gpt2", vocab_size=64, n_embd=32, n_layer=2, n_head=2, n_positions=64)
LLAMA = dict(model_type="llama", vocab_size=64, hidden_size=32, num_hidden_layers=2, num_attention_heads=4,
             num_key_value_heads=2, intermediate_size=48, max_position_embeddings=64)
NEOX = dict(model_type="gpt_neox", vocab_size=64, hidden_size=32, num_hidden_layers=2, num_attention_heads=2,
            intermediate_size=64, max_position_embeddings=64)
T5 = dict(model_type="t5", vocab_size=64, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4)
PEFT = {
    "LORA": dict(peft_type="LORA", r=4, lora_alpha=8, lora_dropout=0.0),
    "PROMPT_TUNING": dict(peft_type="PROMPT_TUNING", num_virtual_tokens=3),
    "PREFIX_TUNING": dict(peft_type="PREFIX_TUNING", num_virtual_tokens=3),
}
WRAPPERS = {"ppo": AutoModelForCausalLMWithHydraValueHead, "ilql": AutoModelForCausalLMWithILQLHeads, "sft": CausalLMWrapper}


def _inputs():
    torch.manual_seed(0)
    ids = torch.randint(2, 60, (2, 6))
    mask = torch.ones(2, 6, dtype=torch.long)
    mask[0, :2] = 0
    return ids, mask


def _perturb_adapter(model):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.requires_grad and ("lora_B" in n or "prompt_embeddings" in n):
                p.add_(torch.randn_like(p) * 0.1)


@pytest.mark.parametrize("kind", list(WRAPPERS))
@pytest.mark.parametrize("peft_type", list(PEFT))
def test_save_load_layout(kind, peft_type, tmp_path):
    model = WRAPPERS[kind].from_config(GPT2, peft_config=PEFT[peft_type]).eval()
    _perturb_adapter(model)
    model.save_pretrained(str(tmp_path))
    files = set(os.listdir(tmp_path))
    assert {"adapter_config.json", "adapter_model.bin", "pytorch_model.bin"} <= files
    heads = torch.load(tmp_path / "pytorch_model.bin")
    assert not any(k.startswith("base_model.") or "transformer." in k for k in heads)  # heads only
    assert os.path.getsize(tmp_path / "pytorch_model.bin") < 1.3e9
    adapter = torch.load(tmp_path / "adapter_model.bin")
    if peft_type == "LORA":
        assert "base_model.model.transformer.h.0.attn.c_attn.lora_A.weight" in adapter
    else:
        assert set(adapter) == {"prompt_embeddings"}
    again = WRAPPERS[kind].from_pretrained(str(tmp_path)).eval()  # adapter is discovered from adapter_config.json
    assert again.peft_type == peft_type
    for k, v in model.base_model.adapter_state_dict().items():
        torch.testing.assert_close(again.base_model.adapter_state_dict()[k], v)
    for k, v in model.state_dict(heads_only=True).items():
        torch.testing.assert_close(again.state_dict(heads_only=True)[k], v)


def test_only_adapter_and_heads_train_and_lora_disable_restores_base():
    torch.manual_seed(0)
    model = AutoModelForCausalLMWithHydraValueHead.from_config(GPT2, peft_config=PEFT["LORA"])
    ids, mask = _inputs()
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert trainable and all(("lora_" in n) or n.startswith("v_head") for n in trainable)
    with torch.no_grad():
        base_logits = model(ids, mask, return_dict=True, ignore_peft_adapter=True).logits
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.5)
    for _ in range(3):
        out = model(ids, mask, return_dict=True)
        (out.logits.float().pow(2).mean() + out.value.pow(2).mean()).backward()
        opt.step()
        opt.zero_grad()
    changed = {n for n, p in model.named_parameters() if not torch.equal(p, before[n])}
    assert changed and changed <= trainable and any("lora_B" in n for n in changed)
    with torch.no_grad():
        tuned = model(ids, mask, return_dict=True).logits
        hydra = model.forward_hydra(ids, mask, return_dict=True).logits
    assert not torch.allclose(tuned, base_logits)
    torch.testing.assert_close(hydra, base_logits)  # adapter off ⇒ original policy = reference policy


@pytest.mark.parametrize("peft_type", ["PROMPT_TUNING", "PREFIX_TUNING"])
def test_prompt_adapters_shapes_and_reference(peft_type):
    model = AutoModelForCausalLMWithHydraValueHead.from_config(GPT2, peft_config=PEFT[peft_type]).eval()
    ids, mask = _inputs()
    out = model(ids, mask, return_dict=True)
    assert out.logits.shape == (2, 6, 64) and out.value.shape == (2, 6)  # virtual tokens are not exposed
    plain = AutoModelForCausalLMWithHydraValueHead.from_config(GPT2).eval()
    plain.base_model.load_state_dict(model.base_model.base_model.state_dict())
    torch.testing.assert_close(model.forward_hydra(ids, mask, return_dict=True).logits, plain(ids, mask, return_dict=True).logits)
    a = model.generate(ids, attention_mask=mask, max_new_tokens=4, do_sample=False, eos_token_id=63, pad_token_id=63)
    b = model.generate(ids, attention_mask=mask, max_new_tokens=4, do_sample=False, eos_token_id=63, pad_token_id=63)
    assert torch.equal(a, b) and a.shape[1] <= 10


@pytest.mark.parametrize("cfg,targets,key", [
    (LLAMA, ["q_proj", "v_proj", "gate_proj"], "base_model.model.model.layers.1.self_attn.v_proj.lora_B.weight"),
    (NEOX, None, "base_model.model.gpt_neox.layers.0.attention.query_key_value.lora_A.weight"),
])
def test_lora_on_fused_projections(cfg, targets, key):
    """LoRA targets that are slices of a fused canonical weight (q/v of QKV, gate of gate|up) or an interleaved QKV."""
    torch.manual_seed(0)
    pc = dict(PEFT["LORA"], target_modules=targets)
    model = AutoModelForCausalLMWithHydraValueHead.from_config(cfg, peft_config=pc).eval()
    _perturb_adapter(model)
    sd = model.base_model.adapter_state_dict()
    assert key in sd
    ids, mask = _inputs()
    with torch.no_grad():
        tuned = model(ids, mask, return_dict=True).logits
        # merging the low-rank update into the base weight must reproduce the adapted forward
        for layer in model.base_model.lora_layers():
            layer.base.weight.copy_(layer.merged_weight())
        model.base_model.disable_adapter_layers()
        merged = model(ids, mask, return_dict=True).logits
    torch.testing.assert_close(merged, tuned, atol=1e-5, rtol=1e-4)
    if cfg is LLAMA:
        assert sd[key].shape == (cfg["hidden_size"] // cfg["num_attention_heads"] * cfg["num_key_value_heads"], 4)


def test_stacked_low_rank_epilogue_form_equals_per_adapter_updates(monkeypatch):
    """The CUDA path writes every adapter of a projection as ONE pair of GEMMs (stacked A, block-structured B with the
    scaling folded in, frozen output as the residual operand).  Same maths on the CPU through the reference ops."""
    from trlx_b200.models.peft import LoRALinear

    torch.manual_seed(0)
    base = torch.nn.Linear(24, 40)
    lin = LoRALinear(base, r=4, alpha=12.0, dropout=0.0)
    lin.add_adapter("q", (0, 16), "q_proj")
    lin.add_adapter("v", (24, 40), "v_proj")
    for k in lin.lora_B:
        torch.nn.init.normal_(lin.lora_B[k], std=0.3)
    x = torch.randn(3, 5, 24, requires_grad=True)
    plain = base(x) + lin.delta(x)
    g_plain = torch.autograd.grad(plain.pow(2).sum(), [x, lin.lora_A["q"], lin.lora_B["v"]])
    monkeypatch.setattr(LoRALinear, "_fused_ok", lambda self, x: True)
    fused = lin(x)
    g_fused = torch.autograd.grad(fused.pow(2).sum(), [x, lin.lora_A["q"], lin.lora_B["v"]])
    torch.testing.assert_close(fused, plain, atol=1e-5, rtol=1e-5)
    for a, b in zip(g_fused, g_plain):
        torch.testing.assert_close(a, b, atol=1e-4, rtol=1e-4)
    both, ref = lin.forward_both(x)
    torch.testing.assert_close(both, plain, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(ref, base(x))


def test_unknown_target_module_and_bad_config():
    with pytest.raises(ValueError):
        AutoModelForCausalLMWithHydraValueHead.from_config(GPT2, peft_config=dict(peft_type="LORA", target_modules=["nope"]))
    with pytest.raises(ValueError):
        get_peft_config(42)
    assert get_peft_config(dict(peft_type="lora", r=2)).peft_type == "LORA"


def test_modules_to_save():
    model = AutoModelForCausalLMWithHydraValueHead.from_config(GPT2, peft_config=dict(PEFT["LORA"], modules_to_save=["ln_f"]))
    assert model.base_model.base_model.transformer.ln_f.weight.requires_grad
    assert any("transformer.ln_f.weight" in k for k in model.base_model.adapter_state_dict())


def test_works_without_external_peft_package():
    with mock.patch.dict(sys.modules, {"peft": None}):
        model = AutoModelForCausalLMWithHydraValueHead.from_config(GPT2, peft_config=PEFT["LORA"])
        ids, mask = _inputs()
        assert model(ids, mask, return_dict=True).logits.shape == (2, 6, 64)


def test_seq2seq_lora(tmp_path):
    model = AutoModelForSeq2SeqLMWithHydraValueHead.from_config(T5, peft_config=PEFT["LORA"]).eval()
    ids, mask = _inputs()
    dec = torch.randint(2, 60, (2, 4))
    _perturb_adapter(model)
    with torch.no_grad():
        tuned = model(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True).logits
        ref = model.forward_hydra(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True).logits
    assert not torch.allclose(tuned, ref)
    model.save_pretrained(str(tmp_path))
    assert any("SelfAttention.q.lora_A" in k for k in torch.load(tmp_path / "adapter_model.bin"))
