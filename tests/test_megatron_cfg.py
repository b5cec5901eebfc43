"""NeMo-schema recipes, optimizer / scheduler aliases and the model-parallel seeding rule."""
import os

import pytest
import torch

from trlx_b200.data.default_configs import default_ppo_config
from trlx_b200.parallel.megatron_cfg import apply_megatron_cfg, arch_from_megatron, parse_megatron_cfg
from trlx_b200.utils import get_optimizer_class, get_scheduler_class, set_seed


@pytest.mark.parametrize("name,tp,pp,layers,hidden", [("megatron_1.3b", 1, 1, 24, 2048), ("megatron_20b.yaml", 4, 1, 44, 6144),
                                                     ("megatron_65b", 8, 4, 80, 8192), ("sft_megatron_20b", 4, 1, 44, 6144)])
def test_recipes_parse(name, tp, pp, layers, hidden):
    rec = parse_megatron_cfg(name)
    assert rec["parallel"]["tensor_parallel"] == tp and rec["parallel"]["pipeline_parallel"] == pp
    arch = rec["arch"]
    assert arch.get("n_layer", arch.get("num_hidden_layers")) == layers and arch.get("n_embd", arch.get("hidden_size")) == hidden
    assert arch["vocab_size"] % 128 == 0  # make_vocab_size_divisible_by


def test_rope_swiglu_recipe_maps_to_llama_layout():
    arch = parse_megatron_cfg("megatron_2b")["arch"]
    assert arch["model_type"] == "llama" and arch["intermediate_size"] == 5440 and arch["tie_word_embeddings"] is False
    neox = arch_from_megatron(dict(hidden_size=64, num_layers=2, num_attention_heads=4, position_embedding_type="rope"))
    assert neox["model_type"] == "gpt_neox" and neox["intermediate_size"] == 256


def test_apply_recipe_to_config_and_dict_recipes():
    cfg = apply_megatron_cfg(default_ppo_config(), "megatron_65b")
    par = cfg.train.parallel
    assert (par.tensor_parallel, par.pipeline_parallel, par.sequence_parallel) == (8, 4, True)
    assert isinstance(cfg.model.model_path, dict) and cfg.model.model_path["n_layer"] == 80
    inline = dict(model=dict(num_layers=2, hidden_size=32, num_attention_heads=4, tensor_model_parallel_size=2))
    assert apply_megatron_cfg(default_ppo_config(), inline).train.parallel.tensor_parallel == 2
    with pytest.raises(FileNotFoundError):
        parse_megatron_cfg("no_such_recipe")


def test_nemo_optimizer_and_scheduler_names():
    from trlx_b200.parallel.optim import FusedAdamW

    assert get_optimizer_class("distributed_fused_adam") is FusedAdamW
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sched = get_scheduler_class("CosineAnnealing")(opt, warmup_steps=2, constant_steps=5, min_lr=0.1, max_steps=15)
    lrs = []
    for _ in range(16):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    assert lrs[0] < lrs[1] <= 1.0                      # warm-up
    assert all(a >= b - 1e-9 for a, b in zip(lrs[2:10], lrs[3:11]))  # cosine decay
    assert abs(lrs[-1] - 0.1) < 1e-6 and abs(lrs[11] - 0.1) < 1e-6   # constant tail at min_lr


def test_seed_is_shared_inside_a_model_parallel_replica(monkeypatch):
    from trlx_b200.data.configs import ParallelConfig

    draws = {}
    for rank in range(8):  # world 8 = pp2 x dp2 x tp2, rank = (pp*2 + dp)*2 + tp
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "8")
        set_seed(1000, ParallelConfig(tensor_parallel=2, pipeline_parallel=2))
        draws[rank] = torch.rand(1).item()
    dp_of = {r: (r // 2) % 2 for r in range(8)}
    for a in range(8):
        for b in range(8):
            assert (draws[a] == draws[b]) == (dp_of[a] == dp_of[b])
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "1")


def test_debug_sized_only_shrinks_huge_recipes_on_a_single_cpu_process(monkeypatch):
    from trlx_b200.parallel.megatron_cfg import debug_sized

    big = dict(model_type="gpt2", n_embd=6144, n_layer=44, n_head=48, n_inner=24576, vocab_size=50304)
    small = dict(model_type="gpt2", n_embd=64, n_layer=2, n_head=4, vocab_size=100)
    monkeypatch.delenv("TRLX_B200_FULL_SIZE", raising=False)
    monkeypatch.setenv("WORLD_SIZE", "1")
    import torch

    if not torch.cuda.is_available():
        with pytest.warns(UserWarning, match="2 x 64"):
            shrunk = debug_sized(big)
        assert (shrunk["n_embd"], shrunk["n_layer"], shrunk["vocab_size"], shrunk["model_type"]) == (64, 2, 50304, "gpt2")
    assert debug_sized(small) == small
    monkeypatch.setenv("TRLX_B200_FULL_SIZE", "1")
    assert debug_sized(big) == big


def test_single_process_runtime_runs_model_parallel_recipes_unsharded(monkeypatch):
    from trlx_b200.data.configs import ParallelConfig
    from trlx_b200.parallel.runtime import Runtime

    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("RANK", "0")
    with pytest.warns(UserWarning, match="without model parallelism"):
        rt = Runtime(ParallelConfig(tensor_parallel=4, pipeline_parallel=2, sequence_parallel=True))
    assert (rt.tp_size, rt.pp_size, rt.dp_size) == (1, 1, 1) and rt.parallel.sequence_parallel is False
