import glob
import os

import pytest

from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import (default_ilql_config, default_nemo_20b_config, default_ppo_config,
                                            default_sft_config)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _config_files():
    return [f for pat in ("configs/*.yml", "examples/**/configs/*.yml") for f in glob.glob(os.path.join(ROOT, pat), recursive=True)]


def test_repo_configs_load():
    files = _config_files()
    assert files, "no config files found"
    for path in files:
        cfg = TRLConfig.load_yaml(path)
        assert cfg.train.entity_name is None, f"{path} must not pin a tracker entity"
        assert TRLConfig.from_dict(cfg.to_dict()).to_dict() == cfg.to_dict()


@pytest.mark.parametrize("factory", [default_ppo_config, default_ilql_config, default_sft_config, default_nemo_20b_config])
def test_default_configs_roundtrip(factory):
    cfg = factory()
    again = TRLConfig.from_dict(cfg.to_dict())
    assert again.to_dict() == cfg.to_dict()
    assert str(cfg).startswith("{")


def test_default_ppo_matches_reference_recipe():
    cfg = default_ppo_config()
    assert (cfg.train.seq_length, cfg.train.batch_size, cfg.model.num_layers_unfrozen) == (1024, 32, 2)
    m = cfg.method
    assert (m.num_rollouts, m.chunk_size, m.ppo_epochs, m.gen_kwargs["max_new_tokens"]) == (128, 128, 4, 40)
    assert cfg.optimizer.kwargs["lr"] == 3e-5 and cfg.train.trainer == "AcceleratePPOTrainer"


def test_evolve_is_functional_and_nested():
    cfg = default_ilql_config()
    new = cfg.evolve(method=dict(gamma=0.5, gen_kwargs=dict(max_new_tokens=100)), train=dict(seed=7))
    assert new.method.gamma == 0.5 and new.method.gen_kwargs["max_new_tokens"] == 100 and new.train.seed == 7
    assert new.method.gen_kwargs["top_k"] == cfg.method.gen_kwargs["top_k"]
    assert cfg.method.gamma == 0.99  # untouched


def test_update_with_dotted_keys_and_typos():
    base = default_ppo_config()
    cfg = TRLConfig.update(base, {"train.batch_size": 8, "method.gen_kwargs.max_new_tokens": 3, "optimizer": {"kwargs": {"lr": 1.0}}})
    assert cfg.train.batch_size == 8 and cfg.method.gen_kwargs["max_new_tokens"] == 3 and cfg.optimizer.kwargs["lr"] == 1.0
    with pytest.raises(ValueError):
        TRLConfig.update(base, {"train.batchsize": 8})
    with pytest.raises(ValueError):
        TRLConfig.update(base.to_dict(), {"trian": {"batch_size": 8}})


def test_parallel_section_accepts_dict():
    cfg = default_ppo_config().evolve(train=dict(parallel=dict(tensor_parallel=4, sequence_parallel=True)))
    assert cfg.train.parallel.tensor_parallel == 4 and cfg.train.parallel.sequence_parallel is True
    assert cfg.to_dict()["train"]["parallel"]["tensor_parallel"] == 4


def test_method_registry_is_case_insensitive():
    from trlx_b200.data.method_configs import get_method

    assert get_method("ppoconfig") is get_method("PPOConfig")
    with pytest.raises(Exception):
        get_method("nope")
