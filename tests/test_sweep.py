"""Sweep CLI pieces (parameter space, trial enumeration, metric readback, report) and the A/B report helpers."""
import json
import random

import pytest
import yaml

from trlx_b200 import reference, sweep


def test_param_space_strategies_sample_in_range():
    rng = random.Random(0)
    specs = {
        "uniform": ([0.1, 0.9], lambda x: 0.1 <= x <= 0.9),
        "quniform": ([0.0, 1.0, 0.25], lambda x: x in (0.0, 0.25, 0.5, 0.75, 1.0)),
        "loguniform": ([1e-5, 1e-3], lambda x: 1e-5 <= x <= 1e-3),
        "qloguniform": ([1e-2, 1.0, 0.01], lambda x: 0.01 - 1e-9 <= x <= 1.0),
        "randn": ([0.0, 1.0], lambda x: abs(x) < 10),
        "qrandn": ([0.0, 1.0, 0.5], lambda x: abs(x * 2 - round(x * 2)) < 1e-9),
        "randint": ([1, 5], lambda x: x in (1, 2, 3, 4)),
        "qrandint": ([0, 10, 5], lambda x: x in (0, 5, 10)),
        "lograndint": ([1, 100], lambda x: 1 <= x <= 100 and isinstance(x, int)),
        "qlograndint": ([10, 100, 10], lambda x: x % 10 == 0),
        "choice": ([1, 5, 10], lambda x: x in (1, 5, 10)),
    }
    for strategy, (values, ok) in specs.items():
        for _ in range(20):
            assert ok(sweep.sample_value({"strategy": strategy, "values": values}, rng)), strategy
    with pytest.raises(ValueError):
        sweep.get_param_space({"a": {"strategy": "nope", "values": [1]}})


def test_trials_grid_times_samples():
    with open("configs/sweeps/ppo_sweep.yml") as fh:
        cfg = yaml.safe_load(fh)
    space = sweep.get_param_space(cfg)
    assert "tune_config" not in space and space
    space = {"a": {"strategy": "grid_search", "values": [1, 2, 3]}, "b": {"strategy": "uniform", "values": [0, 1]}}
    trials = list(sweep.iter_trials(space, {"num_samples": 2}))
    assert len(trials) == 6 and sorted({t["a"] for t in trials}) == [1, 2, 3]
    assert len({t["b"] for t in trials}) == 6


def test_metric_readback_and_report(tmp_path):
    d = tmp_path / "trial"
    d.mkdir()
    with open(d / "run.jsonl", "w") as fh:
        for step, v in enumerate([0.1, 0.7, 0.4]):
            fh.write(json.dumps({"step": step, "reward/mean": v, "junk": "x"}) + "\n")
    assert sweep.read_metric(str(d), "reward/mean", "max") == (0.7, 0.4, 3)
    assert sweep.read_metric(str(d), "reward/mean", "min")[0] == 0.1
    assert sweep.read_metric(str(d), "absent", "max") == (None, None, 0)
    results = [dict(id=0, hparams={"lr": 1e-3}, best=0.7, last=0.4, points=3, returncode=0),
               dict(id=1, hparams={"lr": 1e-4}, best=None, last=None, points=0, returncode=1)]
    sweep.write_report(results, {"lr": {}}, "reward/mean", "max", "ex.py", str(tmp_path / "report.md"))
    text = (tmp_path / "report.md").read_text()
    assert "Best configuration" in text and "| 0 | 0.001 |" in text

    runs = reference.load_runs(str(tmp_path))
    assert runs["trial"]["reward/mean"][-1] == (2, 0.4)
    md = reference.compare(runs, {"trial": {"reward/mean": [(0, 0.2)]}}, "pr", "base")
    assert "| reward/mean | 0.4 | 0.2 | +0.2 |" in md
