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


def test_launcher_dry_run_exports_parallel_preset(capsys):
    from trlx_b200 import launch

    rc = launch.main(["--config_file", "configs/accelerate/zero2-bf16.yaml", "--num_processes", "2", "--dry_run",
                      "examples/ppo_sentiments.py", '{"train.total_steps": 2}'])
    out = capsys.readouterr().out
    assert rc == 0 and "--nproc-per-node=2" in out and "--master-addr 127.0.0.1" in out
    assert '"zero_stage": 2' in out and '"grad_clip": 1.0' in out


def test_parallel_preset_env_overrides_config(monkeypatch, tmp_path):
    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_sft_config

    monkeypatch.setenv("TRLX_B200_PARALLEL", json.dumps({"zero_stage": 0, "grad_clip": 0.5}))
    cfg = default_sft_config().evolve(
        train=dict(total_steps=1, batch_size=2, seq_length=16, tracker=None, checkpoint_interval=100, eval_interval=100,
                   checkpoint_dir=str(tmp_path)),
        model=dict(model_path=dict(model_type="gpt2", vocab_size=257, n_embd=16, n_layer=1, n_head=2, n_positions=32, eos_token_id=256,
                                   bos_token_id=256)),
        tokenizer=dict(tokenizer_path="toy://bytes"), method=dict(gen_kwargs=dict(max_new_tokens=2)))
    trainer = trlx.train(samples=["ab", "cd", "ef", "gh"], eval_prompts=["a"], config=cfg)
    assert trainer.config.train.parallel.zero_stage == 0 and trainer.config.train.parallel.grad_clip == 0.5


def test_sweep_default_config_yaml_is_applied_under_sampled_hparams(tmp_path):
    """``--default_config`` (reference ``trlx/sweep.py``): YAML sections are the defaults, sampled keys win."""
    from trlx_b200 import sweep

    script = tmp_path / "target.py"
    script.write_text(
        "import json, os\n"
        "def main(hparams):\n"
        "    from trlx_b200.data.configs import TRLConfig\n"
        "    from trlx_b200.data.default_configs import default_sft_config\n"
        "    cfg = TRLConfig.update(default_sft_config().to_dict(), hparams)\n"
        "    with open(os.path.join(cfg.train.logging_dir, 'seen.json'), 'w') as fh:\n"
        "        json.dump({'bs': cfg.train.batch_size, 'seq': cfg.train.seq_length, 'lr': cfg.optimizer.kwargs['lr']}, fh)\n")
    default = tmp_path / "default.yml"
    default.write_text("train:\n  batch_size: 3\n  seq_length: 77\noptimizer:\n  kwargs:\n    lr: 0.5\n")
    tdir = tmp_path / "trial"
    proc = sweep.launch_trial(str(script), {"train.batch_size": 5}, str(tdir), [], None, str(default))
    assert proc.wait(timeout=120) == 0, (tdir / "stdout.log").read_text()
    seen = json.loads((tdir / "seen.json").read_text())
    assert seen == {"bs": 5, "seq": 77, "lr": 0.5}


@pytest.mark.parametrize("alg", ["random", "bayesopt", "bohb"])
def test_search_algorithms_propose_inside_the_space_and_use_observations(alg):
    import math

    from trlx_b200 import sweep

    space = {"lr": {"strategy": "loguniform", "values": [1e-6, 1e-2]}, "x": {"strategy": "quniform", "values": [0, 1, 0.05]},
             "n": {"strategy": "randint", "values": [1, 9]}, "c": {"strategy": "choice", "values": ["a", "b"]}}
    searcher = sweep.get_search_alg({"search_alg": alg, "num_samples": 24, "mode": "min"}, space, seed=0)
    seen, best = 0, float("inf")
    while True:
        hp = searcher.suggest()
        if hp is None:
            break
        seen += 1
        assert 1e-6 <= hp["lr"] <= 1e-2 and 0 <= hp["x"] <= 1 and 1 <= hp["n"] < 9 and hp["c"] in ("a", "b")
        assert abs(hp["x"] / 0.05 - round(hp["x"] / 0.05)) < 1e-6
        loss = (math.log10(hp["lr"]) + 4) ** 2 + 10 * (hp["x"] - 0.3) ** 2
        best = min(best, loss)
        searcher.observe(hp, loss)
    assert seen == 24 and len(searcher.history) == 24
    assert best < 1.5  # a 2-D bowl with minimum 0: every searcher gets close within 24 trials
    with pytest.raises(NotImplementedError):
        sweep.get_search_alg({"search_alg": "annealing"}, space)


def test_scheduler_rungs_and_tune_config():
    from trlx_b200 import sweep

    assert sweep.get_scheduler({"scheduler": "fifo"}) == dict(name="fifo", rungs=[None], eta=1)
    hb = sweep.get_scheduler({"scheduler": "hyperbandforbohb", "max_t": 90, "reduction_factor": 3})
    assert hb["rungs"] == [10, 30, 90] and hb["eta"] == 3
    with pytest.raises(NotImplementedError):
        sweep.get_scheduler({"scheduler": "pbt"})
    cfg = sweep.get_tune_config({"search_alg": "bohb", "scheduler": "hyperband", "max_t": 9}, {}, 0)
    assert cfg["metric"] == "reward/mean" and cfg["search_alg"].name == "bohb" and cfg["scheduler"]["rungs"][-1] == 9


def test_run_sweep_end_to_end_with_model_based_search_and_successive_halving(tmp_path):
    """A trivial target script logs ``score`` through the jsonl tracker format; the sweep draws trials lazily from the TPE
    searcher, promotes the best third to the second rung and writes results + report."""
    script = tmp_path / "target.py"
    script.write_text(
        "import json, os\n"
        "def main(hparams):\n"
        "    steps = int(hparams.get('train.total_steps', 4))\n"
        "    x = float(hparams['x'])\n"
        "    os.makedirs(hparams['train.logging_dir'], exist_ok=True)\n"
        "    with open(os.path.join(hparams['train.logging_dir'], 'run.jsonl'), 'w') as fh:\n"
        "        for s in range(steps):\n"
        "            fh.write(json.dumps({'step': s, 'score': -(x - 0.3) ** 2 + 0.01 * s}) + '\\n')\n")
    cfg = {"tune_config": {"metric": "score", "mode": "max", "search_alg": "bohb", "scheduler": "hyperband", "max_t": 9,
                           "reduction_factor": 3, "grace_period": 3, "num_samples": 6, "max_concurrent_trials": 2,
                           "n_initial_points": 3},
           "x": {"strategy": "uniform", "values": [0.0, 1.0]}}
    out = tmp_path / "sweep"
    results = sweep.run_sweep(str(script), cfg, str(out), num_gpus=0, gpu_ids=[], poll=0.05)
    assert len(results) == 6 and all(r["returncode"] == 0 for r in results)
    promoted = [r for r in results if r["budget"] == 9]
    assert 1 <= len(promoted) <= 2 and all("rung_1" in r["dir"] for r in promoted)
    best = results[0]
    assert best["best"] == max(r["best"] for r in results) and best in promoted
    saved = json.loads((out / "sweep_results.json").read_text())
    assert saved["metric"] == "score" and len(saved["trials"]) == 6 and (out / "report.md").exists()


def test_launcher_maps_deepspeed_json_onto_the_parallel_preset(tmp_path, capsys):
    from trlx_b200 import launch

    assert launch.parallel_from_deepspeed({"zero_optimization": {"stage": 3, "reduce_bucket_size": 5e8}, "fp16": {"enabled": True},
                                           "gradient_clipping": 0.5}) == dict(zero_stage=3, precision="fp16", grad_clip=0.5,
                                                                              bucket_mb=5e8 * 2 / (1 << 20))
    ds = tmp_path / "ds.json"
    ds.write_text(json.dumps({"zero_optimization": {"stage": 2}, "bf16": {"enabled": True}}))
    rc = launch.main(["--deepspeed_config", str(ds), "--num_processes", "2", "--dry_run", "examples/ppo_sentiments.py"])
    out = capsys.readouterr().out
    assert rc == 0 and '"zero_stage": 2' in out and '"precision": "bf16"' in out
    # the summarisation preset points at its DeepSpeed-style JSON
    rc = launch.main(["--config_file", "examples/summarize_rlhf/configs/default_accelerate_config.yaml", "--dry_run", "x.py"])
    out = capsys.readouterr().out
    assert rc == 0 and "--nproc-per-node=7" in out and '"grad_clip": 1.0' in out
