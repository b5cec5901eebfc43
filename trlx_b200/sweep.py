"""Hyper-parameter sweeps:  ``python -m trlx_b200.sweep --config configs/sweeps/ppo_sweep.yml examples/ppo_sentiments.py``

Reference counterpart: ``trlx/sweep.py`` (Ray Tune + W&B reports): parameter-space strategies ``:16-98``, search
algorithms / schedulers ``:103-176``, the report ``:178-265``, the CLI ``:268-348``.  Same YAML schema and CLI, but no Ray
and no W&B dependency: trials are plain subprocesses (``python -m torch.distributed.run`` when a trial uses several GPUs)
scheduled over the node's GPUs, each logging through the ``jsonl`` tracker; the sweep reads the target metric back from
those logs, supports ``random`` / grid search, a Gaussian-process ``bayesopt`` searcher and a TPE (``bohb``) searcher
(scikit-learn / pure Python, suggestions are drawn as slots free up) with ``fifo`` or successive-halving (``hyperband``,
``hyperbandforbohb``) scheduling, and writes ``sweep_results.json`` + a Markdown report (best trials, parameter table) instead of a W&B report.

An example script only has to expose ``main(hparams: dict)`` (every script under ``examples/`` does).
"""
from __future__ import annotations

import argparse
import itertools
import json
import math
import os
import random
import subprocess
import sys
import time
from datetime import datetime
from typing import Any, Dict, Iterator, List, Optional, Tuple

import yaml

STRATEGIES = ("uniform", "quniform", "loguniform", "qloguniform", "randn", "qrandn", "randint", "qrandint", "lograndint",
              "qlograndint", "choice", "grid_search", "grid")


def _quantize(x: float, q: float) -> float:
    return round(x / q) * q


def sample_value(spec: Dict[str, Any], rng: random.Random):
    """One draw from a ``{strategy, values}`` entry (grid strategies are expanded by :func:`iter_trials`)."""
    strategy, v = spec["strategy"], spec["values"]
    if strategy not in STRATEGIES:
        raise ValueError(f"unknown search strategy `{strategy}`; expected one of {STRATEGIES}")
    if strategy == "uniform":
        lo, hi = v
        return rng.uniform(lo, hi)
    if strategy == "quniform":
        lo, hi, q = v
        return _quantize(rng.uniform(lo, hi), q)
    if strategy in ("loguniform", "qloguniform"):
        lo, hi = v[0], v[1]
        base = v[2] if strategy == "loguniform" and len(v) > 2 else 10
        x = base ** rng.uniform(math.log(lo, base), math.log(hi, base))
        return _quantize(x, v[2]) if strategy == "qloguniform" else x
    if strategy == "randn":
        mean, sd = v
        return rng.gauss(mean, sd)
    if strategy == "qrandn":
        mean, sd, q = v
        return _quantize(rng.gauss(mean, sd), q)
    if strategy == "randint":
        lo, hi = v
        return rng.randrange(int(lo), int(hi))
    if strategy == "qrandint":
        lo, hi, q = v
        return int(_quantize(rng.randrange(int(lo), int(hi) + 1), q))
    if strategy in ("lograndint", "qlograndint"):
        lo, hi = v[0], v[1]
        x = int(math.exp(rng.uniform(math.log(lo), math.log(hi))))
        return int(_quantize(x, v[2])) if strategy == "qlograndint" else x
    if strategy == "choice":
        return rng.choice(list(v))
    return rng.choice(list(v))  # grid entries sampled at random when a random search touches them


def get_param_space(config: Dict[str, Any]) -> Dict[str, Dict[str, Any]]:
    """Validate and return ``{dotted.key: {strategy, values}}`` (everything but ``tune_config``)."""
    space = {}
    for key, spec in config.items():
        if key == "tune_config":
            continue
        if not isinstance(spec, dict) or "strategy" not in spec or "values" not in spec:
            raise ValueError(f"sweep entry `{key}` must be a mapping with `strategy` and `values`")
        if spec["strategy"] not in STRATEGIES:
            raise ValueError(f"sweep entry `{key}`: unknown strategy `{spec['strategy']}`")
        if not isinstance(spec["values"], list):
            raise ValueError(f"sweep entry `{key}`: `values` must be a list")
        space[key] = spec
    return space


def iter_trials(space: Dict[str, Dict[str, Any]], tune_config: Dict[str, Any], seed: int = 0) -> Iterator[Dict[str, Any]]:
    """Grid entries are enumerated exhaustively; the remaining entries are re-sampled ``num_samples`` times per grid point
    (Ray Tune semantics)."""
    rng = random.Random(seed)
    grid_keys = [k for k, s in space.items() if s["strategy"] in ("grid_search", "grid")]
    other = [k for k in space if k not in grid_keys]
    num_samples = int(tune_config.get("num_samples", 1))
    search = str(tune_config.get("search_alg", "random")).lower()
    if search not in ("random", "grid"):
        print(f"[sweep] search_alg `{search}` needs a Bayesian-optimisation package that is not bundled; using random search")
    grids = itertools.product(*[space[k]["values"] for k in grid_keys]) if grid_keys else [()]
    for point in grids:
        for _ in range(num_samples):
            hp = dict(zip(grid_keys, point))
            for k in other:
                hp[k] = sample_value(space[k], rng)
            yield hp


# ---- search algorithms -------------------------------------------------------------------------------------------------------
_NUMERIC = ("uniform", "quniform", "loguniform", "qloguniform", "randint", "qrandint", "lograndint", "qlograndint")


def _to_unit(spec: Dict[str, Any], value) -> Optional[float]:
    """Position of ``value`` inside the range of a numeric entry, on the scale it is sampled on (``None``: not numeric)."""
    st, v = spec["strategy"], spec["values"]
    if st not in _NUMERIC:
        return None
    lo, hi = float(v[0]), float(v[1])
    if "log" in st:
        lo, hi, value = math.log(lo), math.log(hi), math.log(max(float(value), 1e-300))
    return min(max((float(value) - lo) / max(hi - lo, 1e-12), 0.0), 1.0)


def _from_unit(spec: Dict[str, Any], u: float):
    st, v = spec["strategy"], spec["values"]
    lo, hi = float(v[0]), float(v[1])
    if "log" in st:
        x = math.exp(math.log(lo) + u * (math.log(hi) - math.log(lo)))
    else:
        x = lo + u * (hi - lo)
    x = min(max(x, lo), hi)  # exp(log(hi)) may overshoot by an ulp
    if st in ("quniform", "qloguniform", "qrandint", "qlograndint"):
        x = _quantize(x, v[2])
    if "int" in st:
        x = int(min(max(round(x), int(lo)), int(hi) - (1 if st == "randint" else 0)))
    return x


class RandomSearcher:
    """Grid entries enumerated exhaustively, everything else re-sampled ``num_samples`` times per grid point."""

    name = "random"

    def __init__(self, space, tune_config, seed: int = 0):
        self.space, self.mode = space, tune_config.get("mode", "max")
        self._it = iter_trials(space, tune_config, seed)
        self.history: List[Tuple[Dict[str, Any], float]] = []

    def suggest(self) -> Optional[Dict[str, Any]]:
        return next(self._it, None)

    def observe(self, hparams: Dict[str, Any], score: Optional[float]) -> None:
        if score is not None and math.isfinite(score):
            self.history.append((hparams, score if self.mode == "max" else -score))


class _ModelBasedSearcher(RandomSearcher):
    """Common part of the sequential searchers: ``n_initial`` random trials, then ``_propose`` from the observations.
    Numeric entries are modelled on the unit interval of their sampling scale; ``choice`` / grid entries are drawn at
    random and only enter the model through the numeric coordinates they co-occur with."""

    def __init__(self, space, tune_config, seed: int = 0):
        super().__init__(space, tune_config, seed)
        self.rng = random.Random(seed + 1)
        self.numeric = [k for k, sp in space.items() if sp["strategy"] in _NUMERIC]
        grid = 1
        for sp in space.values():
            if sp["strategy"] in ("grid_search", "grid"):
                grid *= max(len(sp["values"]), 1)
        self.total = int(tune_config.get("num_samples", 1)) * grid
        self.n_initial = int(tune_config.get("n_initial_points", max(4, 2 * len(self.numeric))))
        self.issued = 0

    def _random(self) -> Dict[str, Any]:
        return {k: sample_value(sp, self.rng) for k, sp in self.space.items()}

    def _unit(self, hp) -> List[float]:
        return [_to_unit(self.space[k], hp[k]) for k in self.numeric]

    def suggest(self) -> Optional[Dict[str, Any]]:
        if self.issued >= self.total:
            return None
        self.issued += 1
        if not self.numeric or len(self.history) < self.n_initial:
            return self._random()
        hp = self._random()
        for k, u in zip(self.numeric, self._propose()):
            hp[k] = _from_unit(self.space[k], u)
        return hp

    def _propose(self) -> List[float]:  # pragma: no cover - abstract
        raise NotImplementedError


class BayesOptSearcher(_ModelBasedSearcher):
    """Gaussian-process surrogate (Matérn 5/2, scikit-learn) + expected improvement maximised over random candidates —
    the role of Ray Tune's ``BayesOptSearch`` in the reference (``trlx/sweep.py:103-133``)."""

    name = "bayesopt"

    def _propose(self) -> List[float]:
        try:
            import numpy as np
            from scipy.stats import norm
            from sklearn.gaussian_process import GaussianProcessRegressor
            from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
        except ImportError:  # pragma: no cover
            return [self.rng.random() for _ in self.numeric]
        X = np.array([self._unit(hp) for hp, _ in self.history])
        y = np.array([sc for _, sc in self.history], dtype=float)
        y = (y - y.mean()) / (y.std() + 1e-9)
        gp = GaussianProcessRegressor(ConstantKernel(1.0) * Matern(length_scale=0.3, nu=2.5) + WhiteKernel(1e-3),
                                      normalize_y=False, random_state=self.rng.randrange(2 ** 31))
        import warnings

        with warnings.catch_warnings():  # hyper-parameter fit on a handful of points: convergence chatter is expected
            warnings.simplefilter("ignore")
            gp.fit(X, y)
        cand = np.array([[self.rng.random() for _ in self.numeric] for _ in range(512)])
        mu, sd = gp.predict(cand, return_std=True)
        best = y.max()
        z = (mu - best - 0.01) / np.maximum(sd, 1e-9)
        ei = (mu - best - 0.01) * norm.cdf(z) + sd * norm.pdf(z)
        return [float(v) for v in cand[int(np.argmax(ei))]]


class TPESearcher(_ModelBasedSearcher):
    """Tree-structured Parzen estimator (the model inside BOHB, ``search_alg: bohb``): kernel density of the best quarter of
    the observations against the rest, candidates drawn from the former and ranked by the density ratio."""

    name = "bohb"

    def _propose(self) -> List[float]:
        ranked = sorted(self.history, key=lambda t: t[1], reverse=True)
        n_good = max(len(ranked) // 4, 2)
        good = [self._unit(hp) for hp, _ in ranked[:n_good]]
        bad = [self._unit(hp) for hp, _ in ranked[n_good:]] or good

        def density(points, x, bw):
            return sum(math.exp(-0.5 * ((x - p) / bw) ** 2) for p in points) / (len(points) * bw) + 1e-12

        out = []
        for d in range(len(self.numeric)):
            g, b = [p[d] for p in good], [p[d] for p in bad]
            bw = max(1.06 * (max(g) - min(g) + 0.1) * len(g) ** -0.2, 0.05)
            cands = [min(max(self.rng.gauss(self.rng.choice(g), bw), 0.0), 1.0) for _ in range(24)]
            out.append(max(cands, key=lambda x: density(g, x, bw) / density(b, x, bw)))
        return out


def get_search_alg(tune_config: Dict[str, Any], space: Optional[Dict[str, Dict[str, Any]]] = None, seed: int = 0):
    """``search_alg`` of the sweep YAML → searcher with ``suggest()`` / ``observe(hparams, score)`` (reference
    ``trlx/sweep.py:103-133``: ``bayesopt``, ``bohb``, ``random``; unknown names raise like the reference)."""
    name = str(tune_config.get("search_alg", "random") or "random").lower()
    space = space or {}
    if name in ("random", "grid"):
        return RandomSearcher(space, tune_config, seed)
    if name == "bayesopt":
        return BayesOptSearcher(space, tune_config, seed)
    if name in ("bohb", "tpe"):
        return TPESearcher(space, tune_config, seed)
    raise NotImplementedError(f"search_alg `{name}` is not supported (random, bayesopt, bohb)")


def get_scheduler(tune_config: Dict[str, Any]) -> Dict[str, Any]:
    """``scheduler`` of the sweep YAML → ``dict(name, rungs, eta)`` (reference ``:136-158``: ``hyperband``,
    ``hyperbandforbohb``, ``fifo``).  The hyperband family runs successive halving: rungs of growing step budget
    (``grace_period · eta^i`` up to ``max_t``), the best ``1/eta`` of a rung is promoted."""
    name = str(tune_config.get("scheduler", "fifo") or "fifo").lower()
    if name in ("hyperband", "hyperbandforbohb", "asha", "bohb", "median"):
        max_t = int(tune_config.get("max_t", tune_config.get("max_steps", 0)) or 0)
        eta = int(tune_config.get("reduction_factor", 3))
        grace = int(tune_config.get("grace_period", max(max_t // (eta ** 2), 1))) if max_t else None
        rungs: List[Optional[int]] = []
        t = grace
        while max_t and t < max_t:
            rungs.append(t)
            t *= eta
        rungs.append(max_t if max_t else None)
        return dict(name=name, rungs=rungs, eta=eta)
    if name != "fifo":
        raise NotImplementedError(f"scheduler `{name}` is not supported (fifo, hyperband, hyperbandforbohb)")
    return dict(name="fifo", rungs=[None], eta=1)


def get_tune_config(tune_config: Dict[str, Any], space: Optional[Dict[str, Dict[str, Any]]] = None, seed: int = 0) -> Dict[str, Any]:
    """Normalised ``tune_config`` with the searcher and scheduler objects filled in (reference ``:161-176``)."""
    cfg = dict(tune_config)
    cfg.setdefault("metric", "reward/mean")
    cfg.setdefault("mode", "max")
    cfg.setdefault("num_samples", 1)
    cfg["search_alg"] = get_search_alg(tune_config, space, seed)
    cfg["scheduler"] = get_scheduler(tune_config)
    return cfg


# ---- running trials ------------------------------------------------------------------------------------------------------------
_TRIAL_SNIPPET = """
import importlib.util, json, sys
spec = importlib.util.spec_from_file_location("sweep_target", sys.argv[1])
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
mod.main(json.loads(sys.argv[2]))
"""


def launch_trial(script: str, hparams: Dict[str, Any], trial_dir: str, gpus: List[int], budget_steps: Optional[int],
                 default_config: Optional[str]) -> subprocess.Popen:
    os.makedirs(trial_dir, exist_ok=True)
    hp: Dict[str, Any] = {}
    if default_config:  # YAML TRLConfig sections first, so that the sampled dotted keys override them
        import yaml

        with open(default_config) as fh:
            hp.update({k: v for k, v in (yaml.safe_load(fh) or {}).items() if isinstance(v, dict)})
    hp.update(hparams)
    hp.setdefault("train.tracker", "jsonl")
    hp.setdefault("train.logging_dir", trial_dir)
    hp.setdefault("train.checkpoint_dir", os.path.join(trial_dir, "ckpts"))
    if budget_steps is not None:
        hp["train.total_steps"] = int(budget_steps)
    env = dict(os.environ)
    pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # trials import the same trlx_b200 as the sweep
    env["PYTHONPATH"] = os.pathsep.join(p for p in (pkg_root, env.get("PYTHONPATH", "")) if p)
    if gpus:
        env["CUDA_VISIBLE_DEVICES"] = ",".join(str(g) for g in gpus)
    with open(os.path.join(trial_dir, "hparams.json"), "w") as fh:
        json.dump(hparams, fh, indent=2, default=str)
    runner = os.path.join(trial_dir, "_run_trial.py")
    with open(runner, "w") as fh:
        fh.write(_TRIAL_SNIPPET)
    if len(gpus) > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={len(gpus)}", "--standalone",
               "--local-addr", "127.0.0.1", runner, script, json.dumps(hp)]
    else:
        cmd = [sys.executable, runner, script, json.dumps(hp)]
    log = open(os.path.join(trial_dir, "stdout.log"), "w")
    return subprocess.Popen(cmd, env=env, stdout=log, stderr=subprocess.STDOUT)


def read_metric(trial_dir: str, metric: str, mode: str) -> Tuple[Optional[float], Optional[float], int]:
    """``(best, last, n_points)`` of ``metric`` over every jsonl log of a trial."""
    values: List[float] = []
    for name in sorted(os.listdir(trial_dir)):
        if not name.endswith(".jsonl"):
            continue
        with open(os.path.join(trial_dir, name)) as fh:
            for line in fh:
                try:
                    rec = json.loads(line)
                except json.JSONDecodeError:
                    continue
                if metric in rec and isinstance(rec[metric], (int, float)) and math.isfinite(rec[metric]):
                    values.append(float(rec[metric]))
    if not values:
        return None, None, 0
    return (max(values) if mode == "max" else min(values)), values[-1], len(values)


def run_sweep(script: str, sweep_config: Dict[str, Any], out_dir: str, num_gpus: int = 1, gpu_ids: Optional[List[int]] = None,
              default_config: Optional[str] = None, seed: int = 0, poll: float = 1.0) -> List[Dict[str, Any]]:
    tune_config = dict(sweep_config.get("tune_config", {}))
    metric, mode = tune_config.get("metric", "reward/mean"), tune_config.get("mode", "max")
    space = get_param_space(sweep_config)
    searcher = get_search_alg(tune_config, space, seed)
    sched = get_scheduler(tune_config)
    rungs, eta = sched["rungs"], sched["eta"]
    trials: List[Dict[str, Any]] = []
    os.makedirs(out_dir, exist_ok=True)
    if gpu_ids is None:
        try:
            import torch

            gpu_ids = list(range(torch.cuda.device_count()))
        except Exception:  # pragma: no cover
            gpu_ids = []
    slots: List[List[int]] = ([gpu_ids[i:i + num_gpus] for i in range(0, len(gpu_ids) - num_gpus + 1, num_gpus)]
                              if gpu_ids and num_gpus > 0 else [[]])
    max_conc = int(tune_config.get("max_concurrent_trials", len(slots)))
    if slots == [[]]:  # CPU-only node: concurrency is bounded by the config alone
        slots = [[] for _ in range(max(max_conc, 1))]
    slots = slots[:max(max_conc, 1)]

    alive: List[Dict[str, Any]] = []
    for rung_i, budget in enumerate(rungs):
        # rung 0 draws its trials from the searcher as slots free up (model-based searchers see the results so far);
        # later rungs re-run the promoted trials with a larger step budget
        pending = list(alive)
        exhausted = rung_i > 0
        running: List[Tuple[Dict[str, Any], subprocess.Popen, List[int]]] = []
        free = list(slots)
        while pending or running or not exhausted:
            while free and (pending or not exhausted):
                if pending:
                    tr = pending.pop(0)
                else:
                    hp = searcher.suggest()
                    if hp is None:
                        exhausted = True
                        break
                    tr = dict(id=len(trials), hparams=hp)
                    trials.append(tr)
                slot = free.pop(0)
                tdir = os.path.join(out_dir, f"trial_{tr['id']:04d}", f"rung_{rung_i}")
                tr["dir"] = tdir
                proc = launch_trial(script, tr["hparams"], tdir, slot, budget, default_config)
                running.append((tr, proc, slot))
                print(f"[sweep] trial {tr['id']} (rung {rung_i}, budget {budget}) started on gpus {slot}: {tr['hparams']}")
            if not running:
                continue
            time.sleep(poll)
            for item in list(running):
                tr, proc, slot = item
                rc = proc.poll()
                if rc is None:
                    continue
                running.remove(item)
                free.append(slot)
                best, last, n = read_metric(tr["dir"], metric, mode)
                tr.update(returncode=rc, best=best, last=last, points=n, budget=budget)
                if rung_i == 0:
                    searcher.observe(tr["hparams"], best)
                print(f"[sweep] trial {tr['id']} finished rc={rc} {metric}: best={best} last={last}")
        if rung_i == 0:
            alive = list(trials)
        if rung_i < len(rungs) - 1:
            scored = [t for t in alive if t.get("best") is not None]
            scored.sort(key=lambda t: t["best"], reverse=(mode == "max"))
            alive = scored[:max(len(scored) // eta, 1)]
    results = sorted(trials, key=lambda t: (t.get("best") is None, -(t.get("best") or 0) if mode == "max" else (t.get("best") or 0)))
    with open(os.path.join(out_dir, "sweep_results.json"), "w") as fh:
        json.dump(dict(metric=metric, mode=mode, script=script, trials=results), fh, indent=2, default=str)
    write_report(results, space, metric, mode, script, os.path.join(out_dir, "report.md"))
    return results


def write_report(results, space, metric: str, mode: str, script: str, path: str) -> None:
    """Markdown stand-in for the reference's W&B report: best configuration, full trial table, per-parameter view."""
    keys = list(space)
    lines = [f"# Sweep report — `{script}`", "", f"Target metric: `{metric}` ({mode}); {len(results)} trials; "
             f"generated {datetime.now().isoformat(timespec='seconds')}", ""]
    done = [r for r in results if r.get("best") is not None]
    if done:
        best = done[0]
        lines += ["## Best configuration", "", "```json", json.dumps(best["hparams"], indent=2, default=str), "```",
                  f"`{metric}` = {best['best']:.6g}", ""]
    lines += ["## Trials", "", "| trial | " + " | ".join(keys) + f" | best {metric} | last | points | rc |",
              "|---|" + "---|" * (len(keys) + 4)]
    for r in results:
        vals = " | ".join(f"{r['hparams'].get(k):.4g}" if isinstance(r["hparams"].get(k), float) else str(r["hparams"].get(k))
                          for k in keys)
        b = "-" if r.get("best") is None else f"{r['best']:.5g}"
        la = "-" if r.get("last") is None else f"{r['last']:.5g}"
        lines.append(f"| {r['id']} | {vals} | {b} | {la} | {r.get('points', 0)} | {r.get('returncode')} |")
    if done:
        lines += ["", "## Parameter view (trials sorted by each parameter)", ""]
        for k in keys:
            pts = sorted(((r["hparams"].get(k), r["best"]) for r in done), key=lambda p: str(p[0]))
            lines.append(f"* `{k}`: " + ", ".join(f"{v if not isinstance(v, float) else format(v, '.3g')}→{b:.4g}" for v, b in pts))
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def create_report(target_metric, column_names, entity_name, project_name, group_name, best_config):
    """Publish a Weights & Biases report for a finished sweep whose trials logged to ``project_name`` / ``group_name``
    (reference: ``trlx/sweep.py:177-264``): parallel coordinates and parameter importance over the swept columns, a scatter of the
    target metric per trial, line plots of every logged metric family, and the best configuration.  Sweeps here always get the
    local Markdown report (:func:`write_report`); this is the optional W&B view of the same runs.  Returns the report URL, or
    ``None`` when the ``wandb`` reports API is unavailable (offline boxes)."""
    try:
        import wandb.apis.reports as wb

        needed = ("Report", "Runset", "PanelGrid", "ParallelCoordinatesPlot", "ParameterImportancePlot", "ScatterPlot", "LinePlot",
                  "PCColumn", "H2", "CodeBlock")
        lacking = [n for n in needed if not hasattr(wb, n)]
        if lacking:
            raise ImportError(f"wandb reports API incomplete (no {lacking[0]}; `pip install wandb[workspaces]`)")
    except Exception as err:  # no wandb / stubbed wandb: the Markdown report is the deliverable
        print(f"[sweep] W&B report skipped ({type(err).__name__}: {err}); see report.md")
        return None
    runs = wb.Runset(project=project_name, entity=entity_name).set_filters_with_python_expr(f'group == "{group_name}"')
    swept = [wb.PCColumn(f"c::{c}") for c in column_names]
    overview = wb.PanelGrid(runsets=[runs], panels=[
        wb.ParallelCoordinatesPlot(columns=swept + [wb.PCColumn(target_metric)], layout={"x": 0, "y": 0, "w": 24, "h": 10}),
        wb.ParameterImportancePlot(with_respect_to=target_metric, layout={"x": 0, "y": 10, "w": 12, "h": 10}),
        wb.ScatterPlot(x="Index", y=target_metric, running_ymax=True, font_size="small",
                       layout={"x": 12, "y": 10, "w": 12, "h": 10}),
    ])
    families = {}  # one line plot per logged metric, grouped by prefix ("reward/", "losses/", ...)
    try:
        import wandb

        for run in wandb.Api().runs(f"{entity_name}/{project_name}" if entity_name else project_name, filters={"group": group_name}):
            for key in run.history(samples=1).columns:
                if not key.startswith("_"):
                    families.setdefault(key.split("/")[0], set()).add(key)
            break
    except Exception:
        families = {target_metric.split("/")[0]: {target_metric}}
    curves, row = [], 0
    for _, keys in sorted(families.items()):
        for i, key in enumerate(sorted(keys)):
            curves.append(wb.LinePlot(x="Step", y=[key], title=key, layout={"x": 12 * (i % 2), "y": row, "w": 12, "h": 8}))
            row += 8 * (i % 2)
        row += 8
    report = wb.Report(project=project_name, entity=entity_name, title=f"Hyperparameter sweep: {project_name}",
                       description=group_name)
    report.blocks = [overview, wb.PanelGrid(runsets=[runs], panels=curves),
                     wb.H2(text="Best configuration"), wb.CodeBlock(code=[json.dumps(best_config, indent=2, default=str)], language="json")]
    report.save()
    print(f"[sweep] W&B report: {report.url}")
    return report.url


def main(argv: Optional[List[str]] = None) -> int:
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("script", type=str, help="Path to the example script (must define main(hparams))")
    parser.add_argument("--config", type=str, required=True, help="Param-space YAML (configs/sweeps/*.yml)")
    parser.add_argument("--default_config", type=str, default=None, help="Default TRLConfig YAML for the script")
    parser.add_argument("--num_gpus", type=int, default=1, help="GPUs (ranks) per trial")
    parser.add_argument("--num_cpus", type=int, default=4, help="Accepted for CLI compatibility (unused)")
    parser.add_argument("--output_dir", type=str, default=None, help="Where trial logs and the report go")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("-y", "--assume_yes", action="store_true", help="Don't ask for confirmation")
    parser.add_argument("--server_address", type=str, default=None, help="Accepted for CLI compatibility (no Ray cluster)")
    args = parser.parse_args(argv)
    with open(args.config) as fh:
        sweep_config = yaml.safe_load(fh)
    out = args.output_dir or os.path.join("sweeps", os.path.splitext(os.path.basename(args.script))[0] + "_"
                                          + datetime.now().strftime("%Y%m%d_%H%M%S"))
    print(f'Running `main(hparams)` of "{args.script}" for every trial; results in {out}')
    if not args.assume_yes and sys.stdin.isatty():
        if input("Proceed? [y/N] ").strip().lower() not in ("y", "yes"):
            return 1
    results = run_sweep(args.script, sweep_config, out, num_gpus=args.num_gpus, default_config=args.default_config,
                        seed=args.seed)
    ok = [r for r in results if r.get("best") is not None]
    print(f"[sweep] {len(ok)}/{len(results)} trials reported `{sweep_config.get('tune_config', {}).get('metric')}`; "
          f"report: {os.path.join(out, 'report.md')}")
    return 0 if ok else 2


if __name__ == "__main__":
    sys.exit(main())
