"""A/B learning-curve regression check:  ``python -m trlx_b200.reference <ref> --against <base-ref>``

Reference counterpart: ``trlx/reference.py:1-103`` + ``scripts/benchmark.sh`` (clone two branches from GitHub, run the
example set on each with W&B tags = content hash, build a W&B report of paired line plots).  Here everything is local and
offline: both refs are materialised from *this* git repository (``git archive``), identified by the same content hash
(sha1 over all non-Markdown files), the example set is run through ``scripts/benchmark.sh`` with the ``jsonl`` tracker into
``benchmark_logs/<hash>/``, runs that already exist are reused, and the comparison is a Markdown report: for every
experiment and metric the final / mean values of both sides and their delta, reward & metric curves first.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import tarfile
import tempfile
from typing import Dict, List, Optional, Tuple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def content_hash(tree: str) -> str:
    """sha1 over (path, bytes) of every file except ``.git`` and ``*.md`` (the reference's identity for a code state)."""
    h = hashlib.sha1()
    for base, dirs, files in os.walk(tree):
        dirs[:] = sorted(d for d in dirs if d not in (".git", "__pycache__", "benchmark_logs", "gpurun_out", "build"))
        for name in sorted(files):
            if name.endswith((".md", ".so", ".o", ".pyc", ".stamp")):
                continue
            path = os.path.join(base, name)
            h.update(os.path.relpath(path, tree).encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def materialise(ref: str, dest: str) -> str:
    """Extract ``ref`` of this repository into ``dest``; returns ``<short hash>/<subject>/<date>``."""
    with tempfile.NamedTemporaryFile(suffix=".tar") as tmp:
        subprocess.run(["git", "-C", ROOT, "archive", "--format=tar", "-o", tmp.name, ref], check=True)
        with tarfile.open(tmp.name) as tf:
            tf.extractall(dest)
    return subprocess.run(["git", "-C", ROOT, "log", "--format=%h/%s/%as", "-n1", ref], check=True, capture_output=True,
                          text=True).stdout.strip()


def load_runs(log_dir: str) -> Dict[str, Dict[str, List[Tuple[int, float]]]]:
    """``{experiment: {metric: [(step, value), ...]}}`` from every ``*.jsonl`` under ``log_dir``."""
    runs: Dict[str, Dict[str, List[Tuple[int, float]]]] = {}
    if not os.path.isdir(log_dir):
        return runs
    for base, _, files in os.walk(log_dir):
        for name in files:
            if not name.endswith(".jsonl"):
                continue
            exp = os.path.relpath(base, log_dir).split(os.sep)[0]
            exp = exp if exp != "." else name.split(":")[0].split(".jsonl")[0]
            series = runs.setdefault(exp, {})
            with open(os.path.join(base, name)) as fh:
                for line in fh:
                    try:
                        rec = json.loads(line)
                    except json.JSONDecodeError:
                        continue
                    step = rec.get("step") or 0
                    for k, v in rec.items():
                        if k != "step" and isinstance(v, (int, float)) and math.isfinite(v):
                            series.setdefault(k, []).append((int(step), float(v)))
    return runs


def _summary(points: Optional[List[Tuple[int, float]]]) -> Tuple[Optional[float], Optional[float]]:
    if not points:
        return None, None
    vals = [v for _, v in points]
    return vals[-1], sum(vals) / len(vals)


def compare(a: Dict, b: Dict, name_a: str, name_b: str) -> str:
    lines = [f"# {name_a} v. {name_b}", ""]
    for exp in sorted(set(a) | set(b)):
        ma, mb = a.get(exp, {}), b.get(exp, {})
        metrics = sorted(set(ma) | set(mb), key=lambda m: (not (m.startswith("reward") or m.startswith("metric")), m))
        lines += [f"## {exp}", "", f"| metric | {name_a} final | {name_b} final | Δ final | {name_a} mean | {name_b} mean |",
                  "|---|---|---|---|---|---|"]
        for m in metrics:
            fa, ava = _summary(ma.get(m))
            fb, avb = _summary(mb.get(m))
            fmt = lambda x: "-" if x is None else f"{x:.5g}"  # noqa: E731
            delta = "-" if fa is None or fb is None else f"{fa - fb:+.4g}"
            lines.append(f"| {m} | {fmt(fa)} | {fmt(fb)} | {delta} | {fmt(ava)} | {fmt(avb)} |")
        lines.append("")
    return "\n".join(lines)


def ensure_runs(ref: str, only_tiny: bool, logs_root: str) -> Tuple[str, str, str]:
    with tempfile.TemporaryDirectory(prefix="trlx_ref_") as tmp:
        git_hash = materialise(ref, tmp)
        h = content_hash(tmp)
        log_dir = os.path.join(logs_root, h)
        if os.path.isdir(log_dir) and any(f.endswith(".jsonl") for _, _, fs in os.walk(log_dir) for f in fs):
            print(f"On {ref} @{git_hash} these runs were already made: {sorted(os.listdir(log_dir))}")
        else:
            print(f"Making runs on {ref} @{git_hash}")
            cmd = ["bash", os.path.join(tmp, "scripts", "benchmark.sh"), "--logs", log_dir] + (["--only_tiny"] if only_tiny else [])
            subprocess.run(cmd, cwd=tmp, check=False)
        return h, git_hash, log_dir


def main(argv: Optional[List[str]] = None) -> int:
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("branch", type=str, help="git ref of the change (branch, tag or commit of this repository)")
    parser.add_argument("--against", type=str, default="main", help="git ref to compare against")
    parser.add_argument("--only_tiny", action="store_true", help="only the CPU-sized randomwalks experiments")
    parser.add_argument("--logs", type=str, default=os.path.join(ROOT, "benchmark_logs"))
    parser.add_argument("--public", action="store_true", help="accepted for CLI compatibility (no W&B entity here)")
    args = parser.parse_args(argv)
    ref_hash, ref_git, ref_logs = ensure_runs(args.against, args.only_tiny, args.logs)
    pr_hash, pr_git, pr_logs = ensure_runs(args.branch, args.only_tiny, args.logs)
    print(f"{args.branch}: hash={pr_hash} {pr_git}\n{args.against}: hash={ref_hash} {ref_git}")
    report = compare(load_runs(pr_logs), load_runs(ref_logs), args.branch, args.against)
    out = os.path.join(args.logs, f"report_{pr_hash[:8]}_vs_{ref_hash[:8]}.md")
    with open(out, "w") as fh:
        fh.write(report + f"\n\n{args.branch} @{pr_git}\n\n{args.against} @{ref_git}\n")
    print(f"report: {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
