"""`accelerate launch` counterpart:  ``python -m trlx_b200.launch --config_file configs/accelerate/zero2-bf16.yaml
[--num_processes N] [--main_process_port P] script.py [script args]``

Reference counterpart: the reference is started with `accelerate launch --config_file configs/accelerate/*.yaml …`
(``README.md:95-101``, ``scripts/accelerate_train_example.sh``).  Here a launch preset is a small YAML with ``num_processes``
and a ``parallel:`` block; the launcher exports it as ``TRLX_B200_PARALLEL`` (JSON), which :func:`trlx_b200.trlx.train`
folds into ``config.train.parallel``, and starts one process per GPU with ``torch.distributed.run`` on 127.0.0.1 (or the
SLURM-provided rendezvous for multi-node jobs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from typing import List, Optional

import yaml


def build_command(args, preset: dict) -> List[str]:
    nproc = args.num_processes or preset.get("num_processes") or 1
    try:
        import torch

        if torch.cuda.is_available():
            nproc = min(int(nproc), torch.cuda.device_count()) if not args.num_processes else int(nproc)
    except Exception:  # pragma: no cover
        pass
    nnodes = int(args.num_machines or preset.get("num_machines", 1))
    cmd = [sys.executable, "-m", "torch.distributed.run", f"--nnodes={nnodes}", f"--nproc-per-node={nproc}"]
    if nnodes == 1:
        cmd += ["--master-addr", "127.0.0.1", "--master-port", str(args.main_process_port)]
    else:
        cmd += ["--node-rank", str(args.machine_rank), "--master-addr", args.main_process_ip, "--master-port", str(args.main_process_port)]
    return cmd + [args.script] + list(args.script_args)


def parallel_from_deepspeed(ds: dict) -> dict:
    """The part of a DeepSpeed JSON config (``examples/summarize_rlhf/configs/ds_config_*.json`` in the reference) that has a
    counterpart here: ZeRO stage → ``zero_stage`` of the fused sharded optimizer, ``bf16`` / ``fp16`` → ``precision``,
    ``gradient_clipping`` → ``grad_clip``, an all-gather / reduce bucket size → ``bucket_mb``.  CPU offload sections are
    ignored (optimizer state is sharded across 180 GB GPUs instead)."""
    out = {}
    zero = ds.get("zero_optimization") or {}
    if "stage" in zero:
        out["zero_stage"] = int(zero["stage"])
    if (ds.get("bf16") or {}).get("enabled"):
        out["precision"] = "bf16"
    elif (ds.get("fp16") or {}).get("enabled"):
        out["precision"] = "fp16"
    if ds.get("gradient_clipping"):
        out["grad_clip"] = float(ds["gradient_clipping"])
    bucket = zero.get("reduce_bucket_size") or zero.get("allgather_bucket_size")
    if bucket:
        out["bucket_mb"] = max(float(bucket) * 2 / (1 << 20), 1.0)  # elements (fp16) → MiB
    return out


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(description="start a trlx_b200 training script on every GPU of the node")
    ap.add_argument("--config_file", type=str, default=None, help="launch preset (configs/accelerate/*.yaml)")
    ap.add_argument("--num_processes", type=int, default=None)
    ap.add_argument("--num_machines", type=int, default=None)
    ap.add_argument("--machine_rank", type=int, default=int(os.environ.get("SLURM_NODEID", 0)))
    ap.add_argument("--main_process_ip", type=str, default=os.environ.get("MASTER_ADDR", "127.0.0.1"))
    ap.add_argument("--main_process_port", type=int, default=int(os.environ.get("MASTER_PORT", 29500)))
    ap.add_argument("--deepspeed_config", type=str, default=None,
                    help="DeepSpeed JSON whose ZeRO stage / precision / clipping are mapped onto the parallel preset")
    ap.add_argument("--dry_run", action="store_true", help="print the command and exit")
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)
    preset = {}
    if args.config_file:
        with open(args.config_file) as fh:
            preset = yaml.safe_load(fh) or {}
    env = dict(os.environ)
    ds_path = args.deepspeed_config or (preset.get("deepspeed_config") or {}).get("deepspeed_config_file")
    if ds_path:
        with open(ds_path) as fh:
            preset = dict(preset, parallel={**parallel_from_deepspeed(json.load(fh)), **(preset.get("parallel") or {})})
    if preset.get("parallel"):
        env["TRLX_B200_PARALLEL"] = json.dumps(preset["parallel"])
    cmd = build_command(args, preset)
    if args.dry_run:
        print(" ".join(cmd))
        print("TRLX_B200_PARALLEL=" + env.get("TRLX_B200_PARALLEL", ""))
        return 0
    return subprocess.call(cmd, env=env)


if __name__ == "__main__":
    sys.exit(main())
