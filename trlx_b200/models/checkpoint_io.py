"""Weight-file IO for HF-style model directories (``config.json`` + ``pytorch_model.bin`` / ``model.safetensors`` /
sharded ``*.index.json``).  Parity: the loading branch of ``trlx/models/modeling_base.py:275-315``, minus the hub
download (B200 boxes are offline)."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

WEIGHTS_BIN = "pytorch_model.bin"
WEIGHTS_BIN_INDEX = "pytorch_model.bin.index.json"
WEIGHTS_SAFE = "model.safetensors"
WEIGHTS_SAFE_INDEX = "model.safetensors.index.json"
CONFIG_NAME = "config.json"


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu", weights_only=True)


def has_weights(directory: str) -> bool:
    return any(os.path.exists(os.path.join(directory, f))
               for f in (WEIGHTS_BIN, WEIGHTS_BIN_INDEX, WEIGHTS_SAFE, WEIGHTS_SAFE_INDEX))


def load_state_dict(directory: str) -> Optional[Dict[str, torch.Tensor]]:
    """Merge every weight file of ``directory`` into one CPU state dict (``None`` if there is none)."""
    for single in (WEIGHTS_BIN, WEIGHTS_SAFE):
        p = os.path.join(directory, single)
        if os.path.exists(p):
            return _load_file(p)
    for index in (WEIGHTS_BIN_INDEX, WEIGHTS_SAFE_INDEX):
        p = os.path.join(directory, index)
        if os.path.exists(p):
            with open(p) as fh:
                shards = sorted(set(json.load(fh)["weight_map"].values()))
            merged: Dict[str, torch.Tensor] = {}
            for shard in shards:
                merged.update(_load_file(os.path.join(directory, shard)))
            return merged
    return None


def save_state_dict(directory: str, state_dict: Dict[str, torch.Tensor], max_shard_bytes: Optional[int] = None,
                    safe_serialization: bool = False) -> None:
    """Write ``state_dict`` as one file or as size-bounded shards with an index (tied tensors are stored once per key,
    on CPU, detached)."""
    os.makedirs(directory, exist_ok=True)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in state_dict.items()}
    ext = ".safetensors" if safe_serialization else ".bin"
    stem = "model" if safe_serialization else "pytorch_model"

    def write(path, part):
        if safe_serialization:
            from safetensors.torch import save_file

            save_file({k: v.clone() for k, v in part.items()}, path, metadata={"format": "pt"})
        else:
            torch.save(part, path)

    total = sum(v.numel() * v.element_size() for v in sd.values())
    if not max_shard_bytes or total <= max_shard_bytes:
        write(os.path.join(directory, stem + ext), sd)
        return
    shards, cur, cur_bytes = [], {}, 0
    for k, v in sd.items():
        nbytes = v.numel() * v.element_size()
        if cur and cur_bytes + nbytes > max_shard_bytes:
            shards.append(cur)
            cur, cur_bytes = {}, 0
        cur[k] = v
        cur_bytes += nbytes
    if cur:
        shards.append(cur)
    weight_map = {}
    for i, part in enumerate(shards):
        name = f"{stem}-{i + 1:05d}-of-{len(shards):05d}{ext}"
        write(os.path.join(directory, name), part)
        weight_map.update({k: name for k in part})
    with open(os.path.join(directory, f"{stem}{ext}.index.json"), "w") as fh:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, fh, indent=2)


def save_config(directory: str, config: dict) -> None:
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, CONFIG_NAME), "w") as fh:
        json.dump(config, fh, indent=2, default=str)
