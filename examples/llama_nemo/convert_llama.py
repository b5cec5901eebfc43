"""Checkpoint layout conversion for the model-parallel trainers (reference: examples/llama_nemo/convert_llama_to_nemo.py).

    python examples/llama_nemo/convert_llama.py shard   <hf_dir> <out_dir> --tp 4     # HF → mp_rank_XX/model_weights.ckpt
    python examples/llama_nemo/convert_llama.py unshard <ckpt_dir> <out_hf_dir> --tp 4 # and back to one HF state dict
"""
import argparse
import os

import torch

from trlx_b200.models.checkpoint_io import load_state_dict, save_config, save_state_dict
from trlx_b200.models.modeling_base import build_base_model, export_base_state_dict, hf_config_dict, import_base_state_dict
from trlx_b200.parallel.tensor_parallel import shard_state_dict, unshard_state_dicts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("command", choices=["shard", "unshard"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--tp", type=int, default=4)
    a = ap.parse_args()
    if a.command == "shard":
        model = build_base_model(a.src)
        import_base_state_dict(model, load_state_dict(a.src), strict=False)
        full = {k: v for k, v in model.state_dict().items()}
        for r in range(a.tp):
            sub = os.path.join(a.dst, f"mp_rank_{r:02d}")
            os.makedirs(sub, exist_ok=True)
            torch.save(shard_state_dict(model.config, full, r, a.tp), os.path.join(sub, "model_weights.ckpt"))
        save_config(a.dst, hf_config_dict(model))
    else:
        shards = [torch.load(os.path.join(a.src, f"mp_rank_{r:02d}", "model_weights.ckpt"), map_location="cpu") for r in range(a.tp)]
        model = build_base_model(a.src)
        model.load_state_dict(unshard_state_dicts(model.config, shards), strict=False)
        save_state_dict(a.dst, export_base_state_dict(model))
        save_config(a.dst, hf_config_dict(model))


if __name__ == "__main__":
    main()
