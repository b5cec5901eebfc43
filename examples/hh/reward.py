"""Reward model for the HH examples + three ways to call it (reference: examples/hh/ppo_hh.py:113-200, hh/to_triton.py).

* ``RewardModel``       – causal LM trunk + scalar head read at the first EOS (the reference's GPT-J reward model shape).
* in-process            – the model lives on the last GPU of rank 0's node and scores ``chosen − original`` deltas.
* served                – ``python -m examples.hh.reward serve --port 8000`` exposes ``POST /score`` (FastAPI/uvicorn); set
                          ``REWARD_HOST=host:port`` and every rank scores through HTTP.  This replaces the reference's Triton
                          Inference Server deployment (``TRITON_HOST``) with something that runs in this image.
Without a trained checkpoint (offline) the "model" is a rubric: polite, on-topic, non-dismissive answers score higher.
"""
from __future__ import annotations

import json
import math
import os
import urllib.request
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn


class RewardModel(nn.Module):
    def __init__(self, config_or_path, eos_token_id: int):
        super().__init__()
        from trlx_b200.models.modeling_base import build_base_model

        lm = build_base_model(config_or_path)
        self.transformer = lm
        self.v_head = nn.Linear(lm.config.hidden_size, 1, bias=False)
        self.eos_token_id = eos_token_id

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        states = self.transformer(input_ids=input_ids, compute_logits=False).last_hidden_state
        rewards = self.v_head(states.to(self.v_head.weight.dtype)).squeeze(-1)
        ends = torch.argmax((input_ids == self.eos_token_id).float(), dim=1).view(-1, 1)
        return torch.gather(rewards, 1, ends).squeeze(-1)


def rubric_score(sample: str) -> float:
    answer = sample.split("Assistant:")[-1].lower()
    good = sum(w in answer for w in ("help", "start", "practise", "practice", "step", "try", "basics", "question", "sure"))
    bad = sum(w in answer for w in ("no.", "do not care", "yourself", "stupid", "shut"))
    return float(good) - 2.0 * bad + min(len(answer.split()), 30) / 30.0


def _score_local(model: Optional[RewardModel], tokenizer, samples: Sequence[str], device, batch_size: int = 48) -> torch.Tensor:
    if model is None:
        return torch.tensor([rubric_score(s) for s in samples])
    enc = tokenizer(list(samples), padding=True, truncation=True, max_length=1024, return_tensors="pt")
    out = []
    for i in range(math.ceil(len(samples) / batch_size)):
        out.append(model(enc.input_ids[i * batch_size:(i + 1) * batch_size].to(device)).float().cpu())
    return torch.cat(out)


def create_reward_fn(checkpoint: Optional[str] = None, delta_reward: bool = True) -> Callable:
    """``reward_fn(samples, prompts, outputs, original_output=..., **kw)``; on ranks other than 0 (in-process mode) the
    trainer's rank-0 funnel is used (``trainer_kwargs.rank0_reward``), like the reference's ``reward_fn = True`` placeholder."""
    host = os.environ.get("REWARD_HOST") or os.environ.get("TRITON_HOST")
    if host:
        url = f"http://{host.split('/')[0]}/score"

        def get_reward(samples: List[str]) -> torch.Tensor:
            req = urllib.request.Request(url, data=json.dumps({"samples": samples}).encode(), headers={"Content-Type": "application/json"})
            with urllib.request.urlopen(req, timeout=600) as resp:
                return torch.tensor(json.loads(resp.read())["rewards"])
    else:
        model, tok, device = None, None, "cpu"
        if checkpoint:
            from trlx_b200.utils.tokenizer import load_tokenizer

            tok = load_tokenizer("gpt2")
            tok.pad_token, tok.truncation_side = tok.eos_token, "left"
            model = RewardModel(checkpoint, tok.eos_token_id)
            sd = torch.load(os.path.join(checkpoint, "reward_model.pt"), map_location="cpu") if os.path.isdir(checkpoint) else {}
            model.load_state_dict(sd, strict=False)
            if torch.cuda.is_available():
                device = torch.device("cuda", torch.cuda.device_count() - 1)  # reward model on the last GPU
                model = model.half().to(device)
            model.eval().requires_grad_(False)

        def get_reward(samples: List[str]) -> torch.Tensor:
            return _score_local(model, tok, samples, device)

    def reward_fn(samples, prompts=None, outputs=None, original_output=None, **kwargs):
        eos = "<|endoftext|>"
        rewards = get_reward([s + eos for s in samples])
        if not delta_reward or original_output is None:
            return rewards
        original = get_reward([p + o + eos for p, o in zip(prompts, original_output)])
        return rewards - original

    return reward_fn


def serve(port: int = 8000, checkpoint: Optional[str] = None):  # pragma: no cover - network service
    import uvicorn
    from fastapi import FastAPI

    app = FastAPI()
    fn = create_reward_fn(checkpoint, delta_reward=False)

    @app.post("/score")
    def score(body: dict):
        return {"rewards": [float(x) for x in fn(body["samples"])]}

    uvicorn.run(app, host="0.0.0.0", port=port)


if __name__ == "__main__":  # python -m examples.hh.reward serve --port 8000 [--checkpoint DIR]
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("command", choices=["serve"])
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--checkpoint", type=str, default=None)
    a = ap.parse_args()
    serve(a.port, a.checkpoint)
