"""Pairwise-preference reward model for TL;DR summaries (reference: examples/summarize_rlhf/reward_model/reward_model.py).

Same objective as the reference — `-log σ(r_chosen − r_rejected)` averaged over the positions where the two sequences differ,
scores read at the last non-pad token — but vectorised over the batch (the reference loops over pairs in Python)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200.models.modeling_base import build_base_model


class GPTRewardModel(nn.Module):
    def __init__(self, model_path, pad_id: int):
        super().__init__()
        self.transformer = build_base_model(model_path)
        self.config = self.transformer.config
        self.v_head = nn.Linear(self.config.hidden_size, 1, bias=False)
        self.PAD_ID = pad_id

    def rewards(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        h = self.transformer(input_ids=input_ids, attention_mask=attention_mask, compute_logits=False).last_hidden_state
        return self.v_head(h.to(self.v_head.weight.dtype)).squeeze(-1)

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, **unused) -> Dict[str, torch.Tensor]:
        """`input_ids = cat(chosen, rejected)` along the batch.  Identical halves ⇒ inference (scores only)."""
        rewards = self.rewards(input_ids, attention_mask)
        bs = input_ids.shape[0] // 2
        chosen, rejected = input_ids[:bs], input_ids[bs:]
        rc, rr = rewards[:bs], rewards[bs:]
        T = chosen.shape[1]
        pos = torch.arange(T, device=input_ids.device).unsqueeze(0)

        def first_pad(x):
            is_pad = x == self.PAD_ID
            return torch.where(is_pad.any(1), is_pad.float().argmax(1), torch.full((x.shape[0],), T, device=x.device))

        c_end, r_end = first_pad(chosen), first_pad(rejected)
        same = (chosen == rejected).all(1)
        if bool(same.all()):
            return {"chosen_end_scores": rc.gather(1, (c_end - 1).clamp_min(0).unsqueeze(1)).squeeze(1)}
        end = torch.maximum(c_end, r_end)
        diverge = (chosen != rejected).float().argmax(1)
        window = ((pos >= diverge.unsqueeze(1)) & (pos < end.unsqueeze(1)) & ~same.unsqueeze(1)).float()
        pair_loss = -(F.logsigmoid(rc - rr) * window).sum(1) / window.sum(1).clamp_min(1)
        last = (end - 1).clamp_min(0).unsqueeze(1)
        return {"loss": pair_loss.sum() / bs, "chosen_end_scores": rc.gather(1, last).squeeze(1),
                "rejected_end_scores": rr.gather(1, last).squeeze(1)}
