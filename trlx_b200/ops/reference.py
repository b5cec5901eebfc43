"""Plain-PyTorch (fp32) twins of every kernel: the CPU execution path AND the numerical oracle the GPU
tests compare against.  Formulas follow the reference implementation cited in each docstring."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def activation(x: torch.Tensor, act: str) -> torch.Tensor:
    if act in ("", "none", None):
        return x
    if act in ("gelu_new", "gelu_tanh", "gelu_pytorch_tanh", "gelu_fast"):
        return F.gelu(x, approximate="tanh")
    if act == "gelu":
        return F.gelu(x)
    if act == "relu":
        return F.relu(x)
    if act in ("silu", "swish"):
        return F.silu(x)
    raise ValueError(act)


def linear(x, w, b=None, act: str = "none", residual=None):
    y = F.linear(x, w, b)
    y = activation(y, act)
    return y if residual is None else y + residual


def fused_logprob(h, w, b, labels) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(log p(label), logsumexp)`` of ``logits = h @ w.T + b`` (``trlx/utils/modeling.py:213-219``)."""
    logits = F.linear(h, w, b).float()
    lse = torch.logsumexp(logits, -1)
    safe = labels.clamp_min(0)
    picked = logits.gather(-1, safe.unsqueeze(-1)).squeeze(-1)
    lp = torch.where(labels >= 0, picked - lse, torch.zeros_like(lse))
    return lp, lse


def gae(values: torch.Tensor, rewards: torch.Tensor, width: int, gamma: float, lam: float):
    """Reverse-scan GAE over the first ``width`` columns (``trlx/models/modeling_ppo.py:161-170``)."""
    v, r = values[:, :width].float(), rewards[:, :width].float()
    adv = torch.zeros_like(v)
    last = torch.zeros(v.shape[0], dtype=v.dtype, device=v.device)
    for t in reversed(range(width)):
        nxt = v[:, t + 1] if t < width - 1 else 0.0
        delta = r[:, t] + gamma * nxt - v[:, t]
        last = delta + gamma * lam * last
        adv[:, t] = last
    ret = adv + v
    return adv, ret


def ppo_loss(logprobs, values, old_logprobs, old_values, advantages, returns, mask, cliprange, cliprange_value,
             vf_coef, width_tensor=None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Clipped PPO objective + statistics (``trlx/models/modeling_ppo.py:189-238``), differentiable."""
    from trlx_b200.utils.modeling import get_tensor_stats

    mask = mask.float()
    values_clipped = torch.clamp(values, old_values - cliprange_value, old_values + cliprange_value)
    n = mask.sum().clamp_min(1)  # an all-padding batch yields zero loss/gradients, not NaN
    vf1, vf2 = (values - returns) ** 2, (values_clipped - returns) ** 2
    vf_loss = 0.5 * torch.sum(torch.max(vf1, vf2) * mask) / n
    vf_clipfrac = torch.sum((vf2 > vf1).float() * mask) / n
    log_ratio = (logprobs - old_logprobs) * mask
    ratio = torch.exp(log_ratio)
    with torch.no_grad():
        approx_kl = torch.mean((ratio - 1) - log_ratio)
    pg1 = -advantages * ratio
    pg2 = -advantages * torch.clamp(ratio, 1.0 - cliprange, 1.0 + cliprange)
    pg_loss = torch.sum(torch.max(pg1, pg2) * mask) / n
    pg_clipfrac = torch.sum((pg2 > pg1).float() * mask) / n
    loss = pg_loss + vf_coef * vf_loss
    with torch.no_grad():
        stats = {
            "losses/total_loss": loss.detach(), "losses/policy_loss": pg_loss.detach(), "losses/value_loss": vf_loss.detach(),
            **{f"values/{k}": v for k, v in get_tensor_stats(values.detach(), mask, n).items()},
            "values/values_error": torch.sum(((values - returns) * mask) ** 2) / n,
            "values/values_mape_error": torch.sum((torch.abs(values - returns) * mask) / torch.abs(returns * mask + 1e-2)) / n,
            "values/clipfrac": vf_clipfrac,
            **{f"old_values/{k}": v for k, v in get_tensor_stats(old_values, mask, n).items()},
            **{f"returns/{k}": v for k, v in get_tensor_stats(returns, mask, n).items()},
            "policy/approx_kl": approx_kl, "policy/clipfrac": pg_clipfrac,
            "ratio": (ratio * mask).sum() / n,
            "padding_percentage": 1 - n / mask.numel(),
        }
    return loss, stats


def rollout_rewards(logprobs, ref_logprobs, values, mask, scores, start: int, kl_coef: float):
    """What ``make_experience`` derives from a scored rollout (``trlx/trainer/accelerate_ppo_trainer.py:455-504``):
    ``(rewards, logprobs, values)`` on the response window ``[start, T-1)`` zeroed past ``slice_len``, ``slice_len`` and the
    summed k3 KL estimate over all positions.  ``scores``: one scalar per row, added on the last scored token."""
    R = logprobs.shape[1] - start
    log_ratio = (logprobs - ref_logprobs) * mask[:, :-1]
    kl = log_ratio.exp() - 1 - log_ratio
    slice_len = (mask[:, start:].sum(1) + 1).clamp(max=R)
    cols = torch.arange(R, device=logprobs.device).unsqueeze(0)
    valid = cols < slice_len.unsqueeze(1)
    zero = torch.zeros_like(logprobs[:, start:])
    lp_s = torch.where(valid, logprobs[:, start:], zero)
    v_s = torch.where(valid, values[:, start:], zero)
    rewards = torch.where(valid, -kl_coef * log_ratio[:, start:], zero)
    rewards.scatter_add_(1, (slice_len - 1).clamp_min(0).unsqueeze(1), scores.reshape(-1, 1).to(rewards.dtype))
    return rewards, lp_s, v_s, slice_len, kl.sum()


def adamw_step(w, g, m, v, step: int, lr, beta1, beta2, eps, weight_decay, decoupled=True):
    if not decoupled:
        g = g + weight_decay * w
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    if decoupled:
        w.mul_(1 - lr * weight_decay)
    w.addcdiv_(m / bc1, (v / bc2).sqrt() + eps, value=-lr)
    return w
