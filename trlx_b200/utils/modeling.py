"""Model / statistics helpers (parity: ``trlx/utils/modeling.py``).

B200 notes
* cross-rank statistics use ONE collective carrying ``(n, mean, M2)`` per rank and Chan's
  parallel-variance merge, instead of the reference's two dependent all-reduces
  (``utils/modeling.py:185-197``).
* ``logprobs_of_labels`` dispatches to the fused online-logsumexp kernel on CUDA so the
  ``[B,T,V]`` fp32 log-softmax is never materialised.
"""
from __future__ import annotations

import functools
from collections.abc import MutableMapping
from typing import Dict, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


# ---- heads / freezing --------------------------------------------------------------------------
def make_head(n_embd: int, out: int, dtype: torch.dtype = torch.float32) -> nn.Sequential:
    """``Linear(n, 2n) → ReLU → Linear(2n, out)``; state-dict keys ``0.*`` and ``2.*``."""
    return nn.Sequential(
        nn.Linear(n_embd, n_embd * 2, dtype=dtype),
        nn.ReLU(),
        nn.Linear(n_embd * 2, out, dtype=dtype),
    )


def freeze_bottom_causal_layers(model: nn.Module, num_layers_unfrozen: int = 0) -> None:
    """Freeze all but the top ``num_layers_unfrozen`` blocks.

    ``0``  → every block + input & output embeddings; ``k>0`` → blocks ``[:-k]`` + input embeddings
    (+ output embeddings when tied); ``-1`` → nothing.  The final norm always stays trainable.
    """
    if num_layers_unfrozen < 0:
        return
    blocks = list(hf_get_decoder_blocks(model))
    if num_layers_unfrozen == 0:
        frozen = blocks + [model.get_input_embeddings(), model.get_output_embeddings()]
    else:
        frozen = blocks[:-num_layers_unfrozen] + [model.get_input_embeddings()]
        if getattr(model.config, "tie_word_embeddings", False):
            frozen.append(model.get_output_embeddings())
    for m in frozen:
        if m is not None:
            m.requires_grad_(False)


def freeze_bottom_seq2seq_layers(model: nn.Module, num_layers_unfrozen: int = 0) -> None:
    """Seq2seq: encoder, embeddings and both final norms are frozen; decoder blocks ``[:-k]`` too."""
    if num_layers_unfrozen == -1:
        return
    dec_blocks = list(model.decoder.block)
    dec_frozen = dec_blocks if num_layers_unfrozen == 0 else dec_blocks[:-num_layers_unfrozen]
    for m in (
        list(model.encoder.block)
        + dec_frozen
        + [model.shared, model.encoder.final_layer_norm, model.decoder.final_layer_norm, model.decoder.embed_tokens]
    ):
        m.requires_grad_(False)


# ---- attribute walking -------------------------------------------------------------------------
def rhasattr(obj, attr: str) -> bool:
    """``hasattr`` over a dotted path."""
    for part in attr.split("."):
        if not hasattr(obj, part):
            return False
        obj = getattr(obj, part)
    return True


def rgetattr(obj, attr: str, *default):
    """``getattr`` over a dotted path."""
    return functools.reduce(lambda o, a: getattr(o, a, *default), attr.split("."), obj)


def findattr(obj, attrs: Sequence[str]):
    for path in attrs:
        if rhasattr(obj, path):
            return rgetattr(obj, path)
    raise ValueError(f"Could not find an attribute from `{attrs}` in `{type(obj).__name__}`")


_DECODER_PATHS = ("transformer", "model.decoder", "model", "gpt_neox", "decoder")
_FINAL_NORM_PATHS = (
    "transformer.ln_f",
    "model.decoder.final_layer_norm",
    "model.norm",
    "decoder.final_layer_norm",
    "gpt_neox.final_layer_norm",
)
_BLOCK_PATHS = (
    "h",
    "layers",
    "model.layers",
    "decoder.layers",
    "transformer.h",
    "transformer.blocks",
    "model.decoder.layers",
    "gpt_neox.layers",
    "decoder.block",
)


def hf_get_decoder(model: nn.Module) -> nn.Module:
    """Causal decoder trunk (``transformer`` / ``model.decoder`` / ``gpt_neox`` / ``decoder``)."""
    return findattr(model, _DECODER_PATHS)


def hf_get_decoder_final_norm(model: nn.Module) -> nn.Module:
    return findattr(model, _FINAL_NORM_PATHS)


def hf_get_decoder_blocks(model: nn.Module):
    return findattr(model, _BLOCK_PATHS)


def hf_get_lm_head(model: nn.Module) -> nn.Module:
    return model.get_output_embeddings()


def hf_get_hidden_size(config) -> int:
    # `final_hidden_size`: ArchSpec of a model whose last hidden state is narrower than the residual stream (OPT-350m project_out)
    return findattr(config, ("final_hidden_size", "hidden_size", "n_embd", "d_model"))


def hf_get_num_hidden_layers(config) -> int:
    return findattr(config, ("num_hidden_layers", "n_layer", "num_layers"))


# ---- statistics --------------------------------------------------------------------------------
def _local_moments(xs: torch.Tensor) -> torch.Tensor:
    x = xs.detach().float()
    n = x.numel()
    mean = x.mean() if n else x.new_zeros(())
    m2 = ((x - mean) ** 2).sum() if n else x.new_zeros(())
    return torch.stack([x.new_tensor(float(n)), mean, m2])


def _merge_moments(stats: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Chan et al. merge of per-rank ``(n, mean, M2)`` rows → global ``(mean, biased var, n)``."""
    n = stats[:, 0]
    tot = n.sum()
    mean = (n * stats[:, 1]).sum() / tot.clamp_min(1e-24)
    m2 = stats[:, 2].sum() + (n * (stats[:, 1] - mean) ** 2).sum()
    return mean, m2 / tot.clamp_min(1e-24), tot


_STATS_GROUP = None


def set_statistics_group(group) -> None:
    """Default process group of the global statistics (whitening, running moments).  The runtime points it at the
    data-parallel group when tensor/pipeline-parallel peers (which hold *identical* data, and under a pipeline schedule
    are at different micro-batches at any moment) share the world."""
    global _STATS_GROUP
    _STATS_GROUP = group


def statistics_group(group=None):
    return _STATS_GROUP if group is None else group


def get_global_statistics(xs: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Global ``(mean, biased variance, count)`` of ``xs`` over ``group`` in one collective."""
    group = statistics_group(group)
    local = _local_moments(xs)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        bucket = local.new_empty((dist.get_world_size(group), 3))
        dist.all_gather_into_tensor(bucket, local.unsqueeze(0), group=group)
    else:
        bucket = local.unsqueeze(0)
    return _merge_moments(bucket)


def whiten(xs: torch.Tensor, shift_mean: bool = True, distributed: bool = True, group=None) -> torch.Tensor:
    """``(xs - mean) * rsqrt(var + 1e-8)``.  Across ranks the variance is the biased global one;
    in a single process it is ``torch.var_mean``'s unbiased estimate — both as in the reference
    (``utils/modeling.py:200-210``)."""
    if distributed and dist.is_available() and dist.is_initialized():
        mean, var, _ = get_global_statistics(xs, group=group)
    else:
        var, mean = torch.var_mean(xs.float())
    out = (xs - mean) * torch.rsqrt(var + 1e-8)
    if not shift_mean:
        out = out + mean
    return out.to(xs.dtype)


def logprobs_of_labels(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """``log_softmax(logits)[labels]`` without materialising the log-softmax on CUDA."""
    if logits.is_cuda:
        from trlx_b200 import ops

        if ops.available():
            return ops.logprobs_from_logits(logits, labels)
    lse = torch.logsumexp(logits.float(), dim=-1)
    picked = torch.gather(logits, -1, labels.unsqueeze(-1)).squeeze(-1).float()
    return picked - lse


def flatten_dict(d: Union[dict, MutableMapping], parent_key: str = "", sep: str = "/") -> dict:
    out = {}
    stack = [(parent_key, d)]
    while stack:
        prefix, node = stack.pop()
        for k, v in node.items():
            key = f"{prefix}{sep}{k}" if prefix else str(k)
            if isinstance(v, MutableMapping):
                stack.append((key, v))
            else:
                out[key] = v
    return out


def gather_dict(obj: Dict, grad_state=None, remainder: Optional[int] = None, group=None) -> Dict:
    """Concatenate dict-of-lists across ranks (pickle all-gather); optionally trim dataloader padding — either an explicit
    ``remainder`` or, as in the reference (``utils/modeling.py:238-259``), an object with ``end_of_dataloader`` / ``remainder``
    attributes (Accelerate's ``GradientState``)."""
    if isinstance(grad_state, int) and remainder is None:  # older call sites passed the remainder positionally
        grad_state, remainder = None, grad_state
    if grad_state is not None and getattr(grad_state, "end_of_dataloader", False) and getattr(grad_state, "remainder", 0) > 0:
        remainder = int(grad_state.remainder)
    if not (dist.is_available() and dist.is_initialized()):
        return obj
    shards = [None] * dist.get_world_size(group)
    dist.all_gather_object(shards, obj, group=group)
    merged = {k: list(v) for k, v in shards[0].items()}
    for shard in shards[1:]:
        for k, v in shard.items():
            merged.setdefault(k, []).extend(v)
    if remainder:
        merged = {k: v[:remainder] for k, v in merged.items()}
    return merged


def get_tensor_stats(xs: torch.Tensor, mask: torch.Tensor, n) -> Dict[str, torch.Tensor]:
    """Masked mean/min/max/std (std divides by ``n`` like the reference)."""
    if xs.numel() == 0:
        return dict(mean=0, min=0, max=0, std=0)
    m = mask.bool()
    mean = (xs * mask).sum() / n
    return dict(
        mean=mean,
        min=torch.where(m, xs, torch.full_like(xs, float("inf"))).min(),
        max=torch.where(m, xs, torch.full_like(xs, float("-inf"))).max(),
        std=torch.sqrt((((xs - mean) * mask) ** 2).sum() / n),
    )


class RunningMoments:
    """Streaming mean / std over reward batches (Welford-style merge; global across ranks)."""

    def __init__(self):
        self.mean = 0.0
        self.std = 1.0
        self.var = 1.0
        self.count = 1e-24

    def update(self, xs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Fold a batch in; returns the batch's own ``(mean, unbiased std)``."""
        b_mean, b_var, b_count = get_global_statistics(xs)
        b_mean, b_var, b_count = float(b_mean), float(b_var), float(b_count)
        delta = b_mean - self.mean
        total = self.count + b_count
        m2 = self.var * self.count + b_var * b_count + delta * delta * self.count * b_count / total
        self.mean = self.mean + delta * b_count / total
        self.var = m2 / total
        self.std = (self.var * total / (total - 1)) ** 0.5 if total > 1 else float("nan")
        self.count = total
        b_std = (b_var * b_count / (b_count - 1)) ** 0.5 if b_count > 1 else float("nan")
        return torch.tensor(b_mean), torch.tensor(b_std)

    def state_dict(self):
        return dict(mean=self.mean, std=self.std, var=self.var, count=self.count)

    def load_state_dict(self, sd):
        self.mean, self.std, self.var, self.count = sd["mean"], sd["std"], sd["var"], sd["count"]
