"""Autograd-aware wrappers around the sm_100a kernels (CUDA bf16) with PyTorch fallbacks elsewhere."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from trlx_b200.ops import reference

_PPO_KEYS = [
    "losses/total_loss", "losses/policy_loss", "losses/value_loss", "values/mean", "values/min", "values/max",
    "values/std", "values/values_error", "values/values_mape_error", "values/clipfrac", "old_values/mean",
    "old_values/min", "old_values/max", "old_values/std", "returns/mean", "returns/min", "returns/max", "returns/std",
    "policy/approx_kl", "policy/clipfrac", "ratio", "padding_percentage",
]


def _ops():
    from trlx_b200 import ops

    return ops


def _kernel_ok(x: torch.Tensor, w: torch.Tensor) -> bool:
    if not (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16):
        return False
    if not _ops().available():
        return False
    K = x.shape[-1]
    return K % 8 == 0 and w.stride(-1) == 1 and w.stride(0) % 8 == 0 and w.data_ptr() % 16 == 0


def _as_2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    return x2


def _ex_ok(t: torch.Tensor) -> bool:
    return (t.dim() == 2 and t.dtype == torch.bfloat16 and t.stride(1) == 1 and t.stride(0) % 8 == 0
            and t.data_ptr() % 16 == 0)


_OWN_BACKWARD = os.environ.get("TRLX_B200_BACKWARD_GEMM", "tcgen05") != "cublas"


def grad_input(g2: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dX[M, K] = dY[M, N] · W[N, K]: W is consumed as an MN-major B operand (no transposed copy)."""
    if _OWN_BACKWARD and _ex_ok(g2) and _ex_ok(w):
        return _ops().C.gemm_ex(g2, w, False, True)
    return g2 @ w


def grad_weight(g2: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """dW[N, K] = dYᵀ[N, M] · X[M, K]: both operands are consumed MN-major, straight from their forward layouts."""
    if _OWN_BACKWARD and _ex_ok(g2) and _ex_ok(x2):
        return _ops().C.gemm_ex(g2, x2, True, True)
    return g2.t() @ x2


def col_sum(g2: torch.Tensor) -> torch.Tensor:
    """Bias gradient ``g2.sum(0)`` of a bf16 ``[M, N]`` matrix on the 64-column-per-CTA kernel (``csrc/norm_train.cu``)."""
    C = _ops().C
    if (hasattr(C, "colsum") and g2.is_cuda and g2.dtype == torch.bfloat16 and g2.dim() == 2 and g2.stride(1) == 1
            and g2.shape[1] % 2 == 0 and g2.stride(0) % 2 == 0 and g2.data_ptr() % 4 == 0):
        return C.colsum(g2)
    return g2.sum(0)


class _Norm(torch.autograd.Function):
    """LayerNorm / RMSNorm over the last dimension with the row statistics kept for a one-pass backward."""

    @staticmethod
    def forward(ctx, x, w, b, eps, rms):
        C = _ops().C
        x2 = _as_2d(x)
        y, stats = C.ln_fwd(x2, w, None if rms else b, eps, rms)
        ctx.save_for_backward(x2, w, stats)
        ctx.rms, ctx.has_bias, ctx.x_shape = rms, (b is not None and not rms), x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        x2, w, stats = ctx.saved_tensors
        dx, dg, db = _ops().C.ln_bwd(x2, w, stats, _as_2d(gy), ctx.rms, ctx.has_bias)
        return dx.view(ctx.x_shape), dg, (db if ctx.has_bias else None), None, None


_OWN_NORM = os.environ.get("TRLX_B200_NORM", "own") != "torch"


def norm_ok(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> bool:
    """The training-side norm kernels apply: CUDA bf16, contiguous 16-byte aligned parameters, H % 8 == 0, H <= 4096."""
    if not (_OWN_NORM and x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and _ops().available()):
        return False
    C = _ops().C
    if not hasattr(C, "ln_fwd") or not C.ln_train_ok(x.shape[-1]):
        return False
    ok = w.is_contiguous() and w.data_ptr() % 16 == 0
    if b is not None:
        ok = ok and b.dtype == torch.bfloat16 and b.is_contiguous() and b.data_ptr() % 16 == 0
    return ok


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], eps: float, rms: bool = False) -> torch.Tensor:
    """``LayerNorm(x)`` (or ``RMSNorm`` with ``rms``) on the in-repo forward / backward kernels; callers check
    :func:`norm_ok` first."""
    return _Norm.apply(x, w, b, float(eps), bool(rms))


_WGRAD_ACCUM = os.environ.get("TRLX_B200_WGRAD_ACCUM", "1") == "1"


def _wgrad_sink(w: torch.Tensor, g2: torch.Tensor, x2: torch.Tensor):
    """The optimizer's gradient-ready callback when ``dW`` may be accumulated in place into ``w.grad`` (a view of the fused
    optimizer's flat bf16 buffer), else ``None``.  Parameters that several modules share (tied embeddings) keep the autograd
    path: their hook must fire once per backward, after every contribution."""
    if not (_WGRAD_ACCUM and _OWN_BACKWARD) or not getattr(w, "_b200_inplace_ok", False):
        return None
    sink = getattr(w, "_b200_grad_sink", None)
    g = w.grad
    if (sink is None or g is None or g.dtype != torch.bfloat16 or g.dim() != 2 or g.stride(-1) != 1 or g.stride(0) % 8
            or g.data_ptr() % 16 or not (_ex_ok(g2) and _ex_ok(x2)) or torch.is_grad_enabled()):
        return None
    return sink


def mark_inplace_wgrad(model: torch.nn.Module) -> int:
    """Flag the ``nn.Linear`` weights whose gradient may be accumulated in place by the wgrad GEMM: 2-D, trainable, owned by
    exactly one module (a tied LM head / embedding also receives gradient through other autograd paths and keeps the
    autograd accumulation).  Called by the trainers once the fused optimizer owns the gradients."""
    uses: Dict[int, int] = {}
    for _, p in model.named_parameters(remove_duplicate=False):
        uses[id(p)] = uses.get(id(p), 0) + 1
    n = 0
    for m in model.modules():
        if type(m) is torch.nn.Linear and m.weight.requires_grad and uses.get(id(m.weight), 0) == 1:
            m.weight._b200_inplace_ok = True
            n += 1
    return n


class _Linear(torch.autograd.Function):
    """y = x @ w.T + b (+ residual) — forward and both backward GEMMs on the tcgen05 kernel."""

    @staticmethod
    def forward(ctx, x, w, b, residual):
        C = _ops().C
        x2 = _as_2d(x)
        r2 = None if residual is None else _as_2d(residual)
        y = C.gemm(x2, w, b, r2, "none")
        if getattr(w, "_b200_inplace_ok", False):  # uses of this weight recorded for the coming backward (see backward)
            w._b200_uses = getattr(w, "_b200_uses", 0) + 1
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.has_res = residual is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        g2 = _as_2d(gy)
        gx = gw = gb = gr = None
        if ctx.needs_input_grad[0]:
            gx = grad_input(g2, w).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            sink = _wgrad_sink(w, g2, x2)
            if sink is not None:
                # dW is added into the parameter's slice of the flat gradient buffer by the GEMM epilogue itself: no dW tensor,
                # no separate accumulate kernel; the optimizer's "gradient ready" callback replaces the autograd hook
                _ops().C.gemm_ex(g2, x2, True, True, False, -1, w.grad, True)
                w._b200_uses = getattr(w, "_b200_uses", 1) - 1
                if w._b200_uses <= 0:  # every recorded use of the weight has contributed: the gradient is final
                    sink()
            else:
                gw = grad_weight(g2, x2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(g2)
        if ctx.has_res and ctx.needs_input_grad[3]:
            gr = gy
        return gx, gw, gb, gr


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, act: str = "none",
           residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(x @ w.T + b) + residual``.  Activation is fused into the GEMM epilogue when no gradient is needed;
    under autograd the pre-activation is kept and the activation runs as a separate op."""
    if not _kernel_ok(x, w):
        return reference.linear(x, w, b, act, residual)
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or (b is not None and b.requires_grad)
                                              or (residual is not None and residual.requires_grad))
    if not needs_grad:
        C = _ops().C
        y = C.gemm(_as_2d(x), w, b, None if residual is None else _as_2d(residual), act or "none")
        return y.view(*x.shape[:-1], w.shape[0])
    if act in ("none", "", None):
        return _Linear.apply(x, w, b, residual)
    y = reference.activation(_Linear.apply(x, w, b, None), act)
    return y if residual is None else y + residual


class _ShortAttention(torch.autograd.Function):
    """softmax(scale * q kᵀ + bias) v on the one-CTA-per-(batch, head) kernels of ``csrc/attention.cu`` (T <= 128)."""

    @staticmethod
    def forward(ctx, q, k, v, bias, causal, scale):
        C = _ops().C
        if _ATTENTION_TC and hasattr(C, "attn_tc_fwd") and C.attn_tc_ok(q.shape[2], k.shape[2], q.shape[3]):
            o, stats = C.attn_tc_fwd(q, k, v, bias, causal, scale)   # tcgen05: S and O accumulate in TMEM
        else:
            o, stats = C.attn_short_fwd(q, k, v, bias, causal, scale)
        ctx.save_for_backward(q, k, v, bias, o, stats)
        ctx.causal, ctx.scale = causal, scale
        return o.transpose(1, 2)  # [B, H, Tq, d] view of the [B, Tq, H, d] buffer

    @staticmethod
    def backward(ctx, g):
        q, k, v, bias, o, stats = ctx.saved_tensors
        if g.stride(-1) != 1 or any(st % 8 for st in g.stride()[:3]) or g.data_ptr() % 16:
            g = g.contiguous()
        dq, dk, dv = _ops().C.attn_short_bwd(q, k, v, bias, o, g, stats, ctx.causal, ctx.scale)
        return dq.transpose(1, 2), dk.transpose(1, 2), dv.transpose(1, 2), None, None, None


# "auto": the in-repo kernels serve no-grad calls with tiny score tiles (the rollout prefill), the library SDPA everything else —
# the CUDA-core backward is shared-memory-bandwidth bound and measured 2x slower than cuDNN's tensor-core kernel at
# 32 x 12 x 56 x 56 (run36: 29.6 vs 26.7 ms per 16 optimizer steps), and the forward loses from ~56 x 56 on (run45);
# "own" / "sdpa" force one side for A/B runs.
_ATTENTION_MODE = os.environ.get("TRLX_B200_ATTENTION", "auto")
# forward on the tcgen05 kernel (``attn_tc_fwd_kernel``) instead of the CUDA-core one; opt-in until it has been A/B-measured
_ATTENTION_TC = os.environ.get("TRLX_B200_ATTENTION_TC", "0") == "1"


def _attn_view_ok(t: torch.Tensor) -> bool:
    return t.stride(-1) == 1 and all(st % 8 == 0 for st in t.stride()[:3]) and t.data_ptr() % 16 == 0


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, bias: Optional[torch.Tensor] = None, causal: bool = False,
              scale: Optional[float] = None) -> torch.Tensor:
    """``softmax(scale · q kᵀ + bias) v`` for ``[B, H, T, d]`` operands; ``bias`` is an additive fp32 mask broadcastable to
    ``[B, H, Tq, Tk]`` (or ``None`` with ``causal``).  Sequences of at most 128 tokens run on the in-repo kernels (forward
    and backward, any operand strides with a contiguous head dimension); everything else — CPU tensors, longer sequences —
    goes through ``F.scaled_dot_product_attention``."""
    scale = float(scale) if scale is not None else q.shape[-1] ** -0.5
    need_grad = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)
    # measured (profiles/attention_fwd_bench.jsonl): the one-CTA-per-head kernels win only on tiny score tiles (prefill, 24 x 24:
    # 37 us vs 47 us for cuDNN); from ~56 x 56 on the library's tensor-core kernel is 2-3x faster
    own = _ATTENTION_MODE == "own" or (_ATTENTION_MODE == "auto" and not need_grad and q.shape[2] * k.shape[2] <= 32 * 32)
    if (own and q.is_cuda and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
            and q.dim() == 4 and _ops().available() and hasattr(_ops().C, "attn_short_fwd")
            and _ops().C.attn_short_ok(q.shape[2], k.shape[2], q.shape[3], need_grad)
            and (bias is None or (bias.dim() == 4 and bias.shape[-1] == k.shape[2]))):
        q, k, v = (t if _attn_view_ok(t) else t.contiguous() for t in (q, k, v))
        if bias is not None:
            bias = bias.float()
            if bias.stride(-1) != 1 and bias.shape[-1] != 1:
                bias = bias.contiguous()
        return _ShortAttention.apply(q, k, v, bias, bool(causal) and bias is None, scale)
    import torch.nn.functional as F

    if bias is None:
        return F.scaled_dot_product_attention(q, k, v, is_causal=bool(causal) and q.shape[2] > 1, scale=scale)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype), scale=scale)


class _FusedLogprob(torch.autograd.Function):
    """log p(label | h) through the LM head without materialising logits in the forward.
    Backward: one tcgen05 GEMM recomputes the logits tile by tile and emits d-logits straight from its epilogue
    (``(onehot − softmax) · g``, bf16, row pitch padded to 64 so the two library GEMMs that consume it are aligned)."""

    @staticmethod
    def forward(ctx, h, w, b, labels):
        C = _ops().C
        h2 = _as_2d(h)
        lab = labels.reshape(-1).contiguous()
        lse, lp, _, _ = C.lmhead(h2, w, b, lab)
        ctx.save_for_backward(h2, w, b if b is not None else h2.new_empty(0), lab, lse)
        ctx.has_bias = b is not None
        ctx.h_shape = h.shape
        ctx.mark_non_differentiable(lse)
        return lp.view(labels.shape), lse.view(labels.shape)

    @staticmethod
    def backward(ctx, g_lp, _g_lse):
        C = _ops().C
        h2, w, b, lab, lse = ctx.saved_tensors
        logits = C.lmhead_dlogits(h2, w, b if ctx.has_bias else None, lab, lse, g_lp.reshape(-1).float().contiguous())
        gh = gw = gb = None
        if ctx.needs_input_grad[0]:
            gh = grad_input(logits, w).view(ctx.h_shape)
        if ctx.needs_input_grad[1]:
            gw = grad_weight(logits, h2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(logits)
        return gh, gw, gb, None


def fused_logprob(h, w, b, labels) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(log p(labels), logsumexp)`` of the LM head applied to ``h``; labels < 0 are ignored (0 output)."""
    if not _kernel_ok(h, w):
        return reference.fused_logprob(h, w, b, labels)
    return _FusedLogprob.apply(h, w, b, labels)


class _LogprobsFromLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        lp, lse = _ops().C.logprob_from_logits(logits, labels)
        ctx.save_for_backward(logits, labels, lse)
        return lp

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse = ctx.saved_tensors
        p = torch.exp(logits.float() - lse.unsqueeze(-1))
        onehot = torch.zeros_like(p).scatter_(-1, labels.clamp_min(0).unsqueeze(-1), 1.0)
        return ((onehot - p) * g.unsqueeze(-1)).to(logits.dtype), None


def logprobs_from_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    return _LogprobsFromLogits.apply(logits, labels)


def gae_and_whiten(values, rewards, width: int, gamma: float, lam: float, use_whitening: bool = True, group=None,
                   width_tensor: Optional[torch.Tensor] = None):
    """Advantages (optionally whitened) and returns over the first ``width`` response positions.

    Whitening matches the reference: unbiased variance in a single process, biased global variance across ranks
    (``trlx/utils/modeling.py:200-210``).  On CUDA: one scan kernel + (one small all-reduce) + one whiten kernel."""
    distributed = dist.is_available() and dist.is_initialized()
    if distributed and group is None:
        from trlx_b200.utils.modeling import statistics_group

        group = statistics_group()
    ops = _ops()
    if values.is_cuda and ops.available():
        v = values.float().contiguous()
        r = rewards.float().contiguous()
        # ``width_tensor`` (int32, device) overrides ``width`` inside the kernels: a captured CUDA graph then serves
        # batches whose longest response differs (the tensors keep their static full width)
        if not use_whitening:
            adv, ret, _ = ops.C.gae(v, r, width, gamma, lam, False, True, width_tensor)
        elif not distributed or dist.get_world_size(group) == 1:
            adv, ret, _ = ops.C.gae(v, r, width, gamma, lam, True, not distributed, width_tensor)
        else:
            adv, ret, stats = ops.C.gae(v, r, width, gamma, lam, False, False, width_tensor)
            dist.all_reduce(stats, group=group)
            ops.C.whiten_(adv, width, stats, False, width_tensor)
        return adv[:, :width].detach(), ret[:, :width]
    from trlx_b200.utils.modeling import whiten

    adv, ret = reference.gae(values, rewards, width, gamma, lam)
    if use_whitening:
        adv = whiten(adv, group=group)
    return adv.detach(), ret


class _PPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logprobs, values, old_logprobs, old_values, adv, ret, mask, clip, clip_v, vf_coef, width_tensor):
        C = _ops().C
        args = [t.float().contiguous() for t in (logprobs, values, old_logprobs, old_values, adv, ret, mask)]
        out, dlp, dv = C.ppo_loss(*args, clip, clip_v, vf_coef, width_tensor)
        ctx.save_for_backward(dlp, dv)
        ctx.shapes = (logprobs.shape, values.shape, logprobs.dtype, values.dtype)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        dlp, dv = ctx.saved_tensors
        s_lp, s_v, dt_lp, dt_v = ctx.shapes
        return ((dlp * g_loss).view(s_lp).to(dt_lp), (dv * g_loss).view(s_v).to(dt_v),
                None, None, None, None, None, None, None, None, None)


def ppo_loss(logprobs, values, old_logprobs, old_values, advantages, returns, mask, cliprange: float,
             cliprange_value: float, vf_coef: float, width_tensor: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Fused PPO loss: returns ``(loss, stats)`` where ``stats`` maps the reference's flattened stat keys to 0-dim
    DEVICE tensors (no host sync here; the trainer converts once per optimizer step)."""
    ops = _ops()
    if logprobs.is_cuda and ops.available():
        loss, out = _PPOLoss.apply(logprobs, values, old_logprobs, old_values, advantages, returns, mask,
                                   float(cliprange), float(cliprange_value), float(vf_coef), width_tensor)
        return loss, {k: out[i] for i, k in enumerate(_PPO_KEYS)}
    return reference.ppo_loss(logprobs, values, old_logprobs, old_values, advantages, returns, mask, cliprange,
                              cliprange_value, vf_coef)
