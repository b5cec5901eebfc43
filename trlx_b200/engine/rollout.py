"""CUDA rollout engine: batched sampling for PPO on the sm_100a kernels, one CUDA graph per decode step.

What the reference does per chunk of prompts: HF ``generate`` (python loop, dozens of launches per
token) → a second full forward for log-probs and values → ``forward_hydra`` which runs the trunk a third time for the
reference log-probs → ``[B,T,V]`` fp32 logits twice → ``.cpu()``.  Here:

* prompts are prefilled once (tcgen05 GEMMs) and their K/V scattered into a **paged KV cache**
  (``csrc/bindings.cpp::PagedKVAllocator`` hands out pages; every layer owns a ``[pages, page, kv_heads, d]`` tensor);
* each decode step is ONE captured CUDA graph: embedding → per layer {norm, QKV GEMM, paged attention with fused
  rotary + cache append, out-proj GEMM (+residual), norm, MLP GEMMs (+bias+act, +residual)} → the fused LM-head
  kernel samples by Gumbel-max inside the GEMM epilogue and returns the token and its log-prob without ever
  writing logits → value head → the frozen *reference branch* runs from the shared trunk activation and scores
  the same token → a bookkeeping kernel records (token, log-prob, ref log-prob, value), retires finished rows and
  prepares the next step — all on the device;
* the trunk activation at the branch point is kept per position so PPO updates never recompute the frozen trunk.

The result has exactly the tensors ``AcceleratePPOTrainer.make_experience`` needs (``logprobs``, ``ref_logprobs``,
``values`` over ``T-1`` positions, ``mask``, ``start``), so the reference's re-scoring passes are simply not run.
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200 import ops
from trlx_b200.models.modeling_base import base_lm
from trlx_b200.nn.transformer import alibi_slopes, build_attn_context
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)

PAGE = 16


@dataclass
class _LayerW:
    n1w: torch.Tensor
    n1b: Optional[torch.Tensor]
    n2w: Optional[torch.Tensor]
    n2b: Optional[torch.Tensor]
    qkv_w: torch.Tensor
    qkv_b: Optional[torch.Tensor]
    out_w: torch.Tensor
    out_b: Optional[torch.Tensor]
    up_w: torch.Tensor
    up_b: Optional[torch.Tensor]
    down_w: torch.Tensor
    down_b: Optional[torch.Tensor]
    window: int


def _w(linear):
    """Weight a rollout kernel should read: LoRA-wrapped projections contribute ``W + (alpha / r) B A`` (a merged copy that
    the engine refreshes after optimizer steps), plain ones their own tensor."""
    return linear.merged_weight() if hasattr(linear, "merged_weight") and len(getattr(linear, "lora_A", ())) else linear.weight


def _layer_weights(block, spec, idx: int) -> _LayerW:
    n2 = block.norm2
    return _LayerW(block.norm1.weight, block.norm1.bias, n2.weight if n2 is not None else None,
                   n2.bias if n2 is not None else None, _w(block.attn.qkv), block.attn.qkv.bias, _w(block.attn.out),
                   block.attn.out.bias, _w(block.mlp.up), block.mlp.up.bias, _w(block.mlp.down), block.mlp.down.bias,
                   spec.local_window if idx in spec.local_layers else 0)


def _lora_sources(block):
    """``[(field of _LayerW, LoRA-wrapped linear)]`` for the projections of ``block`` that carry adapters."""
    pairs = (("qkv_w", block.attn.qkv), ("out_w", block.attn.out), ("up_w", block.mlp.up), ("down_w", block.mlp.down))
    return [(f, lin) for f, lin in pairs if hasattr(lin, "merged_weight") and len(getattr(lin, "lora_A", ()))]


@dataclass
class _Fp8W:
    """e4m3 copies of the two weight matrices that follow a norm (QKV and MLP-up), one dequantisation scale per output
    channel; the matching activations are quantised per row inside the norm kernel (``rollout_dtype = "fp8"``)."""

    qkv_w: torch.Tensor
    qkv_s: torch.Tensor
    up_w: torch.Tensor
    up_s: torch.Tensor
    trainable: bool
    # all four GEMMs of a block in e4m3 (``fp8_all``): out-projection and MLP-down copies; their activations (attention output,
    # MLP activation) are quantised per row by ``quant_rows_kernel``
    out_w: Optional[torch.Tensor] = None
    out_s: Optional[torch.Tensor] = None
    down_w: Optional[torch.Tensor] = None
    down_s: Optional[torch.Tensor] = None


def _quant_e4m3(w: torch.Tensor):
    wf = w.float()
    scale = (wf.abs().amax(1).clamp_min(1e-12) / 448.0).contiguous()
    q = (wf / scale[:, None]).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8).contiguous()
    return q, scale


@dataclass
class _FoldW:
    """Per-layer derived weights for norm-folded GEMMs: ``LN(x)·Wᵀ = rstd·(x·(γ⊙W)ᵀ) − rstd·μ·c1 + (W·β + b)``."""

    qkv_w: torch.Tensor
    qkv_b: torch.Tensor
    qkv_c1: torch.Tensor
    up_w: torch.Tensor
    up_b: torch.Tensor
    up_c1: torch.Tensor
    trainable: bool


def _fold(w: torch.Tensor, b: Optional[torch.Tensor], gamma: torch.Tensor, beta: Optional[torch.Tensor]):
    wf = (w.float() * gamma.float()).to(torch.bfloat16)
    c1 = wf.float().sum(1).contiguous()
    bias = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
    if beta is not None:
        bias = bias + w.float() @ beta.float()
    if b is not None:
        bias = bias + b.float()
    return wf.contiguous(), bias.to(torch.bfloat16).contiguous(), c1


class RolloutEngine:
    @staticmethod
    def why_not(model, gen_kwargs: Dict[str, Any], config=None, stop_sequences=None) -> Optional[str]:
        """``None`` when the engine can serve this model / sampling configuration, else the reason it cannot."""
        if not ops.available():
            return "the sm_100a extension is not available"
        peft = getattr(model, "peft_type", None)
        if peft and str(peft).upper().split(".")[-1] != "LORA":
            return f"{peft} adapters are scored through the PyTorch path (only LoRA is merged into the rollout weights)"
        if not peft and getattr(model, "frozen_head", None) is None:
            return "no frozen reference branch (num_layers_unfrozen <= 0 or a separate reference model)"
        if getattr(model, "num_value_layers_unfrozen", 0) != 0 and (peft or not model.can_share_trunk()):
            return "a value branch deeper than the policy branch (no shared trunk activation to score it from)"
        if getattr(model, "tp_context", None) is not None or getattr(base_lm(model.base_model), "tp_context", None) is not None:
            return "tensor-parallel weights are sharded: the engine's kernels assume full (replicated) weight matrices"
        from trlx_b200.parallel import state as pstate

        if pstate.get_tensor_model_parallel_world_size() > 1 or pstate.get_pipeline_model_parallel_world_size() > 1:
            return "model-parallel run (tensor / pipeline parallel groups are active)"
        lm = base_lm(model.base_model)
        spec = lm.config
        if not hasattr(lm, "transformer") or lm.dtype != torch.bfloat16 or not lm.device.type == "cuda":
            return "needs a decoder-only bf16 model on CUDA"
        if spec.head_dim % 8 or spec.hidden_size % 8 or spec.head_dim > 256 or spec.ffn_size % 8:
            return "head_dim / hidden / ffn sizes must be multiples of 8 (head_dim <= 256)"
        if getattr(spec, "post_norm", False) or not getattr(spec, "plain_tail", True):
            return "post-LN blocks / embedding projections (OPT-350m layout) run on the PyTorch path"
        g = gen_kwargs or {}
        if g.get("num_beams", 1) not in (1, None):
            return "beam search"
        if g.get("repetition_penalty") not in (None, 1.0):
            return "repetition_penalty"
        for k in ("typical_p", "penalty_alpha", "no_repeat_ngram_size", "bad_words_ids", "num_return_sequences"):
            if g.get(k) not in (None, 0, 1, 1.0):
                return f"generation option {k}"
        if any(isinstance(v, list) for k, v in g.items() if not str(k).startswith("_")):
            return "a generation option is being swept (list value)"
        return None

    @staticmethod
    def supports(model, gen_kwargs: Dict[str, Any], config=None, stop_sequences=None) -> bool:
        """Decoder-only hydra model with a frozen reference branch, bf16 on CUDA; temperature / top-k / top-p sampling or
        greedy.  With ``stop_sequences`` the engine still generates, but the trainer re-tokenises the trimmed text and
        re-scores it exactly like the reference does (token ids may change at the cut)."""
        return RolloutEngine.why_not(model, gen_kwargs, config, stop_sequences) is None

    def __init__(self, model, pad_token_id: int, eos_token_id: int, gen_kwargs: Dict[str, Any], cache_trunk: bool = True,
                 seed: int = 0, use_cuda_graph: bool = True):
        self.model = model
        self.lm = base_lm(model.base_model)
        self.spec = self.lm.config
        self.pad, self.eos = int(pad_token_id), int(eos_token_id if eos_token_id is not None else -1)
        self.gen = dict(gen_kwargs)
        self.cache_trunk = cache_trunk and not getattr(model, "peft_type", None)  # adapters leave no frozen trunk to share
        self.seed = int(seed)
        self.calls = 0
        self.use_cuda_graph = use_cuda_graph
        self.device = self.lm.device
        self.ilql = bool(getattr(self, "ilql", False))  # set by engine/ilql.py before calling up
        self.branch = model.branch_layer if (hasattr(model, "branch_layer") and not getattr(model, "peft_type", None)) else 0
        spec = self.spec
        # LoRA (the reference toggles adapters per forward, ``trlx/models/modeling_ppo.py:318-324``): rollouts read merged
        # weights W + (alpha / r) B A, refreshed in place after optimizer steps — zero adapter overhead per decoded token and the
        # fp8 / megakernel paths apply unchanged; reference log-probs come from ONE batched adapter-free pass after the loop
        self.lora = bool(getattr(model, "peft_type", None))
        # value *branch* (``num_value_layers_unfrozen > 0``, reference ``modeling_ppo.py:331-343``): sampling never needs the
        # value, so the decode graph skips it and ONE batched pass over the cached trunk activations scores every position
        self.value_branch = int(getattr(model, "num_value_layers_unfrozen", 0) or 0) > 0
        self.layers = [_layer_weights(b, spec, i) for i, b in enumerate(self.lm.transformer.h)]
        self._lora_src = [_lora_sources(b) for b in self.lm.transformer.h] if self.lora else []
        fh = getattr(model, "frozen_head", None) if not self.lora else None
        self.ref_layers = [_layer_weights(b, spec, self.branch + j) for j, b in enumerate(fh.decoder_blocks)] if fh is not None else []
        self.scale = spec.attn_scale if spec.attn_scale is not None else 1.0 / math.sqrt(spec.head_dim)
        self.alibi = alibi_slopes(spec.num_heads).to(self.device) if spec.pos == "alibi" else None
        self.rot_dim = spec.rotary_dim if spec.pos == "rotary" else 0
        self._state = None  # (B, P) → static buffers + graph
        do_sample = bool(self.gen.get("do_sample", False))
        t = self.gen.get("temperature", 1.0)
        self.temperature = float(t if t is not None else 1.0) if do_sample else 0.0
        # top-k / top-p: the LM head then writes fp32 logits once and a radix-select sampling kernel applies HF's
        # temperature -> top-k -> top-p -> multinomial chain (csrc/decode_ops.cu: sample_filtered_kernel)
        self.top_k = int(self.gen.get("top_k") or 0) if do_sample else 0
        tp_ = self.gen.get("top_p")
        self.top_p = float(tp_ if tp_ is not None else 1.0) if do_sample else 1.0
        self.filtered = do_sample and ((0 < self.top_k < spec.vocab_size) or self.top_p < 1.0)
        self.launches_per_step = 0
        import os

        # LayerNorm folding (decode, opt-in: TRLX_B200_FOLD_NORMS=1 — measured neutral once PDL overlaps the norm kernels,
        # run18, and the plain path keeps rollout numerics closest to the training forward): the two norm kernels of every
        # block disappear — their effect is applied in the
        # epilogue of the GEMM that consumes them, from row moments accumulated by the GEMM (or embed) that produced x.
        folded_bytes = sum((W.qkv_w.numel() + W.up_w.numel()) * 2 for W in self.layers + self.ref_layers)
        self.fold_norms = (os.environ.get("TRLX_B200_FOLD_NORMS", "0") == "1" and self.lm.transformer.emb_norm is None
                           and spec.hidden_size % 16 == 0 and all(W.qkv_w.shape[0] % 16 == 0 and W.up_w.shape[0] % 16 == 0
                                                                  for W in self.layers + self.ref_layers)
                           and folded_bytes <= int(os.environ.get("TRLX_B200_FOLD_BUDGET_MB", "4096")) << 20)
        # fp8 rollout (config.train.parallel.rollout_dtype == "fp8" or TRLX_B200_ROLLOUT_FP8=1): the norm → GEMM pairs of every
        # block (QKV, MLP up) — and with `fp8_all` the out- and down-projections too — run e4m3 x e4m3 on the tensor cores with
        # per-row / per-channel scales; half the weight bytes
        self.fp8 = (str(gen_kwargs.get("_rollout_dtype", os.environ.get("TRLX_B200_ROLLOUT_FP8", ""))).lower() in ("fp8", "1", "true")
                    and spec.hidden_size % 16 == 0 and not self.fold_norms)
        # all four GEMMs of a block in e4m3 (+ one row-quantisation kernel in front of the out- and down-projections): pays when the
        # weights dominate a decode step (hidden >= 2048); small models are launch-latency bound and keep those two in bf16.
        # TRLX_B200_FP8_ALL=1 / 0 forces it.
        mode = os.environ.get("TRLX_B200_FP8_ALL", "auto")
        self.fp8_all = self.fp8 and spec.ffn_size % 16 == 0 and (mode == "1" or (mode == "auto" and spec.hidden_size >= 2048))
        self.fp8_w: List[_Fp8W] = []
        self.ref_fp8_w: List[_Fp8W] = []
        self.folded: List[_FoldW] = []
        self.ref_folded: List[_FoldW] = []
        self.dirty = True  # folded copies must be (re)built before the next rollout
        # Deferred reference scoring: the frozen branch only produces log-probs that are consumed AFTER generation, so it does
        # not have to sit on the per-token critical path.  The decode graph then keeps just the trunk activation of every
        # position, and one batched pass (2 frozen blocks + LM head over [B, Q+R] tokens at GEMM-efficient M) scores all
        # positions at the end — instead of 2 latency-bound blocks + a 77 MB LM-head sweep per decoded token.
        self.defer_ref = os.environ.get("TRLX_B200_DEFER_REF", "1") == "1" or self.lora or self.ilql
        # per-position activation at the branch point (no branch with adapters on every layer / in ILQL generation)
        self.keep_trunk = (self.cache_trunk or self.defer_ref) and not self.lora and not self.ilql
        self.parallel_branches = (os.environ.get("TRLX_B200_PARALLEL_BRANCHES", "1") == "1" and self.branch < len(self.layers)
                                  and not self.defer_ref)
        self.side = torch.cuda.Stream(device=self.device) if self.parallel_branches else None
        ops.C.set_pdl(os.environ.get("TRLX_B200_PDL", "1") == "1")
        # Persistent decode megakernel (csrc/decode_mega.cu): the whole policy layer stack of a decode step in ONE launch —
        # clusters of up to 16 CTAs own 16 batch rows each and walk all blocks with cluster barriers between phases.  Opt-in
        # (TRLX_B200_DECODE_MEGA=1): numerically validated, but measured at 1.23 ms / token against 0.64 ms for the
        # kernel-per-op graph on GPT-2 124M, batch 128 — with 16 rows per cluster the 64 x 16 x 16 tcgen05 atoms are bound by
        # the ~68-cycle issue interval of one UTCHMMA (scripts/umma_probe.py), ~1000 of them per layer and CTA
        # (profiles/decode_megakernel.md has the phase-by-phase timeline and what would close the gap).
        self.mega = self._mega_eligible() and os.environ.get("TRLX_B200_DECODE_MEGA", "0") == "1"
        self._mega_tables = None

    # ------------------------------------------------------------------------------------------------ megakernel
    def _mega_eligible(self) -> bool:
        spec = self.spec
        if not hasattr(ops.C, "decode_mega") or self.fp8 or self.fold_norms or not self.defer_ref:
            return False
        if spec.head_dim != 64 or spec.num_kv_heads != spec.num_heads or spec.hidden_size != spec.num_heads * 64:
            return False
        if spec.hidden_size % 64 or spec.ffn_size % 64 or spec.parallel_residual or spec.gated_mlp:
            return False
        if spec.pos not in ("learned", "none", "alibi") or spec.local_layers or self.lm.transformer.emb_norm is not None:
            return False
        if any(W.n2w is None for W in self.layers):
            return False
        try:
            return int(ops.C.decode_mega_stages(spec.hidden_size, spec.ffn_size)) >= 4
        except Exception:
            return False

    def _mega_build(self, st):
        """Pointer table + weight tensor maps of the policy stack for this state's KV caches (addresses are stable: the
        parameters are views of the optimizer's flat buffer, the caches live as long as the state)."""
        rows = []
        for i, W in enumerate(self.layers):
            rows.append([W.n1w, W.n1b, W.n2w, W.n2b, W.qkv_w, W.qkv_b, W.out_w, W.out_b, W.up_w, W.up_b, W.down_w, W.down_b,
                         st["kc"][i], st["vc"][i]])
        table, maps = ops.C.decode_mega_build(rows)
        B, spec = st["B"], self.spec
        import os

        timing = (torch.zeros(16 * len(self.layers) * 16, dtype=torch.int64, device=self.device)
                  if os.environ.get("TRLX_B200_MEGA_TIMING") == "1" else None)  # clock64 stamps of cluster 0 (profiling aid)
        st["mega"] = dict(table=table, maps=maps, timing=timing,
                          a=torch.empty(B, spec.hidden_size, dtype=torch.bfloat16, device=self.device),
                          mid=torch.empty(B, spec.ffn_size, dtype=torch.bfloat16, device=self.device))

    def _mega_stack(self, x, st):
        """All policy blocks on ``x`` (in place) + the trunk capture; returns ``x``."""
        spec = self.spec
        if "mega" not in st:
            self._mega_build(st)
        m = st["mega"]
        want_trunk = self.keep_trunk and st["trunk_decode"] is not None
        ops.C.decode_mega(x, m["a"], m["mid"], st["block_table"], st["seq_lens"], m["table"], m["maps"], spec.num_heads,
                          len(self.layers), spec.activation, spec.norm == "rmsnorm", spec.norm_eps, self.scale, PAGE,
                          st["trunk_decode"] if want_trunk else None, st["step64"] if want_trunk else None,
                          self.branch if want_trunk else -1, self.alibi, m["timing"])
        return x

    # ------------------------------------------------------------------------------------------------ folded weights
    def mark_dirty(self):
        """The trainer calls this after every optimizer step / checkpoint load (weights are updated through raw pointers
        by the fused optimizer, so tensor version counters cannot be relied on)."""
        self.dirty = True

    @torch.no_grad()
    def _refresh_fp8(self):
        first = not self.fp8_w

        def build(W: _LayerW, old: Optional[_Fp8W]) -> _Fp8W:
            trainable = any(t.requires_grad for t in (W.qkv_w, W.up_w, W.out_w, W.down_w)) or self.lora  # merged LoRA copies change every step
            if old is not None and not trainable:
                return old
            q, qs = _quant_e4m3(W.qkv_w)
            u, us = _quant_e4m3(W.up_w)
            extra = (*_quant_e4m3(W.out_w), *_quant_e4m3(W.down_w)) if self.fp8_all else (None,) * 4
            if old is None:
                return _Fp8W(q, qs, u, us, trainable, *extra)
            for dst, src in zip((old.qkv_w, old.qkv_s, old.up_w, old.up_s, old.out_w, old.out_s, old.down_w, old.down_s),
                                (q, qs, u, us, *extra)):
                if dst is not None:
                    dst.copy_(src)
            return old

        self.fp8_w = [build(W, None if first else self.fp8_w[i]) for i, W in enumerate(self.layers)]
        self.ref_fp8_w = [build(W, None if first else self.ref_fp8_w[i]) for i, W in enumerate(self.ref_layers)]

    @torch.no_grad()
    def _refresh_lora(self):
        """Re-merge ``W + (alpha / r) B A`` into the engine's weight copies IN PLACE (graphs / tensor maps keep the addresses)."""
        for W, srcs in zip(self.layers, self._lora_src):
            for field, lin in srcs:
                getattr(W, field).copy_(lin.merged_weight())

    @torch.no_grad()
    def refresh_folded(self):
        """(Re)build γ-scaled copies of the QKV / MLP-up weights in place (the captured CUDA graph keeps their addresses)."""
        if self.lora and self.dirty:
            self._refresh_lora()
            if not self.fp8 and not self.fold_norms:
                self.dirty = False
        if self.fp8 and self.dirty:
            self._refresh_fp8()
            self.dirty = False
            return
        if not self.fold_norms or not self.dirty:
            return
        first = not self.folded

        def build(W: _LayerW, old: Optional[_FoldW]) -> _FoldW:
            n2w, n2b = (W.n2w, W.n2b) if W.n2w is not None else (W.n1w, W.n1b)
            trainable = any(t is not None and t.requires_grad for t in (W.n1w, W.n1b, W.n2w, W.n2b, W.qkv_w, W.qkv_b, W.up_w, W.up_b))
            if old is not None and not trainable:
                return old
            q = _fold(W.qkv_w, W.qkv_b, W.n1w, W.n1b)
            u = _fold(W.up_w, W.up_b, n2w, n2b)
            if old is None:
                return _FoldW(*q, *u, trainable)
            for dst, src in zip((old.qkv_w, old.qkv_b, old.qkv_c1, old.up_w, old.up_b, old.up_c1), q + u):
                dst.copy_(src)
            return old

        self.folded = [build(W, None if first else self.folded[i]) for i, W in enumerate(self.layers)]
        self.ref_folded = [build(W, None if first else self.ref_folded[i]) for i, W in enumerate(self.ref_layers)]
        self.dirty = False

    def _layer_folded(self, x, xs, W: _LayerW, FW: _FoldW, kc, vc, st):
        """One block with both norms folded.  ``xs`` = row moments ``[B, 2]`` of ``x``; returns ``(x_out, moments of x_out)``."""
        C, spec = ops.C, self.spec
        rms, eps = spec.norm == "rmsnorm", spec.norm_eps
        qkv = C.gemm_ln(x, FW.qkv_w, FW.qkv_b, None, "none", xs, FW.qkv_c1, eps, rms, None)
        a = C.decode_attention(qkv, kc, vc, st["block_table"], st["seq_lens"], st["positions"], spec.num_heads,
                               spec.num_kv_heads, spec.head_dim, self.scale, self.rot_dim, spec.rotary_base,
                               spec.rotary_interleaved, self.alibi, W.window)
        out_stats = self._next_stats(st)
        if spec.parallel_residual:
            t = C.gemm(a, W.out_w, W.out_b, x)
            mid = self._mlp_mid_folded(x, xs, FW)
            return C.gemm_ln(mid, W.down_w, W.down_b, t, "none", None, None, eps, rms, out_stats), out_stats
        x = C.gemm_ln(a, W.out_w, W.out_b, x, "none", None, None, eps, rms, out_stats)
        mid = self._mlp_mid_folded(x, out_stats, FW)
        out2 = self._next_stats(st)
        return C.gemm_ln(mid, W.down_w, W.down_b, x, "none", None, None, eps, rms, out2), out2

    def _mlp_mid_folded(self, x, xs, FW: _FoldW):
        C, spec = ops.C, self.spec
        rms, eps = spec.norm == "rmsnorm", spec.norm_eps
        if spec.gated_mlp:
            g, u = C.gemm_ln(x, FW.up_w, FW.up_b, None, "none", xs, FW.up_c1, eps, rms, None).chunk(2, dim=-1)
            return (F.silu(g) * u).contiguous() if spec.activation in ("silu", "swish") else (F.gelu(g, approximate="tanh") * u).contiguous()
        return C.gemm_ln(x, FW.up_w, FW.up_b, None, spec.activation, xs, FW.up_c1, eps, rms, None)

    def _next_stats(self, st):
        i = st["stats_cursor"]
        st["stats_cursor"] = i + 1
        return st["ln_stats"][i]

    # ------------------------------------------------------------------------------------------------ kernels per layer
    def _layer_fp8(self, x, W: _LayerW, Q: _Fp8W, kc, vc, st):
        """Block with the two norm → GEMM pairs in e4m3 (attention, out-proj and MLP-down stay bf16)."""
        C, spec = ops.C, self.spec
        rms, eps = spec.norm == "rmsnorm", spec.norm_eps
        h8, hs = C.norm_quant(x, W.n1w, W.n1b, eps, rms)
        qkv = C.gemm_fp8(h8, Q.qkv_w, hs, Q.qkv_s, W.qkv_b, None, "none")
        a = C.decode_attention(qkv, kc, vc, st["block_table"], st["seq_lens"], st["positions"], spec.num_heads,
                               spec.num_kv_heads, spec.head_dim, self.scale, self.rot_dim, spec.rotary_base,
                               spec.rotary_interleaved, self.alibi, W.window)

        def mlp_mid(h8_, hs_):
            if spec.gated_mlp:
                g, u = C.gemm_fp8(h8_, Q.up_w, hs_, Q.up_s, W.up_b, None, "none").chunk(2, dim=-1)
                return (F.silu(g) * u).contiguous() if spec.activation in ("silu", "swish") else (F.gelu(g, approximate="tanh") * u).contiguous()
            return C.gemm_fp8(h8_, Q.up_w, hs_, Q.up_s, W.up_b, None, spec.activation)

        def out_proj(a_, res):
            if Q.out_w is None:
                return C.gemm(a_, W.out_w, W.out_b, res)
            a8, a_s = C.quant_rows(a_)
            return C.gemm_fp8(a8, Q.out_w, a_s, Q.out_s, W.out_b, res, "none")

        def down_proj(mid, res):
            if Q.down_w is None:
                return C.gemm(mid, W.down_w, W.down_b, res)
            m8, m_s = C.quant_rows(mid)
            return C.gemm_fp8(m8, Q.down_w, m_s, Q.down_s, W.down_b, res, "none")

        if spec.parallel_residual:
            h2 = (h8, hs) if W.n2w is None else tuple(C.norm_quant(x, W.n2w, W.n2b, eps, rms))
            t = out_proj(a, x)
            return down_proj(mlp_mid(*h2), t)
        x = out_proj(a, x)
        h2 = C.norm_quant(x, W.n2w, W.n2b, eps, rms)
        return down_proj(mlp_mid(*h2), x)

    def _layer(self, x, W: _LayerW, kc, vc, st):
        C, spec = ops.C, self.spec
        if self.fp8:
            idx = next((i for i, L_ in enumerate(self.layers) if L_ is W), None)
            Q = self.fp8_w[idx] if idx is not None else self.ref_fp8_w[next(i for i, L_ in enumerate(self.ref_layers) if L_ is W)]
            return self._layer_fp8(x, W, Q, kc, vc, st)
        rms = spec.norm == "rmsnorm"
        h = C.norm(x, W.n1w, W.n1b, spec.norm_eps, rms)
        qkv = C.gemm(h, W.qkv_w, W.qkv_b)
        a = C.decode_attention(qkv, kc, vc, st["block_table"], st["seq_lens"], st["positions"], spec.num_heads,
                               spec.num_kv_heads, spec.head_dim, self.scale, self.rot_dim, spec.rotary_base,
                               spec.rotary_interleaved, self.alibi, W.window)
        if spec.parallel_residual:
            h2 = h if W.n2w is None else C.norm(x, W.n2w, W.n2b, spec.norm_eps, rms)
            t = C.gemm(a, W.out_w, W.out_b, x)
            return C.gemm(self._mlp_mid(h2, W), W.down_w, W.down_b, t)
        x = C.gemm(a, W.out_w, W.out_b, x)
        h2 = C.norm(x, W.n2w, W.n2b, spec.norm_eps, rms)
        return C.gemm(self._mlp_mid(h2, W), W.down_w, W.down_b, x)

    def _mlp_mid(self, h, W: _LayerW):
        C, spec = ops.C, self.spec
        if spec.gated_mlp:
            g, u = C.gemm(h, W.up_w, W.up_b).chunk(2, dim=-1)
            return (F.silu(g) * u).contiguous() if spec.activation in ("silu", "swish") else (F.gelu(g, approximate="tanh") * u).contiguous()
        return C.gemm(h, W.up_w, W.up_b, None, spec.activation)

    def _decode_step(self, st):
        """One token for every running row; pure device work (captured into a CUDA graph)."""
        C, spec, lm, model = ops.C, self.spec, self.lm, self.model
        tr = lm.transformer
        fold = self.fold_norms
        xs = None
        if fold:
            st["ln_stats"].zero_()  # one memset node: every row-moment slot of this step
            st["stats_cursor"] = 0
            xs = self._next_stats(st)
        x = C.embed(st["next_tokens"], st["positions"], tr.wte.weight, tr.wpe.weight if tr.wpe is not None else None,
                    spec.pos_offset, None, xs)
        if tr.emb_norm is not None:
            x = C.norm(x, tr.emb_norm.weight, tr.emb_norm.bias, spec.norm_eps, spec.norm == "rmsnorm")
        trunk_x = x
        rms = spec.norm == "rmsnorm"
        fh = model.frozen_head
        L = len(self.layers)
        main = torch.cuda.current_stream()
        rf = None
        trunk_xs = xs
        mega = self.mega and not fold
        if mega:
            x = self._mega_stack(x, st)
        for i, W in enumerate(() if mega else self.layers):
            if i == self.branch:
                trunk_x = x
                if self.parallel_branches:
                    # the frozen reference branch only depends on the trunk activation: run it on a second stream so the
                    # graph has two independent chains (policy top blocks || reference blocks) instead of one long one
                    self.side.wait_stream(main)
                    with torch.cuda.stream(self.side):
                        y, ys = trunk_x, xs
                        for j, Wr in enumerate(self.ref_layers):
                            if fold:
                                y, ys = self._layer_folded(y, ys, Wr, self.ref_folded[j], st["kc"][L + j], st["vc"][L + j], st)
                            else:
                                y = self._layer(y, Wr, st["kc"][L + j], st["vc"][L + j], st)
                        rf = C.norm(y, fh.final_norm.weight, fh.final_norm.bias, spec.norm_eps, rms)
                        if self.cache_trunk:
                            st["trunk_decode"].index_copy_(1, st["step64"], trunk_x.unsqueeze(1))
            if i == self.branch:
                trunk_xs = xs
            if fold:
                x, xs = self._layer_folded(x, xs, W, self.folded[i], st["kc"][i], st["vc"][i], st)
            else:
                x = self._layer(x, W, st["kc"][i], st["vc"][i], st)
        if self.keep_trunk and not self.parallel_branches and not mega:
            st["trunk_decode"].index_copy_(1, st["step64"], trunk_x.unsqueeze(1))
        hf = C.norm(x, tr.ln_f.weight, tr.ln_f.bias, spec.norm_eps, rms)
        if self.filtered:
            V = spec.vocab_size
            C.gemm(hf, lm.lm_head.weight, lm.lm_head.bias, None, "none", True, st["logits"][:, :V])
            tok, tlp = C.sample_filtered(st["logits"], V, self.top_k, self.top_p, self.temperature, self.seed, st["step"],
                                         self.eos if st["min_new"] > 0 else -1, st["min_new"], st["seed_dev"])
        else:
            _, _, tok, tlp = C.lmhead(hf, lm.lm_head.weight, lm.lm_head.bias, None, True, self.temperature, self.seed,
                                      st["step"], self.eos if st["min_new"] > 0 else -1, st["min_new"], st["ws"], st["seed_dev"])
        if self.value_branch:
            val = None  # the value function has its own transformer branch: all positions are scored in one pass after the loop
        else:
            vh = model.v_head
            h1 = C.gemm(hf, vh[0].weight, vh[0].bias, None, "relu")
            val = C.rowdot(h1, vh[2].weight.view(-1), vh[2].bias)
        if self.defer_ref:
            ref_lp = tlp  # placeholder column; the real reference log-probs come from `_ref_score` after the loop
        elif rf is None:
            y, ys = trunk_x, (trunk_xs if fold else None)
            for j, W in enumerate(self.ref_layers):
                if fold:
                    y, ys = self._layer_folded(y, ys, W, self.ref_folded[j], st["kc"][L + j], st["vc"][L + j], st)
                else:
                    y = self._layer(y, W, st["kc"][L + j], st["vc"][L + j], st)
            rf = C.norm(y, fh.final_norm.weight, fh.final_norm.bias, spec.norm_eps, rms)
        else:
            main.wait_stream(self.side)
            rf.record_stream(main)
        if not self.defer_ref:
            _, ref_lp, _, _ = C.lmhead(rf, fh.lm_head.weight, fh.lm_head.bias, tok, False, 1.0, 0, None, -1, 0, st["ws_ref"])
        C.decode_step(tok, tlp, ref_lp, val, st["step"], st["R"], self.eos, self.pad, st["tokens_out"], st["lp_out"],
                      st["ref_lp_out"], st["val_out"], st["finished"], st["resp_lens"], st["seq_lens"], st["positions"],
                      st["next_tokens"], st["n_running"])
        st["step64"].add_(1)

    # ------------------------------------------------------------------------------------------------ state / graph
    def _build_state(self, B: int, Q: int, R: int):
        spec, dev = self.spec, self.device
        P = (Q + R + PAGE - 1) // PAGE
        n_layers = len(self.layers) + (0 if self.defer_ref else len(self.ref_layers))  # deferred scoring needs no ref KV cache
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        V = spec.vocab_size
        n_tiles = int(ops.C.lmhead_tiles(V))  # partials per row emitted by the fused LM-head epilogue
        st = dict(
            B=B, Q=Q, R=R, P=P,
            allocator=ops.C.PagedKVAllocator(B * P, PAGE),
            kc=[torch.zeros(B * P, PAGE, spec.num_kv_heads, spec.head_dim, dtype=torch.bfloat16, device=dev) for _ in range(n_layers)],
            vc=[torch.zeros(B * P, PAGE, spec.num_kv_heads, spec.head_dim, dtype=torch.bfloat16, device=dev) for _ in range(n_layers)],
            block_table=torch.zeros(B, P, **i32), seq_lens=torch.zeros(B, **i32), positions=torch.zeros(B, **i32),
            next_tokens=torch.zeros(B, dtype=torch.long, device=dev), finished=torch.zeros(B, **i32),
            resp_lens=torch.zeros(B, **i32), step=torch.zeros(1, **i32), step64=torch.zeros(1, dtype=torch.long, device=dev),
            n_running=torch.zeros(1, **i32),
            tokens_out=torch.full((B, R), self.pad, dtype=torch.long, device=dev), lp_out=torch.zeros(B, R, **f32),
            ref_lp_out=torch.zeros(B, R, **f32), val_out=torch.zeros(B, R, **f32),
            ws=torch.empty(5 * B * n_tiles + B, **f32), ws_ref=torch.empty(5 * B * n_tiles + B, **f32),
            trunk_decode=(torch.zeros(B, R, spec.hidden_size, dtype=torch.bfloat16, device=dev)
                          if self.keep_trunk else None),
            seed_dev=torch.zeros(1, dtype=torch.long, device=dev), min_new=0, graph=None,
            ln_stats=torch.zeros(2 * n_layers + 2, B, 2, **f32), stats_cursor=0,
            logits=torch.empty(B, (V + 7) // 8 * 8, **f32) if self.filtered else None,
        )
        return st

    def _ensure_state(self, B, Q, R):
        st = self._state
        if st is None or st["B"] != B or st["R"] != R or st["P"] < (Q + R + PAGE - 1) // PAGE:
            self._state = st = self._build_state(B, Q, R)
        return st

    def _reset(self, st, q_lens: torch.Tensor, last_tokens: torch.Tensor):
        B, P = st["B"], st["P"]
        alloc = st["allocator"]
        alloc.reset()
        for b in range(B):
            alloc.reserve(b, P * PAGE)
        st["block_table"].copy_(alloc.block_table(list(range(B)), P), non_blocking=True)
        st["seq_lens"].copy_(q_lens.to(torch.int32))
        st["positions"].copy_((q_lens - 1).to(torch.int32))
        st["next_tokens"].copy_(last_tokens)
        st["finished"].zero_()
        st["resp_lens"].zero_()
        st["step"].zero_()
        st["step64"].zero_()
        st["n_running"].fill_(B)
        st["tokens_out"].fill_(self.pad)
        for k in ("lp_out", "ref_lp_out", "val_out"):
            st[k].zero_()

    # ------------------------------------------------------------------------------------------------ prefill
    @torch.no_grad()
    def _prefill(self, st, prompt: torch.Tensor, mask: torch.Tensor):
        """Forward the first Q-1 prompt tokens, fill the paged caches, return prompt-position scores."""
        C, lm, model, spec = ops.C, self.lm, self.model, self.spec
        B, Q = prompt.shape
        T = Q - 1
        H = spec.hidden_size
        dev = self.device
        if T == 0:
            z = torch.zeros(B, 0, device=dev)
            return z, z.clone(), z.clone(), torch.zeros(B, 0, H, dtype=torch.bfloat16, device=dev)
        ids, am = prompt[:, :T], mask[:, :T]
        pos = (am.long().cumsum(-1) - 1).clamp_min(0)
        out = lm(input_ids=ids, attention_mask=am, position_ids=pos, use_cache=True, output_hidden_states=True,
                 compute_logits=False)
        first = (Q - mask.long().sum(1)).to(torch.int32)  # left padding
        lens = (mask.long().sum(1) - 1).to(torch.int32)

        def scatter(layer_idx, kv):
            k, v = kv
            k2 = k.transpose(1, 2).reshape(B, T, -1).contiguous()
            v2 = v.transpose(1, 2).reshape(B, T, -1).contiguous()
            C.paged_kv_write(k2, v2, st["kc"][layer_idx], st["vc"][layer_idx], st["block_table"], first, lens,
                             spec.num_kv_heads, spec.head_dim)

        for i, kv in enumerate(out.past_key_values):
            scatter(i, kv)
        trunk = out.hidden_states[self.branch]
        final = out.last_hidden_state
        labels = prompt[:, 1:Q]
        lp, _ = ops.fused_logprob(final, lm.lm_head.weight, lm.lm_head.bias, labels)
        vals = torch.zeros(B, T, device=dev)
        if self.defer_ref:
            return lp.float(), None, vals, trunk
        # reference branch on the shared trunk activation
        fh = model.frozen_head
        ctx = build_attn_context(spec, am, pos, T, 0, trunk.dtype, dev)
        y = trunk
        L = len(self.layers)
        for j, blk in enumerate(fh.decoder_blocks):
            y, present = blk(y, ctx, None, True)
            scatter(L + j, present)
        y = fh.final_norm(y)
        ref_lp, _ = ops.fused_logprob(y, fh.lm_head.weight, fh.lm_head.bias, labels)
        return lp.float(), ref_lp.float(), vals, trunk

    @torch.no_grad()
    def _ref_score(self, prompt: torch.Tensor, mask: torch.Tensor, trunk: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """Reference log-probs of ``labels`` ``[B, T]`` from the trunk activations ``[B, T, H]`` of the T = Q-1+r input
        positions: the frozen blocks and LM head run ONCE over all positions (reference: ``forward_hydra``,
        ``trlx/models/modeling_ppo.py:356-401``, which re-runs the whole model instead)."""
        spec, fh = self.spec, self.model.frozen_head
        B, T = labels.shape
        Q = prompt.shape[1]
        # inputs are prompt[0..Q-1] followed by the sampled tokens: the prompt keeps its (left-)padding mask, every generated
        # position is visible — exactly what the incremental decode attends to
        am = torch.cat([mask, mask.new_ones(B, max(T - Q, 0))], 1)[:, :T]
        pos = (am.long().cumsum(-1) - 1).clamp_min(0)
        ctx = build_attn_context(spec, am, pos, T, 0, trunk.dtype, self.device)
        y = trunk
        for blk in fh.decoder_blocks:
            y, _ = blk(y, ctx, None, False)
        y = fh.final_norm(y)
        ref_lp, _ = ops.fused_logprob(y, fh.lm_head.weight, fh.lm_head.bias, labels)
        return ref_lp.float()

    @torch.no_grad()
    def _ref_score_adapter_free(self, all_tokens: torch.Tensor, full_mask: torch.Tensor, Q: int) -> torch.Tensor:
        """LoRA models: reference log-probs of every next token from ONE batched forward with the adapters switched off (the
        reference toggles them the same way, ``trlx/models/modeling_ppo.py:318-324``, but re-runs the model per chunk after
        ``generate``); there is no frozen branch or shared trunk with adapters on every layer."""
        lm, peft_model = self.lm, self.model.base_model
        ids, am = all_tokens[:, :-1], full_mask[:, :-1]
        # generated positions are all visible to later ones even when a row sampled the pad token id (pad == eos)
        am = torch.cat([am[:, :Q], torch.ones_like(am[:, Q:])], 1)
        pos = (am.long().cumsum(-1) - 1).clamp_min(0)
        ctx = peft_model.disable_adapter() if hasattr(peft_model, "disable_adapter") else contextlib.nullcontext()
        with ctx:
            out = lm(input_ids=ids, attention_mask=am, position_ids=pos, compute_logits=False)
        ref_lp, _ = ops.fused_logprob(out.last_hidden_state, lm.lm_head.weight, lm.lm_head.bias, all_tokens[:, 1:])
        return ref_lp.float()

    @contextlib.contextmanager
    def _static_weights(self):
        """Inside the rollout graphs no kernel writes a weight matrix, so the GEMMs captured here may fetch their first
        weight tiles ahead of the programmatic-dependent-launch wait — while the producer of their activations is still
        running (``b200_set_static_weights``).  ``TRLX_B200_WEIGHT_PREFETCH=0`` disables it."""
        import os

        on = os.environ.get("TRLX_B200_WEIGHT_PREFETCH", "1") == "1" and hasattr(ops.C, "set_static_weights")
        if on:
            ops.C.set_static_weights(True)
        try:
            yield
        finally:
            if on:
                ops.C.set_static_weights(False)

    def _prefill_maybe_graphed(self, st, prompt, mask):
        """The prefill is ~200 small launches for short prompts (CPU-bound when issued eagerly): capture it once per
        (batch, prompt width) and replay it from static input buffers."""
        import os

        if not self.use_cuda_graph or os.environ.get("TRLX_B200_PREFILL_GRAPH", "1") != "1" or st.get("prefill_failed") \
                or prompt.shape[1] < 2:
            return self._prefill(st, prompt, mask)
        graphs = st.setdefault("prefill_graphs", {})  # one per prompt width (the state itself serves every width <= P pages)
        Q = int(prompt.shape[1])
        if Q not in graphs:
            if len(graphs) >= 8:
                return self._prefill(st, prompt, mask)
            try:
                pf_prompt, pf_mask = prompt.clone(), mask.clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                before = ops.launch_count()
                with torch.cuda.stream(side):
                    self._prefill(st, pf_prompt, pf_mask)
                launches = ops.launch_count() - before
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph), self._static_weights():
                    outs = self._prefill(st, pf_prompt, pf_mask)
                graphs[Q] = (graph, pf_prompt, pf_mask, outs, launches)
            except Exception as err:  # keep the eager path if anything in the prefill is not capturable
                logger.warning(f"prefill CUDA graph disabled ({type(err).__name__}: {err})")
                st["prefill_failed"] = True
                torch.cuda.synchronize()
                return self._prefill(st, prompt, mask)
        graph, pf_prompt, pf_mask, outs, launches = graphs[Q]
        pf_prompt.copy_(prompt, non_blocking=True)
        pf_mask.copy_(mask, non_blocking=True)
        graph.replay()
        ops.add_launches(launches)
        return outs  # static buffers: every consumer below copies (cat / index) before the next replay

    # ------------------------------------------------------------------------------------------------ public API
    @torch.no_grad()
    def rollout(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        dev = self.device
        prompt = input_ids.to(dev, non_blocking=True)
        mask = (attention_mask if attention_mask is not None else torch.ones_like(input_ids)).to(dev, non_blocking=True).long()
        B, Q = prompt.shape
        g = self.gen
        R = g.get("max_new_tokens")
        if R is None:
            R = max(int(g.get("max_length", Q + 20)) - Q, 1)
        R = int(R)
        self.refresh_folded()
        st = self._ensure_state(B, Q, R)
        q_lens = mask.sum(1)
        self._reset(st, q_lens, prompt[:, -1])
        self.calls += 1
        st["seed_dev"].fill_((self.calls * 0x9E3779B1) & 0x7FFFFFFFFFFF)
        st["min_new"] = max(int(g.get("min_new_tokens") or 0), int(g.get("min_length") or 0) - Q, 0)

        torch.cuda.nvtx.range_push("engine/prefill")
        lp_p, ref_lp_p, val_p, trunk_p = self._prefill_maybe_graphed(st, prompt, mask)
        torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push("engine/decode")

        if self.use_cuda_graph:
            if st["graph"] is None or st.get("graph_key_min_new") != st["min_new"]:
                # warm-up on a side stream (also loads kernels / sets attributes), then capture
                snap = {k: st[k].clone() for k in ("seq_lens", "positions", "next_tokens", "finished", "resp_lens", "step",
                                                   "step64", "n_running")}
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                before = ops.launch_count()
                with torch.cuda.stream(s):
                    self._decode_step(st)
                self.launches_per_step = ops.launch_count() - before
                torch.cuda.current_stream().wait_stream(s)
                for k, v in snap.items():
                    st[k].copy_(v)
                st["tokens_out"].fill_(self.pad)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph), self._static_weights():
                    self._decode_step(st)
                for k, v in snap.items():
                    st[k].copy_(v)
                st["tokens_out"].fill_(self.pad)
                for k2 in ("lp_out", "ref_lp_out", "val_out"):
                    st[k2].zero_()
                st["graph"], st["graph_key_min_new"] = graph, st["min_new"]
                # the sampling noise varies per step (device step counter) and per call (device seed offset)
            for s_i in range(R):
                st["graph"].replay()
                ops.add_launches(self.launches_per_step)
                if (s_i & 7) == 7 and s_i + 1 < R and int(st["n_running"].item()) == 0:
                    break
        else:
            for s_i in range(R):
                self._decode_step(st)
                if (s_i & 7) == 7 and s_i + 1 < R and int(st["n_running"].item()) == 0:
                    break

        torch.cuda.nvtx.range_pop()
        resp_lens = st["resp_lens"].clone()
        r_max = max(int(resp_lens.max().item()), 1)
        sample_outputs = st["tokens_out"][:, :r_max].clone()
        logprobs = torch.cat([lp_p, st["lp_out"][:, :r_max]], 1)
        values = torch.cat([val_p, st["val_out"][:, :r_max]], 1)
        all_tokens = torch.cat([prompt, sample_outputs], 1)
        # the sampled tokens start their trip to (pinned) host memory now, ahead of the scoring kernels queued below: the
        # trainer detokenises and calls the reward function while the GPU is still scoring the rollout
        host_tokens = torch.empty(all_tokens.shape, dtype=all_tokens.dtype, pin_memory=True)
        host_tokens.copy_(all_tokens, non_blocking=True)
        host_ready = torch.cuda.Event()
        host_ready.record()
        full_mask = all_tokens.not_equal(self.pad).long()
        trunk = None
        if self.keep_trunk:
            trunk = torch.cat([trunk_p, st["trunk_decode"][:, :r_max]], 1)
        if self.value_branch:
            am = torch.cat([mask, mask.new_ones(B, r_max)], 1)[:, :-1]
            pos = (am.long().cumsum(-1) - 1).clamp_min(0)
            _, v_all, _, _ = self.model.score(all_tokens[:, :-1], am, pos, all_tokens[:, 1:], trunk_hidden=trunk, with_ref=False)
            values = v_all.float()
            values[:, :Q - 1] = 0.0  # prompt positions carry no value (matches the in-graph path)
        if self.lora:
            torch.cuda.nvtx.range_push("engine/reference_scoring")
            ref_logprobs = self._ref_score_adapter_free(all_tokens, full_mask, Q)
            torch.cuda.nvtx.range_pop()
        elif self.defer_ref:
            torch.cuda.nvtx.range_push("engine/reference_scoring")
            ref_logprobs = self._ref_score(prompt, mask, trunk, all_tokens[:, 1:Q + r_max])
            torch.cuda.nvtx.range_pop()
            if not self.cache_trunk:
                trunk = None
        else:
            ref_logprobs = torch.cat([ref_lp_p, st["ref_lp_out"][:, :r_max]], 1)
        return dict(samples=all_tokens, prompt_tensors=prompt, sample_outputs=sample_outputs, logprobs=logprobs,
                    ref_logprobs=ref_logprobs, values=values, mask=full_mask, start=Q - 1, trunk=trunk,
                    samples_host=(host_tokens, host_ready))
