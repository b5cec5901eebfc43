"""Fused tensor-parallel GEMM ↔ collective operations over NVLink peer memory.

* :meth:`FusedTP.allgather_gemm` — the sequence-sharded activation of every rank lives in a symmetric buffer; ONE GEMM
  kernel computes ``concat_r(x_r) · Wᵀ`` while the peers' shards are still being copied into the local gathered buffer:
  its TMA producer starts on the local shard and gates each remote shard on a device-side ready flag.
* :meth:`FusedTP.gemm_reduce_scatter` — every rank's partial product tile goes from the GEMM epilogue straight into the
  owner's staging slot over NVLink (bf16, 16-byte stores) as soon as its accumulator completes; the owner sums the slots.
  (Earlier variants — remote TMA loads per tile, fp32 ``red.global.add`` — are kept behind environment switches.)

Ordering between ranks uses the device-side flag barrier of ``csrc/optim.cu`` on the buffers' signal pads.

The backward passes are the same two primitives with the roles swapped (Apex does AG / RS in both directions,
``trlx/models/modeling_nemo_ppo.py:93-120``): column-parallel ``dX = RS(dY · W)`` is a GEMM→reduce-scatter on a transposed
weight copy, row-parallel ``dX = AG(dY) · W`` an all-gather→GEMM; the weight gradients contract over the *gathered* rows,
which arrive by peer copies on a side stream while the dgrad kernel runs (column) or are the gathered buffer the dgrad
all-gather→GEMM just filled (row).  ``TRLX_B200_TP_FUSED_BWD=0`` switches the backward to NCCL + ``torch.matmul``.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from trlx_b200 import ops


class FusedTP:
    def __init__(self, group, rank: int, size: int, device: torch.device):
        self.group, self.rank, self.size, self.device = group, rank, size, device
        self._bufs: Dict[Tuple, Tuple[torch.Tensor, object]] = {}
        self._local: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self._pads = None

    def _symm(self, key: Tuple, shape, dtype):
        if key not in self._bufs:
            import torch.distributed._symmetric_memory as symm

            t = symm.empty(*shape, dtype=dtype, device=self.device)
            name = self.group.group_name if self.group is not None else dist.group.WORLD.group_name
            hdl = symm.rendezvous(t, name)
            self._bufs[key] = (t, hdl)
            if self._pads is None:
                self._pads = list(hdl.signal_pad_ptrs)
        return self._bufs[key]

    def barrier(self):
        ops.C.signal_barrier(self._pads, self.rank, self.epoch)

    @staticmethod
    def usable(rows_per_rank: int, k: int, n: int) -> bool:
        return rows_per_rank % 128 == 0 and k % 8 == 0 and n % 8 == 0

    def allgather_gemm(self, x_local: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: str = "none",
                       slot: str = "fwd"):
        """``x_local`` ``[m, K]`` (this rank's sequence shard) → ``[m·size, N]`` = act(all_gather(x) · wᵀ + bias).

        Every remote byte crosses NVLink exactly once: the peers' shards are pulled into a local gathered buffer by copies
        on a side stream (rank+1 first, …) while ONE GEMM kernel is already running on the main stream — its TMA producer
        starts on this rank's own shard and waits on a per-shard ready flag (device memory, acquire load) before touching the
        rows of a shard that is still in flight.  (The first version read remote tiles straight out of the peers' HBM with
        per-peer tensor maps; remote reads are not cached in the local L2, so every N-tile re-fetched its A rows over the
        link — 5x slower than NCCL + cuBLAS at TP = 4.  ``TRLX_B200_TP_REMOTE_TMA=1`` keeps that variant for comparison.)"""
        m, K = x_local.shape
        buf, hdl = self._symm(("ag", m, K, slot), (m, K), torch.bfloat16)
        buf.copy_(x_local)
        self.barrier()  # every rank's shard is in place
        if os.environ.get("TRLX_B200_TP_REMOTE_TMA") == "1":
            out = ops.C.gemm_allgather(list(hdl.buffer_ptrs), m, K, K, w, bias, act)
            self.barrier()
            return out
        full, flags, ep = self._start_gather(buf, hdl, slot)
        out = ops.C.gemm_flagged(full, w, bias, act, flags, ep, m, self.rank)
        torch.cuda.current_stream().wait_stream(self._copy_stream)
        self.barrier()  # all peers finished reading before the buffer is reused
        self.last_gathered = full  # valid until the next gather with the same (shape, slot): the wgrad of a row-parallel layer reads it
        return out

    def _start_gather(self, x_local: torch.Tensor, hdl, slot: str = "fwd"):
        """Own shard into the gathered buffer now, the peers' shards by copy-engine pulls on the side stream; returns
        ``(gathered buffer, per-shard ready flags, epoch)`` — consumers either gate on the flags (fused GEMM) or wait for the
        side stream."""
        m, K = x_local.shape
        key = ("ag_full", m, K, slot)
        if key not in self._local:
            self._local[key] = (torch.empty(self.size * m, K, dtype=torch.bfloat16, device=self.device),
                                torch.zeros(self.size, dtype=torch.int32, device=self.device))
            self._copy_stream = getattr(self, "_copy_stream", None) or torch.cuda.Stream(device=self.device)
        full, flags = self._local[key]
        self._epoch_host = getattr(self, "_epoch_host", 0) + 1
        ep = self._epoch_host
        main = torch.cuda.current_stream()
        full[self.rank * m:(self.rank + 1) * m].copy_(x_local)
        flags[self.rank:self.rank + 1].fill_(ep)
        side = self._copy_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for d in range(1, self.size):
                r = (self.rank + d) % self.size
                peer = hdl.get_buffer(r, (m, K), torch.bfloat16)
                full[r * m:(r + 1) * m].copy_(peer, non_blocking=True)
                flags[r:r + 1].fill_(ep)
        return full, flags, ep

    def gather_async(self, x_local: torch.Tensor, slot: str = "bwd"):
        """Start an all-gather of ``x_local`` ``[m, K]`` over NVLink peer copies; :meth:`gather_wait` returns the ``[m·size, K]``
        result.  Anything launched in between on the current stream overlaps the transfer."""
        m, K = x_local.shape
        buf, hdl = self._symm(("ag", m, K, slot), (m, K), torch.bfloat16)
        buf.copy_(x_local)
        self.barrier()
        full, _, _ = self._start_gather(buf, hdl, slot)
        return full

    def gather_wait(self, full: torch.Tensor) -> torch.Tensor:
        torch.cuda.current_stream().wait_stream(self._copy_stream)
        self.barrier()  # peers are done reading this rank's staging buffer
        return full

    def gemm_reduce_scatter(self, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                            residual: Optional[torch.Tensor] = None):
        """``x`` ``[M, K_local]``, ``w`` ``[N, K_local]`` → this rank's ``[M/size, N]`` rows of Σ_ranks x·wᵀ (+bias) (+residual).

        The GEMM epilogue stores each bf16 partial tile straight into slot ``rank`` of the OWNER's staging buffer over NVLink
        (16-byte stores, tile by tile as the accumulators complete), and the owner sums its ``size`` slots afterwards.  (The
        first version used fp32 ``red.global.add`` into the owner's accumulator: twice the bytes and remote atomics — 3x slower
        than cuBLAS + NCCL at TP = 4; ``TRLX_B200_TP_RED_ADD=1`` keeps it for comparison.)"""
        M, N = x.shape[0], w.shape[0]
        rows = M // self.size
        if os.environ.get("TRLX_B200_TP_RED_ADD") == "1" or N % 8:
            acc, hdl = self._symm(("rs", rows, N), (rows, N), torch.float32)
            acc.zero_()
            self.barrier()  # accumulators are clean everywhere
            ops.C.gemm_reduce_scatter(x, w, list(hdl.buffer_ptrs), N, bias)
            self.barrier()  # every partial sum has landed
            return ops.C.rs_finalize(acc, None, residual)
        if os.environ.get("TRLX_B200_TP_RS_NVLS", "1") == "1" and N % 8 == 0:
            out = self._gemm_rs_nvls(x, w, bias, residual)
            if out is not None:
                return out
        stage, hdl = self._symm(("rs_stage", rows, N), (self.size, rows, N), torch.bfloat16)
        self.barrier()  # the previous consumer of the staging slots is done
        ops.C.gemm_stage_scatter(x, w, list(hdl.buffer_ptrs), self.rank, N, bias)  # bias: real on one rank, zeros elsewhere
        self.barrier()  # every rank's partial tiles have landed
        return ops.C.stage_reduce(stage, None, residual)


    def _gemm_rs_nvls(self, x, w, bias, residual):
        """GEMM → reduce-scatter through the NVSwitch: the partial product is written LOCALLY (plain TMA-store epilogue, the GEMM
        runs at its stand-alone speed) into a symmetric buffer that is also mapped at one multicast address; after a flag
        barrier the owner of a row block reads it with ``multimem.ld_reduce`` — the switch adds the ``size`` copies in fp32 — and
        adds the residual on the way out.  The columns are processed in ``TRLX_B200_TP_RS_SPLIT`` windows (default 2): the
        reduction of window ``i`` runs on a side stream underneath the GEMM of window ``i + 1``."""
        M, N = x.shape[0], w.shape[0]
        rows = M // self.size
        part, hdl = self._symm(("rs_part", M, N), (M, N), torch.bfloat16)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            return None
        split = int(os.environ.get("TRLX_B200_TP_RS_SPLIT", "2"))
        while split > 1 and (N % (split * 128) or N // split < 1024):
            split -= 1
        out = torch.empty(rows, N, dtype=torch.bfloat16, device=x.device)
        mine = mc + self.rank * rows * N * 2
        main = torch.cuda.current_stream()
        self._rs_stream = getattr(self, "_rs_stream", None) or torch.cuda.Stream(device=self.device)
        side = self._rs_stream
        res = residual
        self.barrier()  # every rank has finished reducing the previous contents of the partial buffer
        w_cols = N // split
        for i in range(split):
            c0 = i * w_cols
            # the bias rides in the partial product of the one rank that was handed it (callers pass it on a single rank)
            ops.C.gemm(x, w[c0:c0 + w_cols], None if bias is None else bias[c0:c0 + w_cols], None, "none", False,
                       part[:, c0:c0 + w_cols])
            self.barrier()  # this window's partials are complete on every rank
            if i + 1 < split:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.C.mc_reduce_rows(mine, out, None, res, c0, w_cols, N, 148)
            else:
                ops.C.mc_reduce_rows(mine, out, None, res, c0, w_cols, N, 0)
        if split > 1:
            main.wait_stream(side)
        return out


def _fused_bwd() -> bool:
    return os.environ.get("TRLX_B200_TP_FUSED_BWD", "1") == "1"


class _ColumnLinearFused(torch.autograd.Function):
    """AG→GEMM forward; backward: ``dX = RS(dY · W)`` on the fused GEMM→reduce-scatter kernel (transposed weight copy) while the
    all-gather of ``x`` for ``dW = dYᵀ · AG(x)`` travels over NVLink on the side stream."""

    @staticmethod
    def forward(ctx, x, w, b, fused: FusedTP):
        B, t, K = x.shape
        y = fused.allgather_gemm(x.reshape(B * t, K).contiguous(), w, b)
        ctx.save_for_backward(x, w)
        ctx.fused, ctx.has_bias = fused, b is not None
        # rows are ordered rank-major ([rank][batch][time]) → back to [B, T, N]
        return y.view(fused.size, B, t, -1).permute(1, 0, 2, 3).reshape(B, fused.size * t, -1)

    @staticmethod
    def backward(ctx, gy):
        from trlx_b200.ops.functional import col_sum, grad_weight

        x, w = ctx.saved_tensors
        fused = ctx.fused
        B, t, K = x.shape
        g = gy.reshape(B, fused.size, t, -1).permute(1, 0, 2, 3).reshape(fused.size * B * t, -1).contiguous()
        gx = gw = gb = None
        use_fused = _fused_bwd() and FusedTP.usable(B * t, g.shape[1], K) and g.dtype == torch.bfloat16
        xs = None
        if ctx.needs_input_grad[1] and use_fused:
            xs = fused.gather_async(x.reshape(B * t, K).contiguous())  # peer copies run behind the dgrad kernel below
        if ctx.needs_input_grad[0]:
            if use_fused:
                gx = fused.gemm_reduce_scatter(g, w.t().contiguous()).view(B, t, K)
            else:
                full = g @ w
                out = torch.empty(B * t, K, dtype=full.dtype, device=full.device)
                dist.reduce_scatter_tensor(out, full, group=fused.group)
                gx = out.view(B, t, K)
        if ctx.needs_input_grad[1]:
            if xs is not None:
                gw = grad_weight(g, fused.gather_wait(xs))
            else:
                xs = torch.empty(fused.size * B * t, K, dtype=x.dtype, device=x.device)
                dist.all_gather_into_tensor(xs, x.reshape(B * t, K).contiguous(), group=fused.group)
                gw = g.t() @ xs
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(g)
        return gx, gw, gb, None


class _RowLinearFused(torch.autograd.Function):
    """GEMM→RS forward; backward: ``dX = AG(dY) · W`` on the fused all-gather→GEMM kernel (transposed weight copy), and
    ``dW = AG(dY)ᵀ · x`` from the gathered buffer that kernel just filled."""

    @staticmethod
    def forward(ctx, x, w, b, fused: FusedTP):
        B, T, K = x.shape
        t = T // fused.size
        xr = x.reshape(B, fused.size, t, K).permute(1, 0, 2, 3).reshape(fused.size * B * t, K).contiguous()  # rank-major rows
        y = fused.gemm_reduce_scatter(xr, w, b)
        ctx.save_for_backward(xr, w)
        ctx.fused, ctx.has_bias, ctx.shape = fused, b is not None, (B, T, K)
        return y.view(B, t, -1)

    @staticmethod
    def backward(ctx, gy):
        from trlx_b200.ops.functional import col_sum, grad_weight

        xr, w = ctx.saved_tensors
        fused = ctx.fused
        B, T, K = ctx.shape
        t = T // fused.size
        g_local = gy.reshape(B * t, -1).contiguous()
        N = g_local.shape[1]
        gx = gw = gb = None
        use_fused = _fused_bwd() and FusedTP.usable(B * t, N, K) and g_local.dtype == torch.bfloat16
        g = None
        if use_fused:
            if ctx.needs_input_grad[0]:
                full = fused.allgather_gemm(g_local, w.t().contiguous(), None, "none", slot="bwd")  # [M, K_local]
                gx = full.view(fused.size, B, t, K).permute(1, 0, 2, 3).reshape(B, T, K)
                g = fused.last_gathered
            else:
                g = fused.gather_wait(fused.gather_async(g_local))
        else:
            g = torch.empty(fused.size * B * t, N, dtype=g_local.dtype, device=g_local.device)
            dist.all_gather_into_tensor(g, g_local, group=fused.group)
            if ctx.needs_input_grad[0]:
                gx = (g @ w).view(fused.size, B, t, K).permute(1, 0, 2, 3).reshape(B, T, K)
        if ctx.needs_input_grad[1]:
            gw = grad_weight(g, xr) if use_fused else g.t() @ xr
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(g) if use_fused else g.sum(0)
        return gx, gw, gb, None


def column_linear(fused: FusedTP, linear, x):
    return _ColumnLinearFused.apply(x, linear.weight, linear.bias, fused)


def row_linear(fused: FusedTP, linear, x):
    # the row-parallel bias is real on one rank and zero on the others (tensor_parallel._shard_cols): every rank may add "its" bias
    return _RowLinearFused.apply(x, linear.weight, linear.bias, fused)
