"""sm_100a kernels vs plain PyTorch fp32 references (run on a B200: ``pytest -m gpu``)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from trlx_b200 import ops

    assert ops.available(), "extension must load on a GPU box"
    return ops.C


def _bf(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 2304, 768), (1600, 3072, 768), (100, 776, 3072), (257, 50257, 768), (32, 768, 64)])
@pytest.mark.parametrize("bn", [0, 32, 64, 128, 256])
def test_gemm_matches_fp32(C, M, N, K, bn):
    torch.manual_seed(0)
    x, w = _bf(M, K), _bf(N, K, scale=K ** -0.5)
    ref = x.float() @ w.float().t()
    out = C.gemm(x, w, None, None, "none", True, None, None, 1.0, bn)
    torch.testing.assert_close(out, ref, atol=2e-3, rtol=2e-3)
    out16 = C.gemm(x, w, force_bn=bn)
    torch.testing.assert_close(out16.float(), ref, atol=3e-2, rtol=3e-2)


@pytest.mark.parametrize("M,N,K,bn", [(2048, 4096, 512, 128), (4096, 8192, 256, 256), (1300, 20000, 192, 0), (9000, 300, 128, 32)])
def test_gemm_persistent_many_tiles(C, M, N, K, bn):
    """More tiles than SMs: every CTA walks several tiles through the double-buffered TMEM accumulator."""
    torch.manual_seed(M + N)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
    y = C.gemm(x, w, None, None, "none", force_bn=bn)
    ref = x.float() @ w.float().t()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), err


@pytest.mark.parametrize("M,N,K", [(1792, 768, 3072), (300, 200, 136), (768, 50257 // 8 * 8, 1280), (2304, 768, 1792), (1280, 768, 50257)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
def test_gemm_mn_major_operands(C, M, N, K, a_mn, b_mn):
    """Backward-pass layouts: operands stored transposed ([K, M] / [K, N]) are consumed without a transposed copy."""
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    if (a_mn and M % 8) or (b_mn and N % 8) or (K % 8 and not (a_mn and b_mn)):
        pytest.skip("row pitches must be multiples of 8 elements")
    a_in = a.t().contiguous() if a_mn else a
    b_in = b.t().contiguous() if b_mn else b
    out = C.gemm_ex(a_in, b_in, a_mn, b_mn)
    assert out.shape == (M, N)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-2, err


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("split_k", [-1, 3, 16])
def test_gemm_split_k(C, a_mn, b_mn, split_k):
    """Skinny output, long contraction: K-ranges are accumulated into an fp32 buffer by red.add from the epilogue."""
    torch.manual_seed(3)
    M, N, K = 256, 192, 8192 + 64
    a = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    out = C.gemm_ex(a.t().contiguous() if a_mn else a, b.t().contiguous() if b_mn else b, a_mn, b_mn, False, split_k)
    out32 = C.gemm_ex(a.t().contiguous() if a_mn else a, b.t().contiguous() if b_mn else b, a_mn, b_mn, True, split_k)
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert (out32 - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("rms", [False, True])
def test_gemm_with_folded_norm_and_row_moments(C, rms):
    """LN(x)·Wᵀ computed as rstd·(x·(γW)ᵀ) − rstd·μ·c1 + (W·β + b) from producer-side row moments; the output's own row
    moments are accumulated for the next folded norm."""
    import torch.nn.functional as F

    torch.manual_seed(11)
    M, K, N, eps = 128, 768, 2304, 1e-5
    x = (torch.randn(M, K, device="cuda") * 1.5 + 0.3).to(torch.bfloat16)
    gamma = (1 + 0.1 * torch.randn(K, device="cuda")).to(torch.bfloat16)
    beta = (0.1 * torch.randn(K, device="cuda")).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    b = (0.1 * torch.randn(N, device="cuda")).to(torch.bfloat16)
    res = (torch.randn(M, N, device="cuda")).to(torch.bfloat16)
    xf = x.float()
    if rms:
        ref_n = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * gamma.float()
        bias_f = b.float()
    else:
        ref_n = F.layer_norm(xf, (K,), gamma.float(), beta.float(), eps)
        bias_f = W.float() @ beta.float() + b.float()
    ref = ref_n @ W.float().t() + b.float() + res.float()
    Wf = (W.float() * gamma.float()).to(torch.bfloat16)
    c1 = Wf.float().sum(1).contiguous()
    stats = torch.stack([xf.sum(1), xf.pow(2).sum(1)], 1).contiguous()
    out_stats = torch.zeros(M, 2, device="cuda")
    y = C.gemm_ln(x, Wf, bias_f.to(torch.bfloat16), res, "none", stats, c1, eps, rms, out_stats)
    err = (y.float() - ref).abs().max().item()
    assert err <= 3e-2 * ref.abs().max().item(), err
    torch.testing.assert_close(out_stats[:, 0], y.float().sum(1), atol=0.5, rtol=1e-3)
    torch.testing.assert_close(out_stats[:, 1], y.float().pow(2).sum(1), atol=0.5, rtol=1e-3)
    # embed produces the first set of moments
    wte = (torch.randn(50, K, device="cuda")).to(torch.bfloat16)
    wpe = (torch.randn(16, K, device="cuda")).to(torch.bfloat16)
    tok = torch.randint(0, 50, (M,), device="cuda")
    pos = torch.randint(0, 16, (M,), device="cuda", dtype=torch.int32)
    st0 = torch.zeros(M, 2, device="cuda")
    e = C.embed(tok, pos, wte, wpe, 0, None, st0)
    torch.testing.assert_close(st0[:, 0], e.float().sum(1), atol=0.1, rtol=1e-3)
    torch.testing.assert_close(st0[:, 1], e.float().pow(2).sum(1), atol=0.1, rtol=1e-3)


@pytest.mark.parametrize("M,N,K", [(512, 512, 256), (2048, 4096, 512), (1000, 1304, 192), (4096, 4096, 4096)])
def test_gemm_cta_pair(C, M, N, K):
    """cta_group::2: a cluster of two CTAs computes 256 x 256 tiles with one UMMA (M = 256) issued by the leader."""
    torch.manual_seed(M + K)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, device="cuda")).to(torch.bfloat16)
    y = C.gemm(x, w, b, None, "none", force_bn=-2)
    ref = x.float() @ w.float().t() + b.float()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), err


@pytest.mark.parametrize("M,N,K", [(128, 2304, 768), (128, 768, 3072), (128, 3072, 768), (128, 768, 768), (100, 776, 3072),
                                   (5, 64, 256), (77, 1000, 8192)])
@pytest.mark.parametrize("act", ["none", "gelu_new"])
def test_gemm_cluster_split_k(C, M, N, K, act):
    """Decode-shaped GEMM on a cluster of S CTAs per tile: K split across the cluster, fp32 partial tiles exchanged through
    distributed shared memory and summed in the leader's epilogue (bias + activation + residual + TMA store)."""
    torch.manual_seed(N + K)
    x, w, b, r = _bf(M, K, scale=0.5), _bf(N, K, scale=0.5), _bf(N), _bf(M, N)
    y = C.gemm(x, w, b, r, act, force_bn=-3)
    pre = x.float() @ w.float().t() + b.float()
    ref = (F.gelu(pre, approximate="tanh") if act == "gelu_new" else pre) + r.float()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    # the automatic plan picks the same kernel for these shapes (or the plain one): results must agree
    y2 = C.gemm(x, w, b, r, act)
    assert (y2.float() - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize("M,N,K", [(128, 2304, 768), (300, 1000 // 16 * 16, 256), (2048, 4096, 1024)])
def test_gemm_fp8_with_row_and_channel_scales(C, M, N, K):
    """e4m3 x e4m3 on the tensor cores (kind::f8f6f4): exact against the dequantised operands, close to the bf16 product."""
    import torch.nn.functional as F

    torch.manual_seed(K)
    x = (torch.randn(M, K, device="cuda") * 2 + 0.5).to(torch.bfloat16)
    g = (1 + 0.1 * torch.randn(K, device="cuda")).to(torch.bfloat16)
    bt = (0.1 * torch.randn(K, device="cuda")).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda").to(torch.bfloat16)
    a8, a_s = C.norm_quant(x, g, bt, 1e-5, False)
    h = F.layer_norm(x.float(), (K,), g.float(), bt.float(), 1e-5)
    a_deq = a8.view(torch.float8_e4m3fn).float() * a_s[:, None]
    assert (a_deq - h).abs().max().item() <= 0.07 * h.abs().max().item()  # e4m3: 3 mantissa bits
    w_s = (W.float().abs().amax(1) / 448.0).contiguous()
    w8 = (W.float() / w_s[:, None]).to(torch.float8_e4m3fn)
    w_deq = w8.float() * w_s[:, None]
    y = C.gemm_fp8(a8, w8.view(torch.uint8), a_s, w_s, bias, None, "none")
    exact = a_deq @ w_deq.t() + bias.float()
    assert (y.float() - exact).abs().max().item() <= 1e-2 * exact.abs().max().item() + 1e-2
    full = h @ W.float().t() + bias.float()
    rel = (y.float() - full).norm() / full.norm()
    assert rel.item() < 0.06, rel.item()


def test_lmhead_dlogits(C):
    torch.manual_seed(5)
    M, V, K = 300, 50257, 256
    h = (torch.randn(M, K, device="cuda") * 0.3).to(torch.bfloat16)
    w = (torch.randn(V, K, device="cuda") * 0.3).to(torch.bfloat16)
    b = (torch.randn(V, device="cuda") * 0.1).to(torch.bfloat16)
    lab = torch.randint(0, V, (M,), device="cuda")
    lab[::7] = -1
    g = torch.randn(M, device="cuda")
    logits = h.float() @ w.float().t() + b.float()
    lse = torch.logsumexp(logits, -1)
    onehot = torch.zeros_like(logits).scatter_(1, lab.clamp_min(0)[:, None], 1.0)
    ref = (onehot - torch.exp(logits - lse[:, None])) * g[:, None]
    ref[lab < 0] = 0
    d = C.lmhead_dlogits(h, w, b, lab, lse, g)
    assert d.shape == (M, V) and d.stride(0) % 64 == 0
    assert (d.float() - ref).abs().max().item() < 2e-2 * max(ref.abs().max().item(), 1e-3) + 1e-3


@pytest.mark.parametrize("act", ["none", "gelu_new", "gelu", "relu", "silu"])
def test_gemm_epilogue(C, act):
    torch.manual_seed(1)
    M, N, K = 300, 1536, 768
    x, w, b, r = _bf(M, K), _bf(N, K, scale=K ** -0.5), _bf(N), _bf(M, N)
    from trlx_b200.ops import reference

    ref = reference.linear(x.float(), w.float(), b.float(), act, r.float())
    out = C.gemm(x, w, b, r, act)
    torch.testing.assert_close(out.float(), ref, atol=5e-2, rtol=3e-2)


def test_gemm_strided_input(C):
    torch.manual_seed(2)
    big = _bf(64, 3 * 768)
    x = big[:, 768:1536]  # row pitch 2304, offset 768 elements (16B aligned)
    w = _bf(256, 768, scale=0.03)
    out = C.gemm(x, w, out_f32=True)
    torch.testing.assert_close(out, x.float() @ w.float().t(), atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("M,V,K", [(128, 50257, 768), (77, 1000, 256), (300, 32000, 128)])
def test_lmhead_logprob(C, M, V, K):
    torch.manual_seed(3)
    h, w, b = _bf(M, K), _bf(V, K, scale=0.05), _bf(V, scale=0.1)
    labels = torch.randint(0, V, (M,), device="cuda")
    labels[0] = -1
    lse, lp, _, _ = C.lmhead(h, w, b, labels)
    logits = h.float() @ w.float().t() + b.float()
    ref_lse = torch.logsumexp(logits, -1)
    ref_lp = logits.gather(-1, labels.clamp_min(0)[:, None]).squeeze(-1) - ref_lse
    ref_lp[0] = 0
    torch.testing.assert_close(lse, ref_lse, atol=2e-3, rtol=1e-4)
    torch.testing.assert_close(lp, ref_lp, atol=3e-3, rtol=1e-3)


def test_lmhead_greedy_and_sampling(C):
    torch.manual_seed(4)
    M, V, K = 128, 5000, 256
    h, w = _bf(M, K), _bf(V, K, scale=0.2)
    logits = h.float() @ w.float().t()
    lse, _, tok, tlp = C.lmhead(h, w, None, None, True, 0.0, 0)
    assert torch.equal(tok, logits.argmax(-1))
    torch.testing.assert_close(tlp, logits.max(-1).values - torch.logsumexp(logits, -1), atol=3e-3, rtol=1e-3)
    # sampling: empirical frequencies of a peaked 8-way distribution
    V2 = 8
    w2 = torch.zeros(V2, K, device="cuda", dtype=torch.bfloat16)
    w2[:, 0] = torch.tensor([2.0, 1.0, 0.5, 0.0, -0.5, -1.0, -2.0, -3.0], dtype=torch.bfloat16)
    h2 = torch.zeros(4096, K, device="cuda", dtype=torch.bfloat16)
    h2[:, 0] = 1.0
    counts = torch.zeros(V2, device="cuda")
    for seed in range(8):
        _, _, t, l = C.lmhead(h2, w2, None, None, True, 1.0, 1234 + seed)
        counts += torch.bincount(t, minlength=V2).float()
    p = torch.softmax(w2[:, 0].float(), 0)
    freq = counts / counts.sum()
    assert (freq - p).abs().max() < 0.01, (freq, p)
    torch.testing.assert_close(l, torch.log(p)[t], atol=2e-3, rtol=1e-3)
    # suppression of a column while step < suppress_until
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    _, _, t, _ = C.lmhead(h2[:256], w2, None, None, True, 1.0, 7, step, 0, 5)
    assert (t != 0).all()


@pytest.mark.parametrize("B,H,Tq,Tk,d", [(4, 12, 56, 56, 64), (2, 3, 128, 128, 64), (3, 2, 17, 17, 128), (2, 4, 8, 24, 64),
                                         (1, 2, 1, 9, 32)])
@pytest.mark.parametrize("mode", ["causal", "bias", "bias_broadcast"])
def test_attention_short_forward_backward(C, B, H, Tq, Tk, d, mode):
    """One-CTA-per-(batch, head) attention vs an fp32 reference: strided q/k/v views of a fused QKV buffer, causal flag or an
    additive (finite-min) mask with left padding, rectangular score tiles; forward, dQ, dK, dV."""
    from trlx_b200 import ops

    torch.manual_seed(Tq * 131 + d)
    Tmax = max(Tq, Tk)
    qkv = (torch.randn(B, Tmax, 3 * H * d, device="cuda") * 0.7).to(torch.bfloat16).requires_grad_(True)
    q, k, v = qkv.split(H * d, dim=-1)
    q = q[:, :Tq].view(B, Tq, H, d).transpose(1, 2)
    k = k[:, :Tk].view(B, Tk, H, d).transpose(1, 2)
    v = v[:, :Tk].view(B, Tk, H, d).transpose(1, 2)
    scale = d ** -0.5
    i = torch.arange(Tq, device="cuda").view(Tq, 1) + (Tk - Tq)
    j = torch.arange(Tk, device="cuda").view(1, Tk)
    allowed = (j <= i).view(1, 1, Tq, Tk)
    if mode == "causal":
        bias = None
    else:
        pad = torch.zeros(B, Tk, dtype=torch.bool, device="cuda")
        for b in range(B):
            pad[b, : (b * 3) % max(Tk - 1, 1)] = True  # left padding of varying length
        ok = allowed & ~pad.view(B, 1, 1, Tk)
        neg = torch.finfo(torch.float32).min
        bias = torch.zeros(ok.shape, device="cuda").masked_fill(~ok, neg)
        if mode == "bias_broadcast" and Tq == 1:
            bias = bias[:, :, :1]
    assert C.attn_short_ok(Tq, Tk, d, True)
    from trlx_b200.ops import functional

    out = functional._ShortAttention.apply(q, k, v, bias, bias is None, scale)  # the autograd node itself (training default: SDPA)
    if Tq * Tk <= 32 * 32:  # the public entry point routes tiny no-grad calls to the same kernel (larger ones go to SDPA)
        with torch.no_grad():
            assert torch.equal(ops.attention(q, k, v, bias, causal=bias is None, scale=scale), out)
    g = (torch.randn(B, Tq, H * d, device="cuda") * 0.5).to(torch.bfloat16)
    (out.transpose(1, 2).reshape(B, Tq, H * d) * g).sum().backward()
    got_grad = qkv.grad.float().clone()
    qkv.grad = None

    qf, kf, vf = q.float(), k.float(), v.float()
    sc = qf @ kf.transpose(-1, -2) * scale + (bias if bias is not None else torch.zeros_like(allowed, dtype=torch.float32)
                                               .masked_fill(~allowed, float("-inf")))
    ref = torch.softmax(sc, dim=-1) @ vf
    (ref.transpose(1, 2).reshape(B, Tq, H * d) * g.float()).sum().backward()
    ref_grad = qkv.grad.float()
    assert (out.float() - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0)
    tol = 3e-2 * max(ref_grad.abs().max().item(), 1.0)
    assert (got_grad - ref_grad).abs().max().item() <= tol, ((got_grad - ref_grad).abs().max().item(), tol)


@pytest.mark.parametrize("T", [24, 56])
def test_public_attention_entry_point_matches_reference(T):
    """``ops.attention`` under no-grad: in-repo kernel for tiny score tiles, library SDPA above — same numbers either way."""
    from trlx_b200 import ops

    torch.manual_seed(T)
    q, k, v = ((torch.randn(3, 4, T, 64, device="cuda") * 0.7).to(torch.bfloat16) for _ in range(3))
    with torch.no_grad():
        out = ops.attention(q, k, v, None, causal=True)
    mask = torch.ones(T, T, dtype=torch.bool, device="cuda").tril()
    sc = (q.float() @ k.float().transpose(-1, -2) * 64 ** -0.5).masked_fill(~mask, float("-inf"))
    ref = torch.softmax(sc, -1) @ v.float()
    assert (out.float() - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize("B,H,Tq,Tk,d", [(4, 12, 56, 56, 64), (2, 3, 128, 128, 64), (3, 2, 17, 17, 128), (2, 4, 8, 24, 64),
                                         (2, 2, 100, 128, 128), (128, 12, 24, 24, 64)])
@pytest.mark.parametrize("mode", ["causal", "bias"])
def test_attention_tcgen05_forward(C, B, H, Tq, Tk, d, mode):
    """tcgen05 forward (S and O accumulate in TMEM, probabilities re-enter as a swizzled bf16 A operand) vs the fp32 reference
    and vs the CUDA-core kernel's saved row statistics."""
    torch.manual_seed(Tq + d)
    Tmax = max(Tq, Tk)
    qkv = (torch.randn(B, Tmax, 3 * H * d, device="cuda") * 0.7).to(torch.bfloat16)
    q, k, v = qkv.split(H * d, dim=-1)
    q = q[:, :Tq].view(B, Tq, H, d).transpose(1, 2)
    k = k[:, :Tk].view(B, Tk, H, d).transpose(1, 2)
    v = v[:, :Tk].view(B, Tk, H, d).transpose(1, 2)
    scale = d ** -0.5
    i = torch.arange(Tq, device="cuda").view(Tq, 1) + (Tk - Tq)
    j = torch.arange(Tk, device="cuda").view(1, Tk)
    allowed = (j <= i).view(1, 1, Tq, Tk)
    bias = None
    if mode == "bias":
        pad = torch.zeros(B, Tk, dtype=torch.bool, device="cuda")
        for b in range(B):
            pad[b, : (b * 3) % max(Tk - 1, 1)] = True
        ok = allowed & ~pad.view(B, 1, 1, Tk)
        bias = torch.zeros(ok.shape, device="cuda").masked_fill(~ok, torch.finfo(torch.float32).min)
    assert C.attn_tc_ok(Tq, Tk, d)
    o, stats = C.attn_tc_fwd(q, k, v, bias, bias is None, scale)
    o2, stats2 = C.attn_short_fwd(q, k, v, bias, bias is None, scale)
    sc = q.float() @ k.float().transpose(-1, -2) * scale + (bias if bias is not None else torch.zeros_like(allowed, dtype=torch.float32)
                                                             .masked_fill(~allowed, float("-inf")))
    ref = (torch.softmax(sc, dim=-1) @ v.float()).transpose(1, 2)  # [B, Tq, H, d]
    assert (o.float() - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0)
    assert (o.float() - o2.float()).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0)
    torch.testing.assert_close(stats[..., 0], stats2[..., 0], atol=2e-2, rtol=1e-2)       # row max
    torch.testing.assert_close(stats[..., 1], stats2[..., 1], atol=1e-3, rtol=2e-2)       # 1 / row sum (bf16-rounded probabilities)


@pytest.mark.parametrize("rms", [False, True])
@pytest.mark.parametrize("M,H", [(1792, 768), (37, 4096), (5, 64), (300, 1000)])
def test_training_norm_forward_backward(C, rms, M, H):
    """LayerNorm / RMSNorm with saved row statistics and the one-pass backward (dx, d-gamma, d-beta) vs fp32 autograd."""
    from trlx_b200 import ops

    torch.manual_seed(H + M)
    x = (_bf(M, H, scale=2.0) + 0.5).requires_grad_(True)
    w = (_bf(H) * 0.3 + 1.0).requires_grad_(True)
    b = None if rms else _bf(H).requires_grad_(True)
    g = _bf(M, H)
    assert ops.norm_ok(x, w, b)
    y = ops.layer_norm(x.view(1, M, H), w, b, 1e-5, rms).view(M, H)
    y.backward(g)
    got = [y.detach().float(), x.grad.float().clone(), w.grad.float().clone()] + ([] if rms else [b.grad.float().clone()])
    x.grad = w.grad = None
    xf, wf = x.float(), w.float()
    if rms:
        ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * wf
    else:
        b.grad = None
        ref = F.layer_norm(xf, (H,), wf, b.float(), 1e-5)
    ref.backward(g.float())
    want = [ref.detach(), x.grad.float(), w.grad.float()] + ([] if rms else [b.grad.float()])
    for name, a, e in zip(("y", "dx", "dgamma", "dbeta"), got, want):
        tol = 2e-2 * max(e.abs().max().item(), 1.0)
        assert (a - e).abs().max().item() <= tol, (name, (a - e).abs().max().item(), tol)


@pytest.mark.parametrize("M,N", [(1792, 3072), (1792, 768), (7, 66), (1280, 50304), (0, 64)])
def test_bias_gradient_column_sum(C, M, N):
    x = _bf(M, N + 2)[:, :N]  # strided view: row pitch != N
    ref = x.float().sum(0)
    for _ in range(3):  # repeated calls share one workspace that the kernel must leave zeroed
        out = C.colsum(x)
        assert out.shape == (N,)
        assert (out.float() - ref).abs().max().item() <= 1e-2 * max(ref.abs().max().item(), 1.0) + 1e-6


@pytest.mark.parametrize("rms", [False, True])
def test_norm(C, rms):
    torch.manual_seed(5)
    x, w, b = _bf(333, 768), _bf(768), _bf(768)
    y = C.norm(x, w, None if rms else b, 1e-5, rms)
    xf = x.float()
    if rms:
        ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
    else:
        ref = F.layer_norm(xf, (768,), w.float(), b.float(), 1e-5)
    torch.testing.assert_close(y.float(), ref, atol=3e-2, rtol=2e-2)


def test_embed_rowdot(C):
    torch.manual_seed(6)
    wte, wpe = _bf(1000, 256), _bf(64, 256)
    tok = torch.randint(0, 1000, (50,), device="cuda")
    pos = torch.randint(0, 60, (50,), device="cuda", dtype=torch.int32)
    x = C.embed(tok, pos, wte, wpe, 2)
    torch.testing.assert_close(x.float(), (wte[tok].float() + wpe[pos.long() + 2].float()), atol=2e-2, rtol=2e-2)
    w, b = _bf(256), _bf(1)
    out = C.rowdot(x, w, b)
    torch.testing.assert_close(out, x.float() @ w.float() + b.float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("nq,nkv,d,rot,interleaved", [(12, 12, 64, 0, False), (8, 2, 128, 128, False), (4, 4, 64, 32, True),
                                                      (4, 4, 96, 24, False), (2, 2, 256, 64, True)])
def test_decode_attention(C, nq, nkv, d, rot, interleaved):
    torch.manual_seed(7)
    from trlx_b200.nn.transformer import apply_rotary

    B, page, max_pages = 9, 16, 8
    lens = torch.randint(1, page * max_pages, (B,), device="cuda", dtype=torch.int32)
    lens[0] = 1
    num_pages = B * max_pages
    kc, vc = _bf(num_pages, page, nkv, d), _bf(num_pages, page, nkv, d)
    bt = torch.randperm(num_pages, device="cuda", dtype=torch.int32).view(B, max_pages).contiguous()
    qkv = _bf(B, (nq + 2 * nkv) * d)
    pos = (lens - 1).to(torch.int32)
    kc0, vc0 = kc.clone(), vc.clone()
    out = C.decode_attention(qkv, kc, vc, bt, lens, pos, nq, nkv, d, 1 / math.sqrt(d), rot, 10000.0, interleaved)
    # reference
    q, k, v = qkv.float().split([nq * d, nkv * d, nkv * d], -1)
    q, k, v = q.view(B, nq, 1, d), k.view(B, nkv, 1, d), v.view(B, nkv, 1, d)
    if rot:
        half = rot // 2
        inv = 1.0 / (10000.0 ** (torch.arange(half, device="cuda").float() / half))
        ang = pos.float()[:, None, None] * inv
        q = apply_rotary(q, ang.cos(), ang.sin(), rot, interleaved)
        k = apply_rotary(k, ang.cos(), ang.sin(), rot, interleaved)
    for b in range(B):
        L = int(lens[b])
        slots = torch.arange(L - 1, device="cuda")
        pg = bt[b, (slots // page).long()].long()
        kk = torch.cat([kc0[pg, (slots % page).long()].float(), k[b].transpose(0, 1).to(torch.bfloat16).float()], 0)  # [L, nkv, d]
        vv = torch.cat([vc0[pg, (slots % page).long()].float(), v[b].transpose(0, 1)], 0)
        kk = kk.repeat_interleave(nq // nkv, 1).transpose(0, 1)  # [nq, L, d]
        vv = vv.repeat_interleave(nq // nkv, 1).transpose(0, 1)
        att = torch.softmax((q[b] @ kk.transpose(1, 2)) / math.sqrt(d), -1)
        ref = (att @ vv).reshape(nq * d)
        torch.testing.assert_close(out[b].float(), ref, atol=3e-2, rtol=3e-2)
        # cache append
        last = L - 1
        torch.testing.assert_close(kc[bt[b, last // page].long(), last % page].float(), k[b, :, 0], atol=2e-2, rtol=2e-2)
        torch.testing.assert_close(vc[bt[b, last // page].long(), last % page].float(), v[b, :, 0], atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_logprob_from_logits(C, dtype):
    torch.manual_seed(8)
    logits = (torch.randn(4, 37, 5003, device="cuda") * 3).to(dtype)
    labels = torch.randint(0, 5003, (4, 37), device="cuda")
    lp, lse = C.logprob_from_logits(logits, labels)
    ref = torch.log_softmax(logits.float(), -1).gather(-1, labels[..., None]).squeeze(-1)
    torch.testing.assert_close(lp, ref, atol=2e-3, rtol=1e-3)


@pytest.mark.parametrize("width", [40, 17])
def test_gae_whiten(C, width):
    torch.manual_seed(9)
    from trlx_b200.ops import reference

    v, r = torch.randn(32, 40, device="cuda"), torch.randn(32, 40, device="cuda")
    adv, ret, stats = C.gae(v, r, width, 0.99, 0.95, True, True)
    radv, rret = reference.gae(v, r, width, 0.99, 0.95)
    var, mean = torch.var_mean(radv)
    torch.testing.assert_close(ret[:, :width], rret, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(adv[:, :width], (radv - mean) * torch.rsqrt(var + 1e-8), atol=1e-3, rtol=1e-3)
    assert int(stats[0].item()) == 32 * width


def test_ppo_loss_and_grads():
    torch.manual_seed(10)
    from trlx_b200 import ops
    from trlx_b200.ops import reference

    B, R = 32, 40
    lp = (torch.randn(B, R, device="cuda") * 0.3 - 2).requires_grad_()
    val = torch.randn(B, R, device="cuda").requires_grad_()
    old_lp = lp.detach() + torch.randn(B, R, device="cuda") * 0.3
    old_v = val.detach() + torch.randn(B, R, device="cuda") * 0.3
    adv, ret = torch.randn(B, R, device="cuda"), torch.randn(B, R, device="cuda")
    mask = (torch.arange(R, device="cuda")[None] < torch.randint(1, R + 1, (B, 1), device="cuda")).float()
    loss, stats = ops.ppo_loss(lp, val, old_lp, old_v, adv, ret, mask, 0.2, 0.2, 1.3)
    loss.backward()
    g_lp, g_v = lp.grad.clone(), val.grad.clone()
    lp.grad = val.grad = None
    rloss, rstats = reference.ppo_loss(lp, val, old_lp, old_v, adv, ret, mask, 0.2, 0.2, 1.3)
    rloss.backward()
    torch.testing.assert_close(loss, rloss, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(g_lp, lp.grad, atol=1e-5, rtol=1e-3)
    torch.testing.assert_close(g_v, val.grad, atol=1e-5, rtol=1e-3)
    for k, v in rstats.items():
        torch.testing.assert_close(stats[k].float(), torch.as_tensor(v, device="cuda").float(), atol=2e-3, rtol=2e-3, msg=k)


def test_rollout_rewards(C):
    """make_experience's fused post-processing (csrc/rl_ops.cu: rollout_rewards_kernel) vs the PyTorch formulation."""
    torch.manual_seed(11)
    from trlx_b200.ops import reference

    B, Q, R = 16, 5, 12
    T = Q + R
    start = Q - 1
    lp, rlp, val = (torch.randn(B, T - 1, device="cuda") * 0.3 for _ in range(3))
    mask = torch.ones(B, T, dtype=torch.long, device="cuda")
    for b in range(B):
        mask[b, : b % 3] = 0                       # left-padded prompt
        mask[b, Q + (b * 5) % (R + 1):] = 0        # response ends early (or never)
    scores = torch.randn(B, device="cuda")
    rw, lp_s, v_s, slen, kl = C.rollout_rewards(lp, rlp, val, mask, scores, start, 0.05)
    ref = reference.rollout_rewards(lp, rlp, val, mask, scores, start, 0.05)
    torch.testing.assert_close(rw, ref[0], atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(lp_s, ref[1])
    torch.testing.assert_close(v_s, ref[2])
    assert torch.equal(slen.long(), ref[3])
    torch.testing.assert_close(kl[0].float(), ref[4].float(), atol=1e-2, rtol=1e-3)


def test_adamw_flat(C):
    torch.manual_seed(12)
    from trlx_b200.ops import reference

    n = 100003
    master = torch.randn(n, device="cuda")
    param = master.to(torch.bfloat16)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    rw, rm, rv = master.clone(), m.clone(), v.clone()
    for step in range(1, 4):
        g = torch.randn(n, device="cuda").to(torch.bfloat16)
        hyper = torch.tensor([1e-2, 1 - 0.9 ** step, 1 - 0.95 ** step, 1.0], device="cuda")
        C.adamw_flat(param, master, g, m, v, 0.9, 0.95, 1e-8, 0.01, True, hyper)
        reference.adamw_step(rw, g.float(), rm, rv, step, 1e-2, 0.9, 0.95, 1e-8, 0.01)
    torch.testing.assert_close(master, rw, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(param.float(), rw, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("warp", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("decoupled", [True, False])
def test_adam8bit_kernel_matches_block_quantised_reference(C, dtype, decoupled, warp, monkeypatch):
    """One launch per tensor (decode, update, block maxima, re-encode) against the PyTorch implementation of the same
    code book, which itself tracks fp32 AdamW (tests/test_utils.py)."""
    from trlx_b200.parallel import optim as O

    monkeypatch.setenv("TRLX_B200_ADAM8BIT_WARP", "1" if warp else "0")  # warp-per-block variant (8 elements per lane)
    torch.manual_seed(3)
    n = 70001  # not a multiple of the 256-element block
    w0 = torch.randn(n, device="cuda")
    cls = O.AdamW8bit if decoupled else O.Adam8bit
    pk = torch.nn.Parameter(w0.to(dtype).clone())
    pr = torch.nn.Parameter(w0.to(dtype).clone())
    ok, orf = cls([pk], lr=1e-2, weight_decay=0.05), cls([pr], lr=1e-2, weight_decay=0.05)
    orf._kernel_ok = lambda p: False  # PyTorch path
    fp = torch.nn.Parameter(w0.clone())
    o32 = (torch.optim.AdamW if decoupled else torch.optim.Adam)([fp], lr=1e-2, weight_decay=0.05)
    from trlx_b200 import ops as _ops

    before = _ops.launch_count()
    for step in range(6):
        g = torch.randn(n, device="cuda") * (0.1 + step)
        pk.grad, pr.grad, fp.grad = g.to(dtype), g.to(dtype), g.to(dtype).float()
        ok.step(); orf.step(); o32.step()
    assert _ops.launch_count() - before >= 6  # the kernel ran (no silent fallback)
    st_k, st_r = ok.state[pk], orf.state[pr]
    assert st_k["m"].dtype == torch.int8 and st_k["v"].dtype == torch.uint8
    # codes may differ by one step where exp2f / log2f round differently: compare decoded moments and parameters
    mk = O._dequantize(st_k["m"], st_k["ms"], n, True, (n,))
    mr = O._dequantize(st_r["m"], st_r["ms"], n, True, (n,))
    vk = O._dequantize(st_k["v"], st_k["vs"], n, False, (n,))
    vr = O._dequantize(st_r["v"], st_r["vs"], n, False, (n,))
    assert ((mk - mr).abs() <= 0.2 * mr.abs() + 1e-6).float().mean() > 0.999
    assert ((vk - vr).abs() <= 0.1 * vr.abs() + 1e-9).float().mean() > 0.999
    tol = 2e-2 if dtype == torch.bfloat16 else 5e-3
    assert (pk.float() - pr.float()).abs().max() < tol
    assert (pk.float() - fp).abs().mean() < 1e-2  # and both stay close to fp32 Adam


def test_linear_autograd():
    torch.manual_seed(13)
    from trlx_b200 import ops

    x = _bf(4, 50, 768).requires_grad_()
    w = _bf(1024, 768, scale=0.03).requires_grad_()
    b = _bf(1024).requires_grad_()
    r = _bf(4, 50, 1024).requires_grad_()
    y = ops.linear(x, w, b, "none", r)
    y.float().pow(2).sum().backward()
    gx, gw, gb, gr = x.grad.clone(), w.grad.clone(), b.grad.clone(), r.grad.clone()
    for t in (x, w, b, r):
        t.grad = None
    yr = F.linear(x, w, b) + r
    yr.float().pow(2).sum().backward()
    torch.testing.assert_close(y.float(), yr.float(), atol=5e-2, rtol=3e-2)
    for a, bb in ((gx, x.grad), (gw, w.grad), (gb, b.grad), (gr, r.grad)):
        scale = bb.float().abs().max().item()
        torch.testing.assert_close(a.float(), bb.float(), atol=2e-2 * scale, rtol=5e-2)


def test_wgrad_accumulates_in_place_into_flat_grad_buffer():
    """``gradient_accumulation_fusion``: with the fused optimizer owning the gradients, the weight-gradient GEMM adds into
    ``.grad`` itself (no dW tensor, no AccumulateGrad add) — over several micro-batches the result equals autograd's."""
    from trlx_b200.ops import functional as Fn
    from trlx_b200.parallel.optim import FusedAdamW

    torch.manual_seed(5)
    lin = torch.nn.Linear(256, 384, bias=True).cuda().to(torch.bfloat16)
    ref = torch.nn.Linear(256, 384, bias=True).cuda().to(torch.bfloat16)
    ref.load_state_dict(lin.state_dict())
    opt = FusedAdamW(list(lin.parameters()), lr=1e-3, process_group=None).prepare()
    assert Fn.mark_inplace_wgrad(lin) == 1 and lin.weight._b200_inplace_ok
    fired = []
    lin.weight._b200_grad_sink = lambda: fired.append(1)
    flat_ptr = lin.weight.grad.data_ptr()
    xs = [(torch.randn(4, 96, 256, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(3)]
    for x in xs:
        xa = x.clone().requires_grad_(True)
        Fn.linear(xa, lin.weight, lin.bias).float().pow(2).sum().backward()
        xb = x.clone().requires_grad_(True)
        torch.nn.functional.linear(xb, ref.weight, ref.bias).float().pow(2).sum().backward()
        torch.testing.assert_close(xa.grad.float(), xb.grad.float(), atol=2e-1, rtol=2e-2)
    assert lin.weight.grad.data_ptr() == flat_ptr and len(fired) == 3  # accumulated in place, callback once per backward
    gw, gr = lin.weight.grad.float(), ref.weight.grad.float()
    assert (gw - gr).abs().max() <= 2e-2 * gr.abs().max() + 1e-2
    torch.testing.assert_close(lin.bias.grad.float(), ref.bias.grad.float(), atol=5e-1, rtol=3e-2)
    # a weight used twice in one graph: the callback fires after the second contribution only
    fired.clear()
    opt.zero_grad()
    x = xs[0].clone().requires_grad_(True)
    (Fn.linear(x, lin.weight, lin.bias).float().sum() + Fn.linear(x * 2, lin.weight, lin.bias).float().sum()).backward()
    assert len(fired) == 1
    expect = torch.ones(4 * 96, 384, device="cuda").t() @ (3 * xs[0].float().reshape(-1, 256))
    assert (lin.weight.grad.float() - expect).abs().max() <= 2e-2 * expect.abs().max() + 1e-2


def test_quant_rows_fp8(C):
    torch.manual_seed(2)
    x = (torch.randn(37, 3072, device="cuda") * torch.rand(37, 1, device="cuda") * 4).to(torch.bfloat16)
    x[5] = 0
    q, s = C.quant_rows(x)
    assert q.dtype == torch.uint8 and q.shape == x.shape and s.shape == (37,)
    amax = x.float().abs().amax(1)
    torch.testing.assert_close(s, amax.clamp_min(1e-12) / 448.0, rtol=1e-6, atol=0)
    back = q.view(torch.float8_e4m3fn).float() * s[:, None]
    err = (back - x.float()).abs()
    assert (err <= 0.0625 * x.float().abs() + s[:, None] * 2 ** -9 + 1e-9).all()  # e4m3: 3 mantissa bits
    assert back[5].abs().max() == 0


def test_lora_update_as_gemm_epilogue(monkeypatch):
    """LoRA on the CUDA path: frozen projection on the tcgen05 GEMM, all adapters of the projection as one skinny GEMM + one GEMM
    whose residual operand is the frozen output (K = R = 8 or 16).  Outputs and gradients against fp32 maths."""
    from trlx_b200.models.peft import LoRALinear

    torch.manual_seed(7)
    base = torch.nn.Linear(256, 768).cuda().to(torch.bfloat16)
    lin = LoRALinear(base, r=8, alpha=32.0, dropout=0.0)
    lin.add_adapter("q", (0, 256), "q_proj")
    lin.add_adapter("v", (512, 768), "v_proj")
    for k in lin.lora_B:
        torch.nn.init.normal_(lin.lora_B[k], std=0.05)
    x = (torch.randn(4, 96, 256, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    params = [x, lin.lora_A["q"], lin.lora_B["q"], lin.lora_A["v"], lin.lora_B["v"]]
    from trlx_b200 import ops as _ops

    before = _ops.launch_count()
    y = lin(x)
    assert _ops.launch_count() - before >= 3  # base GEMM, skinny GEMM, epilogue GEMM
    g = torch.autograd.grad(y.float().pow(2).mean(), params)
    # fp32 oracle
    xf = x.detach().float().requires_grad_(True)
    pf = [p.detach().float().requires_grad_(True) for p in params[1:]]
    yf = xf @ base.weight.float().t() + base.bias.float()
    upd_q = (xf @ pf[0].t()) @ pf[1].t() * lin.scaling
    upd_v = (xf @ pf[2].t()) @ pf[3].t() * lin.scaling
    yf = torch.cat([yf[..., :256] + upd_q, yf[..., 256:512], yf[..., 512:] + upd_v], -1)
    gf = torch.autograd.grad(yf.pow(2).mean(), [xf] + pf)
    assert (y.float() - yf).abs().max() < 3e-2 * yf.abs().max()
    for a, b in zip(g, gf):
        assert (a.float() - b).abs().max() <= 4e-2 * b.abs().max() + 1e-6, (a.shape, (a.float() - b).abs().max(), b.abs().max())
    monkeypatch.setenv("TRLX_B200_LORA_FUSED", "0")
    y0 = lin(x)
    assert (y0.float() - y.float()).abs().max() < 3e-2 * yf.abs().max()


def test_fused_logprob_autograd():
    torch.manual_seed(14)
    from trlx_b200 import ops

    M, V, K = 200, 3000, 256
    h = _bf(M, K).requires_grad_()
    w = _bf(V, K, scale=0.05).requires_grad_()
    labels = torch.randint(0, V, (M,), device="cuda")
    lp, _ = ops.fused_logprob(h, w, None, labels)
    (lp * torch.arange(M, device="cuda")).sum().backward()
    gh, gw = h.grad.clone(), w.grad.clone()
    h.grad = w.grad = None
    ref = torch.log_softmax(F.linear(h.float(), w.float()), -1).gather(-1, labels[:, None]).squeeze(-1)
    (ref * torch.arange(M, device="cuda")).sum().backward()
    torch.testing.assert_close(lp, ref, atol=3e-3, rtol=1e-3)
    torch.testing.assert_close(gh.float(), h.grad.float(), atol=2e-2 * h.grad.float().abs().max().item(), rtol=5e-2)
    torch.testing.assert_close(gw.float(), w.grad.float(), atol=2e-2 * w.grad.float().abs().max().item(), rtol=5e-2)


@pytest.mark.parametrize("top_k,top_p,temperature", [(0, 0.9, 1.0), (20, 1.0, 0.7), (50, 0.8, 1.3), (1, 1.0, 1.0)])
def test_sample_filtered_matches_hf_filtering(C, top_k, top_p, temperature):
    """temperature -> top-k -> top-p -> multinomial (csrc/decode_ops.cu: sample_filtered_kernel) vs the same chain written
    with sort / cumsum in PyTorch (HF ``TemperatureLogitsWarper`` / ``TopKLogitsWarper`` / ``TopPLogitsWarper``)."""
    torch.manual_seed(21)
    V, B = 1003, 16384
    row = torch.randn(V, device="cuda") * 2.0
    logits = torch.zeros(B, 1008, device="cuda")
    logits[:, :V] = row
    tok, lp = C.sample_filtered(logits, V, top_k, top_p, temperature, 1234)
    # reference keep-set and distribution
    scaled = row / temperature
    keep = torch.ones(V, dtype=torch.bool, device="cuda")
    if top_k > 0:
        keep &= scaled >= torch.topk(scaled, top_k).values[-1]
    masked = scaled.masked_fill(~keep, -float("inf"))
    if top_p < 1.0:
        sorted_logits, idx = torch.sort(masked, descending=False)
        cum = sorted_logits.softmax(-1).cumsum(-1)
        remove = cum <= (1 - top_p)
        remove[-1] = False
        keep &= ~torch.zeros(V, dtype=torch.bool, device="cuda").scatter(0, idx, remove)
        masked = scaled.masked_fill(~keep, -float("inf"))
    probs = masked.softmax(-1)
    assert keep[tok].all(), "a filtered-out token was drawn"
    torch.testing.assert_close(lp, torch.log_softmax(row, -1)[tok], atol=1e-4, rtol=1e-4)  # RAW log-prob (what PPO scores)
    freq = torch.bincount(tok, minlength=V).float() / B
    sigma = (probs * (1 - probs) / B).sqrt()
    assert ((freq - probs).abs() <= 5 * sigma + 1e-4).all(), (freq - probs).abs().max()


def test_sample_filtered_suppresses_eos_and_varies_with_step(C):
    torch.manual_seed(22)
    V, B = 517, 256
    logits = torch.randn(B, 520, device="cuda")
    logits[:, 7] = 50.0  # "EOS" dominates
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    tok0, _ = C.sample_filtered(logits, V, 0, 0.95, 1.0, 5, step, 7, 3)
    assert (tok0 != 7).all()  # suppressed while step < 3
    step.fill_(3)
    tok3, _ = C.sample_filtered(logits, V, 0, 0.95, 1.0, 5, step, 7, 3)
    assert (tok3 == 7).all()
    logits[:, 7] = 0.0
    a, _ = C.sample_filtered(logits, V, 40, 1.0, 1.0, 5, step)
    step.fill_(4)
    b, _ = C.sample_filtered(logits, V, 40, 1.0, 1.0, 5, step)
    assert (a != b).float().mean() > 0.5  # fresh noise per step


@pytest.mark.parametrize("two_qs,with_mask", [(True, False), (False, True)])
def test_ilql_sample_kernel(C, two_qs, with_mask):
    """log pi_beta + beta * (min Q - V) -> top-k -> softmax(T) sampling (csrc/decode_ops.cu: ilql_sample_kernel) vs the PyTorch
    formulation used by the model's own generate loop; temperature 0 must give the exact argmax."""
    from trlx_b200.models.modeling_ilql import topk_mask

    torch.manual_seed(31)
    V, B = 777, 8192
    row_logits = torch.randn(V, device="cuda") * 1.5
    q1r, q2r = torch.randn(V, device="cuda"), torch.randn(V, device="cuda")
    logits = torch.zeros(B, 784, device="cuda"); logits[:, :V] = row_logits
    q1 = torch.zeros(B, 784, device="cuda"); q1[:, :V] = q1r
    q2 = torch.zeros(B, 784, device="cuda"); q2[:, :V] = q2r
    vs = torch.full((B,), 0.3, device="cuda")
    beta, top_k, T = 2.0, 12, 0.8
    mask = last = None
    masked_logits = row_logits.clone()
    if with_mask:
        mask = torch.zeros(5, 600, dtype=torch.bool, device="cuda")  # covers only part of the vocabulary
        mask[3, ::2] = True
        last = torch.full((B,), 3, dtype=torch.long, device="cuda")
        masked_logits[:600][mask[3]] = -float("inf")
    tok = C.ilql_sample(logits, q1, q2 if two_qs else None, vs, V, beta, top_k, T, 99, None, None, mask, last)
    q = torch.minimum(q1r, q2r) if two_qs else q1r
    shifted = topk_mask((torch.log_softmax(masked_logits, -1) + beta * (q - 0.3))[None], top_k)[0]
    probs = torch.softmax(shifted / T, -1)
    assert (probs[tok] > 0).all()
    freq = torch.bincount(tok, minlength=V).float() / B
    sigma = (probs * (1 - probs) / B).sqrt()
    assert ((freq - probs).abs() <= 5 * sigma + 1e-4).all()
    greedy = C.ilql_sample(logits, q1, q2 if two_qs else None, vs, V, beta, top_k, 0.0, 99, None, None, mask, last)
    assert (greedy == shifted.argmax()).all()
