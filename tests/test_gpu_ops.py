"""-m gpu: every native operator through the C ABI (include/magb200.h) against a plain torch fp32
restatement of the same op (computed on the CPU) on seeded inputs.  Tolerances: fp32 row-wise kernels
1e-5; anything that passes through an fp16-operand tensor-core GEMM 5e-3 absolute on O(1) values."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bert_multimodal_transformer_b200 import _lib as L
from _gpu_util import keep_mask, lib, rel_err, stream

pytestmark = pytest.mark.gpu
H = 768


class Dev:
    """Keeps the device copies alive for the duration of a test: ctypes only sees raw pointers, so a
    temporary `.cuda()` tensor would be freed (and its memory reused) before the kernel runs."""

    def __init__(self):
        self.keep = []

    def __call__(self, t):
        d = t.detach().cuda().contiguous()
        self.keep.append(d)
        return d


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,epi", [
    (2400, 768, 768, 0, 0, "f32"), (2400, 2304, 768, 0, 0, "f16"), (304, 256, 320, 1, 0, "f32"),
    (300, 256, 320, 0, 1, "f32"), (768, 768, 2400, 1, 1, "atomic"), (48, 768, 768, 0, 0, "f32"),
    (129, 32, 8, 0, 0, "f32"),
    # reductions scheduled with stream-K work ranges (persistent pair kernel): the model's weight-gradient,
    # accumulate-into-residual dgrad and N = 768 forward shapes, a shape with fewer tiles than clusters, ragged M / K
    (3072, 768, 2400, 1, 1, "atomic"), (2304, 768, 2400, 1, 1, "atomic"), (2400, 768, 3072, 0, 1, "atomic"),
    (2400, 768, 3072, 0, 0, "atomic"), (2400, 768, 768, 0, 0, "atomic"), (300, 256, 200, 0, 0, "atomic"),
    (100, 128, 64, 0, 0, "atomic"),
])
def test_gemm(M, N, K, a_mn, b_mn, epi):
    g = _g(1)
    A = torch.randn(M, K, generator=g).half()
    B = torch.randn(N, K, generator=g).half()
    bias = torch.randn(N, generator=g)
    ref = A.double() @ B.double().t() + bias.double()
    Ad = (A.t().contiguous() if a_mn else A).cuda()
    Bd = (B.t().contiguous() if b_mn else B).cuda()
    bias_d = bias.cuda()
    splits = 1
    if epi == "f32":
        out, code = torch.full((M, N), float("nan"), device="cuda"), L.EPI_STORE_F32
    elif epi == "f16":
        out, code = torch.zeros(M, N, device="cuda", dtype=torch.float16), L.EPI_STORE_F16
    else:
        out, code, splits = torch.ones(M, N, device="cuda"), L.EPI_ATOMIC_F32, 4
        ref = ref + 1                                  # accumulates into the output; the bias is added once
    L.check(lib().magb200_gemm(L.ptr(Ad), L.ptr(Bd), M, N, K, Ad.shape[1], Bd.shape[1], L.F16, L.F16, a_mn, b_mn, code,
                               L.ptr(bias_d), L.ptr(out), None, None, N, 1.0, splits, 0, 0, stream()))
    torch.cuda.synchronize()
    tol = 2e-3 * math.sqrt(K / 64) if epi != "f16" else 0.1
    assert (out.double().cpu() - ref).abs().max().item() < tol


@pytest.mark.parametrize("M,N,K", [(2400, 3072, 768), (200, 256, 128)])
def test_gemm_gelu_epilogues(M, N, K):
    """EPI_GELU_F16 writes gelu(v) and (second output) gelu'(v); EPI_DGELU_F16 multiplies by that saved derivative:
    together they are BertIntermediate's activation forward/backward (transformers 3.0.2 ACT2FN["gelu"] = erf form)."""
    g = _g(3)
    A = (torch.randn(M, K, generator=g) * 0.5).half()
    B = (torch.randn(N, K, generator=g) * 0.08).half()
    bias = torch.randn(N, generator=g) * 0.1
    v = A.double() @ B.double().t() + bias.double()
    gelu = 0.5 * v * (1 + torch.erf(v / math.sqrt(2)))
    dgelu = 0.5 * (1 + torch.erf(v / math.sqrt(2))) + v * torch.exp(-0.5 * v * v) / math.sqrt(2 * math.pi)
    Ad, Bd, bias_d = A.cuda(), B.cuda(), bias.cuda()
    out0 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    out1 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_gemm(L.ptr(Ad), L.ptr(Bd), M, N, K, K, K, L.F16, L.F16, 0, 0, L.EPI_GELU_F16, L.ptr(bias_d),
                               L.ptr(out0), L.ptr(out1), None, N, 1.0, 1, 0, 0, stream()))
    torch.cuda.synchronize()
    assert (out0.double().cpu() - gelu).abs().max().item() < 4e-3          # fp16 outputs, |v| up to ~4
    assert (out1.double().cpu() - dgelu).abs().max().item() < 1.5e-3
    # without the second output the first is unchanged
    out0b = torch.zeros_like(out0)
    L.check(lib().magb200_gemm(L.ptr(Ad), L.ptr(Bd), M, N, K, K, K, L.F16, L.F16, 0, 0, L.EPI_GELU_F16, L.ptr(bias_d),
                               L.ptr(out0b), None, None, N, 1.0, 1, 0, 0, stream()))
    torch.cuda.synchronize()
    assert torch.equal(out0, out0b)
    # backward epilogue: dU = (dY W) * saved derivative   (B operand MN-major as in the dgrad GEMM)
    dY = (torch.randn(M, K, generator=g) * 0.5).half()
    W = (torch.randn(K, N, generator=g) * 0.08).half()          # [K][N]: B[n][k] stored MN-major
    ref = (dY.double() @ W.double()) * out1.double().cpu()
    du = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    dYd, Wd = dY.cuda(), W.cuda()
    L.check(lib().magb200_gemm(L.ptr(dYd), L.ptr(Wd), M, N, K, K, N, L.F16, L.F16, 0, 1, L.EPI_DGELU_F16, None, L.ptr(du),
                               None, L.ptr(out1), N, 1.0, 1, 0, 0, stream()))
    torch.cuda.synchronize()
    assert (du.double().cpu() - ref).abs().max().item() < 6e-3 * max(1.0, ref.abs().max().item())


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(128, 64, device="cuda", dtype=torch.float16)
    o = torch.zeros(128, 128, device="cuda")
    rc = lib().magb200_gemm(L.ptr(a), L.ptr(a), 128, 128, 64, 64, 64, L.F16, L.BF16, 0, 0, L.EPI_STORE_F32, None, L.ptr(o),
                            None, None, 128, 1.0, 1, 0, 0, stream())
    assert rc != 0 and b"same element type" in lib().magb200_last_error()
    rc = lib().magb200_gemm(L.ptr(a), L.ptr(a), 128, 100, 64, 64, 64, L.F16, L.F16, 0, 0, L.EPI_STORE_F32, None, L.ptr(o),
                            None, None, 128, 1.0, 1, 0, 0, stream())
    assert rc != 0


def test_dropout_mask_is_a_pure_function_with_the_right_rate():
    m1 = keep_mask((1000, H), 0.1, 1234, 7)
    m2 = keep_mask((1000, H), 0.1, 1234, 7)
    m3 = keep_mask((1000, H), 0.1, 1234, 8)
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    assert abs(m1.mean().item() - 0.9) < 3e-3
    assert keep_mask((10, H), 0.0, 1, 1).min().item() == 1.0


def test_dropout_masks_are_uniform_and_independent():
    """The mask generator (common.cuh: keyed integer hash per element, Philox under -DMAGB200_PHILOX) must behave like
    independent Bernoulli draws: the right rate for several p, no correlation between neighbouring elements, between
    rows, between sites and between seeds, and every column / row of a large mask close to the nominal rate."""
    n_rows = 4096
    for p in (0.1, 0.5):
        keep = 1.0 - p
        m = keep_mask((n_rows, H), p, 987654321, 18)
        n = m.numel()
        tol = 5.0 * math.sqrt(keep * p / n)                       # 5 sigma of the mean of n Bernoulli draws
        assert abs(m.mean().item() - keep) < tol
        flat = m.flatten()
        for lag in (1, 2, 3, 4, 7, 8, 32, H, H + 1, 50 * H):      # neighbours, float4 partners, next row, next sequence
            both = (flat[:-lag] * flat[lag:]).mean().item()
            assert abs(both - keep * keep) < 6.0 * math.sqrt(keep * keep * (1 - keep * keep) / (n - lag)), (p, lag, both)
        other_site = keep_mask((n_rows, H), p, 987654321, 19)
        other_seed = keep_mask((n_rows, H), p, 987654322, 18)
        for o in (other_site, other_seed):
            both = (m * o).mean().item()
            assert abs(both - keep * keep) < 6.0 * math.sqrt(keep * keep * (1 - keep * keep) / n), (p, both)
        # per-column and per-row rates: no stuck bit positions
        assert (m.mean(0) - keep).abs().max().item() < 6.0 * math.sqrt(keep * p / n_rows)
        assert (m.mean(1) - keep).abs().max().item() < 6.0 * math.sqrt(keep * p / H)


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_embed_ln_fwd_bwd(p):
    g = _g(2)
    T, S, V, A, vocab = 150, 50, 47, 74, 1000
    ids = torch.randint(0, vocab, (T,), generator=g)
    tt = torch.randint(0, 2, (T,), generator=g)
    word = torch.randn(vocab, H, generator=g, requires_grad=True)
    pos = torch.randn(512, H, generator=g, requires_grad=True)
    typ = torch.randn(2, H, generator=g, requires_grad=True)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(H, generator=g)).requires_grad_(True)
    vis, aco = torch.randn(T, V, generator=g), torch.randn(T, A, generator=g)
    seed, site = 99, 1
    keep = keep_mask((T, H), p, seed, site).cpu() if p > 0 else torch.ones(T, H)
    ref = F.layer_norm(word[ids] + pos[torch.arange(T) % S] + typ[tt], (H,), gamma, beta, 1e-12) * keep / (1 - p)
    dout = torch.randn(T, H, generator=g)
    ref.backward(dout)
    c = Dev()
    emb = torch.empty(T, H, device="cuda")
    amag = torch.empty(T, 896, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_embed_ln_fwd(L.ptr(c(ids)), L.ptr(c(tt)), L.ptr(c(vis)), L.ptr(c(aco)), L.ptr(c(word)),
                                       L.ptr(c(pos)), L.ptr(c(typ)), L.ptr(c(gamma)), L.ptr(c(beta)), 1e-12, p, seed,
                                       site, V, A, 896, T, S, vocab, L.ptr(emb), L.ptr(amag), stream()))
    assert (emb.cpu() - ref.detach()).abs().max().item() < 2e-5
    am = amag.float().cpu()
    assert (am[:, :H] - ref.detach()).abs().max().item() < 5e-3
    assert (am[:, H:H + V] - vis).abs().max().item() < 4e-3 and (am[:, H + V:H + V + A] - aco).abs().max().item() < 4e-3
    assert am[:, H + V + A:].abs().max().item() == 0
    dw, dp, dt = torch.zeros_like(word).cuda(), torch.zeros_like(pos).cuda(), torch.zeros_like(typ).cuda()
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    L.check(lib().magb200_embed_ln_bwd(L.ptr(c(dout)), L.ptr(c(ids)), L.ptr(c(tt)), L.ptr(c(word)), L.ptr(c(pos)),
                                       L.ptr(c(typ)), L.ptr(c(gamma)), 1e-12, p, seed, site, T, S, vocab, 1.0,
                                       L.ptr(dw), L.ptr(dp), L.ptr(dt), L.ptr(dg), L.ptr(db), stream()))
    for got, want in ((dw, word.grad), (dp, pos.grad), (dt, typ.grad), (dg, gamma.grad), (db, beta.grad)):
        assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_dropout_residual_ln_fwd_bwd(p):
    g = _g(3)
    T = 333
    pre = torch.randn(T, H, generator=g, requires_grad=True)
    res = torch.randn(T, H, generator=g, requires_grad=True)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(H, generator=g)).requires_grad_(True)
    seed, site = 5, 21
    keep = keep_mask((T, H), p, seed, site).cpu() if p > 0 else torch.ones(T, H)
    ref = F.layer_norm(pre * keep / (1 - p) + res, (H,), gamma, beta, 1e-12)
    dy = torch.randn(T, H, generator=g)
    ref.backward(dy)
    c = Dev()
    y32 = torch.empty(T, H, device="cuda")
    y16 = torch.empty(T, H, device="cuda", dtype=torch.float16)
    rstd = torch.empty(T, device="cuda")
    L.check(lib().magb200_dropout_residual_ln_fwd(L.ptr(c(pre)), L.ptr(c(res)), L.ptr(c(gamma)), L.ptr(c(beta)), 1e-12,
                                                  p, seed, site, T, L.ptr(y32), L.ptr(y16), L.ptr(rstd), stream()))
    assert (y32.cpu() - ref.detach()).abs().max().item() < 2e-5
    assert (y16.float().cpu() - ref.detach()).abs().max().item() < 5e-3
    scale = 64.0   # loss scale carried by activation gradients; parameter gradients come back unscaled
    dz = torch.empty(T, H, device="cuda")
    dpre = torch.empty(T, H, device="cuda", dtype=torch.float16)
    dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    L.check(lib().magb200_ln_bwd(L.ptr(c(dy * scale)), L.ptr(y32), L.ptr(c(gamma)), L.ptr(c(beta)), L.ptr(rstd), p, seed,
                                 site, T, 1.0 / scale, L.ptr(dz), L.ptr(dpre), L.ptr(dg), L.ptr(db), L.ptr(dbias),
                                 stream()))
    assert rel_err(dz / scale, res.grad) < 1e-4
    assert rel_err(dpre.float() / scale, pre.grad) < 2e-3
    assert rel_err(dg, gamma.grad) < 1e-4 and rel_err(db, beta.grad) < 1e-4
    assert rel_err(dbias, pre.grad.sum(0)) < 1e-3


@pytest.mark.parametrize("S,p,mma", [(50, 0.0, True), (50, 0.1, True), (17, 0.0, True), (64, 0.1, True), (1, 0.0, True),
                                     (50, 0.1, False), (128, 0.0, False),
                                     # S > 64 with lse: the flash-style kernels (attention_flash.cu)
                                     (65, 0.0, True), (100, 0.1, True), (128, 0.1, True), (200, 0.0, True),
                                     (333, 0.1, True), (512, 0.1, True)])
def test_sdpa_fwd_bwd(S, p, mma):
    g = _g(4)
    B, heads, d = 3, 12, 64
    qkv = (torch.randn(B * S, 3 * heads * d, generator=g) * 0.7).half()
    mask = torch.ones(B, S, dtype=torch.int64)
    mask[1, S // 2:] = 0
    mask[2, min(3, S - 1):] = 0
    seed, site = 77, 16
    x = qkv.float().view(B, S, 3, heads, d).permute(2, 0, 3, 1, 4).requires_grad_(True)   # [3,B,h,S,d]
    q, k, v = x[0], x[1], x[2]
    s = q @ k.transpose(-1, -2) / 8.0 + ((1.0 - mask[:, None, None, :].float()) * -10000.0)
    pr = F.softmax(s, dim=-1)
    if p > 0:
        pr = pr * keep_mask((B, heads, S, S), p, seed, site).cpu() / (1 - p)
    ctx_ref = (pr @ v).permute(0, 2, 1, 3).reshape(B * S, heads * d)
    dctx = torch.randn(B * S, heads * d, generator=g).half()
    ctx_ref.backward(dctx.float())
    dqkv_ref = x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * heads * d)
    ctx = torch.empty(B * S, heads * d, device="cuda", dtype=torch.float16)
    qkv_d, mask_d, dctx_d = qkv.cuda(), mask.cuda(), dctx.cuda()
    # lse + ctx present and S <= 64 -> warp-MMA kernels; lse = NULL -> general kernels
    lse = torch.zeros(B * heads * S, device="cuda") if mma else None
    L.check(lib().magb200_sdpa_fwd(L.ptr(qkv_d), L.ptr(mask_d), B, S, heads, d, p, seed, site, L.ptr(ctx), L.ptr(lse),
                                   stream()))
    assert (ctx.float().cpu() - ctx_ref.detach()).abs().max().item() < 4e-3
    if mma:
        lse_ref = torch.logsumexp(s.detach(), dim=-1).reshape(-1)
        assert (lse.cpu() - lse_ref).abs().max().item() < 2e-3
    dqkv = torch.zeros(B * S, 3 * heads * d, device="cuda", dtype=torch.float16)
    dbias = torch.zeros(3 * heads * d, device="cuda")
    ws_bytes = lib().magb200_sdpa_workspace_bytes(B, S, heads)
    assert ws_bytes == B * heads * S * 4
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda") if mma else None
    L.check(lib().magb200_sdpa_bwd(L.ptr(qkv_d), L.ptr(dctx_d), L.ptr(ctx) if mma else None, L.ptr(lse), L.ptr(mask_d), B,
                                   S, heads, d, p, seed, site, 0.5, L.ptr(dqkv), L.ptr(dbias), L.ptr(ws),
                                   ws_bytes if mma else 0, stream()))
    assert rel_err(dqkv.float(), dqkv_ref) < 6e-3
    assert rel_err(dbias, 0.5 * dqkv_ref.sum(0)) < 1e-2


def test_sdpa_without_mask_and_range_errors():
    B, S, heads, d = 1, 8, 12, 64
    qkv = torch.randn(B * S, 3 * heads * d).half().cuda()
    ctx = torch.empty(B * S, heads * d, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_sdpa_fwd(L.ptr(qkv), None, B, S, heads, d, 0.0, 0, 0, L.ptr(ctx), None, stream()))
    assert torch.isfinite(ctx).all()
    assert lib().magb200_sdpa_fwd(L.ptr(qkv), None, B, 513, heads, d, 0.0, 0, 0, L.ptr(ctx), None, stream()) != 0
    assert lib().magb200_sdpa_fwd(L.ptr(qkv), None, B, S, heads, 32, 0.0, 0, 0, L.ptr(ctx), None, stream()) != 0


@pytest.mark.parametrize("M,K,p", [(2400, 768, 0.0), (2400, 3072, 0.1), (2400, 768, 0.1), (300, 768, 0.0), (1, 3072, 0.1),
                                   (16384, 768, 0.1)])
def test_linear_with_layernorm_epilogue(M, K, p):
    """magb200_linear_ln_fwd (Linear + dropout + residual + LayerNorm in one tcgen05 kernel, row statistics exchanged
    across a 2 x 3 cluster) against torch fp32 on the same fp16 operands with the kernels' dropout mask, and against the
    two-kernel path it replaces (same mask: both use element index m * 768 + n at the same site)."""
    g = _g(21)
    A = (torch.randn(M, K, generator=g) * 0.5).half()
    W = (torch.randn(H, K, generator=g) * 0.05).half()
    bias = torch.randn(H, generator=g) * 0.1
    res = torch.randn(M, H, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(H, generator=g), 0.1 * torch.randn(H, generator=g)
    seed, site = 555, 17
    keep = keep_mask((M, H), p, seed, site).cpu() if p > 0 else torch.ones(M, H)
    z = (A.float() @ W.float().t() + bias) * keep / (1 - p) + res
    ref = F.layer_norm(z, (H,), gamma, beta, 1e-12)
    rstd_ref = 1.0 / torch.sqrt(z.var(dim=1, unbiased=False) + 1e-12)
    c = Dev()
    y32 = torch.full((M, H), float("nan"), device="cuda")
    y16 = torch.zeros(M, H, device="cuda", dtype=torch.float16)
    rstd = torch.zeros(M, device="cuda")
    L.check(lib().magb200_linear_ln_fwd(L.ptr(c(A)), K, L.ptr(c(W)), M, K, L.ptr(c(bias)), L.ptr(c(res)), L.ptr(c(gamma)),
                                        L.ptr(c(beta)), 1e-12, p, seed, site, L.ptr(y32), L.ptr(y16), L.ptr(rstd), stream()))
    torch.cuda.synchronize()
    assert (y32.cpu() - ref).abs().max().item() < 2e-3 * math.sqrt(K / 768)
    assert (y16.float() - y32).abs().max().item() < 4e-3
    assert ((rstd.cpu() - rstd_ref).abs() / rstd_ref).max().item() < 1e-3
    # the path it replaces
    pre = torch.empty(M, H, device="cuda")
    L.check(lib().magb200_gemm(L.ptr(c.keep[0]), L.ptr(c.keep[1]), M, H, K, K, K, L.F16, L.F16, 0, 0, L.EPI_STORE_F32,
                               L.ptr(c.keep[2]), L.ptr(pre), None, None, H, 1.0, 1, 0, 0, stream()))
    y32b = torch.empty(M, H, device="cuda")
    L.check(lib().magb200_dropout_residual_ln_fwd(L.ptr(pre), L.ptr(c.keep[3]), L.ptr(c.keep[4]), L.ptr(c.keep[5]), 1e-12, p,
                                                  seed, site, M, L.ptr(y32b), None, None, stream()))
    torch.cuda.synchronize()
    assert (y32 - y32b).abs().max().item() < 1e-4


@pytest.mark.parametrize("T,S,p_emb,p_mag,gates", [(2400, 50, 0.0, 0.0, True), (2400, 50, 0.1, 0.5, True), (150, 50, 0.0, 0.0, False),
                                                 (1, 1, 0.0, 0.0, True), (20000, 500, 0.1, 0.5, True)])
def test_embed_mag_fused_kernel_equals_the_three_kernel_path(T, S, p_emb, p_mag, gates):
    """K1 (magb200_embed_mag_fwd: embeddings + MAG as one warp-specialised cooperative kernel) against the path it
    replaces (magb200_embed_ln_fwd -> magb200_mag_fwd: packed-gate GEMM + tail kernel) on the same inputs, seeds and
    dropout sites: same fp16 GEMM operands and fp32 accumulation, so only the summation order may differ.  T = 20000
    needs several tiles per CTA (the persistent tile loop); gates = False is the inference form (no gate tensor)."""
    g = _g(12)
    V, A, vocab, KA = 47, 74, 3000, 896
    ids = torch.randint(0, vocab, (T,), generator=g)
    tt = torch.randint(0, 2, (T,), generator=g)
    vis, aco = torch.randn(T, V, generator=g), torch.randn(T, A, generator=g)
    vis[::7] = 0
    aco[::7] = 0                                   # rows with no modality signal: h_m == 0 when the biases are zero
    word = torch.randn(vocab, H, generator=g) * 0.5
    pos = torch.randn(512, H, generator=g) * 0.1
    typ = torch.randn(2, H, generator=g) * 0.1
    eg, eb = 1 + 0.1 * torch.randn(H, generator=g), 0.1 * torch.randn(H, generator=g)
    mg, mb = 1 + 0.1 * torch.randn(H, generator=g), 0.1 * torch.randn(H, generator=g)
    W_hv, W_ha = torch.randn(H, V + H, generator=g) * 0.02, torch.randn(H, A + H, generator=g) * 0.02
    W_v, W_a = torch.randn(H, V, generator=g) * 0.02, torch.randn(H, A, generator=g) * 0.02
    bias4 = torch.randn(4 * H, generator=g) * 0.02
    bias4[2 * H:] = 0
    seed = 4242
    c = Dev()
    d = {k: c(v) for k, v in dict(ids=ids, tt=tt, vis=vis, aco=aco, word=word, pos=pos, typ=typ, eg=eg, eb=eb, mg=mg, mb=mb,
                                  W_hv=W_hv, W_ha=W_ha, W_v=W_v, W_a=W_a, bias4=bias4).items()}
    packed = torch.empty(4 * H * KA, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_mag_pack_weights(L.ptr(d["W_hv"]), L.ptr(d["W_ha"]), L.ptr(d["W_v"]), L.ptr(d["W_a"]), V, A, KA,
                                           L.ptr(packed), stream()))
    # the three-kernel path
    emb_a = torch.empty(T, H, device="cuda")
    amag_a = torch.empty(T, KA, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_embed_ln_fwd(L.ptr(d["ids"]), L.ptr(d["tt"]), L.ptr(d["vis"]), L.ptr(d["aco"]), L.ptr(d["word"]),
                                       L.ptr(d["pos"]), L.ptr(d["typ"]), L.ptr(d["eg"]), L.ptr(d["eb"]), 1e-12, p_emb, seed,
                                       1, V, A, KA, T, S, vocab, L.ptr(emb_a), L.ptr(amag_a), stream()))
    amag_tmp = torch.empty(T, KA, device="cuda", dtype=torch.float16)
    gates_a = torch.empty(T, 4 * H, device="cuda")
    out_a = torch.empty(T, H, device="cuda")
    L.check(lib().magb200_mag_fwd(L.ptr(emb_a), L.ptr(d["vis"]), L.ptr(d["aco"]), L.ptr(packed), L.ptr(d["bias4"]),
                                  L.ptr(d["mg"]), L.ptr(d["mb"]), 1.0, 1e-5, p_mag, seed, 2, V, A, KA, T, L.ptr(amag_tmp),
                                  L.ptr(gates_a), L.ptr(out_a), None, stream()))
    # K1
    emb_b = torch.empty(T, H, device="cuda")
    amag_b = torch.empty(T, KA, device="cuda", dtype=torch.float16)
    gates_b = torch.full((T, 4 * H), float("nan"), device="cuda") if gates else None
    out_b = torch.empty(T, H, device="cuda")
    out16_b = torch.empty(T, H, device="cuda", dtype=torch.float16)
    nscr = lib().magb200_embed_mag_scratch_bytes(T)
    scratch = torch.empty(nscr, dtype=torch.uint8, device="cuda")
    L.check(lib().magb200_embed_mag_fwd(L.ptr(d["ids"]), L.ptr(d["tt"]), L.ptr(d["vis"]), L.ptr(d["aco"]), L.ptr(d["word"]),
                                        L.ptr(d["pos"]), L.ptr(d["typ"]), L.ptr(d["eg"]), L.ptr(d["eb"]), 1e-12, p_emb, seed, 1,
                                        V, A, KA, T, S, vocab, L.ptr(packed), L.ptr(d["bias4"]), L.ptr(d["mg"]), L.ptr(d["mb"]),
                                        1.0, 1e-5, p_mag, 2, L.ptr(emb_b), L.ptr(amag_b), L.ptr(gates_b), L.ptr(scratch), nscr,
                                        L.ptr(out_b), L.ptr(out16_b), stream()))
    torch.cuda.synchronize()
    assert torch.equal(emb_a, emb_b) and torch.equal(amag_a, amag_b)              # phase 1 is the same arithmetic
    if gates:
        assert (gates_a - gates_b).abs().max().item() < 1e-4
    assert (out_a - out_b).abs().max().item() < 2e-4
    assert (out16_b.float() - out_b).abs().max().item() < 4e-3
    assert torch.isfinite(out_b).all()


@pytest.mark.parametrize("name", ["mag_mosi", "mag_mosi_zero_rows_beta1e-3", "mag_mosei"])
def test_mag_op_against_reference_golden(golden_dir, name):
    """BASELINE.json configs[0]: MAG(768, V, A) bs=4 seq=50 against outputs of the verbatim reference.
    The reference tolerance is 1e-5 on CPU fp32; the tensor-core path rounds GEMM operands to fp16, so the
    stated tolerance here is 5e-3 absolute on LayerNorm-ed (unit-scale) outputs."""
    import os
    from oracle import inputs as gen, weights
    from bert_multimodal_transformer_b200.modeling import MAG
    import contextlib, io
    gld = np.load(os.path.join(golden_dir, name + ".npz"))
    V, A, wseed, iseed, zero_rows, reflike = [int(x) for x in gld["args"]]
    with contextlib.redirect_stdout(io.StringIO()):
        m = MAG(768, float(gld["beta_shift"]), 0.5, visual_dim=V, acoustic_dim=A)
    m.load_state_dict(weights.mag_state_dict("", 768, V, A, wseed, bool(reflike)))
    m = m.cuda().eval()
    text, vis, aco = gen.mag_case(4, 50, 768, V, A, iseed, bool(zero_rows))
    text = text.cuda().requires_grad_(True)
    out = m(text, vis.cuda(), aco.cuda())
    assert out.shape == (4, 50, 768)
    assert np.abs(out.detach().cpu().numpy() - gld["out"]).max() < 5e-3
    (out * torch.from_numpy(gld["probe"]).cuda()).sum().backward()
    assert rel_err(text.grad, torch.from_numpy(gld["dtext"])) < 1e-2
    for k, p in m.named_parameters():
        ref_norm = float(gld["gnorm/" + k])
        assert abs(p.grad.norm().item() - ref_norm) <= 2e-2 * ref_norm + 1e-6, k


@pytest.mark.parametrize("name", ["mag_mosi", "mag_mosi_zero_rows_beta1e-3", "mag_mosei"])
def test_mag_exact_forward_meets_the_reference_tolerance(golden_dir, name):
    """BASELINE.json configs[0] at its own tolerance (1e-5 on MAG.forward): split-TF32 gate GEMMs."""
    import os
    import contextlib, io
    from oracle import inputs as gen, weights
    from bert_multimodal_transformer_b200.modeling import MAG
    gld = np.load(os.path.join(golden_dir, name + ".npz"))
    V, A, wseed, iseed, zero_rows, reflike = [int(x) for x in gld["args"]]
    with contextlib.redirect_stdout(io.StringIO()):
        m = MAG(768, float(gld["beta_shift"]), 0.5, visual_dim=V, acoustic_dim=A)
    m.load_state_dict(weights.mag_state_dict("", 768, V, A, wseed, bool(reflike)))
    m = m.cuda().eval()
    m.exact_forward = True
    text, vis, aco = gen.mag_case(4, 50, 768, V, A, iseed, bool(zero_rows))
    with torch.no_grad():
        out = m(text.cuda(), vis.cuda(), aco.cuda())
    assert np.abs(out.cpu().numpy() - gld["out"]).max() < 1e-5
    # with gradients enabled the exact output is kept and backward still works
    text_g = text.cuda().requires_grad_(True)
    out2 = m(text_g, vis.cuda(), aco.cuda())
    assert np.abs(out2.detach().cpu().numpy() - gld["out"]).max() < 1e-5
    (out2 * torch.from_numpy(gld["probe"]).cuda()).sum().backward()
    assert rel_err(text_g.grad, torch.from_numpy(gld["dtext"])) < 1e-2


def test_mag_train_mode_dropout_matches_injected_mask():
    import contextlib, io
    from oracle import inputs as gen, restatement as R, weights
    from bert_multimodal_transformer_b200.modeling import MAG
    with contextlib.redirect_stdout(io.StringIO()):
        m = MAG(768, 1.0, 0.5, visual_dim=47, acoustic_dim=74)
    sd = weights.mag_state_dict("", seed=4)
    m.load_state_dict(sd)
    m = m.cuda().train()
    text, vis, aco = gen.mag_case(2, 50, seed=5)
    torch.manual_seed(123)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    torch.manual_seed(123)
    out = m(text.cuda(), vis.cuda(), aco.cuda())
    keep = keep_mask((100, 768), 0.5, seed, 2).cpu().view(2, 50, 768)
    ref = R.mag_forward(sd, "", text, vis, aco, 1.0, drop=lambda site, x: x * keep / 0.5)
    assert (out.detach().cpu() - ref).abs().max().item() < 1e-2


def test_cls_head_mse_and_colsum():
    g = _g(6)
    B, Lb = 48, 1
    pre = torch.randn(B, H, generator=g, requires_grad=True)
    Wc = (torch.randn(Lb, H, generator=g) * 0.05).requires_grad_(True)
    bc = torch.randn(Lb, generator=g).requires_grad_(True)
    labels = torch.randn(B, generator=g)
    p, seed, site = 0.1, 31, 3
    keep = keep_mask((B, H), p, seed, site).cpu()
    pooled_ref = torch.tanh(pre)
    logits_ref = F.linear(pooled_ref * keep / (1 - p), Wc, bc)
    loss_ref = F.mse_loss(logits_ref.view(-1), labels)
    loss_ref.backward()
    c = Dev()
    pooled, logits = torch.empty(B, H, device="cuda"), torch.empty(B, Lb, device="cuda")
    L.check(lib().magb200_cls_fwd(L.ptr(c(pre)), L.ptr(c(Wc)), L.ptr(c(bc)), B, Lb, p, seed, site, L.ptr(pooled),
                                  L.ptr(logits), stream()))
    assert (logits.cpu() - logits_ref.detach()).abs().max().item() < 1e-5
    loss, dlog = torch.empty(1, device="cuda"), torch.empty(B, Lb, device="cuda")
    scale = 256.0
    L.check(lib().magb200_mse_loss(L.ptr(logits), L.ptr(c(labels)), B, scale, L.ptr(loss), L.ptr(dlog), stream()))
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    dpre = torch.empty(B, H, device="cuda", dtype=torch.float16)
    dWc, dbc = torch.zeros(Lb, H, device="cuda"), torch.zeros(Lb, device="cuda")
    L.check(lib().magb200_cls_bwd(L.ptr(dlog), L.ptr(pooled), L.ptr(c(Wc)), B, Lb, p, seed, site, 1.0 / scale, L.ptr(dpre),
                                  L.ptr(dWc), L.ptr(dbc), stream()))
    assert rel_err(dWc, Wc.grad) < 1e-4 and rel_err(dbc, bc.grad) < 1e-4
    assert rel_err(dpre.float() / scale, pre.grad) < 2e-3
    cs = torch.zeros(H, device="cuda")
    L.check(lib().magb200_colsum_f16(L.ptr(dpre), B, H, 1.0 / scale, L.ptr(cs), stream()))
    assert rel_err(cs, pre.grad.sum(0)) < 5e-3


def test_adamw_matches_transformers_3_0_2_semantics():
    """AdamW.step of transformers 3.0.2 (the optimizer built at multimodal_driver.py:345): bias-corrected
    Adam update, THEN decoupled decay p -= lr*wd*p on the updated value; eps 1e-6."""
    g = _g(8)
    n = 256 * 40
    p0, g0 = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.01
    decay = (torch.arange(n // 256) % 2).to(torch.uint8)
    lr, b1, b2, eps, wd = 1e-3, 0.9, 0.999, 1e-6, 0.01
    p, m, v = p0.clone().double(), torch.zeros(n).double(), torch.zeros(n).double()
    pc, gc = p0.clone().cuda(), g0.clone().cuda()
    mc, vc = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    p16 = torch.zeros(n, device="cuda", dtype=torch.float16)
    dflag = decay.repeat_interleave(256).double()
    decay_d = decay.cuda()
    for step in (1, 2, 3):
        gd = g0.double()
        m = b1 * m + (1 - b1) * gd
        v = b2 * v + (1 - b2) * gd * gd
        step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        p = p - step_size * m / (v.sqrt() + eps)
        p = p - lr * wd * dflag * p
        gc.copy_(g0 * 4.0)   # arrives scaled by 4 (e.g. summed over 4 ranks); grad_unscale undoes it
        L.check(lib().magb200_adamw_step(L.ptr(pc), L.ptr(gc), L.ptr(mc), L.ptr(vc), L.ptr(p16), L.ptr(decay_d), n,
                                         lr, b1, b2, eps, wd, step, 0.25, 1, None, stream()))
        assert gc.abs().max().item() == 0.0
    assert (pc.double().cpu() - p).abs().max().item() < 1e-5
    assert (p16.float().cpu() - pc.cpu()).abs().max().item() < 2e-3
    # device-resident hyper-parameters override the by-value ones (CUDA-graph path)
    hyper = torch.tensor([0.0, 0.0], device="cuda")
    before = pc.clone()
    gc.copy_(g0)
    L.check(lib().magb200_adamw_step(L.ptr(pc), L.ptr(gc), L.ptr(mc), L.ptr(vc), None, L.ptr(decay_d), n, lr, b1, b2,
                                     eps, wd, 4, 1.0, 0, L.ptr(hyper), stream()))
    assert torch.equal(pc, before)


def _ptr_array(*tensors):
    import ctypes
    return (ctypes.c_void_p * 8)(*([ctypes.c_void_p(t.data_ptr()) for t in tensors] + [None] * (8 - len(tensors))))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 3])
def test_peer_optimizer_kernel_equals_allreduce_then_adamw(world):
    """csrc/peer_optim.cu with the 'peers' being `world` buffers on ONE GPU (the kernel only sees pointers): summing the
    gradients of all ranks in the kernel + AdamW + storing the parameters into every rank's buffer must equal
    magb200_adamw_step on the pre-summed gradient -- dense, and with the embedding table split into the early pass over
    untouched rows and the late pass over touched rows; the clean-up range clears gradients and recasts the mirror."""
    g = _g(21)
    rows, H = 40, 768                      # a 'table' of 40 rows in front, 256 * 30 further elements behind it
    n = rows * H + 256 * 30
    p0 = torch.randn(n, generator=g).cuda()
    grads = [(torch.randn(n, generator=g) * 0.01).cuda() for _ in range(world)]
    touched = torch.zeros(rows, dtype=torch.bool)
    touched[[1, 5, 6, 17, 39]] = True
    for gr in grads:                       # untouched rows carry no gradient on any rank
        gr[:rows * H].view(rows, H)[~touched.cuda()] = 0
    decay = (torch.arange(n // 256) % 3 != 0).to(torch.uint8).cuda()
    lr, b1, b2, eps, wd = 1e-3, 0.9, 0.999, 1e-6, 0.01
    m0, v0 = torch.rand(n, generator=g).cuda() * 0.01, torch.rand(n, generator=g).cuda() * 1e-4

    # expected: all-reduce (sum in rank order), then the single-GPU kernel with grad_unscale = 1 / world
    pe, me, ve = p0.clone(), m0.clone(), v0.clone()
    gsum = torch.zeros(n, device="cuda")
    for gr in grads:
        gsum += gr
    e16 = torch.zeros(n, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_adamw_step(L.ptr(pe), L.ptr(gsum), L.ptr(me), L.ptr(ve), L.ptr(e16), L.ptr(decay), n, lr, b1, b2, eps,
                                     wd, 2, 1.0 / world, 1, None, stream()))

    def run(sparse):
        params = [p0.clone() for _ in range(world)]
        gs = [gr.clone() for gr in grads]
        m, v = m0.clone(), v0.clone()
        p16 = torch.zeros(n, device="cuda", dtype=torch.float16)
        # 'rank 0' owns everything (one call = one rank's work); its parameter copy is params[0]
        gp, pp = _ptr_array(*gs), _ptr_array(*params)
        ids = torch.tensor([1, 5, 6, 17, 39, 5, 1], dtype=torch.int64, device="cuda")
        marks = torch.zeros(rows, dtype=torch.int32, device="cuda")
        counter = torch.zeros(1, dtype=torch.int32, device="cuda")
        any_rows = torch.zeros(rows, dtype=torch.uint8, device="cuda")

        def step(owned, clean, mode):
            L.check(lib().magb200_adamw_peer_step(gp, pp, world, 0, L.ptr(m), L.ptr(v), L.ptr(p16), L.ptr(decay), owned[0],
                                                  owned[1], clean[0], clean[1], lr, b1, b2, eps, wd, 2, 1.0 / world, None, 0,
                                                  L.ptr(any_rows) if mode else None, mode, 0, rows if mode else 0,
                                                  H if mode else 0, 3, 0, 0, None, None, stream()))
        if sparse:
            L.check(lib().magb200_peer_mark_rows(L.ptr(ids), ids.numel(), rows, L.ptr(marks), L.ptr(counter), stream()))
            L.check(lib().magb200_peer_combine_marks(_ptr_array(*([marks] * world)), world, L.ptr(counter), rows,
                                                     L.ptr(any_rows), stream()))
            assert torch.equal(any_rows.bool().cpu(), touched) and int(counter.item()) == 1
            step((0, n), (0, 0), 1)
            # the early pass must not have touched anything that still waits for its gradient
            assert torch.equal(params[0][rows * H:], p0[rows * H:])
            assert torch.equal(params[0][:rows * H].view(rows, H)[touched.cuda()], p0[:rows * H].view(rows, H)[touched.cuda()])
            step((0, n), (0, 0), 2)
            step((0, 0), (0, n), 2)
        else:
            step((0, n), (0, 0), 0)
            step((0, 0), (0, n), 0)
        torch.cuda.synchronize()
        return params, gs, m, v, p16

    for sparse in (False, True):
        params, gs, m, v, p16 = run(sparse)
        for q in range(world):
            assert torch.equal(params[q], params[0])                 # replicas are copies of one computation
        assert (params[0] - pe).abs().max().item() < 2e-6, sparse
        assert (m - me).abs().max().item() < 1e-7 and (v - ve).abs().max().item() < 1e-9
        assert gs[0].abs().max().item() == 0.0                       # the calling rank's gradients are cleared
        if world > 1:
            assert gs[1].abs().max().item() > 0.0                    # ... and only its own
        tail = slice(rows * H, n)
        assert (p16[tail].float() - e16[tail].float()).abs().max().item() < 4e-3
        if not sparse:                                               # (nothing reads the table's fp16 mirror)
            assert (p16.float() - e16.float()).abs().max().item() < 4e-3
