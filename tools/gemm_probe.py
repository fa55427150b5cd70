"""GPU probe for the tcgen05 GEMM (run on the B200 box): each case in its own process so a trap
in one variant cannot poison the others.

    python tools/gemm_probe.py list
    python tools/gemm_probe.py <case>
"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from bert_multimodal_transformer_b200 import _lib as L

DT = {"f16": (L.F16, torch.float16), "bf16": (L.BF16, torch.bfloat16), "tf32": (L.F32, torch.float32)}


def run(name, M, N, K, adt, bdt, a_mn, b_mn, epi="f32", splits=1, bias=False, n_split=0, k_start_hi=0, alpha=1.0):
    lib = L.load()
    L.check(lib.magb200_device_check())
    lib.magb200_gemm_set_debug(None, int(os.environ.get("MAGB200_GEMM_CFG", "0")))
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).to(DT[adt][1])
    B = torch.randn(N, K, device="cuda", generator=g).to(DT[bdt][1])
    Am = A.t().contiguous() if a_mn else A          # stored [K][M] when MN-major
    Bm = B.t().contiguous() if b_mn else B
    lda = Am.shape[1]
    ldb = Bm.shape[1]
    Af, Bf = A.double(), B.double()
    if adt == "tf32":
        # tensor cores read the top 19 bits: emulate by truncating the mantissa (round-toward-zero bound)
        pass
    ref = (Af @ Bf.t())
    if n_split:
        ref2 = Af[:, k_start_hi:] @ Bf[:, k_start_hi:].t()
        ref[:, n_split:] = ref2[:, n_split:]
    ref = ref * alpha
    bias_t = None
    if bias:
        bias_t = torch.randn(N, device="cuda", generator=g)
        ref = ref + bias_t.double()[None]
    aux = None
    out1 = None
    if epi == "f32":
        out0 = torch.full((M, N), float("nan"), device="cuda"); code = L.EPI_STORE_F32
    elif epi == "f32_add":
        aux = torch.randn(M, N, device="cuda", generator=g)
        ref = ref + aux.double()
        out0 = torch.full((M, N), float("nan"), device="cuda"); code = L.EPI_STORE_F32
    elif epi == "f16":
        out0 = torch.zeros(M, N, device="cuda", dtype=torch.float16); code = L.EPI_STORE_F16
    elif epi == "bf16":
        out0 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); code = L.EPI_STORE_BF16
    elif epi == "gelu":
        out0 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
        out1 = torch.zeros(M, N, device="cuda", dtype=torch.float16); code = L.EPI_GELU_F16
    elif epi == "dgelu":
        aux = torch.randn(M, N, device="cuda", generator=g).half()      # the saved derivative gelu'(u)
        ref = ref * aux.double()
        out0 = torch.zeros(M, N, device="cuda", dtype=torch.float16); code = L.EPI_DGELU_F16
    elif epi == "atomic":
        out0 = torch.ones(M, N, device="cuda"); ref = ref + 1.0; code = L.EPI_ATOMIC_F32
    elif epi == "f32_f16":
        out0 = torch.zeros(M, N, device="cuda"); out1 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
        code = L.EPI_STORE_F32_F16
    elif epi == "relu":
        out0 = torch.zeros(M, N, device="cuda"); ref = ref.clamp_min(0); code = L.EPI_RELU_F32
    else:
        raise ValueError(epi)
    rc = lib.magb200_gemm(L.ptr(Am), L.ptr(Bm), M, N, K, lda, ldb, DT[adt][0], DT[bdt][0], int(a_mn), int(b_mn), code,
                          L.ptr(bias_t), L.ptr(out0), L.ptr(out1), L.ptr(aux), N, alpha, splits, n_split, k_start_hi,
                          L.current_stream())
    L.check(rc)
    torch.cuda.synchronize()
    got = out0.double()
    if epi == "gelu":
        gp = out1.double()      # second output: gelu'(pre-activation)
        gref = 0.5 * (1 + torch.erf(ref / 2 ** 0.5)) + ref * torch.exp(-0.5 * ref * ref) / (2 * torch.pi) ** 0.5
        e1 = (gp - gref).abs().max().item()
        ref = 0.5 * ref * (1 + torch.erf(ref / 2 ** 0.5))
        print(f"   derivative output max err {e1:.3e}")
        if e1 > 2e-3:
            print("FAIL derivative output"); return False
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    tol = {"f32": 2e-3, "f32_add": 2e-3, "atomic": 2e-3, "relu": 2e-3, "f32_f16": 2e-3}.get(epi, 3e-2)
    if adt == "tf32":
        tol = max(tol, 5e-2)
    tol *= max(1.0, (K / 64) ** 0.5)
    mx = err.max().item()
    ok = bool(torch.isfinite(got).all()) and mx <= tol
    print(f"{'PASS' if ok else 'FAIL'} {name}: M={M} N={N} K={K} max_err={mx:.4e} (tol {tol:.2e}) ref_max={scale:.3f}")
    if not ok:
        bad = (err > tol)
        print("   bad fraction", bad.double().mean().item(), "nan", int((~torch.isfinite(got)).sum()))
        rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        print("   first bad rows", rows, "cols", cols)
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        # which (row%8, 16-col chunk) cells are wrong inside the first tile: swizzle diagnostics
        t = bad[:128, :128].double()
        if t.numel() == 128 * 128:
            print("   bad by row%8   ", t.view(16, 8, 128).mean(dim=(0, 2)).tolist())
            print("   bad by col//8%8", t.view(128, 2, 8, 8).mean(dim=(0, 1, 3)).tolist())
    return ok


CASES = {
    "kk_f16_small": lambda: run("kk_f16_small", 128, 128, 64, "f16", "f16", False, False),
    "kk_f16_k128": lambda: run("kk_f16_k128", 128, 128, 128, "f16", "f16", False, False),
    "kk_f16_multi": lambda: run("kk_f16_multi", 256, 256, 512, "f16", "f16", False, False),
    "kk_f16_bert": lambda: run("kk_f16_bert", 2400, 768, 768, "f16", "f16", False, False, bias=True),
    "kk_f16_qkv": lambda: run("kk_f16_qkv", 2400, 2304, 768, "f16", "f16", False, False, epi="f16", bias=True),
    "kk_f16_ffn": lambda: run("kk_f16_ffn", 2400, 3072, 768, "f16", "f16", False, False, epi="gelu", bias=True),
    "kk_f16_ffn2": lambda: run("kk_f16_ffn2", 2400, 768, 3072, "f16", "f16", False, False, epi="f32_add", bias=True),
    "kk_bf16": lambda: run("kk_bf16", 384, 256, 256, "bf16", "bf16", False, False),
    "amn_f16_small": lambda: run("amn_f16_small", 128, 128, 64, "f16", "f16", True, False),
    "bmn_f16_small": lambda: run("bmn_f16_small", 128, 128, 64, "f16", "f16", False, True),
    "mnmn_f16_small": lambda: run("mnmn_f16_small", 128, 128, 64, "f16", "f16", True, True),
    "amn_f16": lambda: run("amn_f16", 384, 256, 320, "f16", "f16", True, False),
    "bmn_f16": lambda: run("bmn_f16", 384, 256, 320, "f16", "f16", False, True),
    "dgrad_bmn": lambda: run("dgrad_bmn", 2400, 768, 3072, "f16", "f16", False, True, epi="f16"),
    "dgrad_dgelu": lambda: run("dgrad_dgelu", 2400, 3072, 768, "f16", "f16", False, True, epi="dgelu"),
    "wgrad_mnmn": lambda: run("wgrad_mnmn", 768, 768, 2400, "f16", "f16", True, True, epi="atomic", splits=4),
    "wgrad_ffn": lambda: run("wgrad_ffn", 3072, 768, 2400, "f16", "f16", True, True, epi="atomic", splits=1),
    "wgrad_pool": lambda: run("wgrad_pool", 768, 768, 48, "f16", "f16", True, True, epi="atomic", splits=1),
    "kk_tf32_small": lambda: run("kk_tf32_small", 128, 128, 32, "tf32", "tf32", False, False),
    "kk_tf32": lambda: run("kk_tf32", 384, 256, 256, "tf32", "tf32", False, False),
    "nsplit": lambda: run("nsplit", 256, 512, 896, "f16", "f16", False, False, n_split=256, k_start_hi=768),
    "epi_f32_f16": lambda: run("epi_f32_f16", 200, 160, 96, "f16", "f16", False, False, epi="f32_f16", bias=True),
    "epi_relu_alpha": lambda: run("epi_relu_alpha", 200, 160, 96, "f16", "f16", False, False, epi="relu", alpha=0.5),
}

if __name__ == "__main__":
    which = sys.argv[1]
    if which == "list":
        print(" ".join(CASES))
    else:
        ok = CASES[which]()
        sys.exit(0 if ok else 1)
