"""GPU: time the tcgen05 GEMM on the model's shapes for each tile configuration and dump a pipeline trace."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bert_multimodal_transformer_b200 import _lib as L
lib = L.load()
T = 2400
SHAPES = [  # name, M, N, K, a_mn, b_mn, epi, splits
    ("ffn_in  KK gelu", T, 3072, 768, 0, 0, L.EPI_GELU_F16, 1),
    ("ffn_out KK f32 ", T, 768, 3072, 0, 0, L.EPI_STORE_F32, 1),
    ("qkv     KK f16 ", T, 2304, 768, 0, 0, L.EPI_STORE_F16, 1),
    ("attnout KK f32 ", T, 768, 768, 0, 0, L.EPI_STORE_F32, 1),
    ("dgrad f2 K-MN  ", T, 3072, 768, 0, 1, L.EPI_DGELU_F16, 1),
    ("dgrad f1 K-MN  ", T, 768, 3072, 0, 1, L.EPI_STORE_F32, 1),
    ("wgrad f2 MN-MN ", 768, 3072, T, 1, 1, L.EPI_ATOMIC_F32, 1),
    ("wgrad ao MN-MN4", 768, 768, T, 1, 1, L.EPI_ATOMIC_F32, 4),
    ("wgrad qkv MNMN2", 2304, 768, T, 1, 1, L.EPI_ATOMIC_F32, 2),
]
def bench(cfg, hot=False):
    lib.magb200_gemm_set_debug(None, cfg)
    res = []
    for name, M, N, K, amn, bmn, epi, splits in SHAPES:
        nbuf = 1 if hot else 10
        As = [(torch.randn((K, M) if amn else (M, K), device="cuda")).half() for _ in range(nbuf)]
        Bs = [(torch.randn((K, N) if bmn else (N, K), device="cuda") * 0.02).half() for _ in range(nbuf)]
        odt = torch.float32 if epi in (L.EPI_STORE_F32, L.EPI_ATOMIC_F32) else torch.float16
        Os = [torch.zeros(M, N, device="cuda", dtype=odt) for _ in range(nbuf)]
        O1 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
        aux = torch.randn(M, N, device="cuda").half() if epi == L.EPI_DGELU_F16 else None
        bias = torch.zeros(N, device="cuda")
        st = L.current_stream()
        def launch(i):
            j = i % nbuf
            L.check(lib.magb200_gemm(L.ptr(As[j]), L.ptr(Bs[j]), M, N, K, As[j].shape[1], Bs[j].shape[1], L.F16, L.F16, amn, bmn,
                                     epi, L.ptr(bias), L.ptr(Os[j]), L.ptr(O1) if epi == L.EPI_GELU_F16 else None,
                                     L.ptr(aux), N, 1.0, splits, 0, 0, st))
        for i in range(nbuf): launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for i in range(reps): launch(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        res.append((name, us, 2.0 * M * N * K / us / 1e6))
    return res

if __name__ == "__main__":
    for hot in (True,):
        print("== L2-hot (same buffers)" if hot else "== cold (rotating buffers > L2)")
        table = {}
        for cfg in (0, 7, 8):
            try:
                table[cfg] = bench(cfg, hot)
            except Exception as e:
                print("cfg", cfg, "failed:", e); table[cfg] = None
        for i, s in enumerate(SHAPES):
            row = f"{s[0]:18s}"
            for cfg in (0, 7, 8):
                row += "   n/a      " if table[cfg] is None else f"  {table[cfg][i][1]:6.1f}us {table[cfg][i][2]:5.0f}TF"
            print(row)
    print("columns: cfg 0 = 1-CTA 128x128 | cfg 7 = 2-CTA pair tiles 256x128 | cfg 8 = 2-CTA 256x256 where N >= 1024")

    # pipeline trace for attn_out shape, cfg 0 and cfg 2
    for cfg in (0,):
        M, N, K = T, 768, 768
        A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.02).half()
        O = torch.zeros(M, N, device="cuda")
        BN = 256 if cfg >= 3 else 128
        nct = ((N + BN - 1) // BN) * ((M + 127) // 128)
        tr = torch.zeros(nct * 64, dtype=torch.int64, device="cuda")
        lib.magb200_gemm_set_debug(None, cfg)
        for _ in range(3):
            L.check(lib.magb200_gemm(L.ptr(A), L.ptr(B), M, N, K, K, K, L.F16, L.F16, 0, 0, L.EPI_STORE_F32, None, L.ptr(O), None, None, N, 1.0, 1, 0, 0, L.current_stream()))
        torch.cuda.synchronize()
        lib.magb200_gemm_set_debug(L.ptr(tr), cfg)
        L.check(lib.magb200_gemm(L.ptr(A), L.ptr(B), M, N, K, K, K, L.F16, L.F16, 0, 0, L.EPI_STORE_F32, None, L.ptr(O), None, None, N, 1.0, 1, 0, 0, L.current_stream()))
        torch.cuda.synchronize()
        lib.magb200_gemm_set_debug(None, 0)
        t = tr.view(nct, 64).cpu()
        for c in (0, nct // 2, nct - 1):
            r = t[c]; t0 = int(r[0])
            d = lambda i: int(r[i]) - t0
            print(f"cfg{cfg} cta{c}: setup_done={d(1)} tma_issue={[d(2+i) for i in range(12)]}")
            print(f"      full_wait_done={[d(20+i) for i in range(12)]}")
            print(f"      mma_committed={[d(40+i) for i in range(12)]}")
            print(f"      epi_start={d(60)} chunk_ld_done={[d(52+i) for i in range(4)]} staged={d(56)} epi_end={d(61)} exit={d(62)}")
        starts = t[:, 0]; ends = t[:, 62]
        print(f"cfg{cfg}: CTA start spread {int(starts.max()-starts.min())} cyc, kernel span {int(ends.max()-starts.min())} cyc")
