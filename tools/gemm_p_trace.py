"""GPU: pipeline trace and measurement modes of the persistent pair-tile GEMM (cfg 9/10; +100 = MMA issue without
waiting for operands, +200 = operand delivery only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bert_multimodal_transformer_b200 import _lib as L
lib = L.load()
T = 2400
SHAPES = [("qkv KK f16", T, 2304, 768, 0, 0, L.EPI_STORE_F16, 1), ("ffn_in KK gelu", T, 3072, 768, 0, 0, L.EPI_GELU_F16, 1),
          ("ffn_out KK f32", T, 768, 3072, 0, 0, L.EPI_STORE_F32, 1), ("wgrad f2 MNMN", 768, 3072, T, 1, 1, L.EPI_ATOMIC_F32, 2)]
for name, M, N, K, amn, bmn, epi, splits in SHAPES:
    A = torch.randn((K, M) if amn else (M, K), device="cuda").half()
    B = (torch.randn((K, N) if bmn else (N, K), device="cuda") * 0.02).half()
    odt = torch.float32 if epi in (L.EPI_STORE_F32, L.EPI_ATOMIC_F32) else torch.float16
    O = torch.zeros(M, N, device="cuda", dtype=odt)
    O1 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    bias = torch.zeros(N, device="cuda")
    tr = torch.zeros(64 * 160, dtype=torch.int64, device="cuda")
    def launch():
        L.check(lib.magb200_gemm(L.ptr(A), L.ptr(B), M, N, K, A.shape[1], B.shape[1], L.F16, L.F16, amn, bmn, epi, L.ptr(bias),
                                 L.ptr(O), L.ptr(O1) if epi == L.EPI_GELU_F16 else None, None, N, 1.0, splits, 0, 0,
                                 L.current_stream()))
    for cfg in [int(c) for c in os.environ.get("TRACE_CFGS", "9,209").split(",")]:
        lib.magb200_gemm_set_debug(None, cfg)
        for _ in range(3): launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): launch()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        tr.zero_()
        lib.magb200_gemm_set_debug(L.ptr(tr), cfg)
        launch(); torch.cuda.synchronize()
        lib.magb200_gemm_set_debug(None, 0)
        t = tr.view(160, 64).cpu()
        used = int((t[:, 0] != 0).sum())
        print(f"== {name} cfg {cfg}: {us:.1f} us, {used} CTAs")
        for c in (0, 1, used - 2):
            r = t[c]; z = int(r[0])
            d = lambda i: int(r[i]) - z if int(r[i]) else -1
            tma = [d(2 + i) for i in range(12)]
            full = [d(14 + i) for i in range(12)]
            com = [d(26 + i) for i in range(12)]
            dif = lambda v: [v[i + 1] - v[i] for i in range(len(v) - 1) if v[i] >= 0 and v[i + 1] >= 0]
            print(f"  cta{c}: setup={d(1)} tma_issue={tma[:14]}\n        full_done={full[:14]}\n        d_full={dif(full)}\n"
                  f"        epi(tile0 start,end; tile1; tile2)={[d(38 + i) for i in range(6)]} exit={d(63)}\n"
                  f"        tile0 sub-block0 [entry, acc in regs, staged, stored]={[d(44 + i) for i in range(4)]} sub-block1={[d(48 + i) for i in range(4)]}")
        starts = t[:used, 0]; ends = t[:used, 63]
        print(f"  CTA start spread {int(starts.max() - starts.min())} cyc, kernel span {int(ends.max() - starts.min())} cyc")
