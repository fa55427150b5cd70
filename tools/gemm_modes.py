"""GPU: pipeline traces of the 1-CTA GEMM in measurement modes: 0 normal, 1 MMA-issue only, 2 operand delivery only,
for 3 / 4 stages (cfg 0 / 1) and 128x256 (cfg 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bert_multimodal_transformer_b200 import _lib as L
lib = L.load()
M, N, K = 2400, 768, 3072
A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.02).half()
O = torch.zeros(M, N, device="cuda")
def go(cfgcode, label, BN=128):
    nct = ((N + BN - 1) // BN) * ((M + 127) // 128)
    tr = torch.zeros(nct * 64, dtype=torch.int64, device="cuda")
    def launch():
        L.check(lib.magb200_gemm(L.ptr(A), L.ptr(B), M, N, K, K, K, L.F16, L.F16, 0, 0, L.EPI_STORE_F32, None, L.ptr(O), None, None, N, 1.0, 1, 0, 0, L.current_stream()))
    lib.magb200_gemm_set_debug(None, cfgcode)
    for _ in range(3): launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): launch()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    lib.magb200_gemm_set_debug(L.ptr(tr), cfgcode)
    launch(); torch.cuda.synchronize()
    lib.magb200_gemm_set_debug(None, 0)
    r = tr.view(nct, 64).cpu()[nct // 2]; t0 = int(r[0])
    fw = [int(r[20 + i]) - t0 for i in range(16)]; mc = [int(r[40 + i]) - t0 for i in range(16)]; ti = [int(r[2 + i]) - t0 for i in range(16)]
    per = (fw[15] - fw[3]) / 12.0
    print(f"{label:44s} kernel {us:6.1f} us  k-block period {per:6.1f} cyc   tma_issue d={[ti[i+1]-ti[i] for i in range(3,9)]} full d={[fw[i+1]-fw[i] for i in range(3,9)]} issue={[mc[i]-fw[i] for i in range(3,7)]}")
go(0, "cfg0 3 stages normal")
go(100, "cfg0 3 stages MMA-issue only")
go(200, "cfg0 3 stages delivery only")
go(1, "cfg1 4 stages normal")
go(201, "cfg1 4 stages delivery only")
go(4, "cfg4 128x256 4 stages normal", 256)
go(104, "cfg4 128x256 MMA-issue only", 256)
go(204, "cfg4 128x256 delivery only", 256)
