"""GPU: split-K and tile choices for the backward GEMMs after the issue-loop fix (L2-hot and cold)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemm_tune as Tn
from bert_multimodal_transformer_b200 import _lib as L
T = 2400
A = L.EPI_ATOMIC_F32
shapes = []
for sp in (1, 2, 4, 8):
    shapes.append((f"wgrad ao  s{sp}", 768, 768, T, 1, 1, A, sp))
for sp in (1, 2, 3, 4):
    shapes.append((f"wgrad qkv s{sp}", 2304, 768, T, 1, 1, A, sp))
for sp in (1, 2):
    shapes.append((f"wgrad f2  s{sp}", 768, 3072, T, 1, 1, A, sp))
    shapes.append((f"wgrad f1  s{sp}", 3072, 768, T, 1, 1, A, sp))
shapes.append(("dgrad qkv K-MN ", T, 768, 2304, 0, 1, L.EPI_STORE_F32, 1))
shapes.append(("dgrad ao  K-MN ", T, 768, 768, 0, 1, L.EPI_STORE_F16, 1))
shapes.append(("dgrad f1  K-MN ", T, 768, 3072, 0, 1, L.EPI_STORE_F32, 1))
Tn.SHAPES = shapes
os.environ["MAGB200_GEMM_POLICY"] = "0"
cfgs = (5, 7, 10)
for hot in (True, False):
    print("== L2-hot" if hot else "== cold")
    table = {c: Tn.bench(c, hot) for c in cfgs}
    for i, s in enumerate(shapes):
        print(f"{s[0]:18s}" + "".join(f"  {table[c][i][1]:6.1f}us {table[c][i][2]:5.0f}TF" for c in cfgs))
print("columns: cfg 5 = 1-CTA 128x128 | cfg 7 = one-shot 256x128 pairs | cfg 10 = persistent 256x128 pairs")
