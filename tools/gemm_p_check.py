"""GPU: correctness of the persistent pair-tile GEMM (cfg 9 / 10) on the probe cases, in one process, then timing of
the model's shapes for cfg 0 (current policy) vs 9 vs 10."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemm_probe as P
from bert_multimodal_transformer_b200 import _lib as L

P.CASES["kk_f16_big"] = lambda: P.run("kk_f16_big", 4096, 3072, 512, "f16", "f16", False, False, epi="f16", bias=True)
P.CASES["wgrad_ffn_s2"] = lambda: P.run("wgrad_ffn_s2", 3072, 768, 2400, "f16", "f16", True, True, epi="atomic", splits=2)
P.CASES["wgrad_f2_s2"] = lambda: P.run("wgrad_f2_s2", 768, 3072, 2400, "f16", "f16", True, True, epi="atomic", splits=2)
P.CASES["amn_wide"] = lambda: P.run("amn_wide", 640, 1024, 320, "f16", "f16", True, False, epi="f32_f16", bias=True)
names = ["kk_f16_small", "kk_f16_multi", "kk_f16_bert", "kk_f16_qkv", "kk_f16_ffn", "kk_f16_ffn2", "kk_bf16", "amn_f16",
         "bmn_f16", "mnmn_f16_small", "dgrad_bmn", "dgrad_dgelu", "wgrad_mnmn", "wgrad_ffn", "wgrad_pool", "kk_f16_big",
         "wgrad_ffn_s2", "wgrad_f2_s2", "amn_wide"]
bad = 0
for cfg in (9, 10):
    os.environ["MAGB200_GEMM_CFG"] = str(cfg)
    print(f"==== cfg {cfg}")
    for n in names:
        ok = P.CASES[n]()
        bad += 0 if ok else 1
print("FAILED CASES:", bad)
if bad:
    sys.exit(1)

import gemm_tune as Tn   # prints its own table for (0, 7, 8); then ours
print("== persistent kernel, L2-hot and cold")
for hot in (True, False):
    table = {c: Tn.bench(c, hot) for c in (0, 9, 10)}
    print("hot" if hot else "cold")
    for i, s in enumerate(Tn.SHAPES):
        print(f"{s[0]:18s}" + "".join(f"  {table[c][i][1]:6.1f}us {table[c][i][2]:5.0f}TF" for c in (0, 9, 10)))
print("columns: cfg 0 = current policy | cfg 9 = persistent 256x256 (N%256==0, N>=1024) else 256x128 | cfg 10 = persistent 256x128")
