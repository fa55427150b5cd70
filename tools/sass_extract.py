#!/usr/bin/env python
"""cuobjdump -sass bert_multimodal_transformer_b200/libmagb200.so | python tools/sass_extract.py > profiles/rNN_sass_extract.txt

Counts, per kernel, the SASS mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load, UTMAREDG = TMA reduce, UTCBAR = tcgen05.commit,
HMMA = mma.sync (legacy tensor path), LDSM = ldmatrix, LDGSTS = cp.async, UCGABAR = cluster barrier,
LDGMC = multimem.ld_reduce (NVSwitch multicast reducing load)."""
import collections
import re
import subprocess
import sys

txt = sys.stdin.read()
funcs = re.split(r"\n\s*Function : ", txt)
pat = re.compile(r"\b(UTCHMMA[A-Z0-9.]*|UTCQMMA[A-Z0-9.]*|UTMALDG[A-Z0-9.]*|UTMAREDG[A-Z0-9.]*|UTMASTG[A-Z0-9.]*|UBLKCP[A-Z0-9.]*|"
                 r"UBLKRED[A-Z0-9.]*|LDTM[A-Z0-9]*|STTM[A-Z0-9]*|UTCBAR[A-Z0-9.]*|HMMA[A-Z0-9.]*|LDSM[A-Z0-9.]*|LDGSTS[A-Z0-9.]*|"
                 r"UCGABAR_[A-Z]*|REDG?\.E[A-Z0-9.]*|LDGMC[A-Z0-9.x]*)")
rows = []
for f in funcs[1:]:
    name = f.split("\n")[0].strip()
    c = collections.Counter(m.group(1).rstrip(".") for m in pat.finditer(f))
    n = len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/", f, re.M))
    if c:
        rows.append((name, n, c))
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.strip().split("\n")
print(__doc__.strip().replace("\n", "\n# ").join(["# ", ""]))
print()
for (name, n, c), d in sorted(zip(rows, dem), key=lambda x: x[1]):
    d = d.replace("(anonymous namespace)::", "").replace("magb200::", "")
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"{d[:100]:100s} instr={n:6d}  " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
