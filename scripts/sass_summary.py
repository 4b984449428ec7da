#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) tcgen05 / TMEM / TMA use in the shipped library:
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UTMAPF (TMA L2 prefetch), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), HMMA (mma.sync), LDGSTS (cp.async), LDSM (ldmatrix).   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "accelerating-t2i-ar-with-sjd_b200" / "libsjd_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
names = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "LDGSTS", "LDSM", "ATOMG", "REDUX"]
kern = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for n in names:
            if op.startswith(n):
                kern[cur][n] += 1
dem = subprocess.run(["cu++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib.name}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print(f"{'kernel':72s} {'instrs':>7s} " + " ".join(f"{n:>7s}" for n in names))
for (k, c), d in zip(kern.items(), dem):
    short = re.sub(r"\(.*", "", d).replace("void ", "")[:72]
    print(f"{short:72s} {c['_total']:7d} " + " ".join(f"{c[n]:7d}" for n in names))
