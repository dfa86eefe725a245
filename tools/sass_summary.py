#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (no GPU needed):  python tools/sass_summary.py > profiles/r02_sass_summary.txt
Proof that the tcgen05 / TMEM / TMA paths are what the compiled kernels contain (B200_PROFILING.md's mnemonic list)."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dmdqn_b200", "libdmdqn_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMAPF", "UBLKPF", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "USETMAXREG", "FENCE.VIEW.ASYNC",
        "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "MUFU", "FFMA", "HMMA", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"dmdqn::\(anonymous namespace\)::|\(.*", "", kern).replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        total[kern] += 1
        for w in WANT:
            if m.group(1).startswith(w):
                counts[kern][w] += 1
print(f"# cuobjdump -sass {os.path.basename(LIB)} -- instruction counts per kernel (static, sm_100a)")
for k in sorted(total):
    c = counts[k]
    print(f"{k}: {total[k]} instructions; " + ", ".join(f"{w} {c[w]}" for w in WANT if c[w]))
