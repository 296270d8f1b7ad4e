#!/usr/bin/env python
"""Write profiles/sass_blackwell_r02.txt: per kernel of liborbx.so the counts of the Blackwell-specific / characteristic SASS mnemonics
(`cuobjdump -sass`) and the first lines that carry the TMA / tcgen05 / mbarrier instructions.   usage: python tools/sass_evidence.py [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dani_slam_b200", "liborbx.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_blackwell_r02.txt")
MARK = ("UTMALDG", "UTMASTG", "UBLKCP", "UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "ELECT", "FENCE.VIEW.ASYNC",
        "VIMNMX3", "IDP", "POPC", "ATOMS.POPC", "REDUX", "SHFL.UP")
SHOW = ("UTMALDG", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "SYNCS.ARRIVE")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        kernels[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
out = ["SASS evidence: cuobjdump -sass dani_slam_b200/liborbx.so (all cubins sm_100a), final round-2 build (tools/sass_evidence.py).",
       "Per kernel: counts of the Blackwell-specific / characteristic mnemonics, then the first lines that carry the TMA / tcgen05 / mbarrier instructions.",
       "  UTMALDG = cp.async.bulk.tensor (TMA load), UTCIMMA = tcgen05.mma.kind::i8, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS.* = mbarrier ops,",
       "  VIMNMX3 = packed 3-input min/max (DPX), IDP = DP2A/DP4A.", ""]
for name, lines in kernels.items():
    cnt = collections.Counter()
    shown = []
    for l in lines:
        parts = l.split("*/", 1)[1].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        op = op.rstrip(";")
        if any(op.startswith(k) for k in MARK):
            key = ".".join(op.split(".")[:3]) if op.startswith(("SYNCS", "UTMALDG", "VIMNMX3", "IDP", "ATOMS")) else op.split(".")[0] if op.startswith("POPC") else op
            cnt[key] += 1
            if any(op.startswith(k) for k in SHOW) and len(shown) < 6:
                shown.append("      " + l.strip())
    if not cnt:
        continue
    short = re.sub(r"\(.*", "", demangle(name).replace("(int)", "").replace("(bool)", "")).replace("<unnamed>::", "").replace("void ", "")
    out.append("== " + short)
    out.append("   " + ", ".join(f"{k} x{v}" for k, v in cnt.most_common()))
    out += shown
    out.append("")
open(OUT, "w").write("\n".join(out))
print(f"{len(kernels)} kernels, {OUT}")
