#!/usr/bin/env python
"""Blackwell-native instructions (tcgen05 / TMEM / TMA / mbarrier) of one translation unit of libtvc_b200.so from
`cuobjdump -sass`: count per kernel and, for one kernel, every such instruction in program order.
Usage: python tools/sass_blackwell.py <kernel-substring for the counts> <mangled-substring of the kernel to list> > profiles/x.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tvc_ai_b200", "libtvc_b200.so")
pat, listed = sys.argv[1], sys.argv[2]
OPS = ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "UTCATOMSWS", "UTCPMALLOC", "SYNCS")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, counts, lines = None, collections.OrderedDict(), []
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1)
        if pat in fn:
            counts[fn] = collections.Counter()
        continue
    if fn is None or pat not in fn:
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if not m:
        continue
    op = next((o for o in OPS if re.search(r"(^|\s)" + o + r"(\.|\s)", " " + m.group(2))), None)
    if op:
        counts[fn][op] += 1
        if listed in fn:
            lines.append(f"  /*{m.group(1)}*/  {m.group(2).strip()}")
print(f"# cuobjdump -sass of tvc_ai_b200/libtvc_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo); tools/sass_blackwell.py")
print("# Blackwell-native instructions: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UBLKCP = cp.async.bulk (TMA bulk copy),")
print("# UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS / UTCPMALLOC = TMEM allocation, SYNCS = mbarrier operations\n")
print("# count per kernel:")
for f, c in counts.items():
    if c:
        print(f"  {f}\n      " + "  ".join(f"{k} x{v}" for k, v in sorted(c.items())))
print(f"\n# every such instruction of the kernel whose name contains {listed}, in program order:")
print("\n".join(lines))
