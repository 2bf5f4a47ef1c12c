"""Stall samples and executed instructions per CUDA source line of one kernel of an .ncu-rep (needs -lineinfo and
--import-source on).  Usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "step_kernel_v2"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0])
path = fn = hdr = None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and fn and pat in fn and r[2] == "-":   # per-source-line summary rows
        d = dict(zip(hdr, r))
        try:
            key = (path, int(r[0]), r[1].strip()[:90])
            agg[key][0] += int(d["# Samples"])
            agg[key][1] += int(d["Instructions Executed"])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
totx = sum(v[1] for v in agg.values()) or 1
print(f"{pat}: {tot} samples, {totx} warp-instructions attributed to source lines")
for (p, line, text), (sm, ex) in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
    print(f"{100 * sm / tot:5.1f} % samples {100 * ex / totx:5.1f} % instr  {p}:{line}  {text}")
