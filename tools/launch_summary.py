#!/usr/bin/env python
"""Share of the step from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`): picks the
cold-L2 steps of bench.py (256 MiB fill -> step_kernel_v2 -> close_kernel) and prints the median device time of the two
kernels.  Usage: python tools/launch_summary.py gpurun_out/r02_launches.csv > profiles/r02_launches_summary.txt"""
import csv
import statistics
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(r.get("Metric Unit", "us"), 1.0)
        rows.append((r["Kernel Name"], v))
step, close = [], []
for i in range(len(rows) - 2):
    a, b, c = rows[i], rows[i + 1], rows[i + 2]
    if "FillFunctor<unsigned char>" in a[0] and a[1] > 20 and "step_kernel_v2" in b[0] and "close_kernel" in c[0]:
        step.append(b), close.append(c)
if not step:
    sys.exit("no cold-L2 step found in the launch list")
# the headline workload comes first in bench.py; later legs (e.g. contract_r) launch other instantiations
keep = [k for k, s_ in enumerate(step) if s_[0] == step[0][0]]
step, close = [step[k] for k in keep], [close[k] for k in keep]
ms, mc = statistics.median(v for _, v in step), statistics.median(v for _, v in close)
print(f"cold-L2 steps found (256 MiB fill -> step_kernel_v2 -> close_kernel): {len(step)}")
print(f"{step[0][0].split('(')[0].replace('void ', ''):22s} median {ms:.2f} us  (min {min(v for _, v in step):.2f}, max {max(v for _, v in step):.2f})")
print(f"{close[0][0].split('(')[0].replace('void ', ''):22s} median {mc:.2f} us  (min {min(v for _, v in close):.2f}, max {max(v for _, v in close):.2f})")
print(f"share of the step: step kernel {100 * ms / (ms + mc):.1f} %, closing kernel {100 * mc / (ms + mc):.1f} %")
