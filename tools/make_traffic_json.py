#!/usr/bin/env python
"""Digest an `ncu --set full` capture of one steady-state step (step_kernel_v2 + close_kernel) into
profiles/step_kernel_traffic.json (DRAM bytes and warp-instructions per step; bench.py quotes them as roofline.traffic /
issue_roofline only while the kernel sources still hash to `csrc_sha16`) and profiles/<tag>_step_kernel_ncu_summary.json.
Usage: python tools/make_traffic_json.py report.ncu-rep tag"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_hash  # noqa: E402

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = {"gpu__time_duration.sum": "gpu_time_ns", "launch__registers_per_thread": "registers", "launch__grid_size": "grid",
        "smsp__inst_executed.sum": "warp_instructions", "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
        "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct"}
units = dict(zip(hdr, rows[1]))


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


kern = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    short = "step_kernel_v2" if "step_kernel_v2" in name else ("close_kernel" if "close_kernel" in name else None)
    if not short or short in kern:
        continue
    k = {"name": name[:80]}
    for m, out in KEYS.items():
        if m in d and d[m] != "":
            k[out] = to_bytes(d[m], units[m]) if out.startswith("dram_r") or out.startswith("dram_w") else float(d[m].replace(",", ""))
            if out == "gpu_time_ns":
                k[out] = float(d[m].replace(",", "")) * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6}.get(units[m], 1)
    st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: float(d[h].replace(",", "") or 0) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    sel = st.get("selected", 1.0) or 1.0
    k["stalls_per_issue"] = {a: round(b / sel, 2) for a, b in sorted(st.items(), key=lambda t: -t[1]) if b / sel > 0.04}
    kern[short] = k
step = kern["step_kernel_v2"]
traffic = {"csrc_sha16": csrc_hash(),
           "dram_bytes_per_launch": step["dram_read"] + step["dram_write"],
           "warp_instructions_per_step": sum(k["warp_instructions"] for k in kern.values()),
           "kernels": kern,
           "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none --import-source on, one steady-state step of 262144 envs "
                     "after 400 burn-in steps, cold caches): dram__bytes_read.sum + dram__bytes_write.sum of the step kernel, smsp__inst_executed.sum "
                     "of the step's two kernels"}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json"), "w"), indent=1)
json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_step_kernel_ncu_summary.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1)[:1500])
