"""Digest of an .ncu-rep (read here with `ncu -i`): headline metrics, stall reasons per issue and the
SASS regions (100-instruction windows) with their share of samples / executed instructions / live lanes.
Usage: python tools/ncu_digest.py report.ncu-rep [kernel-substring] [--windows]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "step_kernel_v2"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pat not in d.get("Kernel Name", ""):
        continue
    print("kernel:", d["Kernel Name"][:90])
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    iss = float(d["smsp__inst_executed.sum"].replace(",", ""))
    st = {}
    for h in hdr:
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            st[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(d[h].replace(",", "") or 0)
    sel = st.get("selected", 1.0) or 1.0
    print("  stalls per issue:", {k: round(v / sel, 2) for k, v in sorted(st.items(), key=lambda t: -t[1]) if v / sel > 0.04})
if "--windows" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(secs[:-1], secs[1:]):
        if pat not in rows[a][1]:
            continue
        h = rows[a + 1]
        ins = [dict(zip(h, r)) for r in rows[a + 2:b] if len(r) == len(h)]
        tot = sum(int(x["# Samples"]) for x in ins) or 1
        totx = sum(int(x["Instructions Executed"]) for x in ins) or 1
        stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        print("SASS instructions:", len(ins), "samples:", tot, "warp-instr:", totx)
        W = 100
        for s in range(0, len(ins), W):
            blk = ins[s:s + W]
            sm = sum(int(x["# Samples"]) for x in blk)
            ex = sum(int(x["Instructions Executed"]) for x in blk)
            th = sum(int(x["Thread Instructions Executed"]) for x in blk)
            top = sorted(((c[6:], sum(int(x[c]) for x in blk)) for c in stalls), key=lambda t: -t[1])[:4]
            print(f"  {s:5d} samp {100 * sm / tot:5.1f}% exec {100 * ex / totx:5.1f}% lanes {th / max(ex, 1):5.1f} {top}")
        c = collections.Counter(int(x["Instructions Executed"]) for x in ins)
        print("  largest same-count blocks (exec count x instructions = share):")
        for v, n in sorted(c.items(), key=lambda t: -t[0] * t[1])[:8]:
            print(f"    {v:8d} x {n:4d} = {100 * v * n / totx:5.1f}%")
        break
