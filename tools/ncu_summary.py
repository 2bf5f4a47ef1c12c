"""Key metrics of one kernel of an .ncu-rep as JSON (duration, registers, grid, warp-instructions, lanes per instruction, issue
slots, pipes incl. the tensor pipe, DRAM bytes, stalls per issue).  Usage: python tools/ncu_summary.py report.ncu-rep kernel-substring out.json [note]"""
import csv
import io
import json
import subprocess
import sys

rep, pat, out = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], dict(zip(rows[0], rows[1]))
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pat not in d.get("Kernel Name", ""):
        continue
    k = {"kernel": d["Kernel Name"][:100], "report": rep.split("/")[-1], "note": note}
    for m in WANT:
        if d.get(m, "") != "":
            k[m] = {"value": float(d[m].replace(",", "")), "unit": units[m]}
    for m in hdr:      # every tensor-pipe metric the capture holds, whatever this ncu version calls it
        if "tensor" in m and d.get(m, "") not in ("", "n/a") and m not in k:
            try:
                k[m] = {"value": float(d[m].replace(",", "")), "unit": units[m]}
            except ValueError:
                pass
    st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: float(d[h].replace(",", "") or 0) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    sel = st.get("selected", 1.0) or 1.0
    k["stalls_per_issue"] = {a: round(b / sel, 2) for a, b in sorted(st.items(), key=lambda t: -t[1]) if b / sel > 0.04}
    json.dump(k, open(out, "w"), indent=1)
    print(json.dumps(k, indent=1)[:2500])
    break
