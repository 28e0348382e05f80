"""Summarise gpurun_out/prof_*.ncu-rep (ncu --set full captures) as a markdown table: profiles/r02_ncu_summary.md."""
import csv, glob, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.per_cycle_active", "warps active / SM")]
out = ["# ncu --set full captures, round 2 (`scripts/gpu_round2_*.sh`; reports stay in gpurun_out/)", ""]
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out.append(f"## {os.path.basename(rep)}")
    out.append("| metric | value |")
    out.append("|---|---|")
    for k, name in KEYS:
        if k in d:
            v, u = d[k]
            out.append(f"| {name} | {v[:90]} {u} |")
    stalls = []
    for h, (v, u) in d.items():
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                if float(v) > 0.3:
                    stalls.append((float(v), h.split("issue_stalled_")[1].split("_per_issue")[0]))
            except ValueError:
                pass
    out.append("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6]) + " |")
    out.append("")
path = os.path.join(ROOT, "profiles", "r02_ncu_summary.md")
open(path, "w").write("\n".join(out) + "\n")
print(path)
