"""Per-kernel totals of the ncu launch lists (profiles/r02_launches_{hac,sup}.csv: one full step each, cold caches,
serialised launches) -> profiles/r02_launch_summary.md.  The SHARE of a kernel is what compares with bench.py's stage times."""
import csv, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = ["# ncu launch lists, round 2 (`scripts/gpu_final_b.sh`: `ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
       "One full step per workload (`scripts/profile_step.py`), one batch in flight; torch's one-time buffer fills / weight copies",
       "of the plan build are listed under `setup`.", ""]
for which in ("hac", "sup"):
    path = os.path.join(ROOT, "profiles", f"r02_launches_{which}.csv")
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, order = {}, []
    for r in rows[1:]:
        name = r[ik]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
        if short.startswith("at::") or "elementwise" in short:
            short = "setup (torch fills / copies)"
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)   # -> us
        if short not in agg:
            agg[short] = [0, 0.0]
            order.append(short)
        agg[short][0] += 1
        agg[short][1] += v
    total = sum(v for k, (n, v) in agg.items() if not k.startswith("setup"))
    out += [f"## {which}", "", "| kernel | launches | total ms | share of the step |", "|---|---|---|---|"]
    for k in order:
        n, v = agg[k]
        share = "" if k.startswith("setup") else f"{100 * v / total:.1f} %"
        out.append(f"| `{k}` | {n} | {v / 1e3:.3f} | {share} |")
    out += [f"| **step (sum of kernel durations)** | | **{total / 1e3:.2f}** | |", ""]
open(os.path.join(ROOT, "profiles", "r02_launch_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
