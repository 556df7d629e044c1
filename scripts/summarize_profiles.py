"""Turns the raw ncu outputs in gpurun_out/ into the small, tracked summaries under profiles/.

    python scripts/summarize_profiles.py r1      # reads gpurun_out/launches_r1.csv, gpurun_out/prof_r1.ncu-rep
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: every launch with its device time (cold-cache, serialised: compare SHARES) ----
launch_csv = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
lines = []
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = r[ki].split("(")[0]
        a = agg.setdefault(n, [0, 0.0, 0.0, 0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
        if v > 8000:          # launches that did real work (idle launches take 2-7 us)
            a[3] += 1
    tot = sum(a[1] for n, a in agg.items() if n.startswith("nb::"))
    lines.append(f"## Launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` around "
                 "`bench.py --steps 1 --warmup 3`\n")
    lines.append("Per-launch times under ncu are cold-cache and serialised; compare shares, not absolutes.\n")
    lines.append("| kernel | launches | with work (>8 us) | total us | share of engine time | max us |")
    lines.append("|---|---:|---:|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if not n.startswith("nb::"):
            continue
        lines.append(f"| `{n}` | {a[0]} | {a[3]} | {a[1] / 1e3:.1f} | {a[1] / tot:.3f} | {a[2] / 1e3:.1f} |")
    lines.append("")

# ---- full-set capture: one row per profiled launch ----
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time us"),
            ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("smsp__inst_executed.sum", "warp insts"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts")]
    idx = [(hdr.index(k), lab) for k, lab in want if k in hdr]
    units = rows[1]
    lines.append(f"## Full-set capture ({tag}): `ncu --set full --clock-control none --import-source on`\n")
    lines.append("| " + " | ".join(lab for _, lab in idx) + " |")
    lines.append("|" + "---|" * len(idx))
    for r in rows[2:]:
        cells = []
        for i, lab in idx:
            v = r[i]
            if lab == "kernel":
                v = "`" + v.split("(")[0] + "`"
            else:
                try:
                    f = float(v.replace(",", ""))
                    u = units[i]
                    if lab.startswith("dram") and "MB" in lab:
                        f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    if lab == "time us":
                        f = f * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                    v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
                except ValueError:
                    pass
            cells.append(v)
        lines.append("| " + " | ".join(cells) + " |")
    lines.append("")

path = os.path.join(out_dir, f"{tag}_ncu_summary.md")
open(path, "w").write("\n".join(lines) + "\n")
print(path)
if os.path.exists(launch_csv):
    import shutil
    shutil.copyfile(launch_csv, os.path.join(out_dir, f"launches_{tag}.csv"))
