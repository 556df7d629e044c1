"""Turns the round-2 ncu outputs in gpurun_out/ (scripts/profile_r2.sh) into small tracked summaries under profiles/:
r2_ncu_summary.md (launch list + one table row per captured kernel), launches_r2.csv, r2_traffic.json (DRAM bytes of the
dominant kernel per stream-block, read by bench.py for `roofline.traffic`)."""
import collections
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
lines = ["# Round 2: ncu evidence (B200, `--clock-control none`)\n"]

launch_csv = os.path.join(G, "launches_r2.csv")
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv, errors="replace")) if len(r) > 10]
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
        a[3] += v > 8000
    tot = sum(a[1] for n, a in agg.items() if "nb::" in n or n.startswith("k_"))
    lines += ["## Launch list: `ncu --metrics gpu__time_duration.sum` around `bench.py --steps 1 --warmup 3` (headline workload)\n",
              "Per-launch times under ncu are cold-cache and serialised; compare shares, not absolutes.\n",
              "| kernel | launches | with work (>8 us) | total us | share of engine time | max us |", "|---|---:|---:|---:|---:|---:|"]
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if not ("nb::" in n or n.startswith("k_")):
            continue
        lines.append(f"| `{n}` | {a[0]} | {a[3]} | {a[1] / 1e3:.1f} | {a[1] / max(tot, 1):.3f} | {a[2] / 1e3:.1f} |")
    lines.append("")
    shutil.copyfile(launch_csv, os.path.join(P, "launches_r2.csv"))

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "dram rd MB"),
        ("dram__bytes_write.sum", "dram wr MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor insts"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("launch__cluster_size", "cluster"), ("smsp__inst_executed.sum", "warp insts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts")]
traffic = {}
for tag in ("k_stream", "k_l2", "k_am", "k_am_decim", "k_channelize"):
    raw = os.path.join(G, f"prof_r2_{tag}_raw.csv")
    if not os.path.exists(raw) or os.path.getsize(raw) < 100:
        continue
    rows = list(csv.reader(open(raw, errors="replace")))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), lab) for k, lab in WANT if k in hdr]
    lines += [f"## `{tag}`: `ncu --set full --clock-control none`\n", "| " + " | ".join(lab for _, lab in idx) + " |", "|" + "---|" * len(idx)]
    for r in rows[2:]:
        cells, rec = [], {}
        for i, lab in idx:
            v = r[i]
            if lab == "kernel":
                v = "`" + v.split("(")[0] + "`"
            else:
                try:
                    f = float(v.replace(",", ""))
                    u = units[i]
                    if lab.startswith("dram") and "MB" in lab:
                        f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    if lab == "time us":
                        f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                    rec[lab] = f
                    v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
                except ValueError:
                    pass
            cells.append(v)
        lines.append("| " + " | ".join(cells) + " |")
        if tag == "k_stream" and "dram rd MB" in rec and "k_stream" not in traffic:
            traffic["k_stream"] = {"dram_read_bytes": rec["dram rd MB"] * 1e6, "dram_write_bytes": rec.get("dram wr MB", 0) * 1e6,
                                   "time_us": rec.get("time us"), "stream_blocks": 128 * 16}
        if tag == "k_channelize" and "dram rd MB" in rec:
            traffic["k_channelize"] = {"dram_read_bytes": rec["dram rd MB"] * 1e6, "dram_write_bytes": rec.get("dram wr MB", 0) * 1e6,
                                       "time_us": rec.get("time us")}
    lines.append("")
    shutil.copyfile(raw, os.path.join(P, f"r2_{tag}_raw.csv"))
if traffic:
    traffic["source"] = "profiles/r2_ncu_summary.md (ncu --set full, first k_stream launch of a step: 128 streams x 16 blocks)"
    json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
open(os.path.join(P, "r2_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)[:3000])
