#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the text summaries committed under profiles/.

    python profiles/summarize_ncu.py full  gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_full.md
    python profiles/summarize_ncu.py list  gpurun_out/launches.csv      > profiles/rNN_launch_shares.md

`full`: one row per profiled launch from `ncu --set full` (duration, DRAM bytes, throughput %, issue %, occupancy,
instructions, registers).  `list`: per-kernel share of the summed gpu__time_duration of one bench step
(`ncu --metrics gpu__time_duration.sum`; cold-cache, serialised launches: compare SHARES, not absolutes).
"""
import csv
import subprocess
import sys
from collections import OrderedDict

FULL = OrderedDict([
    ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %")])


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [c for c in FULL if c in hdr]
    print("| kernel | " + " | ".join(FULL[c] for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        vals = []
        for c in cols:
            v, u = r[hdr.index(c)], units[hdr.index(c)]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            vals.append((v + " " + u).strip() if u not in ("", "%") else v)
        print("| `%s` | " % short(r[hdr.index("Kernel Name")]) + " | ".join(vals) + " |")


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k[:90], a[0], a[1], 100 * a[1] / tot))
    print("\ntotal %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2])
