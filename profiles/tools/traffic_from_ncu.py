#!/usr/bin/env python
"""DRAM bytes per render of the whole fwd+bwd step from one `ncu --cache-control none` capture (warm L2, as in the real
pipeline) of a bench.py run of the SAME shape (side, views per image, views per launch) -> profiles/rNN_traffic.json, which
bench.py reads for `roofline.traffic`.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --cache-control none --csv --log-file gpurun_out/step.csv python bench.py --warmup 3 --ncu-step
    python profiles/tools/traffic_from_ncu.py gpurun_out/step.csv S P n_views_in_capture > profiles/r02_traffic.json

The input is that CSV log (one row per launch and metric) or an .ncu-rep holding the same metrics.  `n_views_in_capture` =
the views ONE step of the captured run rendered (images x views per image); `--ncu-step` brackets exactly one step.
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
         "nsecond": 1e-3, "msecond": 1e3}


def short(name):
    return name.replace("void ", "").replace("<unnamed>::", "").split("(")[0].split("<")[0]


def from_log(path, S, P, nviews):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ii, ki, mi, ui, vi = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    agg, seen = OrderedDict(), set()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
        except ValueError:
            continue
        a = agg.setdefault(short(r[ki]), [0.0, 0, 0.0])
        if r[mi].startswith("dram__bytes"):
            a[0] += v
        elif r[mi].startswith("gpu__time_duration"):
            a[2] += v
        if r[ii] not in seen:
            seen.add(r[ii])
            a[1] += 1
    return agg


def main():
    rep, S, P, nviews = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    if rep.endswith(".csv"):
        agg = from_log(rep, S, P, nviews)
        emit(agg, rep, S, P, nviews)
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ki, ri, wi, ti = (hdr.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "gpu__time_duration.sum"))
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = OrderedDict()
    for r in rows[2:]:
        name = r[ki].replace("void ", "").replace("<unnamed>::", "").split("(")[0].split("<")[0]
        b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        a = agg.setdefault(name, [0.0, 0, 0.0])
        a[0] += b
        a[1] += 1
        a[2] += float(r[ti])
    emit(agg, rep, S, P, nviews)


def emit(agg, rep, S, P, nviews):
    total = sum(a[0] for a in agg.values())
    print(json.dumps({
        "image_size": S, "views_per_image": P, "views_in_capture": nviews,
        "dram_bytes_per_render": total / nviews,
        "alg_bytes_per_render": 64.0 * S * S + 48.0 * S * S / P,
        "kernels": {k: {"dram_bytes_per_render": a[0] / nviews, "launches": a[1], "us_under_ncu": round(a[2], 1)}
                    for k, a in agg.items()},
        "source": "ncu --clock-control none --cache-control none (warm L2), one step of bench.py at this shape: %s" % rep,
    }, indent=1))


if __name__ == "__main__":
    main()
