#!/usr/bin/env python
"""DRAM bytes per render of the whole fwd+bwd step from one `ncu --set full --cache-control none` capture (warm L2, as in the
real pipeline) of a bench.py run of the SAME shape (side, views per image, views per launch) -> profiles/rNN_traffic.json,
which bench.py reads for `roofline.traffic`.

    python profiles/tools/traffic_from_ncu.py gpurun_out/x.ncu-rep S P n_views_in_capture > profiles/r02_traffic.json

`n_views_in_capture` = the views ONE step of the captured run rendered (images x views per image); the capture must hold
exactly one step's launches of every kernel (use -s / -c to cut it), or a whole number of steps (then pass the total).
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict


def main():
    rep, S, P, nviews = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
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
    total = sum(a[0] for a in agg.values())
    print(json.dumps({
        "image_size": S, "views_per_image": P, "views_in_capture": nviews,
        "dram_bytes_per_render": total / nviews,
        "alg_bytes_per_render": 64.0 * S * S + 48.0 * S * S / P,
        "kernels": {k: {"dram_bytes_per_render": a[0] / nviews, "launches": a[1]} for k, a in agg.items()},
        "source": "ncu --set full --clock-control none --cache-control none (warm L2) of bench.py at this shape: %s" % rep,
    }, indent=1))


if __name__ == "__main__":
    main()
