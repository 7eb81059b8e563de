#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an `ncu --set full --import-source on` report.

    python profiles/ncu_lines.py gpurun_out/x.ncu-rep k_splat [top_n] [launch_skip]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    fname, hdr, lines = "", None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                float(d["Instructions Executed"] or 0), float(d["# Samples"] or 0), float(d["Thread Instructions Executed"] or 0)
            except ValueError:      # a source line with embedded quotes (inline asm) that the CSV export mangles
                continue
            lines.append((fname, int(r[0]), r[1].strip(), float(d["Instructions Executed"] or 0), float(d["# Samples"] or 0),
                          float(d["Thread Instructions Executed"] or 0)))
    ti = sum(l[3] for l in lines) or 1
    ts = sum(l[4] for l in lines) or 1
    print("total warp inst %.3g, samples %d, avg threads/inst %.1f" % (ti, ts, sum(l[5] for l in lines) / ti))
    lines.sort(key=lambda l: -l[4])
    print("%-18s %6s %6s %5s  %s" % ("file:line", "inst%", "smpl%", "thr", "source"))
    for f, n, src, i, s, t in lines[:top]:
        print("%-18s %6.2f %6.2f %5.1f  %s" % ("%s:%d" % (f.replace("g2s_", ""), n), 100 * i / ti, 100 * s / ts, t / i if i else 0, src[:110]))


if __name__ == "__main__":
    main()
