#!/usr/bin/env python
"""Aggregate the warp-stall samples of an .ncu-rep by source line (read here, no GPU needed).

  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > /tmp/mix.csv
  python tools/ncu_src_lines.py /tmp/mix.csv [top] > profiles/<name>.txt
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur, kern, agg = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            kern = r[1]
            continue
        if r[0] in ("", "Line No"):
            continue
        try:
            ln = int(r[0])
            samples = int(r[4]) if r[4] not in ("-", "") else 0
            inst = int(r[7]) if r[7] not in ("-", "") else 0
        except ValueError:
            continue
        agg[(cur, ln)] = (samples, inst, r[1].strip()[:100])
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print("kernel:", kern)
    print("total warp-stall samples %d, warp instructions executed %d" % (tot, toti))
    byfile = {}
    for (f, _), v in agg.items():
        b = byfile.setdefault(f, [0, 0])
        b[0] += v[0]
        b[1] += v[1]
    print("\nper file: samples (share), instructions (share)")
    for f, v in sorted(byfile.items(), key=lambda x: -x[1][0]):
        print("  %-24s %8d %5.1f%%  %11d %5.1f%%" % (f, v[0], 100 * v[0] / tot, v[1], 100 * v[1] / toti))
    print("\ntop lines: file line samples share instructions source")
    for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print("  %-22s %4d %7d %5.1f%% %10d  %s" % (f, l, v[0], 100 * v[0] / tot, v[1], v[2]))


if __name__ == "__main__":
    main()
