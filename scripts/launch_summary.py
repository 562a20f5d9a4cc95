#!/usr/bin/env python
"""Summarise an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...``): kernel time per
(kernel, grid) with shares, for the whole capture and — when the persistent head kernel appears at least twice — for one AR
step (from one head launch up to the next). ncu serialises launches and runs them cold-cache: compare SHARES, not
absolute times.
  python scripts/launch_summary.py gpurun_out/r02_launches.csv > profiles/r02_launch_list_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict


def read(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r and "Metric Value" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*$", "", d["Kernel Name"]).strip()
        rows.append((name, d.get("Grid Size", ""), us))
    return rows


def table(rows, title, top=28):
    tot = sum(r[2] for r in rows)
    agg = OrderedDict()
    for name, grid, us in rows:
        k = (name, grid)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    print(f"== {title}: {len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (ncu: cold-cache, serialised - compare shares)")
    for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  x{n:5d} {us / n:10.1f} us  {name[:70]} {grid}")


def main():
    rows = read(sys.argv[1])
    table(rows, "whole capture")
    heads = [i for i, r in enumerate(rows) if re.search(r"bd_stream_kernel<(\(int\))?0>", r[0]) or
             (r[0].endswith("bd_stream_kernel") and r[2] > 20000)]
    if len(heads) >= 2:
        table(rows[heads[0]:heads[1]], "one AR step (head launch .. next head launch)")
    elif len(heads) == 1:
        table(rows[heads[0]:], "from the head launch to the end of the capture")


if __name__ == "__main__":
    main()
