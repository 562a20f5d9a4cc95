#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page source --csv` (tens of MB for the persistent kernel) to what is worth keeping: per
kernel, the N SASS lines with the most warp-stall samples, each with its dominant stall reasons, plus the kernel's totals
per stall reason. Reads the CSV on stdin.
  ncu -i prof.ncu-rep --page source --csv --print-source sass | python scripts/ncu_source_top.py 80 > top.txt"""
import csv
import sys


def flush(name, hdr, rows, top):
    if not rows or hdr is None:
        return
    def col(n):
        return hdr.index(n) if n in hdr else -1
    ia, isrc, ismp, iex = col("Address"), col("Source"), col("# Samples"), col("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    def num(r, i):
        try:
            return float(r[i].replace(",", "")) if i >= 0 and r[i] != "" else 0.0
        except ValueError:
            return 0.0
    total = sum(num(r, ismp) for r in rows)
    print(f"=== {name[:150]}")
    print(f"    {len(rows)} SASS lines, {total:.0f} stall samples")
    tot = {h: sum(num(r, i) for r in rows) for i, h in stall_cols}
    s = sum(tot.values()) or 1.0
    print("    stall reasons (all lines): " + ", ".join(f"{h[6:]} {100 * v / s:.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    print(f"    top {top} lines by samples: share | executed | top stalls | address source")
    for r in sorted(rows, key=lambda r: -num(r, ismp))[:top]:
        st = sorted(((num(r, i), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        stt = " ".join(f"{h}:{v:.0f}" for v, h in st if v > 0)
        print(f"    {100 * num(r, ismp) / (total or 1):5.2f}% {num(r, iex):10.0f}  {stt:28s} {r[ia]} {r[isrc][:90]}")


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    name, hdr, rows = None, None, []
    for r in csv.reader(sys.stdin):
        if not r:
            continue
        if r[0] == "Kernel Name":
            flush(name, hdr, rows, top)
            name, hdr, rows = r[1], None, []
        elif r[0] == "Address" and hdr is None:
            hdr = r
        elif hdr is not None and len(r) >= len(hdr) - 1:
            rows.append(r)
    flush(name, hdr, rows, top)


if __name__ == "__main__":
    main()
