#!/usr/bin/env python
"""SASS opcode histogram of the built objects (evidence that the hot kernels are Blackwell-native): per object file and per
kernel, the counts of the mnemonics of profiles/../B200_PROFILING.md — UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st),
UTMALDG/UTMASTG/UBLKCP (TMA / bulk copies), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (legacy mma.sync).
Runs on the CPU box (cuobjdump only):  python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "bitdance_b200", "_C")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCCP|SYNCS|HMMA|IMMA|HGMMA|LDGSTS)\b")

print("object / kernel".ljust(78) + "  " + "mnemonic counts")
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    kern, per, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            per[kern] = collections.Counter()
            continue
        m = PAT.search(line)
        if m and kern:
            key = m.group(1)
            per[kern][key] += 1
            total[key] += 1
    if not total:
        continue
    print(f"{f}".ljust(78) + "  " + ", ".join(f"{k} {v}" for k, v in sorted(total.items())))
    for k, c in per.items():
        if c:
            name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
            name = re.sub(r"\(.*", "", name)
            print(("    " + name)[:78].ljust(78) + "  " + ", ".join(f"{a} {b}" for a, b in sorted(c.items())))
