#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here on the CPU box with `ncu -i`) into profiles/<name>.json / .txt: the metrics the
roofline arithmetic uses (duration, DRAM bytes, throughput percentages, tensor-pipe activity, registers, L2 traffic).
  python scripts/ncu_summary.py gpurun_out/r02_prof_stream.ncu-rep profiles/r02_stream_kernel_ncu"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):   # already exported on the GPU box (scripts/gpu_profile.sh): ncu -i X.ncu-rep --page raw --csv
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    launches = []
    for r in rows[hdr_i + 2:]:
        if len(r) != len(hdr):
            continue
        d = {}
        for k in KEYS:
            if k in hdr:
                j = hdr.index(k)
                d[k] = {"value": r[j], "unit": units[j]}
        launches.append(d)
    json.dump({"source": rep, "launches": launches}, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for i, d in enumerate(launches):
            f.write(f"--- launch {i} ---\n")
            for k, v in d.items():
                f.write(f"{k:70s} {v['value']:>18s} {v['unit']}\n")
    print(f"{len(launches)} launches -> {out}.json / .txt")


if __name__ == "__main__":
    main()
