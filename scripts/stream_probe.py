#!/usr/bin/env python
"""HBM->smem streaming ceiling of the GEMM's producer/consumer ring (bd_probe_stream), with and without the L2-resident
activation re-reads every CTA does per k-block. Prints GB/s of the HBM stream per configuration."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib  # noqa: E402
from bitdance_b200._lib import check, ptr, stream_ptr  # noqa: E402


def main():
    lib = _lib.load()
    dev = "cuda"
    total = 2 << 30
    w = torch.empty(total, dtype=torch.uint8, device=dev)
    w.random_(0, 255)
    x = torch.empty(128 * 5120 * 2, dtype=torch.uint8, device=dev).random_(0, 255)

    def run(per_cta, chunk, stages, xchunk, xstages, ctas, reps=5):
        def once():
            check(lib.bd_probe_stream(ptr(w), C.c_longlong(per_cta), chunk, stages, ptr(x), x.numel(), xchunk, xstages,
                                      ctas, stream_ptr()), "probe")
        # rotate through the 2 GiB buffer so that successive launches never hit L2
        span = per_cta * ctas
        once()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            once()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        return span / 1e9 / (ms / 1e3), ms * 1e3

    print("per_cta_MB chunkKB stages xchunkKB xstages ctas   GB/s    us")
    for per_cta_mb in (1.0625, 8.0):
        per = int(per_cta_mb * 1024 * 1024)
        for ctas in (120, 148):
            for chunk, stages in ((16384, 10), (16384, 6), (16384, 12), (32768, 5), (8192, 20), (16384, 3)):
                for xchunk, xstages in ((0, 1), (16384, 3), (8192, 3), (4096, 3)):
                    if stages * chunk + xstages * xchunk > 220 * 1024:
                        continue
                    per_al = per // chunk * chunk
                    if per_al * ctas > w.numel():
                        continue
                    gbs, us = run(per_al, chunk, stages, xchunk, xstages, ctas)
                    print(f"{per_cta_mb:9.3f} {chunk // 1024:7d} {stages:6d} {xchunk // 1024:8d} {xstages:7d} {ctas:4d} {gbs:7.0f} {us:7.1f}")


if __name__ == "__main__":
    main()
