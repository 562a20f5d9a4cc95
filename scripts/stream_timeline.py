#!/usr/bin/env python
"""Timeline of the persistent kernel on chains of identical GEMM ops (bd_stream_set_debug stamps): where an op's time goes
between the grid-wide dependency, the MMA stream, the epilogue and the arrival — for several ring splits / epilogue
experiment modes (bd_stream_set_tuning). Medians over CTAs and over the ops after the first two."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
G = ops.stream_num_ctas()
SHAPES = [(15360, 5120, 1), (5120, 5120, 4), (71680, 5120, 1)]
CONFIGS = [(5, 2, 0), (4, 3, 0)]
if os.environ.get("QUICK"):
    SHAPES, CONFIGS = SHAPES[:1], CONFIGS[:1]
print("shape            ring  mode |  us/op | dep->MMA | MMA phase (GB/s) | acc->epi done | epi->arrived | arrived->next A")
for N, K, ks in SHAPES:
    nbuf = 4 if N * K < 2e8 else 2
    reps = 4
    w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    p0 = ops.stream_pack_weight(w, None, ksplit=ks)
    n = p0.data.numel()
    big = p0.data.repeat(nbuf)
    views = [ops.StreamWeight(big[i * n:(i + 1) * n], p0.bias, N, K, ks, p0.n_ctas, 0) for i in range(nbuf)]
    a = torch.randn(128, K, device="cuda").to(torch.bfloat16)
    epi = "partial" if ks > 1 else "bias"
    nops = nbuf * reps
    dbg = torch.zeros(nops * G * 8, dtype=torch.int64, device="cuda")
    for ws_, as_, mode in CONFIGS:
        lib.bd_stream_set_tuning(ws_, as_, mode)
        ops.stream_gemm(a, views, epi=epi, repeat=reps)
        torch.cuda.synchronize()
        dbg.zero_()
        lib.bd_stream_set_debug(C.c_void_p(dbg.data_ptr()), nops)
        ops.stream_gemm(a, views, epi=epi, repeat=reps)
        torch.cuda.synchronize()
        lib.bd_stream_set_debug(None, 0)
        d = dbg.view(nops, G, 8).cpu().double() / 1e3
        med = lambda q, e: d[q, :, e][d[q, :, e] > 0].median().item()
        mx = lambda q, e: d[q, :, e][d[q, :, e] > 0].max().item()
        mn = lambda q, e: d[q, :, e][d[q, :, e] > 0].min().item()
        qs = range(2, nops - 1)
        per = sum(mn(q + 1, 0) - mn(q, 0) for q in qs) / len(qs)
        dep = sum(med(q, 1) - med(q, 0) for q in qs) / len(qs)
        mma = sum(med(q, 2) - med(q, 1) for q in qs) / len(qs)
        epi_t = sum(med(q, 4) - med(q, 3) for q in qs) / len(qs)
        arr = sum(mx(q, 5) - med(q, 4) for q in qs) / len(qs)
        nxt = sum(mn(q + 1, 0) - mx(q, 5) for q in qs) / len(qs)
        print(f"{N:6d}x{K:5d} ks{ks}  {ws_:2d}+{as_:d}  {mode:3d}  | {per:6.1f} | {dep:8.2f} | {mma:6.1f} ({N * K * 2 / 1e3 / mma:6.0f}) | "
              f"{epi_t:13.2f} | {arr:12.2f} | {nxt:8.2f}", flush=True)
    lib.bd_stream_set_tuning(5, 2, 0)
    del big, views, p0, w
