#!/usr/bin/env python
"""Micro-benchmark of the weight-streaming GEMM at the AR-step shapes (M = 128): GB/s per variant.
Each shape cycles through enough distinct weight buffers to exceed L2 (126 MB), timed with CUDA events."""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import ops  # noqa: E402

SHAPES = {
    "head.wqkv/w1 15360x5120": (15360, 5120),
    "head.wo 5120x5120": (5120, 5120),
    "head.w2 5120x7680": (5120, 7680),
    "head.ada 71680x5120": (71680, 5120),
    "llm.qkv 7168x5120": (7168, 5120),
    "llm.gate_up 34816x5120": (34816, 5120),
    "llm.down 5120x17408": (5120, 17408),
}


def bench(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    M = int(os.environ.get("M", 128))
    dev = "cuda"
    variants = [dict(tiled=t, bn=bn, splits=sp, pdl=pdl)
                for t, bn, sp, pdl in itertools.product([False, True], [0, 128, 256], [0, 1], [False, True])]
    only = os.environ.get("VARIANTS")
    rows = []
    for name, (N, K) in SHAPES.items():
        nbuf = max(2, int(400e6 // (N * K * 2)) + 1)
        raw = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(nbuf)]
        packed = [ops.pack_weight(w) for w in raw]
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for v in variants:
            if only and json.dumps(v) not in only:
                pass
            ws = packed if v["tiled"] else raw

            def run():
                for w in ws:
                    ops.gemm(a, w, out=out, bn=v["bn"], splits=v["splits"], pdl=v["pdl"])
            try:
                ms = bench(run, 5) / len(ws)
            except Exception as e:  # noqa: BLE001
                print(name, v, "ERR", e)
                continue
            gbs = N * K * 2 / 1e9 / (ms / 1e3)
            rows.append(dict(shape=name, us=round(ms * 1e3, 1), gbs=round(gbs), **v))
            print(f"{name:28s} tiled={int(v['tiled'])} bn={v['bn']:3d} splits={v['splits']} pdl={int(v['pdl'])}  "
                  f"{ms * 1e3:8.1f} us  {gbs:7.0f} GB/s", flush=True)
        del raw, packed
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
