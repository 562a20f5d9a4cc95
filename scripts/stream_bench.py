#!/usr/bin/env python
"""GB/s of the persistent kernel on chains of GEMM ops at the AR-step shapes (weights cycled to exceed L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import ops  # noqa: E402

SHAPES = {"head.wqkv/w1 15360x5120": (15360, 5120, 1), "head.wo 5120x5120 ks4": (5120, 5120, 4),
          "head.w2 5120x7680 ks4": (5120, 7680, 4), "head.ada 71680x5120": (71680, 5120, 1),
          "llm.qkv 7168x5120": (7168, 5120, 1), "llm.gate_up 34816x5120": (34816, 5120, 1),
          "llm.down 5120x17408 ks4": (5120, 17408, 4)}


def main():
    dev = "cuda"
    for name, (N, K, ks) in SHAPES.items():
        nbuf = max(2, int(600e6 // (N * K * 2)) + 1)
        nbuf = min(nbuf, 12)
        w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        p0 = ops.stream_pack_weight(w, None, ksplit=ks)
        n = p0.data.numel()
        big = p0.data.repeat(nbuf)
        views = [ops.StreamWeight(big[i * n:(i + 1) * n], p0.bias, N, K, ks, p0.n_ctas, 0) for i in range(nbuf)]
        a = torch.randn(128, K, device=dev).to(torch.bfloat16)
        epi = "partial" if ks > 1 else "bias"
        for reps in (1, 8):
            ops.stream_gemm(a, views, epi=epi, repeat=reps)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.stream_gemm(a, views, epi=epi, repeat=reps)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            per = ms * 1e3 / (reps * nbuf)
            print(f"{name:28s} nbuf={nbuf:2d} reps={reps}  {per:7.1f} us/op  {N * K * 2 / 1e9 / (per / 1e6):7.0f} GB/s", flush=True)
        del big, views, p0, w


if __name__ == "__main__":
    main()
