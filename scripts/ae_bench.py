#!/usr/bin/env python
"""Tokenizer (ae_d16c32) encode / decode throughput at BASELINE.json configs[0..1] sizes: TFLOP/s of the tcgen05 implicit-GEMM
convolutions against the measured bf16 peak (17.02 TFLOP encode, 20.28 TFLOP decode per 1024^2 image).
NOT yet run in round 1 (GPU budget): meant for `gpurun -- python scripts/ae_bench.py` and an ncu capture of bd_conv_kernel.

  python scripts/ae_bench.py [--size 1024] [--bs 1 2 4 8]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200.ae import AERunner, ae_spec  # noqa: E402
from bitdance_b200.synthetic import AE_D16C32, _gpu_state_dict  # noqa: E402

ENC_TFLOP_1024, DEC_TFLOP_1024 = 17.02, 20.28   # per image; scales with H*W


def timed(fn, reps):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--bs", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda")
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    ae = AERunner(_gpu_state_dict(ae_spec(AE_D16C32), 3, dev), AE_D16C32, device=dev)
    scale = (args.size / 1024.0) ** 2
    for bs in args.bs:
        x = torch.rand(bs, 3, args.size, args.size, device=dev) * 2 - 1
        q, _, _, _ = ae.encode(x)
        ms_e = timed(lambda: ae.encode(x), args.reps)
        ms_d = timed(lambda: ae.decode(q), args.reps)
        te = ENC_TFLOP_1024 * scale * bs / (ms_e / 1e3)
        td = DEC_TFLOP_1024 * scale * bs / (ms_d / 1e3)
        print(f"{args.size}x{args.size} bs={bs}: encode {ms_e:8.1f} ms {te:7.0f} TFLOP/s ({te / peak:.2f} of {peak:.0f} sustained)  "
              f"decode {ms_d:8.1f} ms {td:7.0f} TFLOP/s ({td / peak:.2f})  images/s enc {bs / ms_e * 1e3:.2f} dec {bs / ms_d * 1e3:.2f}",
              flush=True)
        del x, q
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
