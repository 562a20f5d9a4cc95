#!/usr/bin/env python
"""Issue rate of mma.sync.m16n8k16 (HMMA, the legacy tensor path) on this GPU as a function of warps per CTA (one CTA per
SM) and independent accumulator chains per warp: cycles per HMMA per warp and the chip-level TFLOP/s it implies."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
sms = torch.cuda.get_device_properties(0).multi_processor_count
iters = 2000
sink = torch.zeros(4, device=dev)
print(f"{sms} SMs, {iters} rounds; cycles per HMMA per warp | HMMA per cycle per SM | TFLOP/s at 1.9 GHz")
for warps in (1, 4, 8, 16):
    for chains in (1, 2, 4, 8, 16):
        out = torch.zeros(sms * warps, dtype=torch.int64, device=dev)
        for _ in range(2):
            _lib.check(lib.bd_probe_hmma(warps, chains, iters, sms, C.c_void_p(out.data_ptr()), C.c_void_p(sink.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "probe")
        torch.cuda.synchronize()
        cyc = out.double().mean().item()
        per = cyc / (iters * chains)
        rate = warps / per
        print(f"warps {warps:2d} chains {chains:2d}: {per:7.2f} | {rate:6.3f} | {rate * 4096 * sms * 1.9e9 / 1e12:7.1f}")
