#!/usr/bin/env python
"""sample() time of the persistent head at the 14B-64x dimensions over engine knobs: L2 prefetch distance, ring split."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib  # noqa: E402
from bitdance_b200.head import HeadRunner, head_spec  # noqa: E402
from bitdance_b200.synthetic import MODELS, _gpu_state_dict  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
hc = MODELS["BitDance-14B-64x"]["head"]
sd = _gpu_state_dict(head_spec(hc["ch_target"], hc["ch_cond"], hc["ch_latent"], hc["depth_latent"], hc["depth_adanln"],
                               hc["use_swiglu"]), 2, dev)
R, pn, S = 2, 64, 50
z = torch.randn(R, pn, 5120, device=dev)


def timed(reps=3):
    global head
    head.sample(z, 7.5, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        head.sample(z, 7.5, S)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


head = HeadRunner(sd, device=dev, tiled=False, **hc)
for ns in (8, 32, 64, 128, 256, 512):
    lib.bd_stream_set_poll_ns(ns)
    ms = timed()
    print(f"barrier poll sleep {ns:4d} ns: {ms:7.2f} ms  {ms / (S + 1) * 1e3:7.1f} us/eval", flush=True)
lib.bd_stream_set_poll_ns(32)
