#!/usr/bin/env python
"""Per-op timeline of the persistent head sampler at the 14B-64x dimensions (one evaluation = 47 ops), plus event timings
of a whole sample() call on both paths and of one LLM block pass."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib, ops  # noqa: E402
from bitdance_b200.head import HeadRunner, head_spec  # noqa: E402
from bitdance_b200.synthetic import MODELS, _gpu_state_dict  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
hc = MODELS["BitDance-14B-64x"]["head"]
sd = _gpu_state_dict(head_spec(hc["ch_target"], hc["ch_cond"], hc["ch_latent"], hc["depth_latent"], hc["depth_adanln"],
                               hc["use_swiglu"]), 2, dev)
head = HeadRunner(sd, device=dev, tiled=bool(int(os.environ.get("TILED", "1"))), **hc)
del sd
torch.cuda.empty_cache()
R, pn, S = 2, 64, 50
z = torch.randn(R, pn, 5120, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for path in (["stream", "tiled"] if head.w is not None else ["stream"]):
    ms = timed(lambda: head.sample(z, 7.5, S, path=path))
    print(f"sample() path={path}: {ms:.2f} ms  = {ms / (S + 1) * 1e3:.1f} us per evaluation; "
          f"{(S + 1) * 3.465e9 / (ms / 1e3) / 1e9:.0f} GB/s of the head's weight bytes")

G = ops.stream_num_ctas()
nops = 7 + 46 * 2
dbg = torch.zeros(nops * G * 8, dtype=torch.int64, device=dev)
lib.bd_stream_set_debug(C.c_void_p(dbg.data_ptr()), nops)
head.sample(z, 7.5, S, path="stream")
torch.cuda.synchronize()
lib.bd_stream_set_debug(None, 0)
d = dbg.view(nops, G, 8).cpu().double() / 1e3
names = ["cast", "tfreq", "init", "time0", "cond", "time2", "silu_add0"]
body = ["input_proj", "ada", "ln0"]
for b in range(6):
    body += [f"b{b}.wqkv", f"b{b}.attn", f"b{b}.wo", f"b{b}.row_wo", f"b{b}.w1", f"b{b}.w2", f"b{b}.row_w2"]
body += ["sde+silu"]
names = names + body + body
t_prev = None
print("op                 done(us)  dur(us) | arrive spread (med->max)")
tot = {}
for q in range(nops):
    v = d[q, :, 5]
    v = v[v > 0]
    done = v.max().item()
    if t_prev is None:
        t_prev = d[q, :, 4][d[q, :, 4] > 0].min().item()
        t0 = t_prev
    dur = done - t_prev
    if q >= 7 + 46:
        key = names[q].split(".")[-1]
        tot[key] = tot.get(key, 0.0) + dur
        print(f"{names[q]:16s} {done - t0:9.1f} {dur:8.1f} | {done - v.median().item():6.1f}")
    t_prev = done
print("second evaluation, per op kind (us):", {k: round(v, 1) for k, v in tot.items()}, "total", round(sum(tot.values()), 1))
