#!/usr/bin/env python
"""A/B timing of DiffHead.sample() at the 14B-64x dimensions: the program policies in MODES are measured round-robin
(ROUNDS rounds of REPS calls each), so clock / temperature drift hits them equally; prints per-mode median ms and the SM
clock / power seen while it ran. With BD_LIB_PATH pointing at an older build only mode "lib" (whatever that build does) runs.
  MODES="0,1:22:8,1:27:0"  ROUNDS=4 REPS=3"""
import os
import statistics
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib  # noqa: E402
from bitdance_b200.head import HeadRunner, head_spec  # noqa: E402
from bitdance_b200.synthetic import MODELS, _gpu_state_dict  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
has_fill = hasattr(lib, "bd_head_set_fillers")
hc = MODELS["BitDance-14B-64x"]["head"]
sd = _gpu_state_dict(head_spec(hc["ch_target"], hc["ch_cond"], hc["ch_latent"], hc["depth_latent"], hc["depth_adanln"],
                               hc["use_swiglu"]), 2, dev)
head = HeadRunner(sd, device=dev, tiled=False, **hc)
del sd
torch.cuda.empty_cache()
R, pn, S = 2, 64, 50
z = torch.randn(R, pn, 5120, device=dev)
modes = os.environ.get("MODES", "0,1:22:8,1:27:0").split(",") if has_fill else ["lib"]
rounds, reps = int(os.environ.get("ROUNDS", "4")), int(os.environ.get("REPS", "3"))


def smi():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap",
                          "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip().splitlines()[0]
    return out


def set_mode(m):
    """mode = fill[:row_kb[:gemm_kb[:tuning[:w_slots]]]] (tuning = bd_stream_set_tuning mode bits, e.g. 8 = no K rotation)"""
    if m == "lib":
        return
    f = [int(x) for x in m.split(":")] + [0, 0, 0, 0]
    lib.bd_head_set_fillers(f[0], f[1] or 22, f[2] or 8)
    w_slots = f[4] or 5
    lib.bd_stream_set_tuning(w_slots, 7 - w_slots, f[3])


res = {m: [] for m in modes}
info = {m: [] for m in modes}
for m in modes:  # warm-up every mode once
    set_mode(m)
    head.sample(z, 7.5, S, path="stream")
torch.cuda.synchronize()
for r in range(rounds):
    for m in modes:
        set_mode(m)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            head.sample(z, 7.5, S, path="stream")
        e1.record()
        info[m].append(smi())   # sampled while the last call is still running
        torch.cuda.synchronize()
        res[m].append(e0.elapsed_time(e1) / reps)
for m in modes:
    med = statistics.median(res[m])
    print(f"mode {m:10s} median {med:7.3f} ms  ({med / (S + 1) * 1e3:6.1f} us/eval, {(S + 1) * 3.465 / med:6.3f} TB/s, "
          f"{(S + 1) * 3.465 / med / 6.568:.3f} of 6.568)  all {[round(x, 2) for x in res[m]]}  smi {info[m][-1]}")
