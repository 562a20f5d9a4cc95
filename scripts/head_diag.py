#!/usr/bin/env python
"""Diagnostic: small head, per-evaluation network outputs of the three implementations (persistent with fillers, persistent
without, multi-kernel) on identical inputs — where do they part?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import ops  # noqa: E402
from bitdance_b200.head import HeadRunner, head_spec  # noqa: E402
from bitdance_b200.synth import synth_state_dict  # noqa: E402

for (D, nb, na, B, pn, cfg, S) in [(256, 4, 2, 1, 64, 3.0, 6), (256, 2, 2, 2, 16, 2.0, 4), (1024, 2, 1, 1, 64, 3.0, 3)]:
    sd = synth_state_dict(head_spec(32, 256, D, nb, na, True), seed=1, std=0.05)
    run = HeadRunner(sd, ch_target=32, ch_cond=256, ch_latent=D, depth_latent=nb, depth_adanln=na, use_swiglu=True)
    torch.manual_seed(0)
    mult = 2 if cfg > 1 else 1
    z = torch.randn(B * mult, pn, 256).cuda()
    noise = torch.randn(S + 1, B, pn, 32).cuda()
    outs = {}
    for name, fill, path in (("fill", 1, "stream"), ("nofill", 0, "stream"), ("tiled", 1, "tiled"), ("fill2", 1, "stream")):
        ops.head_set_fillers(fill)
        x, tr = run.sample(z, cfg, S, noise=noise, trace=True, path=path)
        outs[name] = (x.cpu(), tr.cpu())
    print(f"D={D} blocks={nb} ada={na} B={B} pn={pn} cfg={cfg} S={S}")
    for a, b in (("fill", "tiled"), ("nofill", "tiled"), ("fill", "nofill"), ("fill", "fill2")):
        d = [(outs[a][1][i] - outs[b][1][i]).abs().max().item() for i in range(S + 1)]
        print(f"  {a:7s} vs {b:7s}: per-evaluation max |diff| " + " ".join(f"{v:.4f}" for v in d))
ops.head_set_fillers(1)
