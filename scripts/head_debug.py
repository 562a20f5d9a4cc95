#!/usr/bin/env python
"""Bisect the persistent head path against the multi-kernel path: one evaluation (S=0), compare workspace regions."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib, ops  # noqa: E402
from bitdance_b200.head import HeadRunner, head_spec  # noqa: E402
from bitdance_b200.synth import synth_state_dict  # noqa: E402

cfg = dict(ch_target=32, ch_cond=256, ch_latent=256, depth_latent=int(os.environ.get("DEPTH", 2)), depth_adanln=2)
B, pn, guidance, S = 1, 64, 3.0, int(os.environ.get("S", 0))
spec = head_spec(cfg["ch_target"], cfg["ch_cond"], cfg["ch_latent"], cfg["depth_latent"], cfg["depth_adanln"], True)
sd = synth_state_dict(spec, seed=1, std=0.05)
r = HeadRunner(sd, ch_target=32, ch_cond=256, ch_latent=256, depth_latent=cfg["depth_latent"], depth_adanln=2, use_swiglu=True)
torch.manual_seed(0)
z = torch.randn(2, pn, 256).cuda()
noise = torch.randn(S + 1, B, pn, 32).cuda()
lib = _lib.load()
D, hid, M, C_ = 256, 384, 128, 32
n_mod = 2 * 6 * D + 2 * D


def offsets(w):
    arr = (C.c_size_t * 32)()
    n = lib.bd_head_ws_offsets(C.byref(w), B, pn, 2, S, arr, 32)
    return list(arr[:n])


xs, ts = r.sample(z, guidance, S, noise=noise, trace=True, path="stream")
torch.cuda.synchronize()
ws_s = r._ws[2].clone()
xt, tt = r.sample(z, guidance, S, noise=noise, trace=True, path="tiled")
torch.cuda.synchronize()
ws_t = r._ws[1].clone()
names_s = ["xb", "h", "y", "mod", "a", "qkv", "o", "g", "cemb", "condb", "tfreq", "th", "temb", "part", "pred", "x", "sync", "total"]
names_t = ["xb", "h", "a", "o", "qkv", "g", "y", "mod", "cemb", "tfreq", "th", "temb", "pred", "x", "condb", "tvals", "gemm", "total"]
os_, ot = dict(zip(names_s, offsets(r.w_stream))), dict(zip(names_t, offsets(r.w)))


def reg(ws, off, nbytes, dtype):
    return ws[off:off + nbytes].view(dtype)


def rowmajor(ws, off, rows, cols, dtype=torch.bfloat16):
    return reg(ws, off, rows * cols * (2 if dtype == torch.bfloat16 else 4), dtype).view(rows, cols).float()


def blocked(ws, off, rows, cols):
    nb = (cols + 63) // 64 * 16384
    return ops.from_blocked(reg(ws, off, nb, torch.bfloat16), rows, cols).float()


def cmp(name, a, b):
    d = (a - b).abs().max().item()
    print(f"{name:8s} max|stream-tiled| = {d:.5f}   (|tiled| max {b.abs().max().item():.4f})")


cmp("cemb", rowmajor(ws_s, os_["cemb"], M, D), rowmajor(ws_t, ot["cemb"], M, D))
cmp("temb", rowmajor(ws_s, os_["temb"], S + 1, D), rowmajor(ws_t, ot["temb"], S + 1, D))
cmp("condb", blocked(ws_s, os_["condb"], M, 256), rowmajor(ws_t, ot["condb"], M, 256))
cmp("tfreq", blocked(ws_s, os_["tfreq"], S + 1, 256), rowmajor(ws_t, ot["tfreq"], S + 1, 256))
cmp("th", blocked(ws_s, os_["th"], S + 1, D), rowmajor(ws_t, ot["th"], S + 1, D))
cmp("y", blocked(ws_s, os_["y"], M, D), rowmajor(ws_t, ot["y"], M, D))
cmp("mod", rowmajor(ws_s, os_["mod"], M, n_mod), rowmajor(ws_t, ot["mod"], M, n_mod))
cmp("qkv", blocked(ws_s, os_["qkv"], M, 3 * D), rowmajor(ws_t, ot["qkv"], M, 3 * D))
cmp("o", blocked(ws_s, os_["o"], M, D), rowmajor(ws_t, ot["o"], M, D))
cmp("g", blocked(ws_s, os_["g"], M, hid), rowmajor(ws_t, ot["g"], M, hid))
cmp("h", rowmajor(ws_s, os_["h"], M, D), rowmajor(ws_t, ot["h"], M, D))
cmp("pred", rowmajor(ws_s, os_["pred"], M, C_, torch.float32), rowmajor(ws_t, ot["pred"], M, C_, torch.float32))
cmp("trace0", ts[0], tt[0])
cmp("x", xs, xt)

# per-op references from the stream path's own inputs
sdg = {k: v.cuda() for k, v in sd.items()}
bfr = lambda t: t.to(torch.bfloat16).float()
th_s = blocked(ws_s, os_["th"], S + 1, D)
temb_s = rowmajor(ws_s, os_["temb"], S + 1, D)
W2, b2 = sdg["net.time_embed.mlp.2.weight"].float(), sdg["net.time_embed.mlp.2.bias"].float()
temb_ref = bfr(th_s @ bfr(W2).t() + bfr(b2))
cmp("temb|th", temb_s, temb_ref)
print("temb_s[0,:8]", temb_s[0, :8].tolist())
print("temb_ref[0,:8]", temb_ref[0, :8].tolist())
print("bias2[:8]", bfr(b2)[:8].tolist())
cemb_s = rowmajor(ws_s, os_["cemb"], M, D)
y_s = blocked(ws_s, os_["y"], M, D)
y_ref = bfr(torch.nn.functional.silu(bfr(temb_s[S:S + 1] + cemb_s)))
cmp("y|temb", y_s, y_ref)
