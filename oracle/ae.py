"""Oracle (torch, CPU): the binary tokenizer (conv encoder -> sign quantiser -> conv decoder). TEST INFRASTRUCTURE ONLY.

Functional restatement (state-dict in, tensors out) of modeling/vision_encoder/autoencoder.py:
  ResBlock.forward :41-57 | Encoder.forward :107-127 | Decoder.forward :172-196 | depth_to_space :198-230 |
  Upsampler.forward :243-249 | AdaptiveGroupNorm.forward :260-277 | VQModel.encode :385-390 / decode :514.
(imagenet_gen/src/qae.py is the same network with the GFQ quantiser, see oracle/quant.py.)

``rnd`` = ``ident`` (exact fp32, pinned against the reference) or ``bf16`` = torch.autocast("cuda", bf16) policy:
convolutions / linears round inputs and output to bf16 (fp32 accumulate, bias added before rounding); GroupNorm is
computed and returned in fp32; swish on an fp32 tensor stays fp32; the residual add of two bf16 conv outputs rounds.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .head import bf16, ident  # noqa: F401


def swish(x):
    return x * torch.sigmoid(x)


def _conv(sd, name, x, rnd, stride=1, padding=1):
    w = rnd(sd[name + ".weight"].float())
    b = sd.get(name + ".bias")
    y = F.conv2d(rnd(x), w, None if b is None else rnd(b.float()), stride=stride, padding=padding)
    return rnd(y)


def _gn(sd, name, x, affine=True):
    w = sd[name + ".weight"].float() if affine else None
    b = sd[name + ".bias"].float() if affine else None
    return F.group_norm(x.float(), 32, w, b, eps=1e-6)


def res_block(sd, p, x, rnd):
    h = _conv(sd, p + "conv1", swish(_gn(sd, p + "norm1", x)), rnd)
    h = _conv(sd, p + "conv2", swish(_gn(sd, p + "norm2", h)), rnd)
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(sd, p + "nin_shortcut", x, rnd, padding=0)
    elif (p + "conv_shortcut.weight") in sd:
        x = _conv(sd, p + "conv_shortcut", x, rnd)
    return rnd(h + x)


def _count(sd, prefix):
    return len({k[len(prefix):].split(".")[0] for k in sd if k.startswith(prefix)})


def encoder_forward(sd, x, *, rnd=ident, prefix="encoder."):
    """[B,3,H,W] -> pre-quantisation latent [B,z,H/2^(L-1),W/2^(L-1)]."""
    n_levels = _count(sd, prefix + "down.")
    n_res = _count(sd, prefix + "mid_block.")
    h = _conv(sd, prefix + "conv_in", x, rnd)
    for lv in range(n_levels):
        for b in range(n_res):
            h = res_block(sd, f"{prefix}down.{lv}.block.{b}.", h, rnd)
        if lv < n_levels - 1:
            h = _conv(sd, f"{prefix}down.{lv}.downsample", h, rnd, stride=2, padding=1)
    for b in range(n_res):
        h = res_block(sd, f"{prefix}mid_block.{b}.", h, rnd)
    h = swish(_gn(sd, prefix + "norm_out", h))
    return _conv(sd, prefix + "conv_out", h, rnd, padding=0)


def encode(sd, x, *, rnd=ident):
    h = encoder_forward(sd, x, rnd=rnd)
    return torch.where(h > 0, 1.0, -1.0), h


def depth_to_space(x, bs=2):
    """DCR pixel shuffle: view(-1,bs,bs,C/bs^2,h,w).permute(0,3,4,1,5,2) (autoencoder.py:221-228)."""
    B, C, H, W = x.shape
    c = C // (bs * bs)
    return x.view(B, bs, bs, c, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B, c, H * bs, W * bs)


def ada_group_norm(sd, p, x, style, rnd):
    """GN32(x, affine=False) * gamma(sqrt(var(z)+eps)) + beta(mean(z)); var is torch's default UNBIASED estimator
    over h*w of the +-1 grid (autoencoder.py:264-272; the '#not unbias' comment there is misleading)."""
    B, C = x.shape[:2]
    z = style.reshape(B, style.shape[1], -1).float()
    s = (z.var(dim=-1) + 1e-6).sqrt()
    m = z.mean(dim=-1)
    gamma = rnd(rnd(s) @ rnd(sd[p + "gamma.weight"].float()).t() + rnd(sd[p + "gamma.bias"].float()))
    beta = rnd(rnd(m) @ rnd(sd[p + "beta.weight"].float()).t() + rnd(sd[p + "beta.bias"].float()))
    xn = _gn(sd, p + "gn", x, affine=False)
    return gamma.view(B, C, 1, 1) * xn + beta.view(B, C, 1, 1)


def decoder_forward(sd, z, *, rnd=ident, prefix="decoder."):
    """+-1 grid [B,z,h,w] -> pixels [B,3,h*2^(L-1),w*2^(L-1)]."""
    n_levels = _count(sd, prefix + "up.")
    n_res = _count(sd, prefix + "mid_block.")
    style = z
    h = _conv(sd, prefix + "conv_in", z, rnd)
    for b in range(n_res):
        h = res_block(sd, f"{prefix}mid_block.{b}.", h, rnd)
    for lv in reversed(range(n_levels)):
        h = ada_group_norm(sd, f"{prefix}adaptive.{lv}.", h, style, rnd)
        for b in range(n_res):
            h = res_block(sd, f"{prefix}up.{lv}.block.{b}.", h, rnd)
        if lv > 0:
            h = depth_to_space(_conv(sd, f"{prefix}up.{lv}.upsample.conv1", h, rnd))
    h = swish(_gn(sd, prefix + "norm_out", h))
    return _conv(sd, prefix + "conv_out", h, rnd)
