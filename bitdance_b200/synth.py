"""Deterministic synthetic weights (there is no network for checkpoints).

Every tensor is drawn from its own generator seeded by crc32(name) ^ seed, so a state dict can be regenerated
tensor-by-tensor on any machine with the same torch build, independent of module construction order. The reference's
own initialisation zeroes the head's adaLN / output layers (flow_head_parallel_x.py:315-323), which would make every
parity check vacuous; here every Linear/Conv weight is N(0, std), biases N(0, std), norm scales 1 + N(0, 0.1).
"""
from __future__ import annotations

import zlib

import torch


def synth_tensor(name: str, shape, seed: int = 0, std: float = 0.02, device="cpu", dtype=torch.float32):
    g = torch.Generator(device=device)
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    is_norm_scale = name.endswith("weight") and len(shape) == 1
    t = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    if is_norm_scale:
        t = 1.0 + 0.1 * t
    else:
        t = t * std
    return t.to(dtype)


def synth_state_dict(spec: dict, seed: int = 0, std: float = 0.02, device="cpu", dtype=torch.float32):
    """spec: {name: shape}."""
    return {k: synth_tensor(k, v, seed, std, device, dtype) for k, v in spec.items()}
