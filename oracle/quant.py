"""Oracle (numpy, CPU): binary quantiser and bit packing. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, without copying:
  * VQModel.encode             modeling/vision_encoder/autoencoder.py:385-390  quant = where(h > 0, +1, -1)
  * GFQ.forward (inference)    imagenet_gen/src/gfq.py:221-222 (same sign rule), :225-239 (index packing with
                               weights 2**arange: channel 0 of each codebook group is the LSB)
  * torch.sign on the AR path  modeling/t2i_pipeline.py:248   (0 < x) - (x < 0): sign(0) = 0 and sign(NaN) = 0
Pinned against the reference itself by tests/test_oracle_vs_reference.py (runs where /root/reference exists).
"""
from __future__ import annotations

import numpy as np


def sign_quantize(h: np.ndarray) -> np.ndarray:
    """where(h > 0, 1, -1): zero and NaN map to -1."""
    return np.where(h > 0, 1.0, -1.0).astype(np.float32)


def sign_lfq(x: np.ndarray) -> np.ndarray:
    """torch.sign = (0 < x) - (x < 0): 0 -> 0 and NaN -> 0 (numpy's sign would propagate NaN)."""
    return ((x > 0).astype(np.float32) - (x < 0).astype(np.float32)).astype(np.float32)


def pack_bits_nchw(h: np.ndarray) -> np.ndarray:
    """[B,C,H,W] -> uint32 [B, H*W, C/32]; bit (c % 32) of word (c // 32) is h[b,c,hw] > 0."""
    B, C, H, W = h.shape
    bits = (h > 0).reshape(B, C // 32, 32, H * W).astype(np.uint64)
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64)).reshape(1, 1, 32, 1)
    words = (bits * weights).sum(axis=2)  # [B, C/32, HW]
    return np.ascontiguousarray(words.transpose(0, 2, 1)).astype(np.uint32)


def pack_bits_tokens(x: np.ndarray) -> np.ndarray:
    """[..., C] -> uint32 [..., C/32]."""
    C = x.shape[-1]
    bits = (x > 0).reshape(*x.shape[:-1], C // 32, 32).astype(np.uint64)
    weights = np.uint64(1) << np.arange(32, dtype=np.uint64)
    return (bits * weights).sum(axis=-1).astype(np.uint32)


def unpack_bits_tokens(words: np.ndarray, C: int) -> np.ndarray:
    w = words.astype(np.uint64).reshape(*words.shape[:-1], C // 32, 1)
    bits = (w >> np.arange(32, dtype=np.uint64)) & np.uint64(1)
    return np.where(bits.reshape(*words.shape[:-1], C) > 0, 1.0, -1.0).astype(np.float32)


def gfq_indices(h: np.ndarray, num_codebooks: int) -> np.ndarray:
    """GFQ index per codebook group: int32 [num_codebooks, B*H*W], idx = sum_i [x_i > 0] << i (i within group)."""
    B, C, H, W = h.shape
    cpg = C // num_codebooks
    bits = (h > 0).reshape(B, num_codebooks, cpg, H * W).astype(np.int64)
    weights = (1 << np.arange(cpg, dtype=np.int64)).reshape(1, 1, cpg, 1)
    idx = (bits * weights).sum(axis=2)  # [B, ncb, HW]
    return np.ascontiguousarray(idx.transpose(1, 0, 2).reshape(num_codebooks, B * H * W)).astype(np.int32)


def gfq_indices_to_bits(idx: np.ndarray, cpg: int) -> np.ndarray:
    """Inverse of gfq_indices for one group: [...] int -> [..., cpg] in {-1,+1}."""
    bits = (idx[..., None].astype(np.int64) >> np.arange(cpg, dtype=np.int64)) & 1
    return np.where(bits > 0, 1.0, -1.0).astype(np.float32)
