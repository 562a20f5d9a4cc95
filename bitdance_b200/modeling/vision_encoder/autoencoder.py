"""API mirror of ``modeling/vision_encoder/autoencoder.py``: ``VQModel`` (binary tokenizer).

Same ctor (``ddconfig``), same state-dict keys (``encoder.*`` / ``decoder.*``) and the same public methods as the
reference (autoencoder.py:354-521): ``encode``, ``decode``, ``forward``, ``vt_forward``, ``vt_forward_maxpad``. The
convolutions, GroupNorms and the sign quantiser run natively (bitdance_b200/ae.py); the image-list bucketing of the
``vt_*`` helpers is host logic re-expressed here. ``gan_decoder=True`` (noise-concat decoder, :279-351) and
``checkpoint=`` loading through the HDFS helper are outside the T2I path and raise."""
from __future__ import annotations

import math
from collections import defaultdict

import torch
import torch.nn.functional as F

from ...ae import AERunner, ae_spec
from .._lazy import NativeModule


def _patch_raster(q: torch.Tensor, ps: int) -> torch.Tensor:
    """'c (h p1) (w p2) -> (h w p1 p2) c'"""
    c, H, W = q.shape
    return q.view(c, H // ps, ps, W // ps, ps).permute(1, 3, 2, 4, 0).reshape(-1, c)


class VQModel(NativeModule):
    def __init__(self, ddconfig, checkpoint=None, gan_decoder=False):
        if gan_decoder:
            raise NotImplementedError("GANDecoder is not used by the T2I pipeline (SURVEY.md §2)")
        if checkpoint is not None:
            raise NotImplementedError("load weights with load_state_dict (the pipeline loads ae.safetensors itself)")
        super().__init__(ae_spec(ddconfig))
        self.ddconfig = dict(ddconfig)

    def _build_runner(self, device):
        return AERunner(self.state_dict(), self.ddconfig, device=device)

    @torch.no_grad()
    def encode(self, x):
        """[B,3,H,W] -> +-1 grid [B,z,H/f,W/f] (where(h > 0, 1, -1); bf16, as under autocast)."""
        q, _, _, _ = self.runner.encode(x)
        return q

    @torch.no_grad()
    def encode_packed(self, x, num_codebooks: int = 0):
        """Extension: (quant, packed uint32 bits [B,hw,z/32], GFQ indices | None)."""
        q, packed, idx, _ = self.runner.encode(x, num_codebooks=num_codebooks)
        return q, packed, idx

    @torch.no_grad()
    def decode(self, quant):
        return self.runner.decode(quant)

    def forward(self, input):
        quant = self.encode(input)
        return self.decode(quant), quant

    @torch.no_grad()
    def vt_forward(self, image_list, max_bs=32, ps=1):
        """Group by (H, W), encode in chunks of max_bs, flatten each grid in patch-raster order; [sum tokens, C]."""
        groups = defaultdict(list)
        for i, img in enumerate(image_list):
            groups[tuple(img.shape[-2:])].append(i)
        out = [None] * len(image_list)
        for _, idxs in groups.items():
            for s in range(0, len(idxs), max_bs):
                chunk = idxs[s:s + max_bs]
                quant = self.encode(torch.cat([image_list[i] for i in chunk], dim=0))
                for b, i in enumerate(chunk):
                    out[i] = _patch_raster(quant[b], ps)
        return torch.cat(out, dim=0)

    @torch.no_grad()
    def vt_forward_maxpad(self, image_list, max_bs=32, stride=32, min_size=256, max_size=2048,
                          max_pixels=1024 * 1024, normal_buckets=(384, 512, 768, 1024)):
        """Bucket by longer side, zero-pad each batch to its max (rounded up to ``stride``), crop the token grids."""
        def bucket_of(H, W):
            major, minor = max(H, W), min(H, W)
            if major >= 1024 and minor <= 768 and major / minor >= 1.5:
                return "long"
            for b in normal_buckets:
                if major <= b:
                    return b
            return "long"

        groups = defaultdict(list)
        for i, img in enumerate(image_list):
            groups[bucket_of(*img.shape[-2:])].append(i)
        out = [None] * len(image_list)
        for _, idxs in groups.items():
            for s in range(0, len(idxs), max_bs):
                chunk = idxs[s:s + max_bs]
                Hp = math.ceil(max(image_list[i].shape[-2] for i in chunk) / stride) * stride
                Wp = math.ceil(max(image_list[i].shape[-1] for i in chunk) / stride) * stride
                batch = torch.cat([F.pad(image_list[i], (0, Wp - image_list[i].shape[-1], 0, Hp - image_list[i].shape[-2]))
                                   for i in chunk], dim=0)
                quant = self.encode(batch)
                for b, i in enumerate(chunk):
                    H, W = image_list[i].shape[-2:]
                    q = quant[b][:, :math.ceil(H / stride), :math.ceil(W / stride)]
                    out[i] = q.permute(1, 2, 0).reshape(-1, q.shape[0])
        return torch.cat(out, dim=0)
