"""Host side of the binary tokenizer (VQModel): weight prepack + the op sequence of Encoder/Decoder.forward.

Every arithmetic op is a ``bd_*`` C call (tcgen05 implicit-GEMM convolutions, NHWC GroupNorm kernels, the sign/pack
quantiser); torch only provides the buffers. Op order and dtypes follow modeling/vision_encoder/autoencoder.py under
autocast(bf16): Encoder.forward :107-127, Decoder.forward :172-196, ResBlock.forward :41-57, Upsampler :243-249,
AdaptiveGroupNorm :260-277, VQModel.encode :385-390.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


def ae_spec(ddconfig: dict) -> dict:
    """State-dict spec (names + shapes) of the reference VQModel(ddconfig) with the plain Decoder."""
    ch, ch_mult, nrb = ddconfig["ch"], list(ddconfig["ch_mult"]), ddconfig["num_res_blocks"]
    zc, cin, cout = ddconfig["z_channels"], ddconfig["in_channels"], ddconfig["out_ch"]
    s = {}

    def conv(name, o, i, k, bias):
        s[name + ".weight"] = (o, i, k, k)
        if bias:
            s[name + ".bias"] = (o,)

    def gn(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)

    def resblock(p, i, o):
        gn(p + "norm1", i)
        gn(p + "norm2", o)
        conv(p + "conv1", o, i, 3, False)
        conv(p + "conv2", o, o, 3, False)
        if i != o:
            conv(p + "nin_shortcut", o, i, 1, False)

    L = len(ch_mult)
    conv("encoder.conv_in", ch, cin, 3, False)
    in_mult = [1] + ch_mult
    block_in = ch
    for lv in range(L):
        block_in, block_out = ch * in_mult[lv], ch * ch_mult[lv]
        for b in range(nrb):
            resblock(f"encoder.down.{lv}.block.{b}.", block_in, block_out)
            block_in = block_out
        if lv < L - 1:
            conv(f"encoder.down.{lv}.downsample", block_out, block_out, 3, True)
    for b in range(nrb):
        resblock(f"encoder.mid_block.{b}.", block_in, block_in)
    gn("encoder.norm_out", block_in)
    conv("encoder.conv_out", zc, block_in, 1, True)

    block_in = ch * ch_mult[-1]
    conv("decoder.conv_in", block_in, zc, 3, True)
    for b in range(nrb):
        resblock(f"decoder.mid_block.{b}.", block_in, block_in)
    for lv in reversed(range(L)):
        block_out = ch * ch_mult[lv]
        for n in ("gamma", "beta"):
            s[f"decoder.adaptive.{lv}.{n}.weight"] = (block_in, zc)
            s[f"decoder.adaptive.{lv}.{n}.bias"] = (block_in,)
        for b in range(nrb):
            resblock(f"decoder.up.{lv}.block.{b}.", block_in, block_out)
            block_in = block_out
        if lv > 0:
            conv(f"decoder.up.{lv}.upsample.conv1", block_in * 4, block_in, 3, True)
    gn("decoder.norm_out", block_in)
    conv("decoder.conv_out", cout, block_in, 3, True)
    return s


class _Conv:
    __slots__ = ("w", "b", "cin", "cout", "k")


class AERunner:
    def __init__(self, state_dict: dict, ddconfig: dict, device="cuda"):
        self.dd = dict(ddconfig)
        self.device = dev = torch.device(device)
        self.L = len(ddconfig["ch_mult"])
        self.nrb = ddconfig["num_res_blocks"]
        self.zc = ddconfig["z_channels"]
        self.sd_keys = set(state_dict.keys())
        self.convs, self.norms, self.ada = {}, {}, {}
        for name, t in state_dict.items():
            if name.endswith(".weight") and t.dim() == 4:
                base = name[: -len(".weight")]
                co, ci, k, _ = t.shape
                cip = (ci + 7) // 8 * 8
                w = torch.zeros((co, k, k, cip), dtype=torch.bfloat16, device=dev)
                w[..., :ci] = t.detach().to(dev).permute(0, 2, 3, 1).to(torch.bfloat16)
                c = _Conv()
                c.w, c.cin, c.cout, c.k = w.reshape(co, k * k * cip).contiguous(), cip, co, k
                bname = base + ".bias"
                c.b = state_dict[bname].detach().to(dev, torch.bfloat16).contiguous() if bname in state_dict else None
                self.convs[base] = c
            elif name.endswith(".weight") and t.dim() == 1:
                base = name[: -len(".weight")]
                self.norms[base] = (t.detach().to(dev, torch.float32).contiguous(),
                                    state_dict[base + ".bias"].detach().to(dev, torch.float32).contiguous())
            elif name.endswith(".weight") and t.dim() == 2:  # adaptive.N.gamma / beta Linear
                base = name[: -len(".weight")]
                self.ada[base] = (t.detach().to(dev, torch.bfloat16).contiguous(),
                                  state_dict[base + ".bias"].detach().to(dev, torch.bfloat16).contiguous())
        self._gn_ws = None
        torch.cuda.synchronize(dev)

    # ---- primitive wrappers -------------------------------------------------------------------------------------
    def _conv(self, name, x, B, H, W, *, stride=1, res=None, out_mode=0, out_f32=False, out=None):
        """x: bf16 NHWC [B,H_in,W_in,Cin]; returns NHWC [B,H_out,W_out,Cout] (or d2s / NCHW per out_mode)."""
        lib = _lib.load()
        c = self.convs[name]
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[-1] == c.cin, (name, x.shape, c.cin)
        Ho, Wo = (H // 2, W // 2) if stride == 2 else (H, W)
        if stride == 2:
            assert H % 2 == 0 and W % 2 == 0
            ph = torch.empty((4, B, Ho, Wo, c.cin), dtype=torch.bfloat16, device=self.device)
            check(lib.bd_phase_split_nhwc(ptr(x), ptr(ph), B, Ho, Wo, c.cin, stream_ptr()), "bd_phase_split_nhwc")
            x = ph
        if res is not None:
            out_f32 = res.dtype == torch.float32
        if out is None:
            if out_mode == 0:
                out = torch.empty((B, Ho, Wo, c.cout), dtype=torch.float32 if out_f32 else torch.bfloat16, device=self.device)
            elif out_mode == 1:
                out = torch.empty((B, 2 * Ho, 2 * Wo, c.cout // 4), dtype=torch.bfloat16, device=self.device)
            else:
                out = torch.empty((B, c.cout, Ho, Wo), dtype=torch.float32 if out_f32 else torch.bfloat16, device=self.device)
        check(lib.bd_conv2d_nhwc(ptr(x), ptr(c.w), ptr(c.b), ptr(res), 1 if (res is not None and res.dtype == torch.float32) else 0,
                                 ptr(out), 1 if out.dtype == torch.float32 else 0, out_mode, B, Ho, Wo, c.cin, c.cout,
                                 c.k, stride, 0, stream_ptr()), f"bd_conv2d_nhwc({name})")
        return out

    def _gn(self, x, B, HW, Cc, weight, bias, mode):
        lib = _lib.load()
        lib.bd_groupnorm_workspace_bytes.restype = C.c_size_t
        need = lib.bd_groupnorm_workspace_bytes(B, C.c_longlong(HW))
        if self._gn_ws is None or self._gn_ws.numel() < need:
            self._gn_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(x.shape, dtype=torch.bfloat16 if mode == 0 else torch.float32, device=self.device)
        check(lib.bd_groupnorm_nhwc(ptr(x), 1 if x.dtype == torch.float32 else 0, B, C.c_longlong(HW), Cc, ptr(weight),
                                    ptr(bias), mode, ptr(out), ptr(self._gn_ws), C.c_size_t(self._gn_ws.numel()),
                                    C.c_float(1e-6), 0, stream_ptr()), "bd_groupnorm_nhwc")
        return out

    def _to_bf16(self, x):
        if x.dtype == torch.bfloat16:
            return x
        lib = _lib.load()
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=self.device)
        check(lib.bd_cast_f32_bf16(ptr(x), ptr(out), C.c_longlong(x.numel()), stream_ptr()), "bd_cast_f32_bf16")
        return out

    def _resblock(self, p, x, B, H, W):
        Cin = x.shape[-1]
        a = self._gn(x, B, H * W, Cin, *self.norms[p + "norm1"], 0)
        t = self._conv(p + "conv1", a, B, H, W)
        a = self._gn(t, B, H * W, t.shape[-1], *self.norms[p + "norm2"], 0)
        res = x
        if (p + "nin_shortcut") in self.convs:
            res = self._conv(p + "nin_shortcut", self._to_bf16(x), B, H, W)
        return self._conv(p + "conv2", a, B, H, W, res=res)

    # ---- public -------------------------------------------------------------------------------------------------
    def encoder_forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,3,H,W] float -> pre-quant latent, bf16 NHWC [B,h,w,z]."""
        lib = _lib.load()
        B, Ci, H, W = x.shape
        xf = x.to(torch.float32).contiguous()
        cin = self.convs["encoder.conv_in"].cin
        h = torch.empty((B, H, W, cin), dtype=torch.bfloat16, device=self.device)
        check(lib.bd_nchw_to_nhwc_bf16(ptr(xf), ptr(h), B, Ci, H, W, cin, stream_ptr()), "bd_nchw_to_nhwc_bf16")
        h = self._conv("encoder.conv_in", h, B, H, W)
        for lv in range(self.L):
            for b in range(self.nrb):
                h = self._resblock(f"encoder.down.{lv}.block.{b}.", h, B, H, W)
            if lv < self.L - 1:
                h = self._conv(f"encoder.down.{lv}.downsample", h, B, H, W, stride=2)
                H, W = H // 2, W // 2
        for b in range(self.nrb):
            h = self._resblock(f"encoder.mid_block.{b}.", h, B, H, W)
        a = self._gn(h, B, H * W, h.shape[-1], *self.norms["encoder.norm_out"], 0)
        return self._conv("encoder.conv_out", a, B, H, W)

    def encode(self, x: torch.Tensor, *, num_codebooks: int = 0):
        """VQModel.encode: returns (quant NCHW +-1 bf16, packed uint32-as-int32 [B,hw,z/32], gfq indices|None, latent NCHW bf16)."""
        lib = _lib.load()
        lat = self.encoder_forward(x)
        B, h, w, zc = lat.shape
        lat_nchw = torch.empty((B, zc, h, w), dtype=torch.bfloat16, device=self.device)
        check(lib.bd_nhwc_to_nchw(ptr(lat), ptr(lat_nchw), 0, B, zc, h, w, stream_ptr()), "bd_nhwc_to_nchw")
        q, packed, idx = ops.sign_pack_nchw(lat_nchw, num_codebooks=num_codebooks)
        return q, packed, idx, lat_nchw

    def decode_grid(self, z: torch.Tensor, out_f32: bool = False) -> torch.Tensor:
        """z: bf16 NHWC +-1 grid [B,h,w,zc] -> image NCHW [B,3,H,W] (bf16, like the reference under autocast)."""
        lib = _lib.load()
        B, H, W, zc = z.shape
        L = self.L
        h = self._conv("decoder.conv_in", z, B, H, W)
        for b in range(self.nrb):
            h = self._resblock(f"decoder.mid_block.{b}.", h, B, H, W)
        for lv in reversed(range(L)):
            Cc = h.shape[-1]
            gw, gb = self.ada[f"decoder.adaptive.{lv}.gamma"]
            bw, bb = self.ada[f"decoder.adaptive.{lv}.beta"]
            gamma = torch.empty((B, Cc), dtype=torch.float32, device=self.device)
            beta = torch.empty((B, Cc), dtype=torch.float32, device=self.device)
            check(lib.bd_adagn_params(ptr(z), B, z.shape[1] * z.shape[2], zc, ptr(gw), ptr(gb), ptr(bw), ptr(bb), Cc,
                                      ptr(gamma), ptr(beta), stream_ptr()), "bd_adagn_params")
            h = self._gn(h, B, H * W, Cc, gamma, beta, 1)          # fp32 stream from here (AdaGN returns fp32)
            for b in range(self.nrb):
                h = self._resblock(f"decoder.up.{lv}.block.{b}.", h, B, H, W)
            if lv > 0:
                h = self._conv(f"decoder.up.{lv}.upsample.conv1", self._to_bf16(h), B, H, W, out_mode=1)
                H, W = 2 * H, 2 * W
        a = self._gn(h, B, H * W, h.shape[-1], *self.norms["decoder.norm_out"], 0)
        return self._conv("decoder.conv_out", a, B, H, W, out_mode=2, out_f32=out_f32)

    def decode(self, quant: torch.Tensor) -> torch.Tensor:
        """VQModel.decode(quant NCHW +-1)."""
        lib = _lib.load()
        B, zc, h, w = quant.shape
        q = quant.to(torch.float32).contiguous()
        z = torch.empty((B, h, w, zc), dtype=torch.bfloat16, device=self.device)
        check(lib.bd_nchw_to_nhwc_bf16(ptr(q), ptr(z), B, zc, h, w, zc, stream_ptr()), "bd_nchw_to_nhwc_bf16")
        return self.decode_grid(z)

    def decode_tokens(self, tokens: torch.Tensor, h: int, w: int, ps: int) -> torch.Tensor:
        """BitDanceT2IPipeline.decode_image: tokens fp32 [B, h*w, zc] in patch-raster order -> image."""
        lib = _lib.load()
        B, n, zc = tokens.shape
        assert n == h * w and tokens.dtype == torch.float32 and tokens.is_contiguous()
        z = torch.empty((B, h, w, zc), dtype=torch.bfloat16, device=self.device)
        check(lib.bd_tokens_to_grid(ptr(tokens), ptr(z), B, h, w, zc, ps, stream_ptr()), "bd_tokens_to_grid")
        return self.decode_grid(z)
