"""Host side of the binary-diffusion vision head: weight prepack + ``bd_head_sample`` launch.

Mirrors ``DiffHead.sample`` (modeling/vision_head/flow_head_parallel_x.py:107-120): the only torch work left on the
path is drawing the noise with the SAME call sequence as the reference sampler (one ``randn`` for x0, then one per
stochastic step; sampling_x.py:60,40) so that a seeded run consumes the CUDA Philox stream identically.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops
from ._lib import check, note, ptr, stream_ptr

MAX_BLOCKS = 16



class HeadBlock(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "norm1_w", "norm1_b", "norm2_w", "norm2_b",
        "wqkv_w", "wqkv_b", "wo_w", "wo_b", "w1_w", "w1_b", "w2_w", "w2_b")]


class HeadWeights(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("D", "Dz", "C", "hidden", "n_blocks", "n_ada", "head_dim", "use_swiglu", "w_tiled",
                                "out_sigmoid", "stream_ctas", "reserved_")]
        + [(n, C.c_void_p) for n in ("input_proj_w", "input_proj_b", "time0_w", "time0_b", "time2_w", "time2_b",
                                     "cond_w", "cond_b", "ada_w", "ada_b", "final_w", "final_b")]
        + [("blocks", HeadBlock * MAX_BLOCKS)]
    )


def head_spec(ch_target, ch_cond, ch_latent, depth_latent, depth_adanln, use_swiglu=True, prefix="net."):
    """State-dict spec of the reference DiffHead (names as in flow_head_parallel_x.py:254-300)."""
    D, hid = ch_latent, int(ch_latent * 1.5)
    s = {}

    def lin(name, o, i):
        s[prefix + name + ".weight"] = (o, i)
        s[prefix + name + ".bias"] = (o,)

    lin("time_embed.mlp.0", D, 256)
    lin("time_embed.mlp.2", D, D)
    lin("cond_embed", D, ch_cond)
    lin("input_proj", D, ch_target)
    for i in range(depth_latent):
        b = f"res_blocks.{i}."
        s[prefix + b + "norm1.weight"] = (D,)
        s[prefix + b + "norm1.bias"] = (D,)
        lin(b + "attn.wqkv", 3 * D, D)
        lin(b + "attn.wo", D, D)
        s[prefix + b + "norm2.weight"] = (D,)
        s[prefix + b + "norm2.bias"] = (D,)
        if use_swiglu:
            lin(b + "w1", 2 * hid, D)
            lin(b + "w2", D, hid)
        else:
            lin(b + "mlp.0", hid, D)
            lin(b + "mlp.2", D, hid)
    for i in range(depth_adanln):
        lin(f"ada_ln_blocks.{i}", 6 * D, D)
    lin("final_layer.ada_ln_modulation", 2 * D, D)
    lin("final_layer.linear", ch_target, D)
    return s


def sampler_schedule(num_sampling_steps: int, time_shift: float = 1.0, last_step_size: float = 0.05, device="cpu"):
    """Per-evaluation fp32 scalars [S+1, 8] = {t, dt, clamp(1-t,.05), var, 1-t, sqrt(2(1-t)dt), 0, 0}, produced with
    the same torch fp32 tensor ops (and on the same device) as sampling_x.py:62-68,82,6-14,39 — t is the running sum
    of dt, not linspace[i]."""
    S = num_sampling_steps
    t_all = torch.linspace(0, 1 - last_step_size, S + 1, device=device, dtype=torch.float32)
    t_all = (1 / time_shift) / ((1 / time_shift) + (1 / t_all - 1) ** 1.0)
    dt = t_all[1:] - t_all[:-1]
    t = torch.tensor(0.0, device=device, dtype=torch.float32)
    rows = torch.zeros(S + 1, 8, dtype=torch.float32, device=device)
    for i in range(S):
        sigma = 1 - t
        var = sigma ** 2 - (t / 1) * (-1) * sigma
        rows[i, 0] = t
        rows[i, 1] = dt[i]
        rows[i, 2] = (1 - t).clamp_min(0.05)
        rows[i, 3] = var
        rows[i, 4] = 1 - t
        rows[i, 5] = (2.0 * (1.0 - t) * dt[i]) ** 0.5
        t = t + dt[i]
    tl = torch.full((), 1 - last_step_size, device=device, dtype=torch.float32)
    rows[S, 0] = tl
    rows[S, 1] = last_step_size
    rows[S, 2] = (1 - tl).clamp_min(0.05)
    rows[S, 3] = 1.0
    rows[S, 4] = 1 - tl
    return rows.cpu().contiguous()


class HeadRunner:
    """Prepacked weights + workspace for one DiffHead on one device."""

    def __init__(self, state_dict: dict, *, ch_target, ch_cond, ch_latent, depth_latent, depth_adanln, use_swiglu,
                 head_dim=128, out_sigmoid=True, time_shift=1.0, device="cuda", prefix="net.", stream=True, tiled=True,
                 stream_ctas: int | None = None):
        """stream: also pack the weights stream-major for the persistent kernel (used whenever B*cfg_mult*pn <= 128);
        tiled: keep the tile-major copy for the multi-kernel path (larger batches);
        stream_ctas: grid of one persistent engine (default: every SM). A small model whose weights live in L2 is better
        served by several engines side by side, each on its share of the SMs and its own 128-row tile (``sample(...,
        slot=i)`` on CUDA stream i): the ImageNet class-conditional sampler."""
        assert depth_latent <= MAX_BLOCKS
        self.device = torch.device(device)
        self.cfg = dict(C=ch_target, Dz=ch_cond, D=ch_latent, n_blocks=depth_latent, n_ada=depth_adanln)
        self.time_shift = time_shift
        self.hidden = int(ch_latent * 1.5)
        self._keep = []  # prepacked tensors (owned here; C side sees raw pointers)
        self.stream_ctas = int(stream_ctas) if stream_ctas else ops.stream_num_ctas()
        assert 1 <= self.stream_ctas <= ops.stream_num_ctas()
        dev = self.device
        D, hidden = ch_latent, self.hidden
        # the persistent kernel needs 16-row units everywhere and one 64-column k-block for the latent bits
        fits = D % 64 == 0 and ch_target <= 64 and hidden % 8 == 0 and ch_cond % 8 == 0 and D <= 6144
        if stream and not fits:
            note(f"head dims outside the persistent engine's limits (ch_latent={D}: multiple of 64 and <= 6144; "
                  f"ch_target={ch_target} <= 64): DiffHead.sample runs on the multi-kernel path")
        stream = bool(stream) and fits
        assert stream or tiled
        self.w = self._build(state_dict, prefix, "tiled", ch_target, ch_cond, depth_latent, depth_adanln, use_swiglu,
                             head_dim, out_sigmoid) if tiled else None
        self.w_stream = self._build(state_dict, prefix, "stream", ch_target, ch_cond, depth_latent, depth_adanln,
                                    use_swiglu, head_dim, out_sigmoid) if stream else None
        torch.cuda.synchronize(dev)
        self._ws = {}
        self._sched = {}

    def _build(self, state_dict, prefix, kind, ch_target, ch_cond, depth_latent, depth_adanln, use_swiglu, head_dim,
               out_sigmoid) -> HeadWeights:
        dev, D, hidden = self.device, self.cfg["D"], self.hidden
        keep = self._keep

        def raw(name, dtype=torch.bfloat16):
            return state_dict[prefix + name].detach().to(device=dev, dtype=dtype).contiguous()

        def kept(t):
            keep.append(t)
            return t.data_ptr()

        n_ctas = self.stream_ctas if kind == "stream" else 0

        def lin(wt, bias, *, ksplit=1, swiglu=False):
            """-> (weight ptr, bias ptr) in the layout of `kind`; the row-major copy is dropped"""
            if kind == "stream":
                sw = ops.stream_pack_weight(wt, bias, ksplit=ksplit, swiglu=swiglu, n_ctas=n_ctas)
                return kept(sw.data), kept(sw.bias)
            if swiglu:
                h = wt.shape[0] // 2
                wt, bias = ops.interleave16(wt[:h], wt[h:], bias[:h], bias[h:])
            return kept(ops.pack_weight(wt).data), kept(bias)

        w = HeadWeights()
        w.w_tiled = 2 if kind == "stream" else 1
        w.stream_ctas = n_ctas
        w.D, w.Dz, w.C, w.hidden = D, ch_cond, ch_target, hidden
        w.n_blocks, w.n_ada, w.head_dim = depth_latent, depth_adanln, head_dim
        w.use_swiglu, w.out_sigmoid = int(use_swiglu), int(out_sigmoid)
        for field, name in (("input_proj", "input_proj"), ("time0", "time_embed.mlp.0"), ("time2", "time_embed.mlp.2"),
                            ("cond", "cond_embed")):
            wp, bp = lin(raw(name + ".weight"), raw(name + ".bias"))
            setattr(w, field + "_w", wp)
            setattr(w, field + "_b", bp)
        w.final_w, w.final_b = kept(raw("final_layer.linear.weight")), kept(raw("final_layer.linear.bias"))
        ada_names = [f"ada_ln_blocks.{i}" for i in range(depth_adanln)] + ["final_layer.ada_ln_modulation"]
        ada_w = torch.cat([raw(n + ".weight") for n in ada_names], dim=0).contiguous()
        ada_b = torch.cat([raw(n + ".bias") for n in ada_names], dim=0).contiguous()
        w.ada_w, w.ada_b = lin(ada_w, ada_b)
        del ada_w
        ks_wo = ops.stream_ksplit(D, D, n_ctas) if kind == "stream" else 1
        ks_w2 = ops.stream_ksplit(D, hidden, n_ctas) if kind == "stream" else 1
        for i in range(depth_latent):
            b = f"res_blocks.{i}."
            blk = w.blocks[i]
            blk.norm1_w, blk.norm1_b = kept(raw(b + "norm1.weight", torch.float32)), kept(raw(b + "norm1.bias", torch.float32))
            blk.norm2_w, blk.norm2_b = kept(raw(b + "norm2.weight", torch.float32)), kept(raw(b + "norm2.bias", torch.float32))
            blk.wqkv_w, blk.wqkv_b = lin(raw(b + "attn.wqkv.weight"), raw(b + "attn.wqkv.bias"))
            # wo / w2 go through fp32 partials (split-K) on the persistent path: their bias is applied by the row op
            wp, _ = lin(raw(b + "attn.wo.weight"), None if kind == "stream" else raw(b + "attn.wo.bias"), ksplit=ks_wo)
            blk.wo_w, blk.wo_b = wp, kept(raw(b + "attn.wo.bias"))
            n1, n2 = ("w1", "w2") if use_swiglu else ("mlp.0", "mlp.2")
            blk.w1_w, blk.w1_b = lin(raw(b + n1 + ".weight"), raw(b + n1 + ".bias"), swiglu=bool(use_swiglu))
            wp, _ = lin(raw(b + n2 + ".weight"), None if kind == "stream" else raw(b + n2 + ".bias"), ksplit=ks_w2)
            blk.w2_w, blk.w2_b = wp, kept(raw(b + n2 + ".bias"))
        return w

    def schedule(self, S: int) -> torch.Tensor:
        if S not in self._sched:
            self._sched[S] = sampler_schedule(S, self.time_shift, device=self.device)
        return self._sched[S]

    def _workspace(self, w, B, pn, mult, S, slot=0):
        lib = _lib.load()
        lib.bd_head_workspace_bytes.restype = C.c_size_t
        need = lib.bd_head_workspace_bytes(C.byref(w), B, pn, mult, S)
        key = (int(w.w_tiled), int(slot))
        if key not in self._ws or self._ws[key].numel() < need:
            self._ws[key] = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws[key]

    def weights_for(self, rows: int, S: int, path: str | None = None) -> HeadWeights:
        """The persistent single-kernel path when the stream-packed weights exist and the batch is one 128-row tile."""
        if path == "tiled":
            assert self.w is not None
            return self.w
        ok = self.w_stream is not None and rows <= 128 and S + 1 <= 104
        if self.w_stream is not None and not ok and path is None:
            note(f"DiffHead.sample with {rows} rows / {S} sampling steps is outside the persistent engine's limits (one "
                  f"128-row tile, <= 103 steps): multi-kernel path")
        if path == "stream":
            assert ok, "persistent path unavailable for this shape"
        if ok:
            return self.w_stream
        assert self.w is not None, "batch needs the tile-major weights (HeadRunner(tiled=True))"
        return self.w

    def draw_noise(self, B, pn, S):
        """Same generator consumption as sampling_x.py: randn(x_shape) then S x randn_like(x)."""
        Cc = self.cfg["C"]
        noise = torch.empty((S + 1, B, pn, Cc), dtype=torch.float32, device=self.device)
        for i in range(S + 1):
            noise[i].normal_()
        return noise

    def sample(self, z: torch.Tensor, cfg: float, num_sampling_steps: int, noise: torch.Tensor | None = None,
               trace: bool = False, pdl: bool = True, path: str | None = None, slot: int = 0):
        """z: [R, pn, Dz] fp32 (cond rows then uncond rows when cfg > 1). Returns x [B, pn, C] fp32 (+ trace).
        path: None = automatic, "stream" = the persistent kernel, "tiled" = the multi-kernel path.
        slot: which private workspace to use (calls in flight at the same time on different CUDA streams need their own)."""
        lib = _lib.load()
        assert z.is_cuda and z.dim() == 3
        mult = 2 if cfg > 1.0 else 1
        R, pn, Dz = z.shape
        assert R % mult == 0 and Dz == self.cfg["Dz"]
        B = R // mult
        S = num_sampling_steps
        Cc = self.cfg["C"]
        if noise is None:
            noise = self.draw_noise(B, pn, S)
        assert noise.shape == (S + 1, B, pn, Cc) and noise.dtype == torch.float32 and noise.is_contiguous()
        zc = z.to(torch.float32).contiguous()
        out = torch.empty((B, pn, Cc), dtype=torch.float32, device=self.device)
        tr = torch.empty((S + 1, R * pn, Cc), dtype=torch.float32, device=self.device) if trace else None
        w = self.weights_for(R * pn, S, path)
        ws = self._workspace(w, B, pn, mult, S, slot)
        sched = self.schedule(S)
        st = lib.bd_head_sample(C.byref(w), ptr(zc), ptr(noise), C.c_void_p(sched.data_ptr()), B, pn, mult,
                                C.c_float(cfg), S, ptr(out), ptr(tr), ptr(ws), C.c_size_t(ws.numel()),
                                1 if pdl else 0, stream_ptr())
        check(st, "bd_head_sample")
        return (out, tr) if trace else out
