"""The ImageNet class-conditional generator on the B200-native ops: ``BitDance.sample``
(imagenet_gen/src/model_parallel.py:372-419), SURVEY.md section 8 rows a16 / f1.

  reference, per AR position i                                  here
  ----------------------------------------------------------    -----------------------------------------------------------
  i == 0: forward_model(cat[cls tokens, query tokens])  :383    two passes of ``bd_llm_forward`` (variant BD_LLM_ROPE_PAIRS):
          under the block-causal additive mask :90-101          the cls_token_num - 1 leading tokens causally, then the first
                                                                 block (last cls token + parallel_num - 1 queries) bidirectionally
                                                                 — the same arithmetic as the masked single pass
  i > 0:  proj_in(last_pred) (SwiGLU MLPConnector :67-77)       two tcgen05 GEMMs (SwiGLU in the first epilogue)
          forward_model(x, mask, start, end) :386-392           one ``bd_llm_forward`` over all 2B sequences: emb_norm, then per
          static KV cache, 2-D interleaved RoPE                  layer RMSNorm / fused qkv GEMM / pair-RoPE + KV append / paged
          (layers_parallel.py:94-168,255-290)                    attention / wo (+res) / RMSNorm / w1 SwiGLU / w2 (+res), final
                                                                 RMSNorm with ``+ pos_for_diff[block]`` fused
  head_sample: linear CFG ramp 1 + (cfg - 1) i / n  :352-369    ``bd_head_sample`` per chunk of sequences (head_dim 64, no output
          DiffHead.sample -> sign                                sigmoid); un-guided (cfg_iter <= 1, all 2B rows independent) at
                                                                 i == 0 exactly as the reference does
  unpatchify_raster -> vae.decode :405-416                      ``bd_tokens_to_grid`` + the tcgen05 decoder

Precision: the reference runs fp32 weights under ``torch.amp.autocast("cuda", bfloat16)``
(sample_ddp_parallel.py:158): Linears in bf16 with fp32 accumulation, RMSNorm / residual stream / RoPE in fp32. Here the
RMSNorm weights are stored in bf16 (<= 1 bf16 ulp before the rounding the consuming Linear applies anyway).
"""
from __future__ import annotations

import ctypes as C
import math

import os

import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr
from .head import HeadRunner
from .llm import LlmRunner

MODELS = {  # imagenet_gen/src/model_parallel.py:437-471
    "BitDance-B": dict(n_layer=24, n_head=12, dim=768, diff_layers=6, diff_dim=768, diff_adanln_layers=2),
    "BitDance-L": dict(n_layer=32, n_head=16, dim=1024, diff_layers=8, diff_dim=1024, diff_adanln_layers=2),
    "BitDance-H": dict(n_layer=40, n_head=20, dim=1280, diff_layers=12, diff_dim=1280, diff_adanln_layers=3),
}


def ffn_hidden(dim: int) -> int:
    """FeedForward hidden size (layers_parallel.py:171-180): find_multiple(int(2 * 4 * dim / 3), 256)."""
    h = int(2 * 4 * dim / 3)
    return h if h % 256 == 0 else h + 256 - (h % 256)


def imagenet_spec(cfg: dict) -> dict:
    """State-dict spec of the reference ``BitDance`` module without its ``vae.`` entries."""
    dim, L, pn, cls = cfg["dim"], cfg["n_layer"], cfg["parallel_num"], cfg["cls_token_num"]
    lat = cfg["latent_dim"] * cfg["patch_size"] ** 2
    hw = (cfg["resolution"] // (cfg["down_size"] * cfg["patch_size"])) ** 2
    hid, ph = ffn_hidden(dim), int(dim * 1.5)
    s = {"cls_embedding.weight": (cfg["num_classes"] + 1, dim * cls), "query_token": (1, pn - 1, dim),
         "proj_in.w1.weight": (2 * ph, lat), "proj_in.w1.bias": (2 * ph,), "proj_in.w2.weight": (dim, ph),
         "proj_in.w2.bias": (dim,), "emb_norm.weight": (dim,), "norm.weight": (dim,), "pos_for_diff.weight": (hw, dim)}
    for i in range(L):
        p = f"layers.{i}."
        s[p + "attention.wqkv.weight"] = (3 * dim, dim)
        s[p + "attention.wo.weight"] = (dim, dim)
        s[p + "feed_forward.w1.weight"] = (2 * hid, dim)
        s[p + "feed_forward.w2.weight"] = (dim, hid)
        s[p + "attention_norm.weight"] = (dim,)
        s[p + "ffn_norm.weight"] = (dim,)
    from .head import head_spec
    for k, v in head_spec(lat, dim, cfg["diff_dim"], cfg["diff_layers"], cfg["diff_adanln_layers"], True, prefix="head.net.").items():
        s[k] = v
    return s


def _get_2d_pos(resolution: int, patch: int) -> torch.Tensor:
    P = max(resolution // patch, 1)
    centers = (torch.arange(P, dtype=torch.float32) + 0.5) * (float(resolution // patch) / P)
    gy, gx = torch.meshgrid(centers, centers, indexing="ij")
    return torch.stack([gx.reshape(-1), gy.reshape(-1)], dim=1)


def rope_tables_2d(cfg: dict):
    """precompute_freqs_cis_2d + the patch-raster reorder of BitDance.__init__ (layers_parallel.py:257-272,
    model_parallel.py:202-211): cos / sin fp32 [cls + h*w - 1, head_dim / 2]."""
    dim, n_head, pn, cls = cfg["dim"], cfg["n_head"], cfg["parallel_num"], cfg["cls_token_num"]
    h = w = cfg["resolution"] // (cfg["down_size"] * cfg["patch_size"])
    n_elem = dim // n_head
    half = n_elem // 2
    freqs = 1.0 / (10000 ** (torch.arange(0, half, 2)[: half // 2].float() / half))
    t = _get_2d_pos(cfg["resolution"], cfg["down_size"] * cfg["patch_size"]) + 1.0
    t = torch.cat([torch.zeros((cls + pn - 1, 2)), t], dim=0)
    fr = torch.outer(t.flatten(), freqs).view(t.shape[0], -1)          # [pos, n_elem / 2]
    if cfg.get("parallel_mode", "patch") == "patch":
        p = int(pn ** 0.5)
        tail = fr[-h * w:].reshape(h // p, p, w // p, p, -1).permute(0, 2, 1, 3, 4).reshape(h * w, -1)
        fr = torch.cat([fr[:-h * w], tail], dim=0)
    fr = fr[:-pn]
    return torch.cos(fr), torch.sin(fr), h, w


class ImageNetEngine:
    """Prepacked weights + KV pool for one class-conditional BitDance model on one device."""

    MAX_SEQ_PER_PASS = 256   # bd_llm_forward: device-side sequence lengths for up to 256 sequences per call

    def __init__(self, state_dict: dict, cfg: dict, ae=None, device="cuda", head_rows_per_call: int = 8192,
                 head_engines: int | None = None):
        """state_dict: the reference ``BitDance`` keys (``layers.N.attention.wqkv.weight`` ...), ``vae.*`` ignored;
        cfg: dim, n_layer, n_head, diff_layers, diff_dim, diff_adanln_layers, latent_dim, down_size, patch_size, resolution,
        cls_token_num, num_classes, parallel_num, parallel_mode, time_shift; ae: an ``AERunner`` (or None: sample() then
        returns the latent grid).
        head_engines: how many persistent head engines run side by side, each on num_SMs / head_engines SMs and one 128-row
        tile of the batch (0 / 1: the multi-kernel path for the whole batch). Default 16 (BD_IMAGENET_HEAD_ENGINES): the
        sampler's weights (80 MB for BitDance-B) live in L2, the work per Linear is a few microseconds, and the multi-kernel
        path spends its time between kernels."""
        self.cfg, self.device, self.ae = dict(cfg), torch.device(device), ae
        dev = self.device
        dim, L, H = cfg["dim"], cfg["n_layer"], cfg["n_head"]
        hid = ffn_hidden(dim)
        sd = state_dict
        cos, sin, self.h, self.w = rope_tables_2d(cfg)
        self.pn, self.cls = cfg["parallel_num"], cfg["cls_token_num"]
        self.ps = int(self.pn ** 0.5)
        self.total = self.h * self.w + self.cls
        pages = (self.total + 63) // 64
        pad = pages * 64 - cos.shape[0]
        cos = torch.cat([cos, torch.ones(pad, cos.shape[1])], dim=0)      # positions beyond the last one are never used
        sin = torch.cat([sin, torch.zeros(pad, sin.shape[1])], dim=0)
        hf = {}
        for i in range(L):
            s, d = f"layers.{i}.", f"model.layers.{i}."
            q, k, v = sd[s + "attention.wqkv.weight"].chunk(3, dim=0)
            hf[d + "self_attn.q_proj.weight"], hf[d + "self_attn.k_proj.weight"], hf[d + "self_attn.v_proj.weight"] = q, k, v
            hf[d + "self_attn.o_proj.weight"] = sd[s + "attention.wo.weight"]
            g, u = sd[s + "feed_forward.w1.weight"].chunk(2, dim=0)      # silu(h1) * h2
            hf[d + "mlp.gate_proj.weight"], hf[d + "mlp.up_proj.weight"] = g, u
            hf[d + "mlp.down_proj.weight"] = sd[s + "feed_forward.w2.weight"]
            hf[d + "input_layernorm.weight"] = sd[s + "attention_norm.weight"]
            hf[d + "post_attention_layernorm.weight"] = sd[s + "ffn_norm.weight"]
        hf["model.norm.weight"] = sd["norm.weight"]
        llm_cfg = dict(hidden_size=dim, intermediate_size=hid, num_hidden_layers=L, num_attention_heads=H,
                       num_key_value_heads=H, head_dim=dim // H, rms_norm_eps=1e-6, rope_theta=1e4)
        self.llm = LlmRunner(hf, llm_cfg, device=dev, stream=False, qk_norm=False, rope_pairs=(cos, sin),
                             emb_norm=sd["emb_norm.weight"])
        del hf
        lat = cfg["latent_dim"] * cfg["patch_size"] ** 2
        head_sd = {"net." + k[len("head.net."):]: v for k, v in sd.items() if k.startswith("head.net.")}
        if head_engines is None:
            head_engines = int(os.environ.get("BD_IMAGENET_HEAD_ENGINES", "16"))
        n_sm = ops.stream_num_ctas()
        self.head_engines = max(0, min(int(head_engines), n_sm))
        eng_ctas = n_sm // self.head_engines if self.head_engines > 1 else None
        self.head = HeadRunner(head_sd, ch_target=lat, ch_cond=dim, ch_latent=cfg["diff_dim"], depth_latent=cfg["diff_layers"],
                               depth_adanln=cfg["diff_adanln_layers"], use_swiglu=True, head_dim=64, out_sigmoid=False,
                               time_shift=cfg.get("time_shift", 1.0), device=dev, tiled=True,
                               stream=self.head_engines > 1, stream_ctas=eng_ctas)
        if self.head.w_stream is None:
            self.head_engines = 0
        self._side = [torch.cuda.Stream(device=dev) for _ in range(self.head_engines if self.head_engines > 1 else 0)]
        self.lat = lat
        bf = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()
        ph = int(dim * 1.5)
        w1, b1 = ops.interleave16(bf(sd["proj_in.w1.weight"][:ph]), bf(sd["proj_in.w1.weight"][ph:]),
                                  bf(sd["proj_in.w1.bias"][:ph]), bf(sd["proj_in.w1.bias"][ph:]))
        self.p_w1, self.p_b1 = ops.pack_weight(w1), b1          # tile-major (K = latent bits is padded to one k-block)
        self.p_w2, self.p_b2 = ops.pack_weight(bf(sd["proj_in.w2.weight"])), bf(sd["proj_in.w2.bias"])
        self.cls_embedding = sd["cls_embedding.weight"].detach().to(dev, torch.float32).contiguous()
        self.query_token = sd["query_token"].detach().to(dev, torch.float32).contiguous()
        self.pos_for_diff = sd["pos_for_diff.weight"].detach().to(dev, torch.float32).contiguous()
        self.head_rows_per_call = head_rows_per_call
        self._kv = None

    # ---- pieces ----------------------------------------------------------------------------------------------------
    def _cache(self, R):
        if self._kv is None or self._kv.page_table.shape[0] < R:
            self._kv = self.llm.new_cache(R, self.total)
        c = self._kv
        c.seq_lens.zero_()
        c.host_lens = [0] * c.page_table.shape[0]
        return c

    def _forward(self, x, cache, *, causal, out_add=None):
        """x fp32 [R, S, dim] (overwritten) -> final-norm output fp32 [R, S, dim] (+ out_add rows)."""
        R = x.shape[0]
        outs = []
        for r0 in range(0, R, self.MAX_SEQ_PER_PASS):
            r1 = min(R, r0 + self.MAX_SEQ_PER_PASS)
            outs.append(self.llm.forward(x[r0:r1].contiguous(), cache, r0, r1 - r0, causal=causal, out_add=out_add,
                                         out_add_mod=self.pn if out_add is not None else 0))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def _proj_in(self, tok_bf):
        """MLPConnector (model_parallel.py:67-77): w2(silu(h1) * h2) with biases; [rows, lat] bf16 -> fp32 [rows, dim]
        (the bf16 Linear output, widened: the residual stream is fp32)."""
        h = ops.gemm(tok_bf, self.p_w1, bias=self.p_b1, swiglu=True)
        return ops.gemm(h, self.p_w2, bias=self.p_b2, out_dtype=torch.float32)

    def _head(self, z, cfg_iter, S, noise):
        """z fp32 [R, pn, dim]; guided (cfg_iter > 1): rows = [cond | uncond], returns [R/2, pn, lat]; else [R, pn, lat].
        Sequences are chunked so that one ``bd_head_sample`` call sees at most head_rows_per_call token rows."""
        guided = cfg_iter > 1.0
        R = z.shape[0]
        n = R // 2 if guided else R
        rows_per_seq = self.pn * (2 if guided else 1)
        if self._side and rows_per_seq <= 128 and S + 1 <= 104:
            # persistent engines side by side: 128-row tiles of the batch, tile i on CUDA stream / workspace i mod engines
            per = 128 // rows_per_seq
            cur = torch.cuda.current_stream(self.device)
            outs, used = [], set()
            for i, b0 in enumerate(range(0, n, per)):
                b1 = min(n, b0 + per)
                k = i % len(self._side)
                st = self._side[k]
                if k not in used:
                    st.wait_stream(cur)
                    used.add(k)
                with torch.cuda.stream(st):   # (z and noise were complete before the first wait_stream above)
                    nz = None if noise is None else noise[:, b0:b1].contiguous()
                    zc = torch.cat([z[b0:b1], z[n + b0:n + b1]], dim=0) if guided else z[b0:b1]
                    o = self.head.sample(zc.contiguous(), cfg_iter if guided else 1.0, S, noise=nz, path="stream", slot=k)
                o.record_stream(cur)
                outs.append(o)
            for k in used:
                cur.wait_stream(self._side[k])
            return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
        per = max(1, self.head_rows_per_call // rows_per_seq)
        outs = []
        for b0 in range(0, n, per):
            b1 = min(n, b0 + per)
            zc = torch.cat([z[b0:b1], z[n + b0:n + b1]], dim=0) if guided else z[b0:b1]
            nz = None if noise is None else noise[:, b0:b1].contiguous()
            outs.append(self.head.sample(zc.contiguous(), cfg_iter if guided else 1.0, S, noise=nz, path="tiled"))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    # ---- BitDance.sample --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample_tokens(self, class_ids: torch.Tensor, sample_steps: int, cfg_scale: float = 1.0, cfg_schedule: str = "linear",
                      noise=None):
        """-> (tokens fp32 [B, h*w, lat] in generation order, packed bits int32 [B, h*w, lat/32] when lat % 32 == 0).
        noise: optional list over AR positions of fp32 [S+1, rows, pn, lat] (rows = B when that position is guided, else
        all sequences) — tests; by default drawn like the reference sampler does."""
        lib = _lib.load()
        dev, pn, cls, dim, lat = self.device, self.pn, self.cls, self.cfg["dim"], self.lat
        cond = class_ids.to(dev).long()
        guided_run = cfg_scale > 1.0
        if guided_run:
            cond = torch.cat([cond, torch.full_like(cond, self.cfg["num_classes"])])
        R = cond.shape[0]
        act = R // 2 if guided_run else R
        steps = self.h * self.w // pn
        cache = self._cache(R)
        c = self.cls_embedding[cond].view(R, cls, dim)
        tokens = torch.zeros((act, self.h * self.w, lat), dtype=torch.float32, device=dev)
        last = None
        for i in range(steps):
            pos = self.pos_for_diff[i * pn:(i + 1) * pn].contiguous()
            if i == 0:
                if cls > 1:
                    self._forward(c[:, :cls - 1].contiguous(), cache, causal=True)
                x0 = torch.cat([c[:, cls - 1:], self.query_token.expand(R, -1, -1)], dim=1).contiguous()
                z = self._forward(x0, cache, causal=False, out_add=pos)
            else:
                x = self._proj_in(last.view(R * pn, lat)).view(R, pn, dim)
                z = self._forward(x, cache, causal=False, out_add=pos)
            if guided_run:
                cfg_iter = cfg_scale if cfg_schedule == "constant" else 1.0 + (cfg_scale - 1.0) * i / steps
                if cfg_schedule not in ("constant", "linear"):
                    raise NotImplementedError(f"unknown cfg_schedule {cfg_schedule}")
            else:
                cfg_iter = 1.0
            pred = self._head(z, cfg_iter, sample_steps, None if noise is None else noise[i])
            if cfg_iter > 1.0:
                pred = torch.cat([pred, pred], dim=0)      # euler_maruyama returns cat([x] * cfg_mult)
            last = torch.sign(pred).to(torch.bfloat16)     # LFQ (head_sample :367): the next proj_in input, all R rows
            tokens[:, i * pn:(i + 1) * pn] = last[:act].float()
        packed = None
        if lat % 32 == 0:
            _, packed = ops.sign_tokens(tokens.contiguous(), want_tokens=False)
        return tokens, packed

    def tokens_to_grid(self, tokens):
        """unpatchify_raster / unpatchify (utils.py:82-94, model_parallel.py:263-272): [B, h*w, lat] -> [B, C, H, W]."""
        B = tokens.shape[0]
        h, w, p = self.h, self.w, self.ps
        if self.cfg.get("parallel_mode", "patch") == "patch":
            return tokens.view(B, h // p, w // p, p, p, self.lat).permute(0, 5, 1, 3, 2, 4).reshape(B, self.lat, h, w)
        q = self.cfg["patch_size"]
        return tokens.reshape(B, h, w, self.cfg["latent_dim"], q, q).permute(0, 3, 1, 4, 2, 5).reshape(
            B, self.cfg["latent_dim"], h * q, w * q)

    @torch.no_grad()
    def sample(self, class_ids, sample_steps, cfg_scale=1.0, cfg_schedule="linear", chunk_size=0):
        tokens, _ = self.sample_tokens(class_ids, sample_steps, cfg_scale, cfg_schedule)
        grid = self.tokens_to_grid(tokens).contiguous()
        if self.ae is None:
            return grid
        if chunk_size and chunk_size > 0:   # decode_in_chunks (:421-431): chunks land on the host
            return torch.cat([self.ae.decode(grid[i:i + chunk_size]).float().cpu() for i in range(0, grid.shape[0], chunk_size)])
        return self.ae.decode(grid)
