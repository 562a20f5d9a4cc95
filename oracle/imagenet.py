"""Oracle (torch, CPU): the ImageNet class-conditional generator ``BitDance.sample``. TEST INFRASTRUCTURE ONLY.

Functional restatement (state-dict in, tensors out) of SURVEY.md §8 row a16 — the reference files are
  * ``imagenet_gen/src/model_parallel.py``: ``BitDance.sample`` :372-419, ``forward_model`` :343-350, ``head_sample``
    :352-369 (linear CFG ramp), ``MLPConnector`` :67-77 (SwiGLU), ``get_block_causal_mask`` :90-101, buffers :202-217;
  * ``imagenet_gen/src/layers_parallel.py``: ``Attention.forward`` / ``naive_attention`` :120-168 (static KV cache
    :94-118), ``FeedForward`` :171-185, ``TransformerBlock.forward_onestep`` :226-238, ``get_2d_pos`` :241-254,
    ``precompute_freqs_cis_2d`` :257-272, ``apply_rotary_emb`` :275-290 (interleaved pairs, fp32);
  * ``imagenet_gen/src/utils.py``: ``patchify_raster_2d`` :96-112, ``unpatchify_raster`` :82-94.
The diffusion head is the same network as the T2I one with ``head_dim=64`` and no output sigmoid
(``imagenet_gen/src/diff_head_parallel.py``), i.e. ``oracle/head.py`` with those two switches.

``rnd=ident``: exact fp32 math, pinned bit-exactly (token grid) against the unmodified reference in
``tests/test_oracle_vs_reference.py``. ``rnd=bf16``: the policy of the reference's deployment — fp32 weights under
``torch.amp.autocast("cuda", bfloat16)`` (sample_ddp_parallel.py:158): every Linear / matmul rounds its inputs and output
to bf16 (fp32 accumulate), RMSNorm, the residual stream, RoPE, the additive mask and softmax run in fp32, ``q * scale``
and ``apply_rotary_emb(...).type_as(x)`` round to bf16; the KV cache holds bf16 values.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import head as oh


def get_2d_pos(resolution: int, patch_size: int) -> torch.Tensor:
    """layers_parallel.get_2d_pos (num_scales = 1): centres (x, y) of the P x P token grid, row-major."""
    P = max(resolution // patch_size, 1)
    max_pos = resolution // patch_size
    centers = (torch.arange(P, dtype=torch.float32) + 0.5) * (float(max_pos) / P)
    gy, gx = torch.meshgrid(centers, centers, indexing="ij")
    return torch.stack([gx.reshape(-1), gy.reshape(-1)], dim=1)


def freqs_cis_2d(pos_2d: torch.Tensor, n_elem: int, base: float, cls_token_num: int) -> torch.Tensor:
    """precompute_freqs_cis_2d: half of head_dim rotates with x, half with y; the leading cls/query positions are 0.
    Returns [cls_token_num + P*P, n_elem // 2, 2] (cos, sin)."""
    half = n_elem // 2
    freqs = 1.0 / (base ** (torch.arange(0, half, 2)[: half // 2].float() / half))
    t = pos_2d + 1.0
    if cls_token_num > 0:
        t = torch.cat([torch.zeros((cls_token_num, 2)), t], dim=0)
    fr = torch.outer(t.flatten(), freqs).view(*t.shape[:-1], -1)
    return torch.stack([torch.cos(fr), torch.sin(fr)], dim=-1)


def patchify_raster_2d(x: torch.Tensor, p: int, H: int, W: int) -> torch.Tensor:
    """[H*W, C1, C2] in row-major token order -> p x p patch-raster order (utils.patchify_raster_2d)."""
    N, C1, C2 = x.shape
    y = x.reshape(H // p, p, W // p, p, C1 * C2).permute(0, 2, 1, 3, 4).reshape(N, C1, C2)
    return y


def unpatchify_raster(x: torch.Tensor, p: int, hw) -> torch.Tensor:
    """[B, N, C] in patch-raster order -> [B, C, H, W] (utils.unpatchify_raster)."""
    B, N, C = x.shape
    H, W = hw
    return x.view(B, H // p, W // p, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)


def apply_rotary(x: torch.Tensor, fc: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb: x [B, S, H, hd] as interleaved (even, odd) pairs; fc [S, hd/2, 2]; fp32."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    f = fc.view(1, xs.size(1), 1, xs.size(3), 2)
    out = torch.stack([xs[..., 0] * f[..., 0] - xs[..., 1] * f[..., 1],
                       xs[..., 1] * f[..., 0] + xs[..., 0] * f[..., 1]], dim=-1)
    return out.flatten(3).type_as(x)


def block_causal_mask(total: int, causal: int, block: int) -> torch.Tensor:
    """get_block_causal_mask: additive mask, causal everywhere, full inside each block of `block` tokens after the
    first `causal` tokens."""
    assert (total - causal) % block == 0
    m = torch.zeros(total, total)
    m.masked_fill_(torch.triu(torch.ones(total, total), diagonal=1).bool(), float("-inf"))
    for i in range(causal, total, block):
        m[i:i + block, i:i + block] = 0
    return m


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """torch.nn.RMSNorm(dim, eps, elementwise_affine=True)."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def make_buffers(cfg: dict):
    """The registered buffers of BitDance.__init__ (:202-217): freqs_cis (patch-raster order, last block dropped) and the
    additive block-causal mask."""
    h = w = cfg["resolution"] // (cfg["down_size"] * cfg["patch_size"])
    pn, cls = cfg["parallel_num"], cfg["cls_token_num"]
    fc = freqs_cis_2d(get_2d_pos(cfg["resolution"], cfg["down_size"] * cfg["patch_size"]), cfg["dim"] // cfg["n_head"],
                      10000, cls + pn - 1)
    if cfg.get("parallel_mode", "patch") == "patch":
        fc[-h * w:] = patchify_raster_2d(fc[-h * w:], int(pn ** 0.5), h, w)
    mask = block_causal_mask(h * w + cls - 1, cls - 1, pn)
    return fc[:-pn], mask, h, w


def _lin(x, w, b, rnd):
    y = rnd(x) @ rnd(w.float()).t()
    return rnd(y + rnd(b.float())) if b is not None else rnd(y)


def block_onestep(sd, prefix, x, mask, fc, cache, start, end, n_head, rnd=oh.ident):
    """TransformerBlock.forward_onestep with the static KV cache (keys / values written at [start, end))."""
    B, S, dim = x.shape
    hd = dim // n_head
    a = rmsnorm(x, sd[prefix + "attention_norm.weight"])
    q, k, v = _lin(a, sd[prefix + "attention.wqkv.weight"], None, rnd).chunk(3, dim=-1)
    q, k, v = (t.view(B, S, n_head, hd) for t in (q, k, v))
    q, k = rnd(apply_rotary(q, fc)), rnd(apply_rotary(k, fc))
    q, k, v = (t.transpose(1, 2) for t in (q, k, v))
    cache[0][:, :, start:end] = k
    cache[1][:, :, start:end] = v
    keys, values = cache[0][:, :, :end], cache[1][:, :, :end]
    att = rnd(rnd(q * hd ** -0.5) @ keys.transpose(-1, -2))
    if att.shape[-2] > 1:
        att = att + mask
    o = rnd(rnd(torch.softmax(att.float(), dim=-1)) @ values).transpose(1, 2).reshape(B, S, dim)
    h = x + _lin(o, sd[prefix + "attention.wo.weight"], None, rnd)
    f = rmsnorm(h, sd[prefix + "ffn_norm.weight"])
    h1, h2 = _lin(f, sd[prefix + "feed_forward.w1.weight"], None, rnd).chunk(2, dim=-1)
    return h + _lin(rnd(rnd(F.silu(h1)) * h2), sd[prefix + "feed_forward.w2.weight"], None, rnd)


def forward_model(sd, cfg, x, mask, fc, caches, start, end, rnd=oh.ident):
    x = rmsnorm(x.float(), sd["emb_norm.weight"])
    for i in range(cfg["n_layer"]):
        x = block_onestep(sd, f"layers.{i}.", x, mask, fc[start:end], caches[i], start, end, cfg["n_head"], rnd)
    return rmsnorm(x, sd["norm.weight"])


def proj_in(sd, x, rnd=oh.ident):
    """MLPConnector (SwiGLU, with biases)."""
    h1, h2 = _lin(x, sd["proj_in.w1.weight"], sd["proj_in.w1.bias"], rnd).chunk(2, dim=-1)
    return _lin(rnd(rnd(F.silu(h1)) * h2), sd["proj_in.w2.weight"], sd["proj_in.w2.bias"], rnd)


def sample(sd: dict, cfg: dict, class_ids: torch.Tensor, sample_steps: int, cfg_scale: float, noise, *,
           cfg_schedule: str = "linear", rnd=oh.ident, trace: list | None = None):
    """BitDance.sample up to (not including) ``vae.decode``.

    sd: the reference state dict (keys as in BitDance, without the ``vae.`` entries); cfg: dim, n_layer, n_head,
    resolution, down_size, patch_size, cls_token_num, parallel_num, num_classes, latent_dim, parallel_mode.
    noise: per AR position a list [x0, eps_0 .. eps_{S-1}] in the reference's torch.randn call order.
    Returns (tokens [B, h*w, C] in generation order, +-1 / 0, and the latent grid [B, C, h*p, w*p])."""
    fc, mask, h, w = make_buffers(cfg)
    pn, cls, dim = cfg["parallel_num"], cfg["cls_token_num"], cfg["dim"]
    cond = class_ids
    if cfg_scale > 1.0:
        cond = torch.cat([class_ids, torch.ones_like(class_ids) * cfg["num_classes"]])
    B = cond.shape[0]
    act = B // 2 if cfg_scale > 1.0 else B
    total = h * w + cls
    hd = dim // cfg["n_head"]
    caches = [[torch.zeros(B, cfg["n_head"], total, hd), torch.zeros(B, cfg["n_head"], total, hd)]
              for _ in range(cfg["n_layer"])]
    c = sd["cls_embedding.weight"][cond].view(B, cls, dim)
    head_sd = {"net." + k[len("head.net."):]: v for k, v in sd.items() if k.startswith("head.net.")}
    steps = h * w // pn
    preds = []
    last = None
    for i in range(steps):
        if i == 0:
            n0 = cls + pn - 1
            x = forward_model(sd, cfg, torch.cat([c, sd["query_token"].repeat(B, 1, 1)], dim=1), mask[:n0, :n0], fc,
                              caches, 0, n0, rnd)
        else:
            start = pn * (i - 1) + cls + pn - 1
            x = forward_model(sd, cfg, proj_in(sd, last, rnd), mask[start:start + pn, :start + pn], fc, caches, start,
                              start + pn, rnd)
        z = x[:, -pn:, :] + sd["pos_for_diff.weight"][i * pn:(i + 1) * pn, :]
        if trace is not None:
            trace.append(dict(z=z.clone()))
        if cfg_scale > 1.0:
            cfg_iter = cfg_scale if cfg_schedule == "constant" else 1.0 + (cfg_scale - 1.0) * i / steps
        else:
            cfg_iter = 1.0
        # head.sample -> euler_maruyama: with cfg_iter <= 1 the reference runs all B rows un-guided
        pred = oh.euler_maruyama(head_sd, z, cfg_iter, sample_steps, noise[i], head_dim=64, out_sigmoid=False, rnd=rnd)
        if cfg_iter > 1.0:
            pass  # euler_maruyama returns cat[x] * 2 already
        last = torch.sign(pred)
        if trace is not None:
            trace[-1]["last"] = last.clone()
        preds.append(last)
    tokens = torch.cat(preds, dim=-2)[:act]
    if cfg.get("parallel_mode", "patch") == "patch":
        grid = unpatchify_raster(tokens, int(pn ** 0.5), (h, w))
    else:
        p = cfg["patch_size"]
        grid = tokens.reshape(act, h, w, cfg["latent_dim"], p, p).permute(0, 3, 1, 4, 2, 5).reshape(
            act, cfg["latent_dim"], h * p, w * p)
    return tokens, grid
