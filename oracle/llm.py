"""Oracle (torch, CPU): Qwen3 decoder stack with a KV cache. TEST INFRASTRUCTURE ONLY.

The LLM arithmetic of the reference lives in a third-party dependency: ``transformers==4.57.0`` (requirements.txt:1),
class ``transformers.models.qwen3.modeling_qwen3.Qwen3Model``; call sites modeling/t2i_pipeline.py:199,211,224,229,
261,266. This file restates the published algorithm (RMSNorm in fp32; q/k RMSNorm over head_dim before RoPE;
rotate_half RoPE with inv_freq = theta^(-2k/d); GQA by repeating KV heads; softmax(QK^T/sqrt(d)) V; SwiGLU MLP;
pre-norm residuals; final RMSNorm) and is pinned against the installed transformers (5.5.0) Qwen3Model on CPU in
tests/test_oracle_vs_reference.py.

Mask semantics of the reference's calls: the first prefill call is causal; every later call passes an all-ones
boolean mask (t2i_pipeline.py:206-210,256-260), i.e. NO masking: the ``parallel_num`` new tokens attend to the whole
cache and to each other in both directions ("block-bidirectional").

``rnd`` selects exact fp32 math (``ident``) or the autocast(cuda, bf16) rounding policy (``bf16``); ``stream_f32``
says whether the residual stream is fp32 (AR steps: inputs_embeds = bf16 MLP output + fp32 pos-embed, t2i_pipeline.py:253)
or bf16 (prefill: embed_tokens output of a bf16 model). With an fp32 stream RoPE runs in fp32 (cos/sin are cast to
the hidden dtype) and q/K are rounded to bf16 at the attention call; with a bf16 stream every op rounds.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .head import bf16, ident  # noqa: F401


def rmsnorm(x, w, eps, rnd, stream_f32):
    xf = x.float()
    h = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    if stream_f32:
        return rnd(w.float()) * h  # fp32 * bf16 weight -> fp32
    return rnd(rnd(w.float()) * rnd(h))


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float):
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = positions.float()[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def _rot_half(x):
    d = x.shape[-1] // 2
    return torch.cat([-x[..., d:], x[..., :d]], dim=-1)


def apply_rope(x, cos, sin, rnd, stream_f32):
    # x [B, H, S, D]; cos/sin [S, D]
    if stream_f32:
        return x * cos + _rot_half(x) * sin  # fp32; rounded to bf16 where attention consumes it
    c, s = rnd(cos), rnd(sin)
    return rnd(rnd(x * c) + rnd(_rot_half(x) * s))


def _lin(x, w, rnd):
    return rnd(rnd(x) @ rnd(w.float()).t())


def decoder_forward(sd, cfg, x, cache, *, causal, rnd=ident, stream_f32=True, prefix="model."):
    """x: [B, S, hidden] input embeddings (fp32 values). cache: list (len = layers) of [K, V] with K,V
    [B, Hkv, L, D] or None; updated in place (concatenation). Returns last_hidden_state [B, S, hidden]."""
    H, Hkv, D = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    eps, theta = cfg["rms_norm_eps"], cfg["rope_theta"]
    B, S, _ = x.shape
    past = 0 if cache[0] is None else cache[0][0].shape[2]
    cos, sin = rope_cos_sin(torch.arange(past, past + S), D, theta)
    h = x if stream_f32 else rnd(x)
    for li in range(cfg["num_hidden_layers"]):
        p = f"{prefix}layers.{li}."
        a = rmsnorm(h, sd[p + "input_layernorm.weight"], eps, rnd, stream_f32)
        q = _lin(a, sd[p + "self_attn.q_proj.weight"], rnd).view(B, S, H, D)
        k = _lin(a, sd[p + "self_attn.k_proj.weight"], rnd).view(B, S, Hkv, D)
        v = _lin(a, sd[p + "self_attn.v_proj.weight"], rnd).view(B, S, Hkv, D).transpose(1, 2)
        q = rmsnorm(q, sd[p + "self_attn.q_norm.weight"], eps, rnd, False).transpose(1, 2)
        k = rmsnorm(k, sd[p + "self_attn.k_norm.weight"], eps, rnd, False).transpose(1, 2)
        q = rnd(apply_rope(q, cos, sin, rnd, stream_f32))
        k = rnd(apply_rope(k, cos, sin, rnd, stream_f32))
        if cache[li] is None:
            cache[li] = [k, v]
        else:
            cache[li] = [torch.cat([cache[li][0], k], dim=2), torch.cat([cache[li][1], v], dim=2)]
        K, V = cache[li]
        L = K.shape[2]
        Kr = K.repeat_interleave(H // Hkv, dim=1)
        Vr = V.repeat_interleave(H // Hkv, dim=1)
        s = (q @ Kr.transpose(-1, -2)) * (D ** -0.5)
        if causal:
            qpos = torch.arange(past, past + S)[:, None]
            kpos = torch.arange(L)[None, :]
            s = s.masked_fill(kpos > qpos, float("-inf"))
        pr = rnd(torch.softmax(s.float(), dim=-1))
        o = rnd(pr @ Vr).transpose(1, 2).reshape(B, S, H * D)
        o = _lin(o, sd[p + "self_attn.o_proj.weight"], rnd)
        h = h + o if stream_f32 else rnd(h + o)
        a = rmsnorm(h, sd[p + "post_attention_layernorm.weight"], eps, rnd, stream_f32)
        g = _lin(a, sd[p + "mlp.gate_proj.weight"], rnd)
        u = _lin(a, sd[p + "mlp.up_proj.weight"], rnd)
        m = _lin(rnd(rnd(F.silu(g)) * u), sd[p + "mlp.down_proj.weight"], rnd)
        h = h + m if stream_f32 else rnd(h + m)
    return rmsnorm(h, sd[prefix + "norm.weight"], eps, rnd, stream_f32)
