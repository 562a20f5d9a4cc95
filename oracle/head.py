"""Oracle (torch, CPU): the binary-diffusion vision head and its Euler–Maruyama sampler. TEST INFRASTRUCTURE ONLY.

Functional restatement (state-dict in, tensors out) of
  * TransEncoder.forward      modeling/vision_head/flow_head_parallel_x.py:325-342
      TimestepEmbedder        :126-143, timestep_embedding :12-27
      TransBlock.forward      :242-252, Attention.forward :192-220, FinalLayer.forward :169-173
  * euler_maruyama            modeling/vision_head/sampling_x.py:44-97 (+ :6-41 step functions)
The ImageNet variant (imagenet_gen/src/diff_head_parallel.py: head_dim 64, no output sigmoid) is the same code with
``head_dim=64, out_sigmoid=False``.

Two numeric modes, selected by ``rnd``:
  * ``rnd=ident``  exact fp32 math — pinned against the reference run in fp32 (tests/test_oracle_vs_reference.py);
  * ``rnd=bf16``   the rounding policy of torch.autocast("cuda", bfloat16), which is what the reference runs under
                   (modeling/t2i_pipeline.py:130) and what the CUDA kernels implement: every Linear rounds its inputs
                   and its output (fp32 accumulate + bias) to bf16; elementwise ops on bf16 tensors round their result;
                   LayerNorm and softmax are computed and returned in fp32; the sampler state is fp32 (autocast off,
                   sampling_x.py:25,34). Attention follows flash-attn semantics (fp32 scores/softmax, P and output in
                   bf16) for every parallel_num (the reference's seqlen<=32 eager branch rounds q*scale and the
                   scores to bf16 as well; that difference is inside the stated tolerance).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def ident(x: torch.Tensor) -> torch.Tensor:
    return x


def bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _linear(sd, name, x, rnd):
    w = sd[name + ".weight"].float()
    b = sd.get(name + ".bias")
    y = rnd(x) @ rnd(w).t()
    if b is not None:
        y = y + rnd(b.float())
    return rnd(y)


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0, time_factor: float = 1000.0):
    """cos || sin of 1000*t*exp(-ln(1e4) k / half)  (flow_head_parallel_x.py:12-27), fp32."""
    half = dim // 2
    t = time_factor * t.float()
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def time_embed(sd, t: torch.Tensor, rnd, prefix="net."):
    h = _linear(sd, prefix + "time_embed.mlp.0", timestep_embedding(t), rnd)
    h = rnd(F.silu(h))
    return _linear(sd, prefix + "time_embed.mlp.2", h, rnd)


def _layernorm(x, w=None, b=None, eps=1e-6):
    return F.layer_norm(x.float(), (x.shape[-1],), None if w is None else w.float(), None if b is None else b.float(), eps)


def _modulate(h_ln, scale, shift, rnd):
    # norm(x) * (1 + scale) + shift : fp32 * bf16(1 + scale) + bf16 shift -> fp32   (TransBlock.forward :243,246)
    return h_ln * rnd(1.0 + scale) + shift


def attention(sd, prefix, h, head_dim, rnd):
    B, S, D = h.shape
    nh = D // head_dim
    qkv = _linear(sd, prefix + "wqkv", h, rnd)
    q, k, v = qkv.chunk(3, dim=-1)
    q = q.view(B, S, nh, head_dim).transpose(1, 2)
    k = k.view(B, S, nh, head_dim).transpose(1, 2)
    v = v.view(B, S, nh, head_dim).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (head_dim ** -0.5)
    p = rnd(torch.softmax(s, dim=-1))
    o = rnd(p @ v).transpose(1, 2).reshape(B, S, D)
    return _linear(sd, prefix + "wo", o, rnd)


def trans_block(sd, prefix, x, mod, head_dim, rnd):
    scale1, shift1, gate1, scale2, shift2, gate2 = mod
    h = _modulate(_layernorm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"]), scale1, shift1, rnd)
    h = attention(sd, prefix + "attn.", h, head_dim, rnd)
    x = rnd(x + rnd(h * gate1))
    h = _modulate(_layernorm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"]), scale2, shift2, rnd)
    if (prefix + "w1.weight") in sd:  # use_swiglu=True
        h1, h2 = _linear(sd, prefix + "w1", h, rnd).chunk(2, dim=-1)
        h = _linear(sd, prefix + "w2", rnd(rnd(F.silu(h1)) * h2), rnd)
    else:
        h = _linear(sd, prefix + "mlp.2", rnd(F.silu(_linear(sd, prefix + "mlp.0", h, rnd))), rnd)
    return rnd(x + rnd(h * gate2))


def head_forward(sd, x, t, c, *, rnd=ident, head_dim=128, out_sigmoid=True, prefix="net.", c_emb=None):
    """x [R,pn,C] fp32, t [R] fp32, c [R,pn,Dz] fp32 -> x-prediction [R,pn,C] in (-1,1)."""
    n_blocks = len({k.split(".")[2] for k in sd if k.startswith(prefix + "res_blocks.")})
    n_ada = len({k.split(".")[2] for k in sd if k.startswith(prefix + "ada_ln_blocks.")})
    switch = max(1, n_blocks // n_ada)
    h = _linear(sd, prefix + "input_proj", x, rnd)
    t_emb = time_embed(sd, t, rnd, prefix).unsqueeze(1)
    if c_emb is None:
        c_emb = _linear(sd, prefix + "cond_embed", c, rnd)
    y = rnd(F.silu(rnd(t_emb + c_emb)))
    mod = _linear(sd, prefix + "ada_ln_blocks.0", y, rnd).chunk(6, dim=-1)
    for i in range(n_blocks):
        if i > 0 and i % switch == 0:
            mod = _linear(sd, f"{prefix}ada_ln_blocks.{i // switch}", y, rnd).chunk(6, dim=-1)
        h = trans_block(sd, f"{prefix}res_blocks.{i}.", h, mod, head_dim, rnd)
    scale, shift = _linear(sd, prefix + "final_layer.ada_ln_modulation", y, rnd).chunk(2, dim=-1)
    h = _layernorm(h) * rnd(1.0 + scale) + shift
    out = _linear(sd, prefix + "final_layer.linear", h, rnd)
    if out_sigmoid:
        out = rnd(rnd(2.0 * rnd(torch.sigmoid(out))) - 1.0)
    return out


def sampler_schedule(num_sampling_steps: int, last_step_size: float = 0.05, time_shift: float = 1.0):
    """fp32 scalars exactly as sampling_x.py:62-68,82 produces them: t_i is the RUNNING fp32 sum of dt (t += dt[i]),
    not t_all[i]. Returns (t_list [S], dt_list [S])."""
    t_all = torch.linspace(0, 1 - last_step_size, num_sampling_steps + 1, dtype=torch.float32)
    t_all = (1 / time_shift) / ((1 / time_shift) + (1 / t_all - 1) ** 1.0)
    dt = t_all[1:] - t_all[:-1]
    t = torch.tensor(0.0, dtype=torch.float32)
    ts = []
    for i in range(num_sampling_steps):
        ts.append(t.clone())
        t = t + dt[i]
    return ts, [dt[i] for i in range(num_sampling_steps)]


def _cfg_combine(v, cfg, mult):
    if mult == 2:
        vc, vu = v.chunk(2, dim=0)
        return vu + cfg * (vc - vu)
    return v


def sde_step(x, out, t, dt, cfg, mult, eps):
    """One euler_maruyama_step (sampling_x.py:33-41) given the network output ``out`` for cat([x]*mult) at time t."""
    comb = torch.cat([x] * mult, dim=0)
    tb = torch.full((comb.shape[0],), float(t), dtype=torch.float32)
    v = (out - comb) / (1 - tb.view(-1, 1, 1)).clamp_min(0.05)
    v = _cfg_combine(v.float(), cfg, mult)
    # get_score_from_velocity (:6-14): alpha=t, sigma=1-t, var = sigma^2 + t*sigma
    sigma = 1 - t
    var = sigma ** 2 - (t / 1) * (-1) * sigma
    score = ((t / 1) * v - x) / var
    drift = v + (1 - t) * score
    noise_scale = (2.0 * (1.0 - t) * dt) ** 0.5
    return x + drift * dt + noise_scale * eps.float()


def final_step(x, out, cfg, mult, last_step_size=0.05):
    """The deterministic last Euler step (sampling_x.py:84-95) at t = 1 - last_step_size."""
    comb = torch.cat([x] * mult, dim=0)
    tb = torch.full((comb.shape[0],), 1 - last_step_size, dtype=torch.float32)
    v = (out - comb) / (1 - tb.view(-1, 1, 1)).clamp_min(0.05)
    v = _cfg_combine(v.float(), cfg, mult)
    return x + v * last_step_size


def euler_maruyama(sd, c, cfg, num_sampling_steps, noise, *, rnd=ident, head_dim=128, out_sigmoid=True,
                   last_step_size=0.05, time_shift=1.0, ch_target=None, trace=None, forced_out=None):
    """sampling_x.euler_maruyama with the noise supplied: noise[0] is x0 (torch.randn at :60), noise[1+i] is the
    randn_like of step i (:40). c: [R,pn,Dz] (cond rows first, then uncond rows when cfg > 1). Returns cat[x]*mult.
    ``forced_out`` (list of S+1 tensors) teacher-forces the sampler with externally computed network outputs while the
    oracle's own output for the same input is still recorded in ``trace``."""
    mult = 2 if cfg > 1.0 else 1
    x = noise[0].float().clone()
    ts, dts = sampler_schedule(num_sampling_steps, last_step_size, time_shift)
    c_emb = _linear(sd, "net.cond_embed", c, rnd)  # constant across evaluations; hoisting is exact

    def net(xx, tval):
        tb = torch.full((c.shape[0],), float(tval), dtype=torch.float32)
        return head_forward(sd, xx, tb, c, rnd=rnd, head_dim=head_dim, out_sigmoid=out_sigmoid, c_emb=c_emb)

    for i in range(num_sampling_steps + 1):
        last = i == num_sampling_steps
        out = net(torch.cat([x] * mult, dim=0), (1 - last_step_size) if last else ts[i])
        if trace is not None:
            trace.append(dict(out=out.clone(), x_in=x.clone()))
        if forced_out is not None:
            out = forced_out[i].float()
        x = final_step(x, out, cfg, mult, last_step_size) if last else sde_step(x, out, ts[i], dts[i], cfg, mult,
                                                                                 noise[1 + i])
    return torch.cat([x] * mult, dim=0)
