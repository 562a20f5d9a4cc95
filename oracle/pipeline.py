"""Oracle (torch, CPU): BitDanceT2IPipeline.gen_image end to end. TEST INFRASTRUCTURE ONLY.

Restates modeling/t2i_pipeline.py:158-283 (prefill :195-236, AR loop :241-268, decode_image :274-283, pos-embed
:79-107) and MLPconnector.forward (modeling/utils.py:16-20) on top of oracle/{llm,head,ae}.py. Pinned against the
reference pipeline (fp32, CPU, shims of oracle/ref_harness.py) in tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ae as oae
from . import head as ohead
from . import llm as ollm
from .head import bf16, ident  # noqa: F401


def sincos_1d(dim, max_len):
    omega = torch.arange(dim // 2, dtype=torch.float32) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = torch.arange(max_len, dtype=torch.float32)[:, None] * omega[None, :]
    return torch.cat([out.sin(), out.cos()], dim=1)


def pos_embed_2d(hidden, h, w, ps, max_len=256):
    p1 = sincos_1d(hidden // 2, max_len)
    gv = p1[:h][:, None, :].expand(h, w, -1)
    gh = p1[:w][None, :, :].expand(h, w, -1)
    pe = torch.cat([gh, gv], dim=-1)
    return pe.reshape(h // ps, ps, w // ps, ps, hidden).permute(0, 2, 1, 3, 4).reshape(h * w, hidden)


def projector(sd, x, rnd):
    """MLPconnector: fc2(gelu_tanh(fc1(x)))."""
    hdn = rnd(rnd(x) @ rnd(sd["fc1.weight"].float()).t() + rnd(sd["fc1.bias"].float()))
    hdn = rnd(F.gelu(hdn, approximate="tanh"))
    return rnd(rnd(hdn) @ rnd(sd["fc2.weight"].float()).t() + rnd(sd["fc2.bias"].float()))


def gen_image(*, sd_llm, cfg_llm, embed, sd_head, sd_proj, sd_ae, cond_ids, uncond_ids, start_ids, h, w, pn,
              num_images, guidance, S, noise, rnd=ident, head_dim=128, decode=True, trace=None, num_steps=None,
              cond_emb=None, uncond_emb=None):
    """noise: list over AR steps of lists [x0, eps_0 .. eps_{S-1}] (the torch.randn sequence of one sample() call).
    embed: [vocab, hidden] embedding table. start_ids: pn + 2 token ids (<|vision_start|>, <|res_h|>, <|res_w|>,
    <|query_1..pn-1|>). ``cond_emb`` / ``uncond_emb`` [L, hidden]: context EMBEDDINGS instead of ``embed[cond_ids]`` — the
    interleaved inference's image item (modeling/mllm.py:745-783), whose context also holds projected image tokens.
    Returns (tokens [B, h*w, C], image | None)."""
    B, L = num_images, cfg_llm["num_hidden_layers"]
    hidden_size = embed.shape[1]
    ps = int(pn ** 0.5)
    pos = pos_embed_2d(hidden_size, h, w, ps)
    groups = [cond_ids] + ([uncond_ids] if guidance > 1.0 else [])
    ctx_emb = [cond_emb] + ([uncond_emb] if guidance > 1.0 else [])
    caches, hid = [], []
    for ids, ce in zip(groups, ctx_emb):
        if ce is None:
            emb = embed[torch.tensor(list(ids) + list(start_ids))].float()
        else:
            emb = torch.cat([ce.float(), embed[torch.tensor(list(start_ids))].float()], dim=0)
        emb = rnd(emb).unsqueeze(0).repeat(B, 1, 1)
        cache = [None] * L
        ollm.decoder_forward(sd_llm, cfg_llm, emb[:, :-pn], cache, causal=True, rnd=rnd, stream_f32=(rnd is ident))
        o = ollm.decoder_forward(sd_llm, cfg_llm, emb[:, -pn:], cache, causal=False, rnd=rnd, stream_f32=(rnd is ident))
        caches.append(cache)
        hid.append(o)
    steps = (h * w) // pn
    if num_steps is not None:
        steps = min(steps, num_steps)
    out_tokens = []
    for step in range(steps):
        blk = pos[step * pn:(step + 1) * pn][None]
        h_fused = torch.cat(hid, dim=0) + blk
        pred = ohead.euler_maruyama(sd_head, h_fused, guidance, S, noise[step], rnd=rnd, head_dim=head_dim)
        tok = torch.sign(pred)
        out_tokens.append(tok[:B])
        if trace is not None:
            trace.append(dict(h_fused=h_fused.clone(), pred=pred[:B].clone()))
        if step == steps - 1 and num_steps is None:
            break  # the reference computes one more (discarded) LLM pass
        x = projector(sd_proj, tok, rnd) + blk
        hid = []
        for gi in range(len(groups)):
            hid.append(ollm.decoder_forward(sd_llm, cfg_llm, x[gi * B:(gi + 1) * B], caches[gi], causal=False, rnd=rnd,
                                            stream_f32=True))
    tokens = torch.cat(out_tokens, dim=1)
    img = None
    if decode and tokens.shape[1] == h * w:
        C = tokens.shape[-1]
        grid = tokens.view(B, h // ps, w // ps, ps, ps, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, h, w)
        img = oae.decoder_forward(sd_ae, grid, rnd=rnd)
    return tokens, img
