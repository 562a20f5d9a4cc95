"""Oracle (torch, CPU): token-by-token text decoding around the Qwen3 decoder — the text half of the interleaved
text+image inference. TEST INFRASTRUCTURE ONLY.

Restates the evident algorithm of ``MLLModel.forward_inference_block_causal``'s text branch (modeling/mllm.py:757-870):
causal pass over the accumulated context on top of the persistent KV cache, ``hidden_state = last_hidden_state[:, -1]``
(:829), ``pred_logits = lm_head(hidden_state)`` (:845-846), ``sample_codebook(..., top_k=1200, top_p=0.95)`` (:852-858,
modeling/utils.py:95-124), the sampled token's embedding is the next input, stop on ``<|im_end|>`` (:863-866).

Pinning. The reference's own text branch cannot run: without a cache it subscripts ``past_key_values = None`` (:798), after
a generated image it feeds the sampled token's 2-D ``(1, hidden)`` embedding back as ``inputs_embeds`` (:857) and the decoder
fails on the second token (tests/test_interleaved_vs_reference.py asserts both on the unmodified reference). So there is no
reference output to pin the LOOP against: **the loop is parity-unpinned**. What is pinned: the decoder (``oracle/llm.py``, against transformers' Qwen3Model,
tests/test_oracle_vs_reference.py), and the sampler — ``tests/test_interleaved_cpu.py`` checks the mirrored
``top_k_top_p_filtering`` / ``sample_codebook`` / ``remove_first_user_block`` bit-exactly against the unmodified reference
functions; this file calls the mirrored sampler through the ``sampler`` argument or plain argmax.
"""
from __future__ import annotations

import torch

from . import llm as ol


def lm_head(hidden: torch.Tensor, weight: torch.Tensor, rnd=ol.bf16) -> torch.Tensor:
    """bias-free Linear hidden [.., D] x weight [V, D] -> logits [.., V] (fp32 accumulate, one rounding: a bf16 Linear)"""
    return rnd(rnd(hidden.float()) @ rnd(weight.float()).t())


def decode_text(sd, cfg, lm_head_w, embed_w, context, cache, *, end_id: int, max_length: int, rnd=ol.bf16, forced=None,
                sampler=None):
    """context [L, D] embeddings (fp32 values). cache: the oracle KV list of ONE sequence (``[None] * layers`` when empty),
    updated in place. ``forced``: teacher-forced token ids (the logits are still returned for every step).
    Returns (token ids list, logits [steps, V], ended?)."""
    x = context[None]
    ids, logits = [], []
    for step in range(max_length):
        h = ol.decoder_forward(sd, cfg, x, cache, causal=True, rnd=rnd, stream_f32=False)[:, -1]   # [1, D]
        lg = lm_head(h, lm_head_w, rnd)
        logits.append(lg[0])
        if forced is not None:
            t = int(forced[step])
        elif sampler is not None:
            t = int(sampler(lg))
        else:
            t = int(lg[0].argmax())
        ids.append(t)
        if t == end_id:
            return ids, torch.stack(logits), True
        x = rnd(embed_w[t].float())[None, None]
    return ids, torch.stack(logits), False
