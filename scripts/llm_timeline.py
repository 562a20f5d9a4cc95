#!/usr/bin/env python
"""Per-op timeline of the one-launch Qwen3 AR block (csrc/bd_llm.cu::llm_stream_all) at the 14B dimensions with LAYERS
layers (default 6) and a CTX-token cache (default 2100): %globaltimer stamps of bd_stream_kernel (bd_stream_set_debug)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BD_LLM_STREAM"] = "1"
from bitdance_b200 import _lib, ops  # noqa: E402
from bitdance_b200.llm import LlmRunner  # noqa: E402
from bitdance_b200.synthetic import QWEN3_14B  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
L = int(os.environ.get("LAYERS", "6"))
ctx = int(os.environ.get("CTX", "2100"))
cfg = {k: v for k, v in QWEN3_14B.items() if k != "vocab_size"}
cfg["num_hidden_layers"] = L
run = LlmRunner(None, cfg, device=dev, synthetic_seed=1, max_positions=8192, stream=True)
R, pn, D = 2, 64, cfg["hidden_size"]
cache = run.new_cache(R, 4224)
# fill the cache: prefill in chunks (bf16 stream, causal)
left = ctx
while left > 0:
    n = min(left, 512)
    run.forward((torch.randn(R, n, D, device=dev) * 0.5).to(torch.bfloat16), cache, 0, R, causal=True)
    left -= n
x = torch.randn(R, pn, D, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def block():
    cache.seq_lens.fill_(ctx)
    cache.host_lens = [ctx] * R
    return run.forward(x.clone(), cache, 0, R, causal=False, sk_bound=cache.max_tokens)


ms = timed(block)
wbytes = L * 2 * (5120 * 7168 + 5120 * 5120 + 3 * 5120 * 17408)
print(f"one AR block, {L} layers, context {ctx}: {ms * 1e3:.1f} us = {ms * 1e3 / L:.1f} us per layer; weights {wbytes / 1e9:.2f} GB -> "
      f"{wbytes / 1e9 / (ms / 1e3):.0f} GB/s ({wbytes / 1e9 / (ms / 1e3) / 6568:.3f} of 6568)")
G = ops.stream_num_ctas()
names_body = ["rope", "attn", "combine", "wo", "row_wo", "gate_up", "down", "row_down", "qkv_next"]
nops = 2 + 9 * L + 1
names = ["rms0", "qkv0"] + names_body * L + ["final"]


def timeline(label, verbose):
    dbg = torch.zeros(nops * G * 8, dtype=torch.int64, device=dev)
    lib.bd_stream_set_debug(C.c_void_p(dbg.data_ptr()), nops)
    block()
    torch.cuda.synchronize()
    lib.bd_stream_set_debug(None, 0)
    d = dbg.view(nops, G, 8).cpu().double()
    t_prev = None
    tot = {}
    for q in range(nops):
        v = d[q, :, 5] / 1e3
        v = v[v > 0]
        if v.numel() == 0:
            continue
        done = v.max().item()
        if t_prev is None:
            t4 = d[q, :, 4] / 1e3
            t_prev = t4[t4 > 0].min().item()
        dur = done - t_prev
        layer = (q - 2) // 9
        if verbose and 2 <= q < nops - 1 and layer == L // 2:
            print(f"  layer {layer} {names[q]:10s} {dur:7.1f} us   (arrive spread {done - v.median().item():.1f})")
        if 2 <= q < nops - 1 and 1 <= layer < L - 1:
            tot[names[q]] = tot.get(names[q], 0.0) + dur / (L - 2)
        t_prev = done
    qa = 2 + 9 * (L // 2) + 1  # the attention op of the middle layer: slots 6 / 7 = cycles waiting for K/V, cycles of work
    w, m = d[qa, :, 6], d[qa, :, 7]
    print(f"{label}: per op (us) {({k: round(v, 1) for k, v in tot.items()})} sum {sum(tot.values()):.1f}; attention of layer "
          f"{L // 2}: median wait {w.median().item() / 1e3:.1f} kcycles, work {m.median().item() / 1e3:.1f} kcycles per CTA")


timeline("default", True)
for label, mode, poll, pf in (("no K/V loads (mode 64)", 64, 32, 0), ("no tensor work (mode 128)", 128, 32, 0),
                              ("neither (mode 192)", 192, 32, 0), ("poll 500 ns", 0, 500, 0), ("poll 2000 ns", 0, 2000, 0),
                              ("L2 prefetch 8 steps", 0, 32, 8)):
    lib.bd_stream_set_tuning(5, 2, mode)
    lib.bd_stream_set_poll_ns(poll)
    lib.bd_stream_set_prefetch(pf)
    timeline(label, False)
lib.bd_stream_set_tuning(5, 2, 0)
lib.bd_stream_set_poll_ns(32)
lib.bd_stream_set_prefetch(0)
