"""Text decode step of the interleaved inference at the Qwen3-14B dimensions, random-init:
one token = one causal [1, 1, D] pass of the 40-layer decoder over the paged KV (bd_llm_forward, chained kernels) +
lm_head (bd_gemm_bf16, 151 936 x 5 120) + the sampler (top-k 1200 / top-p 0.95, torch ops on the logits).
HBM-bound: algorithmic bytes per token = 2 * P_llm (26.42 GB) + 2 * vocab * D (1.556 GB) + the KV read.
usage: python scripts/text_decode_bench.py [--tokens 32] [--context 64]   -> one JSON line"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=32)
    ap.add_argument("--context", type=int, default=64)
    ap.add_argument("--model", default="BitDance-14B-64x")
    args = ap.parse_args()
    from bitdance_b200.llm import LmHead
    from bitdance_b200.modeling.utils import sample_codebook
    from bitdance_b200.synthetic import MODELS, build_synthetic_engine
    dev = torch.device("cuda", 0)
    eng, embed = build_synthetic_engine(args.model, dev, seed=0, with_ae=False, llm_stream=False)
    cfg = MODELS[args.model]["llm"]
    V, D = cfg["vocab_size"], cfg["hidden_size"]
    g = torch.Generator(device=dev).manual_seed(7)
    head = LmHead((torch.randn((V, D), generator=g, device=dev) * 0.02).to(torch.bfloat16), device=dev)
    emb = torch.nn.Embedding.from_pretrained(embed, freeze=True)
    llm = eng.llm
    res = {}
    for do_sample in (False, True):
        cache = llm.new_cache(1, args.context + 2 * args.tokens + 64)
        x = embed[torch.randint(0, V, (args.context,), device=dev)].view(1, -1, D).contiguous()
        torch.manual_seed(1)
        times = []
        with torch.no_grad():
            for step in range(args.tokens + 4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                hid = llm.forward(x, cache, 0, 1, causal=True)[:, -1]
                tok, nxt = sample_codebook(head(hid).float(), "text", emb, do_sample=do_sample, top_k=1200, top_p=0.95)
                x = nxt.view(1, 1, D).contiguous()
                e1.record()
                torch.cuda.synchronize()
                if step >= 4:           # step 0 is the context pass; 1-3 warm-up
                    times.append(e0.elapsed_time(e1))
        times.sort()
        ms = times[len(times) // 2]
        kv = (args.context + args.tokens / 2) * 163840
        gb = (2 * 13.2125e9 + 2.0 * V * D + kv) / 1e9
        peak = 6568.0
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peak = json.load(open(pk)).get("hbm_gbs", peak)
        res["sample" if do_sample else "greedy"] = {"ms_per_token": ms, "tokens_per_s": 1e3 / ms, "algorithmic_gb": gb,
                                                    "achieved_gbs": gb / (ms / 1e3), "frac_of_hbm_peak": gb / (ms / 1e3) / peak}
    print(json.dumps({"workload": f"{args.model} text decode, bs=1, context {args.context}, {args.tokens} timed tokens (median)",
                      **res}))


if __name__ == "__main__":
    main()
