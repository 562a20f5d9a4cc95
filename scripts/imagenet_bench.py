#!/usr/bin/env python
"""ImageNet class-conditional generation throughput (the only throughput the authors published:
~24 / ~90 img/s charts for BitDance-B at 256 px): random-init weights of the published architecture, the API mirror's
``create_model(...).sample(...)`` exactly as imagenet_gen/sample_ddp_parallel.py:159-164 calls it.

  python scripts/imagenet_bench.py [--model BitDance-B] [--parallel-num 16] [--bs 64 256] [--steps 50] [--cfg 3.9]
Prints one JSON line per batch size: images/s, ms per AR position, achieved TFLOP/s of the decoder + head Linears against
the measured sustained bf16 peak (compute-bound regime: M = 2 * bs * parallel_num rows)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200.imagenet import MODELS, ImageNetEngine, ffn_hidden, imagenet_spec  # noqa: E402
from bitdance_b200.synthetic import _gpu_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="BitDance-B", choices=list(MODELS))
    ap.add_argument("--parallel-num", type=int, default=16)
    ap.add_argument("--bs", type=int, nargs="+", default=[64, 256])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cfg", type=float, default=3.9)
    ap.add_argument("--latent-dim", type=int, default=32)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--no-warmup", action="store_true", help="profiling (ncu launch list): no warm-up pass")
    ap.add_argument("--head-engines", type=int, default=None, help="persistent head engines side by side (0: multi-kernel path)")
    ap.add_argument("--modes", type=int, nargs="+", default=[0],
                    help="bd_stream_set_tuning mode bits per run (256: block-wide attention units, 512: block-wide row ops)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    cfg = dict(MODELS[args.model], latent_dim=args.latent_dim, down_size=16, patch_size=1, resolution=256, cls_token_num=64,
               num_classes=1000, parallel_num=args.parallel_num, parallel_mode="patch", time_shift=1.0)
    sd = _gpu_state_dict(imagenet_spec(cfg), 7, dev)
    eng = ImageNetEngine(sd, cfg, ae=None, device=dev, head_engines=args.head_engines)
    del sd
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    dim, L, hid = cfg["dim"], cfg["n_layer"], ffn_hidden(cfg["dim"])
    p_dec = L * (4 * dim * dim + 3 * dim * hid)                                    # decoder Linear parameters
    Dh, nb, na = cfg["diff_dim"], cfg["diff_layers"], cfg["diff_adanln_layers"]
    p_head = nb * (4 * Dh * Dh + 3 * Dh * int(Dh * 1.5) * 1) + na * 6 * Dh * Dh + 2 * Dh * Dh + dim * Dh + Dh * Dh
    hw, pn = eng.h * eng.w, eng.pn
    from bitdance_b200 import _lib
    lib = _lib.load()
    for bs, mode in [(b, m) for b in args.bs for m in args.modes]:
        lib.bd_stream_set_tuning(5, 2, mode)
        ids = torch.randint(0, 1000, (bs,), device=dev)
        if not args.no_warmup:
            eng.sample_tokens(ids, args.steps, args.cfg)   # warm-up (allocations, first launches)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            eng.sample_tokens(ids, args.steps, args.cfg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        rows = 2 * bs * pn                                                          # token rows per AR position (CFG)
        flops = (hw // pn) * (2.0 * p_dec * rows + (args.steps + 1) * 2.0 * p_head * rows)
        print(json.dumps({"model": args.model, "parallel_num": pn, "bs": bs, "head_engines": len(eng._side), "mode": mode, "sampling_steps": args.steps, "cfg": args.cfg,
                          "images_per_s": bs / (ms / 1e3), "ms_per_batch": ms, "ms_per_ar_position": ms / (hw // pn),
                          "rows_per_position": rows, "tflops": flops / 1e12 / (ms / 1e3),
                          "frac_of_sustained_bf16": flops / 1e12 / (ms / 1e3) / peak,
                          "note": "token generation only (tokenizer decode measured by scripts/ae_bench.py)"}), flush=True)


if __name__ == "__main__":
    main()
