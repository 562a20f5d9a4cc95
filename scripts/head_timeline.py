#!/usr/bin/env python
"""Per-op timeline of the persistent head sampler at the 14B-64x dimensions, for each program policy given in MODES
(default "1:22:8,0": fillers with 22 / 8 k-block slots, then the in-line round-1 program), plus event timings of a whole
sample() call.  MODES="fill:row_kb:gemm_kb,..."; TILED=0 skips the multi-kernel path."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitdance_b200 import _lib, ops  # noqa: E402
from bitdance_b200.head import HeadRunner, head_spec  # noqa: E402
from bitdance_b200.synthetic import MODELS, _gpu_state_dict  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
hc = MODELS["BitDance-14B-64x"]["head"]
sd = _gpu_state_dict(head_spec(hc["ch_target"], hc["ch_cond"], hc["ch_latent"], hc["depth_latent"], hc["depth_adanln"],
                               hc["use_swiglu"]), 2, dev)
head = HeadRunner(sd, device=dev, tiled=bool(int(os.environ.get("TILED", "0"))), **hc)
del sd
torch.cuda.empty_cache()
R, pn, S = 2, 64, 50
NB, D = hc["depth_latent"], hc["ch_latent"]
z = torch.randn(R, pn, 5120, device=dev)
G = ops.stream_num_ctas()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def program_names(fill, row_kb, gemm_kb):
    """op names in program order (pre, body) and which ops are fillers — mirrors csrc/bd_head.cu::head_sample_stream"""
    pre = ["cast", "tfreq", "init", "time0", "cond", "time2"] + (["silu_all", "ada0"] if fill else ["silu_add0"])
    slots = {}
    if fill:
        want = [row_kb]
        for b in range(NB):
            want += [row_kb, gemm_kb] + ([row_kb] if b + 1 < NB else [])
        n_mod = hc["depth_adanln"] * 6 * D + 2 * D
        P = -(-(-(-(n_mod // 16) // G)) // 8)
        out = (C.c_int * 1024)()
        n = lib.bd_head_plan_pieces(P, D // 64, (C.c_int * len(want))(*want), len(want), out, 256)
        for i in range(n):
            slots.setdefault(out[4 * i], []).append(f"pc{out[4 * i + 1]}.{out[4 * i + 2]}+{out[4 * i + 3]}")
    body = ["input_proj"] + ([] if fill else ["ada"]) + ["ln0"]
    body += [("F", s) for s in slots.get(0, [])]
    for b in range(NB):
        body += [f"b{b}.wqkv", f"b{b}.attn", f"b{b}.wo", f"b{b}.row_wo"] + [("F", s) for s in slots.get(1 + 3 * b, [])]
        body += [f"b{b}.w1"] + [("F", s) for s in slots.get(2 + 3 * b, [])] + [f"b{b}.w2", f"b{b}.row_w2"]
        if b + 1 < NB:
            body += [("F", s) for s in slots.get(3 + 3 * b, [])]
    body += ["sde"]
    return pre, body


modes = os.environ.get("MODES", "1:22:8,0")
for mode in modes.split(","):
    f = [int(x) for x in mode.split(":")] + [0, 0]
    fill, row_kb, gemm_kb = f[0], f[1] or 22, f[2] or 8
    ops.head_set_fillers(fill, row_kb, gemm_kb)
    ms = timed(lambda: head.sample(z, 7.5, S, path="stream"))
    print(f"=== fillers={fill} row_kb={row_kb} gemm_kb={gemm_kb}: sample() {ms:.2f} ms = {ms / (S + 1) * 1e3:.1f} us per "
          f"evaluation; {(S + 1) * 3.465e9 / (ms / 1e3) / 1e9:.0f} GB/s of the head's weight bytes "
          f"({(S + 1) * 3.465e9 / (ms / 1e3) / 1e9 / 6568:.3f} of 6568)")
    if os.environ.get("TIMELINE", "1") == "0":
        continue
    pre, body = program_names(fill, row_kb, gemm_kb)
    nops = len(pre) + 3 * len(body)
    dbg = torch.zeros(nops * G * 8, dtype=torch.int64, device=dev)
    lib.bd_stream_set_debug(C.c_void_p(dbg.data_ptr()), nops)
    head.sample(z, 7.5, S, path="stream")
    torch.cuda.synchronize()
    lib.bd_stream_set_debug(None, 0)
    d = dbg.view(nops, G, 8).cpu().double() / 1e3
    names = pre + body * 3
    t_prev = t0 = None
    tot = {}
    print("op                 done(us)  dur(us) | arrive spread (med->max)   [fillers: MMA window of the slowest CTA]")
    for q in range(nops):
        nm = names[q]
        if isinstance(nm, tuple):  # filler: no arrival stamps; show when its MMAs ran
            m1, m2 = d[q, :, 1], d[q, :, 2]
            ok = (m1 > 0) & (m2 > 0)
            if q >= len(pre) + 2 * len(body) and ok.any() and t0 is not None:
                print(f"    ~{nm[1]:14s} MMA {m1[ok].min().item() - t0:8.1f} .. {m2[ok].max().item() - t0:8.1f}")
            continue
        v = d[q, :, 5]
        v = v[v > 0]
        done = v.max().item()
        if t_prev is None:
            t_prev = d[q, :, 4][d[q, :, 4] > 0].min().item()
            t0 = t_prev
        dur = done - t_prev
        if q >= len(pre) + 2 * len(body):
            key = nm.split(".")[-1]
            tot[key] = tot.get(key, 0.0) + dur
            print(f"{nm:16s} {done - t0:9.1f} {dur:8.1f} | {done - v.median().item():6.1f}")
        t_prev = done
    print("third evaluation, per op kind (us):", {k: round(v, 1) for k, v in tot.items()}, "total",
          round(sum(tot.values()), 1))
    # phases of the GEMM ops (median over CTAs with work): A loads start -> first MMA -> last MMA issued -> accumulator ready
    # -> first chunk loaded -> first chunk stored -> epilogue done -> arrival published; and the op's last arrival
    print("GEMM op phases, third evaluation (us, medians over CTAs): dep->A  A->MMA1  MMA  ->acc  ->ld1  chunk1  ->epi  ->arrive | last arrival - median arrival")
    for q in range(len(pre) + 2 * len(body), nops):
        nm = names[q]
        if isinstance(nm, tuple) or nm.split(".")[-1] not in ("wqkv", "wo", "w1", "w2", "input_proj", "ada"):
            continue
        st = d[q]
        ok = (st[:, 0] > 0) & (st[:, 5] > 0) & (st[:, 1] > 0)
        if not ok.any():
            continue
        st = st[ok]
        prev_done = d[q - 1, :, 5]
        k = q - 1
        while isinstance(names[k], tuple):
            k -= 1
        prev_done = d[k, :, 5][d[k, :, 5] > 0].max().item()
        med = lambda a: a.median().item()
        seg = [med(st[:, 0]) - prev_done, med(st[:, 1] - st[:, 0]), med(st[:, 2] - st[:, 1]), med(st[:, 3] - st[:, 2]),
               med(st[:, 6] - st[:, 3]) if (st[:, 6] > 0).all() else float("nan"),
               med(st[:, 7] - st[:, 6]) if (st[:, 7] > 0).all() else float("nan"), med(st[:, 4] - st[:, 3]), med(st[:, 5] - st[:, 4])]
        print(f"  {nm:12s} " + " ".join(f"{x:7.2f}" for x in seg) + f" | {st[:, 5].max().item() - med(st[:, 5]):5.2f}")
ops.head_set_fillers(1)
