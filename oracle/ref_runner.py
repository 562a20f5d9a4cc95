"""Run the UNMODIFIED reference pipeline (``BitDanceT2IPipeline.gen_image``, modeling/t2i_pipeline.py:158-272) at a named
model configuration with random-init weights, on the host cores or on the GPU. BASELINE INFRASTRUCTURE ONLY
(``bench.py --impl reference``, bench.py's ``gpu_eager_reference`` leg, tests): the product never imports this.

What is the reference's and what is the harness's:
  * reference, untouched: gen_image's body (prompt embedding, causal prefill + block-bidirectional first block for the
    cond and uncond prompts, the AR loop: vision_head.sample -> sign -> embed_vision_mlp -> + pos-embed -> two
    Qwen3Model passes), DiffHead / TransEncoder / euler_maruyama, MLPconnector, VQModel, the transformers Qwen3 model;
  * harness: the object is assembled with ``object.__new__`` (no checkpoint files exist offline; the constructor only
    loads files), a stub tokenizer returns fixed ids, weights are N(0, 0.02) (norm scales 1), the three shims of
    SURVEY.md section 8c (``oracle/ref_harness.py``), and — to BOUND the run — ``max_length`` is set to ``n_ar * parallel_num``
    so that the unmodified loop runs ``n_ar`` AR steps, with ``decode_image`` replaced by a no-op for that call (the
    tokenizer decode is timed separately). Per-step wall times come from the loop's own progress hook (``tqdm``'s
    ``update``), nothing inside the loop is touched.
  * precision: the reference runs under ``torch.autocast(device_type, bfloat16)`` with a bf16 LLM and fp32
    head / projector / tokenizer weights — exactly what ``generate()`` sets up on CUDA (t2i_pipeline.py:130); on CPU the
    same policy through CPU autocast (the hard-coded ``autocast("cuda")`` is a no-op there).
"""
from __future__ import annotations

import time
import types

import torch

from . import ref_harness as rh

QWEN3_14B = dict(hidden_size=5120, intermediate_size=17408, num_hidden_layers=40, num_attention_heads=40,
                 num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=151936)
AE_D16C32 = dict(double_z=False, z_channels=32, in_channels=3, out_ch=3, ch=256, ch_mult=[1, 1, 2, 2, 4], num_res_blocks=4)
CONFIGS = {
    "BitDance-14B-64x": dict(llm=QWEN3_14B, ae=AE_D16C32, parallel_num=64,
                             head=dict(ch_target=32, ch_cond=5120, ch_latent=5120, depth_latent=6, depth_adanln=2,
                                       use_swiglu=True)),
    "BitDance-14B-16x": dict(llm=QWEN3_14B, ae=AE_D16C32, parallel_num=16,
                             head=dict(ch_target=32, ch_cond=5120, ch_latent=5120, depth_latent=6, depth_adanln=2,
                                       use_swiglu=True)),
    "tiny": dict(llm=dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                          num_key_value_heads=2, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=512),
                 ae=dict(double_z=False, z_channels=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2], num_res_blocks=1),
                 parallel_num=16,
                 head=dict(ch_target=32, ch_cond=256, ch_latent=256, depth_latent=2, depth_adanln=2, use_swiglu=True)),
}


class StubTokenizer:
    """Fixed synthetic prompt: 64 cond ids, 3 uncond ids, fixed special ids."""

    def __init__(self, vocab: int, n_cond: int = 64, n_uncond: int = 3, seed: int = 1):
        g = torch.Generator().manual_seed(seed)
        hi = max(8, min(vocab - 300, 151000))
        self.cond = torch.randint(0, hi, (n_cond,), generator=g).tolist()
        self.uncond = torch.randint(0, hi, (n_uncond,), generator=g).tolist()
        self.base = vocab - 280

    def encode(self, s):
        return self.cond if s == "cond" else self.uncond

    def convert_tokens_to_ids(self, tk):
        if tk == "<|vision_start|>":
            return self.base
        if tk.startswith("<|res_"):
            return self.base + 1 + int(tk[6:-2]) % 100
        return self.base + 110 + int(tk[8:-2])


_TILE = 1 << 22
_TILES = {}


def _fill_normal(p: torch.Tensor, std: float, g=None):
    """N(0, std) values. torch's CPU normal_ is a serial generator (~10 ns / element: minutes for 14.8 B parameters, all of
    it box time spent before the measurement starts), so a large HOST tensor is filled with one 4 M-element random tile
    repeated along the flattened tensor — rows are shifted copies, not equal (4 194 304 is not a multiple of any row
    length here); the timing baseline does not depend on the values. Small tensors and device tensors draw every element."""
    if p.device.type != "cpu" or p.numel() <= _TILE:
        if p.device.type == "cpu":
            p.copy_((torch.randn(p.shape, generator=g, dtype=torch.float32) * std).to(p.dtype))
        else:
            p.normal_(0.0, std)
        return
    key = (p.dtype, float(std))
    if key not in _TILES:       # one tile per (dtype, std) for the whole model: drawing 4 M values per parameter adds up too
        _TILES[key] = (torch.randn(_TILE, generator=torch.Generator().manual_seed(12345), dtype=torch.float32) * std).to(p.dtype)
    tile = _TILES[key]
    flat = p.view(-1)
    n = flat.numel()
    full = n // _TILE
    if full:
        flat[:full * _TILE].view(full, _TILE).copy_(tile.unsqueeze(0).expand(full, _TILE))
    if n - full * _TILE:
        flat[full * _TILE:].copy_(tile[:n - full * _TILE])


def _randomize(module, seed: int, std: float = 0.02):
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() == 1 and ("norm" in name or name.endswith("gn.weight")) and name.endswith("weight"):
                p.fill_(1.0)
            elif p.dim() == 1:
                p.zero_()
            else:
                _fill_normal(p, std, g)


def build_pipeline(model: str = "BitDance-14B-64x", device: str = "cpu", seed: int = 0, with_ae: bool = True):
    """-> (pipe, info). The reference's own classes with random-init weights (bf16 LLM; fp32 head / projector / AE)."""
    from transformers import Qwen3Config, Qwen3ForCausalLM
    from transformers.initialization import no_init_weights
    ref = rh.import_reference()
    m = CONFIGS[model]
    t0 = time.perf_counter()
    lc = dict(m["llm"])
    cfg = Qwen3Config(max_position_embeddings=8192, tie_word_embeddings=False, **lc)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with no_init_weights(), torch.device(device):
            hf = Qwen3ForCausalLM(cfg).eval()   # parameters allocated uninitialised in bf16 on `device`; buffers computed
    finally:
        torch.set_default_dtype(old)
    with torch.no_grad():
        for name, p in hf.named_parameters():
            if p.dim() == 1:
                p.fill_(1.0)
            else:
                _fill_normal(p, 0.02)
    # on the host the constructors' own (serial) initialisers cost ~1 min for the 1.76 B-parameter head: build on the meta
    # device and materialise empty — _randomize below writes EVERY parameter, and these modules have no buffers
    with torch.device("meta" if device == "cpu" else device):
        head = ref.fh.DiffHead(parallel_num=m["parallel_num"], **m["head"]).eval()
        proj = ref.mu.MLPconnector(m["ae"]["z_channels"], lc["hidden_size"], "gelu_pytorch_tanh").eval()
        ae = ref.ae.VQModel(m["ae"]).eval() if with_ae else None
    if device == "cpu":
        assert not list(head.buffers()) and not list(proj.buffers()) and (ae is None or not list(ae.buffers()))
        head, proj = head.to_empty(device=device), proj.to_empty(device=device)
        ae = ae.to_empty(device=device) if ae is not None else None
    _randomize(head, seed + 2)   # incl. the tensors the reference zero-initialises
    _randomize(proj, seed + 4)
    if ae is not None:
        _randomize(ae, seed + 3)
    P = ref.t2i.BitDanceT2IPipeline
    pipe = object.__new__(P)
    pipe.device, pipe.tokenizer, pipe.llm_model = device, StubTokenizer(lc["vocab_size"]), hf
    pipe.hidden_size, pipe.ae, pipe.vision_head, pipe.embed_vision_mlp = lc["hidden_size"], ae, head, proj
    pipe.vae_patch_size = 2 ** (len(m["ae"]["ch_mult"]) - 1)
    pipe.parallel_num = m["parallel_num"]
    pipe.ps = int(m["parallel_num"] ** 0.5)
    pipe.build_pos_embed()
    if hasattr(pipe, "pos_embed_1d") and isinstance(pipe.pos_embed_1d, torch.Tensor):
        pipe.pos_embed_1d = pipe.pos_embed_1d.to(device)
    return pipe, dict(model=model, ref=ref, build_s=time.perf_counter() - t0,
                      params=sum(p.numel() for p in hf.parameters()) + sum(p.numel() for p in head.parameters()))


class _StepClock:
    """Stands in for ``tqdm`` inside the reference module: records a timestamp at every ``update`` (= top of every AR
    step); the GPU arm synchronises first so that the stamps are device-complete times."""

    def __init__(self, sync):
        self.sync = sync
        self.stamps = []

    def __call__(self, *a, **k):
        return self

    def update(self, n=1):
        if self.sync:
            torch.cuda.synchronize()
        self.stamps.append(time.perf_counter())


def run_bounded(pipe, info, *, n_ar: int, image_px: int = 1024, guidance: float = 7.5, S: int = 50, num_images: int = 1,
                seed: int = 1234):
    """One call of the UNMODIFIED gen_image bounded to ``n_ar`` AR steps. Returns dict(prefill_s, ar_s=[...], total_s)."""
    ref = info["ref"]
    dev = pipe.device
    is_cuda = str(dev).startswith("cuda")
    clock = _StepClock(is_cuda)
    old_tqdm, old_decode = ref.t2i.tqdm, pipe.__dict__.get("decode_image")
    ref.t2i.tqdm = clock
    pipe.decode_image = types.MethodType(lambda self, *a, **k: None, pipe)
    torch.manual_seed(seed)
    try:
        with torch.no_grad(), torch.autocast("cuda" if is_cuda else "cpu", dtype=torch.bfloat16):
            if is_cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe.gen_image("cond", "uncond", guidance_scale=guidance, num_sampling_steps=S,
                           max_length=n_ar * pipe.parallel_num, num_images=num_images, image_size=[image_px, image_px],
                           show_progress=True)
            if is_cuda:
                torch.cuda.synchronize()
            t1 = time.perf_counter()
    finally:
        ref.t2i.tqdm = old_tqdm
        if old_decode is None:
            del pipe.__dict__["decode_image"]
        else:
            pipe.decode_image = old_decode
    st = clock.stamps + [t1]
    return dict(prefill_s=st[0] - t0, ar_s=[st[i + 1] - st[i] for i in range(len(st) - 1)], total_s=t1 - t0)


def time_decode(pipe, *, image_px: int = 1024, num_images: int = 1, reps: int = 1):
    """The tokenizer decode of one random token grid through the reference's own decode_image."""
    dev = pipe.device
    is_cuda = str(dev).startswith("cuda")
    h = image_px // pipe.vae_patch_size
    tok = torch.sign(torch.randn(num_images, h * h, pipe.ae.encoder.conv_out.out_channels if hasattr(pipe.ae.encoder, "conv_out") else 32,
                                 device=dev))
    best = None
    with torch.no_grad(), torch.autocast("cuda" if is_cuda else "cpu", dtype=torch.bfloat16):
        for _ in range(reps):
            if is_cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe.decode_image(tok, [h, h], ps=pipe.ps)
            if is_cuda:
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return best



# ---------------------------------------------------------------------------------------------------------------------
# Time-bounded worker (bench.py's host-core arms). One AR step of the 14B reference is ~26 TFLOP of bf16 GEMM on the host:
# seconds on an AMX box, many minutes on others (round 2 measured both), so the CPU arm runs in a SUBPROCESS that logs every
# event as it happens — built, every Qwen3Model pass, every DiffHead network evaluation, every AR step — and the parent
# kills it at its deadline and derives the number from whatever finished (bench.py::cpu_reference_sample).
# ---------------------------------------------------------------------------------------------------------------------
from bitdance_b200.hostinfo import usable_cpus  # noqa: E402,F401  (host facts live with the product; re-exported here)


def _worker(argv=None):
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="BitDance-14B-64x")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--n-ar", type=int, default=2)
    ap.add_argument("--S", type=int, default=50)
    ap.add_argument("--guidance", type=float, default=7.5)
    ap.add_argument("--px", type=int, default=1024)
    ap.add_argument("--bs", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--decode", type=int, default=0)
    ap.add_argument("--progress", required=True)
    ap.add_argument("--go-file", default="", help="build with a few threads, then wait for this file before computing "
                                                  "(bench.py overlaps import + build with its GPU arm, not the compute)")
    a = ap.parse_args(argv)
    out = open(a.progress, "a", buffering=1)

    def emit(**kw):
        kw["t"] = time.perf_counter()
        out.write(json.dumps(kw) + "\n")
        out.flush()

    threads = a.threads or usable_cpus()
    import os
    if a.device == "cpu":
        torch.set_num_threads(min(threads, 4) if a.go_file else threads)
    emit(ev="start", threads=threads, cpu_count=os.cpu_count())
    pipe, info = build_pipeline(a.model, a.device, with_ae=bool(a.decode))
    emit(ev="built", s=info["build_s"], pn=pipe.parallel_num)
    if a.go_file:
        parent, t_wait = os.getppid(), time.perf_counter()
        while not os.path.exists(a.go_file):
            if os.getppid() != parent or time.perf_counter() - t_wait > 1800:   # the bench died / forgot us: do not linger
                emit(ev="abandoned")
                return
            time.sleep(0.2)
        if a.device == "cpu":
            torch.set_num_threads(threads)
        emit(ev="go")

    def timed(name, fn, rows_of):
        def wrapper(*args, **kw):
            t0 = time.perf_counter()
            r = fn(*args, **kw)
            emit(ev=name, s=time.perf_counter() - t0, rows=rows_of(args, kw))
            return r
        return wrapper

    # harness-side stopwatches around two of the reference's own modules (nothing inside them is touched)
    net = pipe.vision_head.net
    net.forward = timed("eval", net.forward, lambda args, kw: int(args[0].shape[0] * args[0].shape[1]) if args else 0)
    lm = pipe.llm_model.model
    lm.forward = timed("llm", lm.forward, lambda args, kw: int(kw["inputs_embeds"].shape[0] * kw["inputs_embeds"].shape[1])
                       if "inputs_embeds" in kw else 0)

    base_update = _StepClock.update

    def update(self, n=1):
        base_update(self, n)
        emit(ev="step")

    _StepClock.update = update          # this process exists only for this run
    emit(ev="gen_start")
    run = run_bounded(pipe, info, n_ar=a.n_ar, image_px=a.px, guidance=a.guidance, S=a.S, num_images=a.bs)
    emit(ev="gen_done", prefill_s=run["prefill_s"], ar_s=run["ar_s"], total_s=run["total_s"])
    if a.decode:
        emit(ev="decode", s=time_decode(pipe, image_px=a.px, num_images=a.bs))
    emit(ev="done")


if __name__ == "__main__":
    _worker()
