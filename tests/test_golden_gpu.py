"""The CUDA path against the committed golden fixtures (outputs of the reference itself, fp32 CPU). The GPU path runs
the autocast-bf16 policy the reference uses on GPUs, so tolerances are bf16-level; the quantiser is exact given its input."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return {k: torch.from_numpy(v) if v.ndim else v.item() for k, v in np.load(os.path.join(G, name)).items()}


@pytest.fixture(scope="module")
def tiny():
    from bitdance_b200.synthetic import engine_from_state_dicts, tiny_state_dicts
    s = tiny_state_dicts()
    return engine_from_state_dicts(s, "tiny", "cuda"), s


def test_tokenizer_golden_gpu(tiny):
    eng, _ = tiny
    g = load("ae_tiny.npz")
    q, packed, idx, lat = eng.ae.encode(g["image"].cuda())
    lat_err = (lat.float().cpu() - g["latent"]).abs().max().item()
    assert lat_err < 5e-2 * g["latent"].abs().max().item()
    safe = g["latent"].abs() > lat_err + 1e-3
    assert torch.equal(q.float().cpu()[safe], g["quant"].float()[safe])
    assert (q.float().cpu() == g["quant"].float()).float().mean().item() > 0.97
    dec = eng.ae.decode(g["quant"].float().cuda()).float().cpu()
    assert (dec - g["decoded"]).abs().max().item() < 5e-2 * g["decoded"].abs().max().item() + 2e-2


def test_head_golden_gpu(tiny):
    eng, _ = tiny
    g = load("head_tiny.npz")
    S, cfg = int(g["S"]), float(g["cfg"])
    B = g["c"].shape[0] // 2
    x, trace = eng.head.sample(g["c"].cuda(), cfg, S, noise=g["noise"].cuda().contiguous(), trace=True)
    ref = g["sample"][:B]
    d = (x.cpu() - ref).abs()
    agree = (torch.sign(x.cpu()) == torch.sign(ref)).float().mean().item()
    print(f"head vs reference golden (fp32): max {d.max().item():.3f} mean {d.mean().item():.4f} sign agreement {agree:.4f}")
    assert d.mean().item() < 0.1 and agree > 0.93


def test_llm_golden_gpu(tiny):
    eng, _ = tiny
    g = load("llm_tiny.npz")
    cache = eng.llm.new_cache(2, 256)
    h0 = eng.llm.forward(g["x0"].to(torch.bfloat16).cuda(), cache, 0, 2, causal=True).float().cpu()
    h1 = eng.llm.forward(g["x1"].to(torch.bfloat16).cuda(), cache, 0, 2, causal=False).float().cpu()
    h2 = eng.llm.forward(g["x2"].cuda().clone(), cache, 0, 2, causal=False).cpu()
    for a, b in ((h0, g["h0"]), (h1, g["h1"]), (h2, g["h2"])):
        assert (a - b).abs().max().item() < 6e-2 * b.abs().max().item()
