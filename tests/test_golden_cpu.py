"""The CPU oracle against the committed golden fixtures (outputs of the reference itself, tests/golden/make_goldens.py).
Runs anywhere (no /root/reference needed): this is what keeps the oracle pinned on the GPU box."""
import os

import numpy as np
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return {k: torch.from_numpy(v) if v.ndim else v.item() for k, v in np.load(os.path.join(G, name)).items()}


def sds():
    from bitdance_b200.synthetic import tiny_state_dicts
    return tiny_state_dicts()


def test_tokenizer_golden():
    from oracle import ae as oa
    g = load("ae_tiny.npz")
    sd = sds()["ae"]
    with torch.no_grad():
        q, lat = oa.encode(sd, g["image"])
        assert (lat - g["latent"]).abs().max().item() < 1e-4
        assert torch.equal(q, g["quant"].float())                   # token grid: bit-exact
        assert (oa.decoder_forward(sd, g["quant"].float()) - g["decoded"]).abs().max().item() < 1e-3


def test_head_golden():
    from oracle import head as oh
    g = load("head_tiny.npz")
    sd = sds()["head"]
    with torch.no_grad():
        assert (oh.head_forward(sd, g["x"], g["t"], g["c"]) - g["xpred"]).abs().max().item() < 1e-4
        out = oh.euler_maruyama(sd, g["c"], g["cfg"], int(g["S"]), list(g["noise"]))
        assert (out - g["sample"]).abs().max().item() < 1e-3


def test_llm_golden():
    from bitdance_b200.synthetic import MODELS
    from oracle import llm as ol
    g = load("llm_tiny.npz")
    sd, cfg = sds()["llm"], MODELS["tiny"]["llm"]
    cache = [None] * cfg["num_hidden_layers"]
    with torch.no_grad():
        assert (ol.decoder_forward(sd, cfg, g["x0"], cache, causal=True) - g["h0"]).abs().max().item() < 1e-4
        assert (ol.decoder_forward(sd, cfg, g["x1"], cache, causal=False) - g["h1"]).abs().max().item() < 1e-4
        assert (ol.decoder_forward(sd, cfg, g["x2"], cache, causal=False) - g["h2"]).abs().max().item() < 1e-4


def test_pipeline_golden():
    from bitdance_b200.synthetic import MODELS
    from oracle import pipeline as op
    g = load("pipeline_tiny.npz")
    s = sds()
    m = MODELS["tiny"]
    pn, S = m["parallel_num"], int(g["S"])
    steps = 64 // pn
    noise = [list(g["noise"][i * (S + 1):(i + 1) * (S + 1)]) for i in range(steps)]
    start = [400, 401, 401] + [410 + i for i in range(1, pn)]
    with torch.no_grad():
        tok, img = op.gen_image(sd_llm=s["llm"], cfg_llm=m["llm"], embed=s["llm"]["model.embed_tokens.weight"],
                                sd_head=s["head"], sd_proj=s["proj"], sd_ae=s["ae"], cond_ids=[5, 17, 33, 2, 90],
                                uncond_ids=[3, 4], start_ids=start, h=8, w=8, pn=pn, num_images=int(g["B"]),
                                guidance=g["guidance"], S=S, noise=noise)
    assert (img - g["image"]).abs().max().item() < 2e-3 * max(1.0, g["image"].abs().max().item())
