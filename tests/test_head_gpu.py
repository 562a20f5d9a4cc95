"""bd_head_sample (x-prediction transformer + Euler-Maruyama sampler) vs the CPU oracle on identical weights/noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(cfg, seed=1, std=0.05):
    from bitdance_b200.head import HeadRunner, head_spec
    from bitdance_b200.synth import synth_state_dict
    spec = head_spec(cfg["ch_target"], cfg["ch_cond"], cfg["ch_latent"], cfg["depth_latent"], cfg["depth_adanln"],
                     cfg.get("use_swiglu", True))
    sd = synth_state_dict(spec, seed=seed, std=std)
    runner = HeadRunner(sd, ch_target=cfg["ch_target"], ch_cond=cfg["ch_cond"], ch_latent=cfg["ch_latent"],
                        depth_latent=cfg["depth_latent"], depth_adanln=cfg["depth_adanln"],
                        use_swiglu=cfg.get("use_swiglu", True), head_dim=cfg.get("head_dim", 128),
                        out_sigmoid=cfg.get("out_sigmoid", True))
    return sd, runner


@pytest.mark.parametrize("cfg,B,pn,guidance,S", [
    (dict(ch_target=32, ch_cond=256, ch_latent=256, depth_latent=4, depth_adanln=2), 1, 64, 3.0, 6),
    (dict(ch_target=32, ch_cond=256, ch_latent=256, depth_latent=4, depth_adanln=2), 2, 16, 7.5, 4),
    (dict(ch_target=32, ch_cond=384, ch_latent=512, depth_latent=2, depth_adanln=1, use_swiglu=False), 3, 16, 1.0, 3),
    (dict(ch_target=16, ch_cond=256, ch_latent=256, depth_latent=2, depth_adanln=2, head_dim=64, out_sigmoid=False),
     2, 4, 2.0, 5),
])
@pytest.mark.parametrize("path", ["stream", "tiled"])
def test_head_sample_vs_oracle(cfg, B, pn, guidance, S, path):
    """path "stream": the whole sampler as ONE persistent kernel (bd_stream.cuh); "tiled": the multi-kernel path."""
    from oracle import head as oh
    sd, runner = make(cfg)
    torch.manual_seed(0)
    mult = 2 if guidance > 1.0 else 1
    z = torch.randn(B * mult, pn, cfg["ch_cond"])
    noise = torch.randn(S + 1, B, pn, cfg["ch_target"])
    x, trace = runner.sample(z.cuda(), guidance, S, noise=noise.cuda(), trace=True, path=path)
    torch.cuda.synchronize()
    hd, osig = cfg.get("head_dim", 128), cfg.get("out_sigmoid", True)
    # (1) teacher-forced parity: drive the oracle sampler with the GPU's own network outputs. Then
    #   * at every evaluation the oracle network sees exactly the state the GPU saw: outputs must agree to a few
    #     bf16 ulps (accumulation order, fast exp);
    #   * the fp32 sampler arithmetic itself must reproduce the GPU state essentially exactly.
    tr = []
    forced = [trace[i].cpu().view(B * mult, pn, -1) for i in range(S + 1)]
    x_tf = oh.euler_maruyama(sd, z, guidance, S, list(noise), rnd=oh.bf16, head_dim=hd, out_sigmoid=osig, trace=tr,
                             forced_out=forced)[:B]
    scale = max(1.0, max(t["out"].abs().max().item() for t in tr))
    e_net = max((forced[i] - tr[i]["out"]).abs().max().item() for i in range(S + 1))
    e_sde = (x.cpu() - x_tf).abs().max().item()
    # (2) free-running: chaotic amplification of those ulps through the SDE (reported, loosely bounded)
    ref = oh.euler_maruyama(sd, z, guidance, S, list(noise), rnd=oh.bf16, head_dim=hd, out_sigmoid=osig)[:B]
    d = (x.cpu() - ref).abs()
    agree = (torch.sign(x.cpu()) == torch.sign(ref)).float().mean().item()
    print(f"head parity [{path}]: teacher-forced net max err {e_net:.4f} (scale {scale:.2f}), sampler err {e_sde:.2e}; "
          f"free-running max {d.max().item():.4f} mean {d.mean().item():.5f} sign agreement {agree:.4f}")
    assert e_net < 4.7e-2 * scale, f"network output err {e_net}"   # 6 bf16 ulps at |out| ~ 1 (2^-8 spacing)
    assert e_sde < 1e-4 * max(1.0, x_tf.abs().max().item()), f"sampler arithmetic err {e_sde}"
    assert agree > 0.95, f"free-running sign agreement {agree}"


def test_head_sampler_deterministic_and_seeded_noise():
    cfg = dict(ch_target=32, ch_cond=256, ch_latent=256, depth_latent=2, depth_adanln=2)
    sd, runner = make(cfg)
    z = torch.randn(2, 16, 256, device="cuda")
    for path in ("stream", "tiled"):
        torch.manual_seed(7)
        a = runner.sample(z, 2.0, 3, path=path)
        torch.manual_seed(7)
        b = runner.sample(z, 2.0, 3, path=path)
        assert torch.equal(a, b), path
    # the noise consumption matches the reference sampler's call sequence: randn(x0) then randn_like per step
    torch.manual_seed(7)
    n0 = torch.randn(1, 16, 32, device="cuda")
    n1 = torch.randn_like(n0)
    torch.manual_seed(7)
    nz = runner.draw_noise(1, 16, 3)
    assert torch.equal(nz[0], n0) and torch.equal(nz[1], n1)


def test_head_full_size_properties():
    """BASELINE size (BitDance-14B-64x head: D=5120, 6 blocks, 1.76 B parameters; R=2, pn=64, S=50): too big for the CPU
    oracle, so size-independent properties — the persistent kernel is run-to-run bit-identical, its output does not depend
    on the engine's ring split, and it agrees with the multi-kernel path where agreement is meaningful (see below)."""
    from bitdance_b200 import _lib
    from bitdance_b200.head import HeadRunner, head_spec
    from bitdance_b200.synthetic import MODELS, _gpu_state_dict
    hc = MODELS["BitDance-14B-64x"]["head"]
    dev = torch.device("cuda")
    sd = _gpu_state_dict(head_spec(hc["ch_target"], hc["ch_cond"], hc["ch_latent"], hc["depth_latent"], hc["depth_adanln"],
                                   hc["use_swiglu"]), 2, dev)
    head = HeadRunner(sd, device=dev, **hc)
    lib = _lib.load()
    R, pn, S = 2, 64, 50
    torch.manual_seed(0)
    z = torch.randn(R, pn, hc["ch_cond"], device=dev)
    noise = head.draw_noise(1, pn, S)
    a = head.sample(z, 7.5, S, noise=noise, path="stream")
    b = head.sample(z, 7.5, S, noise=noise, path="stream")
    assert torch.equal(a, b), "persistent kernel not deterministic"
    assert torch.isfinite(a).all() and a.abs().max().item() < 50.0
    lib.bd_stream_set_tuning(4, 3, 0)
    c = head.sample(z, 7.5, S, noise=noise, path="stream")
    lib.bd_stream_set_tuning(5, 2, 0)
    assert torch.equal(a, c), "result depends on the ring split"
    # BASELINE-size parity against the CPU oracle (bf16 rounding policy) on the FIRST evaluation, where every implementation
    # sees identical inputs (x = cat[noise0, noise0], t = 0): one oracle evaluation of the 1.76 B-parameter head takes
    # ~10-20 s on the host cores. Both GPU paths are checked; after that the sampler is chaotic with random-init weights
    # (CFG 7.5 multiplies every difference), so later evaluations are only compared teacher-forced on small models above.
    from oracle import head as oh
    xs, ts = head.sample(z, 7.5, S, noise=noise, path="stream", trace=True)
    xt, tt = head.sample(z, 7.5, S, noise=noise, path="tiled", trace=True)
    assert torch.equal(xs, a)
    sd_cpu = {k: v.float().cpu() for k, v in sd.items()}
    x0 = torch.cat([noise[0], noise[0]], dim=0).cpu()
    with torch.no_grad():
        ref = oh.head_forward(sd_cpu, x0, torch.zeros(R), z.cpu(), rnd=oh.bf16).reshape(R * pn, -1)
    del sd_cpu
    for name, tr in (("stream", ts), ("tiled", tt)):
        d0 = (tr[0].cpu() - ref).abs()
        print(f"14B-64x head [{name}] vs oracle, first evaluation: |diff| max {d0.max().item():.4f} mean {d0.mean().item():.5f} "
              f"(outputs in (-1, 1), |ref| mean {ref.abs().mean().item():.3f})")
        # 4096 outputs of a 6-block, D = 5120 bf16 network: mean at the bf16 rounding level, max a few tens of ulps
        assert d0.mean().item() < 2e-2 and d0.max().item() < 0.25, name
    dst = (ts[0] - tt[0]).abs()
    print(f"  stream vs tiled: max {dst.max().item():.4f} mean {dst.mean().item():.5f}")
    S2 = 3
    noise2 = head.draw_noise(2, pn, S2)
    a2 = head.sample(z, 1.0, S2, noise=noise2, path="stream")
    t2 = head.sample(z, 1.0, S2, noise=noise2, path="tiled")
    agree = (torch.sign(a2) == torch.sign(t2)).float().mean().item()
    # reported, loosely bounded: with random-init weights the x-prediction is ~0, so the signs of the sample hang on rounding
    print(f"  3 steps without guidance: sign agreement {agree:.4f}, mean |diff| {(a2 - t2).abs().mean().item():.5f}")
    assert agree > 0.6


@pytest.mark.parametrize("path", ["stream", "tiled"])
def test_head_saturating_weights_exact_token_grid(path):
    """SURVEY.md section 7, contract (c): a TRAINED BitDance head predicts bits — its x-prediction is decided by the
    condition c (the LLM hidden state) and saturates to +-1; a random-init head is instead a chaotic map of x (fp32 and
    bf16 runs of the reference algorithm itself then agree on only ~60 % of the signs). Trained-like weights here = the
    usual synthetic network with input_proj x 1e-3 (the prediction hangs on c, through adaLN and cond_embed) and the
    final Linear x 30 (saturation). With them the FREE-RUNNING token grid of the GPU sampler (10 steps, CFG 3) agrees with
    the oracle's on >= 99 % of the entries (measured: 4 of 1024 differ on the multi-kernel path; the oracle's own fp32 vs
    bf16 modes differ on 1 of 1024) — the ~15 % of outputs that are NOT saturated still hang on bf16 rounding."""
    from oracle import head as oh
    from bitdance_b200.head import HeadRunner, head_spec
    from bitdance_b200.synth import synth_state_dict
    sd = synth_state_dict(head_spec(32, 256, 256, 4, 2, True), seed=1, std=0.05)
    sd["net.input_proj.weight"] = sd["net.input_proj.weight"] * 1e-3
    sd["net.final_layer.linear.weight"] = sd["net.final_layer.linear.weight"] * 30.0
    runner = HeadRunner(sd, ch_target=32, ch_cond=256, ch_latent=256, depth_latent=4, depth_adanln=2, use_swiglu=True)
    torch.manual_seed(0)
    B, pn, S, guidance = 2, 16, 10, 3.0
    z = torch.randn(2 * B, pn, 256)
    noise = torch.randn(S + 1, B, pn, 32)
    x, trace = runner.sample(z.cuda(), guidance, S, noise=noise.cuda(), trace=True, path=path)
    with torch.no_grad():
        ref = oh.euler_maruyama(sd, z, guidance, S, list(noise), rnd=oh.bf16)[:B]
    sat = (trace[-1].abs() > 0.99).float().mean().item()
    tok, tok_ref = torch.sign(x.cpu()), torch.sign(ref)
    agree = (tok == tok_ref).float().mean().item()
    safe = ref.abs() > 0.25
    print(f"saturating head [{path}]: {sat:.3f} of the last evaluation's outputs saturated, token-grid agreement "
          f"{agree:.5f} ({int((tok != tok_ref).sum())} of {tok.numel()} differ), safe fraction {safe.float().mean().item():.4f}, "
          f"max |x - ref| {(x.cpu() - ref).abs().max().item():.4f}")
    assert sat > 0.8
    assert agree >= 0.99
