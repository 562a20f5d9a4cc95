"""Binary tokenizer on the GPU (tcgen05 implicit-GEMM convs + NHWC GroupNorm kernels) vs the CPU oracle / torch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DD_SMALL = dict(double_z=False, z_channels=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2], num_res_blocks=2)


def bf(x):
    return x.to(torch.bfloat16).float()


def make(dd, seed=2, std=0.05):
    from bitdance_b200.ae import AERunner, ae_spec
    from bitdance_b200.synth import synth_state_dict
    sd = synth_state_dict(ae_spec(dd), seed=seed, std=std)
    return sd, AERunner(sd, dd)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,bias", [
    (1, 16, 16, 64, 128, 3, 1, False),
    (2, 8, 12, 32, 64, 3, 1, True),       # ragged tile (W=12), Cin < 64 (TMA zero fill of channels)
    (1, 64, 64, 256, 256, 3, 1, False),   # full tiles, several k-blocks per tap
    (1, 20, 36, 128, 256, 3, 2, True),    # stride 2 via the 4-phase split
    (2, 9, 7, 64, 32, 1, 1, True),        # 1x1 (nin_shortcut / conv_out)
    (1, 130, 5, 8, 64, 3, 1, False),      # padded 3->8 channel image conv, tall thin image
])
def test_conv_vs_torch(B, H, W, Cin, Cout, k, stride, bias):
    from bitdance_b200.ae import AERunner
    torch.manual_seed(0)
    w = torch.randn(Cout, Cin, k, k) * 0.05
    b = torch.randn(Cout) * 0.1 if bias else None
    sd = {"c.weight": w}
    if bias:
        sd["c.bias"] = b
    run = AERunner(sd, DD_SMALL)
    x = torch.randn(B, Cin, H, W)
    xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    out = run._conv("c", xn, B, H, W, stride=stride)
    torch.cuda.synchronize()
    ref = F.conv2d(bf(x), bf(w), None if b is None else bf(b), stride=stride, padding=k // 2)
    ref = bf(ref).permute(0, 2, 3, 1)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 1.5e-2 * ref.abs().max().item() + 1e-3, f"conv err {err} (scale {ref.abs().max().item()})"


def test_conv_epilogues():
    from bitdance_b200.ae import AERunner
    torch.manual_seed(1)
    B, H, W, C = 2, 8, 16, 64
    w = torch.randn(4 * C, C, 3, 3) * 0.05
    b = torch.randn(4 * C) * 0.1
    w3 = torch.randn(3, C, 3, 3) * 0.05
    b3 = torch.randn(3) * 0.1
    wr = torch.randn(C, C, 3, 3) * 0.05
    run = AERunner({"up.weight": w, "up.bias": b, "o.weight": w3, "o.bias": b3, "r.weight": wr}, DD_SMALL)
    x = torch.randn(B, C, H, W)
    xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    # depth-to-space scatter (Upsampler)
    out = run._conv("up", xn, B, H, W, out_mode=1).float().cpu()
    from oracle.ae import depth_to_space
    ref = depth_to_space(bf(F.conv2d(bf(x), bf(w), bf(b), padding=1))).permute(0, 2, 3, 1)
    assert (out - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-3
    # NCHW 3-channel output (decoder conv_out)
    out = run._conv("o", xn, B, H, W, out_mode=2).float().cpu()
    ref = bf(F.conv2d(bf(x), bf(w3), bf(b3), padding=1))
    assert (out - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-3
    # residual add, fp32 stream (decoder) and bf16 stream (encoder)
    res32 = torch.randn(B, H, W, C)
    out = run._conv("r", xn, B, H, W, res=res32.cuda()).cpu()
    ref = res32 + bf(F.conv2d(bf(x), bf(wr), None, padding=1)).permute(0, 2, 3, 1)
    assert out.dtype == torch.float32 and (out - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    out = run._conv("r", xn, B, H, W, res=res32.to(torch.bfloat16).cuda()).float().cpu()
    ref = bf(bf(res32) + bf(F.conv2d(bf(x), bf(wr), None, padding=1)).permute(0, 2, 3, 1))
    assert (out - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("C,dtype", [(32, torch.bfloat16), (256, torch.float32), (1024, torch.bfloat16)])
def test_groupnorm_swish(C, dtype):
    _, run = make(DD_SMALL)
    torch.manual_seed(3)
    B, HW = 2, 1500
    x = (torch.randn(B, HW, C) * 2 + 0.5).to(dtype)
    w, b = torch.randn(C), torch.randn(C)
    out = run._gn(x.cuda(), B, HW, C, w.cuda(), b.cuda(), 0).float().cpu()
    xr = x.float().permute(0, 2, 1)
    g = F.group_norm(xr, 32, w, b, eps=1e-6)
    ref = bf(g * torch.sigmoid(g)).permute(0, 2, 1)
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item() + 1e-3


def test_tokenizer_roundtrip_vs_oracle():
    from oracle import ae as oa
    sd, run = make(DD_SMALL)
    torch.manual_seed(0)
    x = torch.rand(2, 3, 32, 48) * 2 - 1
    q, packed, idx, lat = run.encode(x.cuda(), num_codebooks=4)
    torch.cuda.synchronize()
    q_ref, lat_ref = oa.encode(sd, x, rnd=oa.bf16)
    e_lat = (lat.float().cpu() - lat_ref).abs().max().item()
    assert e_lat < 3e-2 * lat_ref.abs().max().item(), f"latent err {e_lat}"
    # token grid: bit-exact wherever the oracle latent is not within rounding noise of zero
    safe = lat_ref.abs() > e_lat + 1e-3
    assert torch.equal(q.float().cpu()[safe], q_ref[safe])
    agree = (q.float().cpu() == q_ref).float().mean().item()
    print(f"AE parity: token agreement {agree:.4f}, safe fraction {safe.float().mean().item():.3f}")
    assert agree > 0.97
    # quantiser itself is bit-exact on the GPU latent
    assert torch.equal(q.float().cpu(), torch.where(lat.float().cpu() > 0, 1.0, -1.0))
    # decoder from the SAME token grid on both sides
    dec = run.decode(q_ref.cuda()).float().cpu()
    dec_ref = oa.decoder_forward(sd, q_ref, rnd=oa.bf16)
    e = (dec - dec_ref).abs().max().item()
    print(f"AE parity: latent err {e_lat:.4f}, decode err {e:.4f} (scale {dec_ref.abs().max().item():.2f})")
    assert e < 4e-2 * dec_ref.abs().max().item() + 2e-2, f"decode err {e}"
    # patch-raster token path (decode_image)
    ps = 2
    B, C, h, w = q_ref.shape
    tok = q_ref.view(B, C, h // ps, ps, w // ps, ps).permute(0, 2, 4, 3, 5, 1).reshape(B, h * w, C).contiguous()
    dec2 = run.decode_tokens(tok.cuda(), h, w, ps).float().cpu()
    assert torch.equal(dec2, dec)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,bias", [
    (1, 16, 16, 1024, 1024, 3, 1, False),   # ae_d16c32 level 4 / mid ResBlock conv (autoencoder.py:41-57)
    (1, 16, 16, 512, 1024, 3, 1, False),    # channel-changing ResBlock conv1
    (1, 16, 16, 512, 1024, 1, 1, False),    # its 1x1 nin_shortcut
    (1, 16, 16, 1024, 32, 1, 1, True),      # encoder conv_out (1x1, 1024 -> z)
    (1, 16, 16, 32, 1024, 3, 1, True),      # decoder conv_in
    (1, 32, 32, 512, 512, 3, 2, True),      # Downsample (stride 2, bias) at 512 channels
])
def test_conv_baseline_channels_vs_torch(B, H, W, Cin, Cout, k, stride, bias):
    """The channel counts the BASELINE tokenizer (ae_d16c32: ch 256 -> 1024) actually runs: 16 k-blocks per tap, 9 taps."""
    test_conv_vs_torch(B, H, W, Cin, Cout, k, stride, bias)


def test_conv_upsampler_1024_to_4096_depth_to_space():
    """Upsampler of the deepest decoder level: conv 1024 -> 4096 + depth-to-space (autoencoder.py:198-249)."""
    from bitdance_b200.ae import AERunner
    from oracle.ae import depth_to_space
    torch.manual_seed(5)
    B, H, W, C = 1, 16, 16, 1024
    w = torch.randn(4 * C, C, 3, 3) * 0.02
    b = torch.randn(4 * C) * 0.1
    run = AERunner({"up.weight": w, "up.bias": b}, DD_SMALL)
    x = torch.randn(B, C, H, W)
    xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    out = run._conv("up", xn, B, H, W, out_mode=1).float().cpu()
    ref = depth_to_space(bf(F.conv2d(bf(x), bf(w), bf(b), padding=1))).permute(0, 2, 3, 1)
    assert (out - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-3


def test_tokenizer_ae_d16c32_256px_vs_oracle():
    """BASELINE.json configs[0]: the real ae_d16c32 tokenizer (ch=256, ch_mult=[1,1,2,2,4], 4 ResBlocks per level,
    460 M parameters) on one 256x256 image: encode -> sign -> pack -> decode against the CPU oracle (oracle/ae.py, bf16
    rounding policy; ~10 s on the host). Token grid bit-exact wherever |latent| exceeds the latent error, packed bits
    equal to the oracle's packing of the same signs, decode compared from the SAME token grid."""
    import numpy as np
    from bitdance_b200.synthetic import AE_D16C32
    from oracle import ae as oa
    from oracle import quant as oq
    sd, run = make(AE_D16C32, seed=3, std=0.02)
    torch.manual_seed(2)
    x = torch.rand(1, 3, 256, 256) * 2 - 1
    q, packed, idx, lat = run.encode(x.cuda(), num_codebooks=4)
    torch.cuda.synchronize()
    with torch.no_grad():
        q_ref, lat_ref = oa.encode(sd, x, rnd=oa.bf16)
    assert tuple(q.shape) == (1, 32, 16, 16)
    e_lat = (lat.float().cpu() - lat_ref).abs().max().item()
    scale = lat_ref.abs().max().item()
    assert e_lat < 3e-2 * scale + 1e-3, f"latent err {e_lat} (scale {scale})"
    safe = lat_ref.abs() > e_lat + 1e-3 * scale
    assert torch.equal(q.float().cpu()[safe], q_ref[safe])
    agree = (q.float().cpu() == q_ref).float().mean().item()
    print(f"ae_d16c32 256px: latent err {e_lat:.4f} (scale {scale:.3f}), token agreement {agree:.4f}, "
          f"safe fraction {safe.float().mean().item():.3f}")
    assert agree > 0.97
    # quantise / pack / GFQ indices of the GPU latent: bit-exact against the oracle's integer code
    lat_np = lat.float().cpu().numpy()
    assert np.array_equal(q.float().cpu().numpy(), oq.sign_quantize(lat_np))
    assert np.array_equal(packed.cpu().numpy().view(np.uint32), oq.pack_bits_nchw(lat_np))
    assert np.array_equal(idx.cpu().numpy(), oq.gfq_indices(lat_np, 4))
    # decode from the SAME token grid
    dec = run.decode(q_ref.cuda()).float().cpu()
    with torch.no_grad():
        dec_ref = oa.decoder_forward(sd, q_ref, rnd=oa.bf16)
    e = (dec - dec_ref).abs().max().item()
    print(f"ae_d16c32 256px: decode err {e:.4f} mean {(dec - dec_ref).abs().mean().item():.5f} "
          f"(scale {dec_ref.abs().max().item():.2f})")
    assert dec.shape == (1, 3, 256, 256)
    # 57 convolutions deep (5 levels x 4 ResBlocks + mid + 4 upsamplers): max error of a few bf16 ulps of the output range,
    # mean at the bf16 rounding level (measured on B200: max 0.24 = 5.5 % of scale, mean 0.013 = 0.3 %)
    assert e < 8e-2 * dec_ref.abs().max().item() + 2e-2, f"decode err {e}"
    assert (dec - dec_ref).abs().mean().item() < 1e-2 * dec_ref.abs().max().item()
