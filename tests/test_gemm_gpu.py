"""tcgen05 weight-streaming GEMM vs a plain torch fp32 reference of the same op (same rounding points)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16).float()


def ref_linear(a, w, bias=None, act=None, swiglu=False, gate=None, res=None, out_dtype=torch.bfloat16):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    y = bf(y)
    if swiglu:
        N = w.shape[0]
        y = y.view(y.shape[0], N // 32, 2, 16)
        g, u = y[:, :, 0], y[:, :, 1]
        y = bf(bf(torch.nn.functional.silu(g)) * u).reshape(a.shape[0], N // 2)
    if act == "silu":
        y = bf(torch.nn.functional.silu(y))
    if act == "gelu_tanh":
        y = bf(torch.nn.functional.gelu(y, approximate="tanh"))
    if gate is not None:
        y = bf(y * gate.float())
    if res is not None:
        y = y + res.float()
    return y.to(out_dtype)


def check_close(out, ref, K, tag):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    # bf16 output rounding (2^-8 relative) + fp32 accumulation-order noise
    assert err <= 1.5e-2 * scale + 1e-3, f"{tag}: max err {err} vs scale {scale}"


@pytest.mark.parametrize("M,N,K,bn,splits", [
    (128, 256, 128, 128, 1),
    (128, 512, 512, 64, 1),
    (128, 512, 1024, 256, 1),
    (128, 5120, 5120, 0, 0),      # auto plan: split-K path
    (128, 15360, 5120, 0, 0),
    (128, 1024, 5120, 128, 4),
    (100, 384, 200, 128, 1),      # ragged M, K not a multiple of 64, TMA zero fill
    (2, 5120, 256, 0, 0),         # time-embed shape
    (128, 5120, 32, 0, 0),        # input_proj / fc1 shape (K < BK)
    (300, 640, 320, 128, 2),      # several M tiles + split
    (1024, 2048, 512, 256, 1),    # bs=8 rows
])
@pytest.mark.parametrize("tiled", [False, True])
def test_gemm_plain(M, N, K, bn, splits, tiled):
    from bitdance_b200 import ops
    torch.manual_seed(0)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16)
    wk = ops.pack_weight(w) if tiled else w      # tile-major HBM layout (bd_pack_weight_tiles) or nn.Linear layout
    out = ops.gemm(a, wk, bias=bias, bn=bn, splits=splits)
    torch.cuda.synchronize()
    check_close(out, ref_linear(a, w, bias), K, f"plain {M}x{N}x{K}")
    out32 = ops.gemm(a, wk, out_dtype=torch.float32, bn=bn, splits=splits)
    torch.cuda.synchronize()
    check_close(out32, ref_linear(a, w, out_dtype=torch.float32), K, "f32 out")
    if tiled and bn == 0:   # same tiles, same accumulation order: the two layouts must agree bit for bit
        assert torch.equal(out, ops.gemm(a, w, bias=bias, bn=bn, splits=splits))


@pytest.mark.parametrize("splits", [1, 3])
def test_gemm_epilogues(splits):
    from bitdance_b200 import ops
    torch.manual_seed(1)
    M, N, K = 128, 1024, 768
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16)
    gate_full = (torch.randn(M, 3 * N, device="cuda")).to(torch.bfloat16)
    gate = gate_full[:, N:2 * N]                      # strided view, like an adaLN chunk
    res_bf = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    res_f32 = torch.randn(M, N, device="cuda")
    for act in ("silu", "gelu_tanh"):
        out = ops.gemm(a, w, bias=bias, act=act, splits=splits)
        check_close(out, ref_linear(a, w, bias, act=act), K, act)
    out = ops.gemm(a, w, bias=bias, gate=gate, res=res_bf, splits=splits)
    check_close(out, ref_linear(a, w, bias, gate=gate, res=res_bf), K, "gated res bf16")
    out = ops.gemm(a, w, res=res_f32, out_dtype=torch.float32, splits=splits)
    check_close(out, ref_linear(a, w, res=res_f32, out_dtype=torch.float32), K, "res f32")
    # SwiGLU with interleave16 prepack
    F = N // 2
    wg, wu = w[:F].contiguous(), w[F:].contiguous()
    bg, bu = bias[:F].contiguous(), bias[F:].contiguous()
    wi, bi = ops.interleave16(wg, wu, bg, bu)
    out = ops.gemm(a, wi, bias=bi, swiglu=True, splits=splits)
    assert torch.equal(out, ops.gemm(a, ops.pack_weight(wi), bias=bi, swiglu=True, splits=splits))
    g = bf(a.float() @ wg.float().t() + bg.float())
    u = bf(a.float() @ wu.float().t() + bu.float())
    ref = bf(bf(torch.nn.functional.silu(g)) * u)
    check_close(out, ref, K, "swiglu")


def test_gemm_pdl_chain():
    """Back-to-back PDL launches (weights prefetched before griddepcontrol.wait) stay correct."""
    from bitdance_b200 import ops
    torch.manual_seed(2)
    M, D = 128, 2048
    x = (torch.randn(M, D, device="cuda") * 0.5).to(torch.bfloat16)
    ws = [(torch.randn(D, D, device="cuda") * (1.0 / D ** 0.5)).to(torch.bfloat16) for _ in range(6)]
    y = x
    for w in ws:
        y = ops.gemm(y, w, pdl=True)
    torch.cuda.synchronize()
    r = x
    for w in ws:
        r = ref_linear(r, w)
    check_close(y, r, D, "pdl chain")


def test_gemm_deterministic():
    from bitdance_b200 import ops
    torch.manual_seed(3)
    a = torch.randn(128, 5120, device="cuda").to(torch.bfloat16)
    w = (torch.randn(5120, 5120, device="cuda") * 0.02).to(torch.bfloat16)
    o1 = ops.gemm(a, w)
    o2 = ops.gemm(a, w)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("M,N,K", [
    (256, 256, 64),          # one cluster, one k-block
    (1024, 5120, 5120),      # bs=8 head / LLM Linear
    (1000, 768, 768 + 32),   # ragged M (TMA zero fill in the second CTA of the last pair), ragged K
    (384, 15360, 512),       # M = 1.5 pairs: the odd CTA of the last pair has no valid rows
    (4096, 2304, 768),       # ImageNet-B qkv at batch 256 x 16 tokens
])
def test_gemm_cta_pair_kernel(M, N, K):
    """bd_gemm2_kernel (tcgen05.mma.cta_group::2, clusters of two CTAs on 256 x 256 tiles; chosen automatically for
    tile-major weights, M >= 256, N % 256 == 0) against torch and — bit for bit — against the 1-CTA kernel (same K order,
    same fp32 accumulation), with the bias / residual / SwiGLU epilogues."""
    from bitdance_b200 import ops
    torch.manual_seed(0)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16)
    res = torch.randn(M, N, device="cuda")
    pw = ops.pack_weight(w)
    out2 = ops.gemm(a, pw, bias=bias)                     # auto plan -> CTA-pair kernel
    out1 = ops.gemm(a, pw, bias=bias, bn=256, splits=1)   # explicit tile: the 1-CTA kernel
    check_close(out2, ref_linear(a, w, bias), K, "pair")
    assert torch.equal(out2, out1), "CTA-pair and 1-CTA kernels differ"
    o32 = ops.gemm(a, pw, bias=bias, res=res, out_dtype=torch.float32)
    check_close(o32, ref_linear(a, w, bias, res=res, out_dtype=torch.float32), K, "pair+res")
    if N % 32 == 0:
        wi, bi = ops.interleave16(w[:N // 2].contiguous(), w[N // 2:].contiguous(), bias[:N // 2].contiguous(),
                                  bias[N // 2:].contiguous())
        osw = ops.gemm(a, ops.pack_weight(wi), bias=bi, swiglu=True)
        y = bf(a.float() @ w.float().t() + bias.float())
        ref = bf(bf(torch.nn.functional.silu(y[:, :N // 2])) * y[:, N // 2:])
        check_close(osw, ref, K, "pair+swiglu")
