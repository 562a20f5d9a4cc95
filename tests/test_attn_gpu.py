"""Flash-style attention kernel vs a plain torch fp32 reference (softmax(QK^T/sqrt(d)) V)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_attn(q, k, v, causal):
    B, Sq, Hq, D = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(Hq // Hkv, dim=1)
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(Hq // Hkv, dim=1)
    s = qf @ kf.transpose(-1, -2) * D ** -0.5
    if causal:
        qi = torch.arange(Sq, device=q.device)[:, None] + (Sk - Sq)
        kj = torch.arange(Sk, device=q.device)[None, :]
        s = s.masked_fill(kj > qi, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,Sq,Sk,Hq,Hkv,D,causal,splits", [
    (2, 64, 64, 4, 4, 128, False, 1),      # head block attention, 64x
    (2, 16, 16, 4, 4, 128, False, 1),      # 16x: partial tile
    (3, 16, 16, 8, 8, 64, False, 1),       # imagenet head_dim 64
    (2, 64, 300, 10, 2, 128, False, 1),    # GQA decode over a ragged cache
    (2, 64, 1500, 10, 2, 128, False, 0),   # auto split-KV
    (1, 64, 4352, 40, 8, 128, False, 0),   # Qwen3-14B heads, full 1024px context
    (2, 77, 77, 8, 2, 128, True, 1),       # causal prefill, ragged
    (1, 200, 200, 4, 4, 64, True, 2),      # causal with splits, several q tiles
    (1, 64, 140, 4, 2, 128, True, 1),      # causal with offset (Sk > Sq)
])
def test_attention_strided(B, Sq, Sk, Hq, Hkv, D, causal, splits):
    from bitdance_b200 import ops
    torch.manual_seed(0)
    q = torch.randn(B, Sq, Hq, D, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Sk, Hkv, D, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Sk, Hkv, D, device="cuda").to(torch.bfloat16)
    out = ops.attention(q, k, v, causal=causal, splits=splits)
    torch.cuda.synchronize()
    ref = ref_attn(q, k, v, causal)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, f"max err {err}"


def test_attention_qkv_packed_layout():
    """q/k/v as strided views of one [M, 3D] buffer (the head's wqkv output)."""
    from bitdance_b200 import ops
    torch.manual_seed(1)
    R, pn, H, D = 2, 64, 6, 128
    qkv = torch.randn(R, pn, 3, H, D, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    out = ops.attention(q, k, v)
    ref = ref_attn(q, k, v, False)
    assert (out.float() - ref).abs().max().item() < 2e-2


def test_attention_paged():
    from bitdance_b200 import ops
    torch.manual_seed(2)
    B, Sq, Hq, Hkv, D = 2, 64, 10, 2, 128
    Sk = 500
    n_tiles = (Sk + 63) // 64
    pool_pages = 40
    kp = torch.randn(pool_pages, Hkv, 64, D, device="cuda").to(torch.bfloat16)
    vp = torch.randn(pool_pages, Hkv, 64, D, device="cuda").to(torch.bfloat16)
    perm = torch.randperm(pool_pages)[: B * n_tiles].view(B, n_tiles).to(torch.int32).cuda()
    q = torch.randn(B, Sq, Hq, D, device="cuda").to(torch.bfloat16)
    out = ops.attention(q, kp, vp, page_table=perm, sk=Sk, splits=2)
    # gather the logical K/V
    k = kp[perm.long()].permute(0, 1, 3, 2, 4).reshape(B, n_tiles * 64, Hkv, D)[:, :Sk]
    v = vp[perm.long()].permute(0, 1, 3, 2, 4).reshape(B, n_tiles * 64, Hkv, D)[:, :Sk]
    ref = ref_attn(q, k, v, False)
    assert (out.float() - ref).abs().max().item() < 2e-2
