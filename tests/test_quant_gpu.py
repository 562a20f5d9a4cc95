"""Binary quantiser / bit packing: bit-exact against the oracle (and torch) on the GPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 32, 16, 16), (2, 32, 7, 5), (1, 256, 8, 8), (3, 64, 1, 1)])
def test_sign_pack_nchw(dtype, shape):
    from bitdance_b200 import ops
    from oracle import quant as oq
    torch.manual_seed(0)
    h = torch.randn(*shape).to(dtype)
    h[0, 0, 0, 0] = 0.0
    h[0, 1, 0, 0] = float("nan")
    h[0, 2, 0, 0] = -0.0
    ncb = 4 if shape[1] == 32 else 0
    q, p, idx = ops.sign_pack_nchw(h.cuda(), num_codebooks=ncb)
    torch.cuda.synchronize()
    q_ref = oq.sign_quantize(h.float().numpy())
    assert np.array_equal(q.float().cpu().numpy(), q_ref)
    p_ref = oq.pack_bits_nchw(h.float().numpy())
    assert np.array_equal(p.cpu().numpy().view(np.uint32), p_ref)
    if ncb:
        i_ref = oq.gfq_indices(h.float().numpy(), ncb)
        assert np.array_equal(idx.cpu().numpy(), i_ref)


def test_sign_tokens_roundtrip():
    from bitdance_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(2, 64, 32)
    x[0, 0, 0] = 0.0
    x[0, 0, 1] = float("nan")
    t, p = ops.sign_tokens(x.cuda())
    torch.cuda.synchronize()
    from oracle import quant as oq
    ref = torch.sign(x)                       # torch: sign(NaN) == 0
    assert torch.equal(t.cpu(), ref)
    assert np.array_equal(t.cpu().numpy(), oq.sign_lfq(x.numpy()))
    bits = (x > 0).numpy().astype(np.uint32)
    words = (bits << np.arange(32, dtype=np.uint32)).sum(-1).astype(np.uint32)
    assert np.array_equal(p.cpu().numpy().view(np.uint32)[..., 0], words)
    for dt in (torch.float32, torch.bfloat16):
        u = ops.unpack_tokens(p, 32, dt)
        assert torch.equal(u.float().cpu(), torch.where(x > 0, 1.0, -1.0))


def test_sign_pack_full_size_property():
    """BASELINE size (8 x 256ch x 32x32 latent of a 1024^2 d32c256 image batch): pack -> unpack round trip."""
    from bitdance_b200 import ops
    torch.manual_seed(2)
    h = torch.randn(8, 256, 32, 32, device="cuda")
    q, p, _ = ops.sign_pack_nchw(h)
    u = ops.unpack_tokens(p, 256, torch.float32)          # [B, HW, C]
    assert torch.equal(u.permute(0, 2, 1).reshape(q.shape), q)
    assert torch.equal(q, torch.where(h > 0, 1.0, -1.0))
