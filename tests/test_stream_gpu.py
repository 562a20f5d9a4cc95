"""The persistent weight-streaming engine (csrc/bd_stream.cuh): layout helpers on CPU, GEMM ops vs torch on the GPU.

Tolerances: bf16 output rounding + fp32 accumulation order (the engine rotates the K loop per CTA), as for bd_gemm_bf16."""
import pytest
import torch


def test_blocked_layout_roundtrip_cpu():
    """to_blocked / from_blocked are inverse and follow the 128B-swizzle rule chunk' = chunk ^ (row % 8)."""
    from bitdance_b200 import ops
    a = torch.arange(100 * 200, dtype=torch.float32).view(100, 200).to(torch.bfloat16)
    b = ops.to_blocked(a)
    assert b.numel() == 4 * 8192
    assert torch.equal(ops.from_blocked(b, 100, 200), a)
    # element (row 3, col 8*5+2) of k-block 1 sits in 16-byte chunk 5^3 = 6 of its 128-byte row
    r, c = 3, 64 + 8 * 5 + 2
    assert b[8192 + r * 64 + (5 ^ 3) * 8 + 2] == a[r, c]
    # padding is zero
    assert float(b.float().abs().sum()) == float(a.float().abs().sum())


def _ref_linear(a, w, bias):
    return a.float() @ w.float().t() + (bias.float() if bias is not None else 0.0)


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,act,blocked", [
    (128, 2048 + 16, 512, None, False),
    (128, 7680, 1024, "silu", True),
    (64, 1024, 256 + 32, None, False),     # ragged K (zero padded k-block), M < 128
    (16, 160, 64, "gelu_tanh", False),     # fewer 16-row units than CTAs
    (128, 148 * 144, 320, None, True),     # 9 units per CTA: ONE wide pass of 144 rows (1 k-block per ring step), odd k-blocks
    (128, 148 * 160 - 32, 256, "silu", False),  # 9-10 units per CTA: passes of 144 / 160 rows
])
def test_stream_gemm_bias(M, N, K, act, blocked):
    from bitdance_b200 import ops
    torch.manual_seed(0)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    b = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16)
    sw = ops.stream_pack_weight(w, b)
    out = ops.stream_gemm(a, sw, epi="bias", act=act, out_blocked=blocked).float()
    ref = _bf(_ref_linear(a, w, b))
    if act == "silu":
        ref = _bf(torch.nn.functional.silu(ref))
    if act == "gelu_tanh":
        ref = _bf(torch.nn.functional.gelu(ref, approximate="tanh"))
    err = (out - ref).abs().max().item()
    assert err <= 1.5e-2 * ref.abs().max().item() + 1e-3, err


@pytest.mark.gpu
@pytest.mark.parametrize("M,hidden,K", [(128, 1536, 512), (48, 384, 256), (128, 148 * 72, 256)])
def test_stream_gemm_swiglu(M, hidden, K):
    from bitdance_b200 import ops
    torch.manual_seed(1)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(2 * hidden, K, device="cuda") * 0.05).to(torch.bfloat16)
    b = (torch.randn(2 * hidden, device="cuda") * 0.1).to(torch.bfloat16)
    sw = ops.stream_pack_weight(w, b, swiglu=True)
    out = ops.stream_gemm(a, sw, epi="swiglu", out_blocked=True).float()
    y = _bf(_ref_linear(a, w, b))
    ref = _bf(_bf(torch.nn.functional.silu(y[:, :hidden])) * y[:, hidden:])
    err = (out - ref).abs().max().item()
    assert out.shape == (M, hidden)
    assert err <= 1.5e-2 * ref.abs().max().item() + 1e-3, err


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,ksplit", [(128, 1024, 1024, 4), (128, 5120, 2048, 4), (32, 256, 1024, 1)])
def test_stream_gemm_partials(M, N, K, ksplit):
    from bitdance_b200 import ops
    torch.manual_seed(2)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    sw = ops.stream_pack_weight(w, None, ksplit=ksplit)
    part = ops.stream_gemm(a, sw, epi="partial")
    assert part.shape == (ksplit, M, N)
    ref = _ref_linear(a, w, None)
    err = (part.sum(0) - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, err
    # each split holds exactly its K range
    ks = K // ksplit
    for s in range(ksplit):
        rs = a[:, s * ks:(s + 1) * ks].float() @ w[:, s * ks:(s + 1) * ks].float().t()
        assert (part[s] - rs).abs().max().item() <= 2e-3 * rs.abs().max().item() + 1e-4


@pytest.mark.gpu
def test_stream_gemm_chain_of_weights_and_repeat():
    """several ops and iterations in one launch: the W producer runs across op boundaries; results stay those of the last op"""
    from bitdance_b200 import ops
    torch.manual_seed(3)
    M, N, K = 128, 3072, 768
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    ws = [(torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16) for _ in range(3)]
    packed = [ops.stream_pack_weight(w, None) for w in ws]
    big = torch.cat([p.data for p in packed])
    n = packed[0].data.numel()
    views = [ops.StreamWeight(big[i * n:(i + 1) * n], packed[0].bias, N, K, 1, packed[0].n_ctas, 0) for i in range(3)]
    out = ops.stream_gemm(a, views, epi="bias", repeat=5).float()
    ref = _bf(_ref_linear(a, ws[-1], None))
    assert (out - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,n_slices,repeat", [
    (128, 148 * 128 * 2 + 160, 1024, 3, 2),   # two passes per CTA (+ a ragged tail), 3 pieces per pass
    (128, 4096, 512 + 64, 2, 3),              # odd number of k-blocks: the last slice is short
    (64, 2048, 256, 1, 1),                    # one piece = the whole pass
    (128, 1024, 2048, 16, 2),                 # one ring step per piece
])
def test_stream_filler_pieces(M, N, K, n_slices, repeat):
    """A GEMM executed as FILLER pieces (k-slices of each pass accumulated in the third TMEM buffer, outside the grid
    barrier) between the ops of a dependent chain gives the result of the whole-op GEMM; the chain is undisturbed."""
    from bitdance_b200 import ops
    torch.manual_seed(4)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w1 = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    w2 = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    b2 = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16)
    p1, p2 = ops.stream_pack_weight(w1, None), ops.stream_pack_weight(w2, b2)
    o1, o2 = ops.stream_gemm_filler(a, p1, p2, n_slices=n_slices, repeat=repeat)
    r1, r2 = _bf(_ref_linear(a, w1, None)), _bf(_ref_linear(a, w2, b2))
    assert (o1.float() - r1).abs().max().item() <= 1.5e-2 * r1.abs().max().item() + 1e-3
    assert (o2.float() - r2).abs().max().item() <= 1.5e-2 * r2.abs().max().item() + 1e-3
    # and exactly the whole-op result when the slices keep the K order (no rotation inside one step)
    whole = ops.stream_gemm(a, p2, epi="bias")
    assert (o2.float() - whole.float()).abs().max().item() <= 1.5e-2 * r2.abs().max().item() + 1e-3


def test_head_filler_plan_cpu():
    """The filler plan of the persistent sampler (csrc/bd_head.cu::plan_head_pieces): the pieces tile every (pass, k-block)
    of the adaLN GEMM exactly once, in order, never span two passes, start on ring-step boundaries, and follow the slots'
    wanted sizes — for the 14B head (4 passes x 80 k-blocks, 18 slots) and degenerate shapes."""
    import ctypes as C
    import __graft_entry__ as ge
    ge.build()
    from bitdance_b200 import _lib
    lib = _lib.load()
    cases = [(4, 80, [22] + [22, 8, 22] * 5 + [22, 8]), (1, 4, [22] + [22, 8, 22] * 3 + [22, 8]), (3, 5, [4, 4]),
             (2, 12, [1] * 30), (7, 80, [22, 8] * 4), (1, 1, [5]), (4, 80, [0, 0, 10])]
    for P, KB, want in cases:
        out = (C.c_int * (4 * 256))()
        n = lib.bd_head_plan_pieces(P, KB, (C.c_int * len(want))(*want), len(want), out, 256)
        assert n > 0, (P, KB, want)
        pcs = [tuple(out[4 * i + j] for j in range(4)) for i in range(n)]
        pos = 0
        for slot, p, kb0, kbn in pcs:
            assert 0 <= slot < len(want) and kbn > 0
            assert p * KB + kb0 == pos and kb0 + kbn <= KB, (P, KB, pcs)
            assert kb0 % 2 == 0
            pos += kbn
        assert pos == P * KB
        assert [s for s, *_ in pcs] == sorted(s for s, *_ in pcs)
    # the path shape: sizes follow the wanted proportions within one ring step (+ the pass-end rule)
    want = [22] + [22, 8, 22] * 5 + [22, 8]
    out = (C.c_int * (4 * 256))()
    n = lib.bd_head_plan_pieces(4, 80, (C.c_int * len(want))(*want), len(want), out, 256)
    per_slot = [0] * len(want)
    for i in range(n):
        per_slot[out[4 * i]] += out[4 * i + 3]
    scale = 320 / sum(want)
    assert all(abs(g - w * scale) <= 4 for g, w in zip(per_slot, want)), per_slot


def test_stream_host_policy_cpu():
    """host-side policy functions of the engine (no GPU): packed size and the small-N k-split rule"""
    import ctypes as C
    import __graft_entry__ as ge
    ge.build()
    from bitdance_b200 import _lib
    lib = _lib.load()
    lib.bd_stream_packed_elems.restype = C.c_size_t
    assert lib.bd_stream_packed_elems(5120, 5120) == 5120 * 5120
    assert lib.bd_stream_packed_elems(5120, 32) == 5120 * 64          # K padded to one 64-wide k-block
    assert lib.bd_stream_packed_elems(5121, 64) == 0                  # N must be a multiple of 16
    assert lib.bd_stream_ksplit(5120, 5120, 148) == 4                 # wo: 320 units < 4 * 148 CTAs -> 4 k-ranges
    assert lib.bd_stream_ksplit(5120, 7680, 148) == 4                 # w2
    assert lib.bd_stream_ksplit(15360, 5120, 148) == 1                # wqkv / w1: enough units
    assert lib.bd_stream_ksplit(256, 256, 148) == 1                   # tiny models: too few k-blocks to split
    assert lib.bd_stream_set_tuning(5, 2, 0) == 0 and lib.bd_stream_set_tuning(6, 2, 0) != 0   # 7 ring slots in total


def test_stream_partition_covers_every_weight_byte_once_cpu():
    """The work split the kernel, the packer and the host share (csrc/bd_stream.cuh::stream_partition): for every GEMM
    shape on the path — and the ImageNet / tiny-model shapes — the (k-split, 16-row unit) space is dealt to the CTAs
    exactly once, passes are <= 128 rows (a share of 9 or 10 units stays one pass of 144 / 160 rows), and the per-pass slot offsets tile the packed weight without gaps or overlap."""
    import ctypes as C
    import __graft_entry__ as ge
    ge.build()
    from bitdance_b200 import _lib
    lib = _lib.load()
    shapes = [(15360, 5120), (5120, 5120), (5120, 7680), (71680, 5120), (5120, 32), (5120, 256), (7168, 5120),
              (34816, 5120), (5120, 17408), (3584, 256), (768, 256), (256, 384), (160, 64), (1536, 256), (2304, 768),
              (4096, 1024), (3840, 1280), (16, 64), (2368, 64)]
    for G in (148, 132, 8):
        for N, K in shapes:
            KB = (K + 63) // 64
            for S in (1, 2, 4):
                if KB % S or G < S:
                    continue
                U = N // 16
                seen = [[0] * U for _ in range(S)]
                spans = []
                out = (C.c_longlong * 2048)()
                for c in range(G):
                    assert lib.bd_stream_partition_info(N, K, S, G, c, out, 2048) == 0
                    split, unit0, units, kb0, kbs, npass = (int(out[i]) for i in range(6))
                    assert kbs == KB // S and (units == 0 or kb0 == split * kbs)
                    assert npass == (0 if units == 0 else 1 if units <= 10 else (units + 7) // 8)
                    covered = 0
                    for i in range(npass):
                        u0, width, off = int(out[6 + 3 * i]), int(out[7 + 3 * i]), int(out[8 + 3 * i])
                        assert 16 <= width <= (160 if npass == 1 else 128) and width % 16 == 0 and u0 == covered
                        assert off == (split * U + unit0 + u0) * kbs
                        spans.append((off, off + (width // 16) * kbs))
                        covered += width // 16
                    assert covered == units
                    for u in range(unit0, unit0 + units):
                        seen[split][u] += 1
                assert all(v == 1 for row in seen for v in row), (N, K, S, G)
                spans.sort()
                assert spans[0][0] == 0 and spans[-1][1] == S * U * (KB // S)
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])), (N, K, S, G)
