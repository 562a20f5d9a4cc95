"""Python wrappers over the C ABI: one function per ``bd_*`` entry point (device pointers in, status out)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmEpilogue, check, ptr, require_cuda, stream_ptr

ACT = {None: 0, "none": 0, "silu": 1, "gelu_tanh": 2, "gelu_pytorch_tanh": 2}


class Workspace:
    """Grow-only device scratch buffer (owned by torch) handed to C calls that need one."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes: int) -> torch.Tensor | None:
        if nbytes <= 0:
            return None
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=self.device)
        return self.buf


_workspaces: dict = {}


def default_workspace(device) -> Workspace:
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _workspaces:
        _workspaces[key] = Workspace(torch.device("cuda", key))
    return _workspaces[key]


def gemm(
    a: torch.Tensor,
    w: torch.Tensor,
    *,
    bias: torch.Tensor | None = None,
    act: str | None = None,
    swiglu: bool = False,
    gate: torch.Tensor | None = None,
    res: torch.Tensor | None = None,
    res_row_mod: int = 0,
    out: torch.Tensor | None = None,
    out_dtype: torch.dtype = torch.bfloat16,
    bn: int = 0,
    splits: int = 0,
    pdl: bool = False,
    workspace: Workspace | None = None,
) -> torch.Tensor:
    """``epilogue(a @ w.T)`` — a [M,K] bf16, w [N,K] bf16 (nn.Linear layout); see ``bd_gemm_bf16``."""
    lib = _lib.load()
    tiled = isinstance(w, PackedWeight)
    require_cuda(a, w.data if tiled else w, bias, gate, res, out)
    assert a.dtype == torch.bfloat16 and a.dim() == 2 and a.stride(1) == 1
    M, K = a.shape
    if tiled:
        assert w.K == K
        N, ldw = w.N, 0
    else:
        assert w.dtype == torch.bfloat16 and w.dim() == 2 and a.shape[1] == w.shape[1] and w.stride(1) == 1
        N, ldw = w.shape[0], w.stride(0)
    n_out = N // 2 if swiglu else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    assert out.dtype in (torch.bfloat16, torch.float32)
    epi = GemmEpilogue()
    epi.bias = bias.data_ptr() if bias is not None else 0
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.numel() == N and bias.is_contiguous()
    epi.gate = gate.data_ptr() if gate is not None else 0
    epi.ld_gate = gate.stride(0) if gate is not None else 0
    if gate is not None:
        assert gate.dtype == torch.bfloat16 and gate.shape == (M, N) and gate.stride(1) == 1
    epi.res = res.data_ptr() if res is not None else 0
    epi.ld_res = res.stride(0) if res is not None else 0
    epi.res_f32 = 0
    if res is not None:
        assert res.shape == ((res_row_mod or M), N) and res.stride(1) == 1
        assert res.dtype in (torch.bfloat16, torch.float32)
        epi.res_f32 = 1 if res.dtype == torch.float32 else 0
    epi.out = out.data_ptr()
    epi.ld_out = out.stride(0)
    epi.act = ACT[act]
    epi.swiglu = 1 if swiglu else 0
    epi.out_f32 = 1 if out.dtype == torch.float32 else 0
    epi.res_row_mod = res_row_mod
    ws_bytes = lib.bd_gemm_workspace_bytes(M, N, K, bn, splits)
    ws = (workspace or default_workspace(a.device)).get(ws_bytes)
    st = lib.bd_gemm_bf16(
        ptr(a), C.c_int64(a.stride(0)), C.c_void_p(w.data_ptr()), C.c_int64(ldw), M, N, K, C.byref(epi), ptr(ws),
        C.c_size_t(ws.numel() if ws is not None else 0), bn, splits, (1 if pdl else 0) | (2 if tiled else 0), stream_ptr(),
    )
    check(st, "bd_gemm_bf16")
    return out


class PackedWeight:
    """A Linear weight in the tile-major HBM layout of ``bd_pack_weight_tiles`` (+ its logical [N, K])."""

    __slots__ = ("data", "N", "K")

    def __init__(self, data: torch.Tensor, N: int, K: int):
        self.data, self.N, self.K = data, N, K

    def data_ptr(self):
        return self.data.data_ptr()

    @property
    def device(self):
        return self.data.device


def pack_weight(w: torch.Tensor) -> PackedWeight:
    """[N, K] bf16 row-major -> tile-major (one-time prepack; see ``bd_pack_weight_tiles``)."""
    lib = _lib.load()
    require_cuda(w)
    assert w.dtype == torch.bfloat16 and w.dim() == 2 and w.stride(1) == 1
    N, K = w.shape
    lib.bd_packed_weight_elems.restype = C.c_size_t
    n = lib.bd_packed_weight_elems(N, K)
    out = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    check(lib.bd_pack_weight_tiles(ptr(w), C.c_int64(w.stride(0)), N, K, ptr(out), stream_ptr()), "bd_pack_weight_tiles")
    return PackedWeight(out, N, K)


def interleave16(gate_w: torch.Tensor, up_w: torch.Tensor, gate_b=None, up_b=None):
    """One-time SwiGLU weight re-layout (see ``bd_interleave16``). Returns (w_interleaved, bias_interleaved|None)."""
    lib = _lib.load()
    require_cuda(gate_w, up_w)
    assert gate_w.shape == up_w.shape and gate_w.dtype == torch.bfloat16 and up_w.dtype == torch.bfloat16
    gate_w, up_w = gate_w.contiguous(), up_w.contiguous()
    F, K = gate_w.shape
    out = torch.empty((2 * F, K), dtype=torch.bfloat16, device=gate_w.device)
    bout = None
    if gate_b is not None:
        gate_b, up_b = gate_b.to(torch.bfloat16).contiguous(), up_b.to(torch.bfloat16).contiguous()
        bout = torch.empty((2 * F,), dtype=torch.bfloat16, device=gate_w.device)
    check(lib.bd_interleave16(ptr(gate_w), ptr(up_w), ptr(out), F, K, ptr(gate_b), ptr(up_b), ptr(bout), stream_ptr()),
          "bd_interleave16")
    return out, bout


def sign_pack_nchw(h: torch.Tensor, *, want_quant=True, want_packed=True, num_codebooks: int = 0):
    """``VQModel.encode`` quantiser (+ packed bits, + GFQ indices). h: [B,C,H,W] fp32/bf16 contiguous."""
    lib = _lib.load()
    require_cuda(h)
    assert h.dim() == 4 and h.is_contiguous() and h.dtype in (torch.float32, torch.bfloat16)
    B, Cc, H, W = h.shape
    HW = H * W
    quant = torch.empty_like(h) if want_quant else None
    packed = torch.empty((B, HW, Cc // 32), dtype=torch.int32, device=h.device) if want_packed else None
    idx = torch.empty((num_codebooks, B * HW), dtype=torch.int32, device=h.device) if num_codebooks else None
    check(lib.bd_sign_pack_nchw(ptr(h), 1 if h.dtype == torch.float32 else 0, B, Cc, HW, ptr(quant), ptr(packed),
                                ptr(idx), num_codebooks, stream_ptr()), "bd_sign_pack_nchw")
    return quant, packed, idx


def sign_tokens(x: torch.Tensor, *, want_tokens=True, want_packed=True):
    """``torch.sign`` on the AR path + packed bits. x: [..., C] fp32 contiguous."""
    lib = _lib.load()
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    tokens = torch.empty_like(x) if want_tokens else None
    packed = torch.empty((*x.shape[:-1], Cc // 32), dtype=torch.int32, device=x.device) if want_packed else None
    check(lib.bd_sign_tokens(ptr(x), C.c_longlong(rows), Cc, ptr(tokens), ptr(packed), stream_ptr()), "bd_sign_tokens")
    return tokens, packed


def unpack_tokens(packed: torch.Tensor, C_bits: int, dtype=torch.float32):
    lib = _lib.load()
    require_cuda(packed)
    assert packed.dtype == torch.int32 and packed.is_contiguous() and packed.shape[-1] == C_bits // 32
    rows = packed.numel() // (C_bits // 32)
    out = torch.empty((*packed.shape[:-1], C_bits), dtype=dtype, device=packed.device)
    check(lib.bd_unpack_tokens(ptr(packed), C.c_longlong(rows), C_bits, ptr(out), 1 if dtype == torch.float32 else 0,
                               stream_ptr()), "bd_unpack_tokens")
    return out


def attention(q, k, v, *, causal=False, scale=None, page_table=None, sk=None, splits=0, out=None, pdl=False,
              workspace: Workspace | None = None):
    """Flash-style attention (see ``bd_attention_bf16``).

    strided: q [B,Sq,Hq,D], k/v [B,Sk,Hkv,D] (any strides with contiguous last dim);
    paged:   k/v pools [n_pages, Hkv, 64, D] contiguous, page_table int32 [B, max_pages], ``sk`` = valid keys.
    Returns [B,Sq,Hq,D] bf16."""
    lib = _lib.load()
    require_cuda(q, k, v)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16 and q.stride(-1) == 1
    B, Sq, Hq, D = q.shape
    if page_table is None:
        assert k.stride(-1) == 1 and v.stride() == k.stride()
        Sk, Hkv = k.shape[1], k.shape[2]
        ksb, kss, ksh = k.stride(0), k.stride(1), k.stride(2)
        pt, max_pages = None, 0
    else:
        assert k.is_contiguous() and v.is_contiguous() and k.shape[2] == 64 and page_table.dtype == torch.int32
        Sk, Hkv = int(sk), k.shape[1]
        ksb = kss = ksh = 0
        pt, max_pages = page_table.contiguous(), page_table.shape[1]
    if out is None:
        out = torch.empty((B, Sq, Hq, D), dtype=torch.bfloat16, device=q.device)
    if scale is None:
        scale = D ** -0.5
    lib.bd_attention_workspace_bytes.restype = C.c_size_t
    need = lib.bd_attention_workspace_bytes(B, Hq, Sq, Sk, D, splits)
    ws = (workspace or default_workspace(q.device)).get(need)
    st = lib.bd_attention_bf16(
        ptr(q), C.c_int64(q.stride(0)), C.c_int64(q.stride(1)), C.c_int64(q.stride(2)), ptr(k), ptr(v),
        C.c_int64(ksb), C.c_int64(kss), C.c_int64(ksh), ptr(pt), max_pages, ptr(out), C.c_int64(out.stride(0)),
        C.c_int64(out.stride(1)), C.c_int64(out.stride(2)), B, Sq, Sk, Hq, Hkv, D, 1 if causal else 0,
        C.c_float(scale), splits, ptr(ws), C.c_size_t(ws.numel() if ws is not None else 0), 1 if pdl else 0,
        stream_ptr())
    check(st, "bd_attention_bf16")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# persistent weight-streaming engine (csrc/bd_stream.cuh)
# ---------------------------------------------------------------------------------------------------------------------
class StreamWeight:
    """A Linear weight in the stream-major layout of ``bd_stream_pack_weight`` (+ packed bias, logical shape)."""

    __slots__ = ("data", "bias", "N", "K", "ksplit", "n_ctas", "perm")

    def __init__(self, data, bias, N, K, ksplit, n_ctas, perm):
        self.data, self.bias, self.N, self.K, self.ksplit, self.n_ctas, self.perm = data, bias, N, K, ksplit, n_ctas, perm

    def data_ptr(self):
        return self.data.data_ptr()


def stream_num_ctas() -> int:
    return int(_lib.load().bd_stream_num_ctas())


def stream_ksplit(N: int, K: int, n_ctas: int | None = None) -> int:
    return int(_lib.load().bd_stream_ksplit(N, K, n_ctas or stream_num_ctas()))


def stream_pack_weight(w: torch.Tensor, bias: torch.Tensor | None = None, *, ksplit: int = 1, swiglu: bool = False,
                       n_ctas: int | None = None) -> StreamWeight:
    """[N, K] bf16 row-major -> stream-major (one-time prepack). swiglu: rows [0, N/2) = x1, [N/2, N) = x2 of a SwiGLU
    Linear, packed as 8+8 row units so that one accumulator chunk holds matching gate/up columns."""
    lib = _lib.load()
    require_cuda(w, bias)
    assert w.dtype == torch.bfloat16 and w.dim() == 2 and w.stride(1) == 1
    N, K = w.shape
    n_ctas = n_ctas or stream_num_ctas()
    lib.bd_stream_packed_elems.restype = C.c_size_t
    n = lib.bd_stream_packed_elems(N, K)
    assert n > 0, "N must be a multiple of 16"
    out = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    bout = torch.empty(N, dtype=torch.bfloat16, device=w.device)
    if bias is not None:
        bias = bias.to(torch.bfloat16).contiguous()
    check(lib.bd_stream_pack_weight(ptr(w), C.c_int64(w.stride(0)), N, K, ksplit, n_ctas, 1 if swiglu else 0,
                                    N // 2 if swiglu else 0, ptr(bias), ptr(out), ptr(bout), stream_ptr()),
          "bd_stream_pack_weight")
    return StreamWeight(out, bout, N, K, ksplit, n_ctas, 1 if swiglu else 0)


def _blocked_index(rows: int, K: int, device):
    """element offsets (in bf16 elements) of (row, col) inside a blocked activation [ceil(K/64)][128][64], 128B swizzle"""
    r = torch.arange(rows, device=device).view(-1, 1)
    c = torch.arange(K, device=device).view(1, -1)
    return (c // 64) * 8192 + r * 64 + ((((c // 8) % 8) ^ (r % 8)) * 8) + (c % 8)


def to_blocked(a: torch.Tensor) -> torch.Tensor:
    """[M <= 128, K] bf16 -> the blocked, swizzled image the stream engine reads as a GEMM operand (zero padded)."""
    M, K = a.shape
    assert M <= 128 and a.dtype == torch.bfloat16
    out = torch.zeros(((K + 63) // 64) * 8192, dtype=torch.bfloat16, device=a.device)
    out[_blocked_index(M, K, a.device).reshape(-1)] = a.reshape(-1)
    return out


def from_blocked(b: torch.Tensor, M: int, K: int) -> torch.Tensor:
    return b[_blocked_index(M, K, b.device).reshape(-1)].view(M, K)


def stream_gemm(a: torch.Tensor, ws: "StreamWeight | list[StreamWeight]", *, epi: str = "bias", act: str | None = None,
                out_blocked: bool = False, repeat: int = 1, a_is_blocked: bool = False, M: int | None = None):
    """One (or several identical-shape) GEMM op(s) through the persistent kernel. epi: 'bias' | 'swiglu' | 'partial'.
    Returns the output of the LAST weight: [M, N] bf16 ('bias'), [M, N/2] bf16 ('swiglu') or fp32 [ksplit, M, N]."""
    lib = _lib.load()
    wl = ws if isinstance(ws, (list, tuple)) else [ws]
    w0 = wl[0]
    if len(wl) > 1:
        stride = wl[1].data_ptr() - wl[0].data_ptr()
        assert all(wl[i].data_ptr() - wl[0].data_ptr() == i * stride for i in range(len(wl)))
    else:
        stride = 0
    if not a_is_blocked:
        M = a.shape[0]
        ab = to_blocked(a)
    else:
        ab = a
    N, K, dev = w0.N, w0.K, w0.data.device
    kind = {"bias": 0, "swiglu": 1, "partial": 2}[epi]
    if kind == 2:
        out = torch.zeros((w0.ksplit, M, N), dtype=torch.float32, device=dev)
        ld = 0
    else:
        n_out = N // 2 if kind == 1 else N
        if out_blocked:
            out = torch.zeros(((n_out + 63) // 64) * 8192, dtype=torch.bfloat16, device=dev)
            ld = 0
        else:
            out = torch.zeros((M, n_out), dtype=torch.bfloat16, device=dev)
            ld = n_out
    sync = torch.zeros(16, dtype=torch.int32, device=dev)
    check(lib.bd_stream_gemm(ptr(ab), ptr(w0.data), C.c_int64(stride), len(wl), ptr(w0.bias) if kind != 2 else None,
                             ptr(out), C.c_int64(ld), M, N, K, w0.ksplit, kind, ACT[act], 1 if out_blocked else 0,
                             w0.n_ctas, repeat, ptr(sync), stream_ptr()), "bd_stream_gemm")
    if kind != 2 and out_blocked:
        return from_blocked(out, M, n_out)
    return out


def stream_gemm_filler(a: torch.Tensor, w_main: StreamWeight, w_fill: StreamWeight, *, n_slices: int, repeat: int = 1):
    """tests: ``repeat`` x {main GEMM; the second GEMM as filler pieces (n_slices k-ranges per pass, third TMEM buffer,
    outside the grid-barrier protocol)} in one persistent launch. Returns (out_main, out_fill) [M, N] bf16."""
    lib = _lib.load()
    M = a.shape[0]
    ab = to_blocked(a)
    N, K, dev = w_main.N, w_main.K, w_main.data.device
    assert (w_fill.N, w_fill.K) == (N, K) and w_main.ksplit == 1 and w_fill.ksplit == 1
    out_main = torch.zeros((M, N), dtype=torch.bfloat16, device=dev)
    out_fill = torch.zeros((M, N), dtype=torch.bfloat16, device=dev)
    sync = torch.zeros(16, dtype=torch.int32, device=dev)
    check(lib.bd_stream_gemm_filler(ptr(ab), ptr(w_main.data), ptr(w_fill.data), ptr(w_fill.bias), ptr(out_main),
                                    ptr(out_fill), M, N, K, w_main.n_ctas, n_slices, repeat, ptr(sync), stream_ptr()),
          "bd_stream_gemm_filler")
    return out_main, out_fill


def head_set_fillers(mode: int, row_kb: int = 0, gemm_kb: int = 0) -> None:
    """Program policy of the persistent head sampler (``bd_head_set_fillers``)."""
    check(_lib.load().bd_head_set_fillers(int(mode), int(row_kb), int(gemm_kb)), "bd_head_set_fillers")
