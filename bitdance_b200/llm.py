"""Host side of the Qwen3 decoder: weight prepack, paged KV cache, RoPE tables and the ``bd_llm_forward`` launch.

Stands in for ``Qwen3ForCausalLM(...).model(inputs_embeds=..., past_key_values=..., attention_mask=...)`` as used by
the reference (modeling/t2i_pipeline.py:199-266). Accepts the HF state-dict key names (``model.layers.N.self_attn...``).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib, ops
from ._lib import check, note, ptr, stream_ptr

PAGE = 64


class LlmLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "ln2_w", "q_norm_w", "k_norm_w", "wqkv", "wo", "w_gate_up", "w_down",
                                          "wqkv_s", "wo_s", "w_gate_up_s", "w_down_s")]


class LlmWeights(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("D", "I", "n_layers", "Hq", "Hkv", "head_dim")] + [
        ("eps", C.c_float), ("w_tiled", C.c_int), ("stream_ctas", C.c_int), ("variant", C.c_int),
        ("final_norm_w", C.c_void_p), ("layers", C.POINTER(LlmLayer)), ("emb_norm_w", C.c_void_p),
        ("layer_tab", C.c_void_p)]

ROPE_PAIRS = 1  # BD_LLM_ROPE_PAIRS (include/bitdance_b200.h)


def llm_config_dict(cfg) -> dict:
    """Accepts a transformers Qwen3Config or a plain dict."""
    g = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    hidden, heads = g("hidden_size"), g("num_attention_heads")
    theta = g("rope_theta")
    if theta is None:
        rp = g("rope_parameters") or {}
        theta = rp.get("rope_theta", 1e6)
    return dict(hidden_size=hidden, intermediate_size=g("intermediate_size"), num_hidden_layers=g("num_hidden_layers"),
                num_attention_heads=heads, num_key_value_heads=g("num_key_value_heads"),
                head_dim=g("head_dim") or hidden // heads, rms_norm_eps=g("rms_norm_eps", 1e-6), rope_theta=float(theta))


def llm_spec(cfg: dict, prefix="model.") -> dict:
    """State-dict spec of the HF Qwen3 decoder layers + final norm (embeddings / lm_head excluded)."""
    D, I, hd = cfg["hidden_size"], cfg["intermediate_size"], cfg["head_dim"]
    Hq, Hkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    s = {}
    for i in range(cfg["num_hidden_layers"]):
        p = f"{prefix}layers.{i}."
        s[p + "input_layernorm.weight"] = (D,)
        s[p + "post_attention_layernorm.weight"] = (D,)
        s[p + "self_attn.q_proj.weight"] = (Hq * hd, D)
        s[p + "self_attn.k_proj.weight"] = (Hkv * hd, D)
        s[p + "self_attn.v_proj.weight"] = (Hkv * hd, D)
        s[p + "self_attn.o_proj.weight"] = (D, Hq * hd)
        s[p + "self_attn.q_norm.weight"] = (hd,)
        s[p + "self_attn.k_norm.weight"] = (hd,)
        s[p + "mlp.gate_proj.weight"] = (I, D)
        s[p + "mlp.up_proj.weight"] = (I, D)
        s[p + "mlp.down_proj.weight"] = (D, I)
    s[prefix + "norm.weight"] = (D,)
    return s


@dataclass
class KVCache:
    pool: torch.Tensor        # bf16 [L, 2, n_pages, Hkv, 64, hd]
    page_table: torch.Tensor  # int32 [R, max_pages]
    seq_lens: torch.Tensor    # int32 [R] (device; bumped by the kernels)
    host_lens: list           # python mirror of seq_lens (planning only)
    max_tokens: int


class LlmRunner:
    def __init__(self, state_dict: dict | None, cfg, device="cuda", prefix="model.", synthetic_seed: int | None = None,
                 max_positions: int = 8192, stream: bool | None = None, qk_norm: bool = True,
                 rope_pairs: tuple | None = None, emb_norm: torch.Tensor | None = None):
        """stream: also keep stream-major copies of the four Linears of every layer, so that AR blocks of one 128-row tile
        run their GEMMs / residual adds / RMSNorms as persistent bd_stream_kernel segments (costs a second copy of the
        decoder weights in HBM: 26 GB for Qwen3-14B).
        qk_norm=False / rope_pairs=(cos, sin) fp32 [positions, head_dim/2] / emb_norm=[D]: the ImageNet class-conditional
        decoder variant (no q/k RMSNorm, interleaved-pair 2-D RoPE from a table, RMSNorm on the input embeddings)."""
        self.cfg = c = llm_config_dict(cfg)
        self.device = dev = torch.device(device)
        D, I, hd = c["hidden_size"], c["intermediate_size"], c["head_dim"]
        Hq, Hkv, L = c["num_attention_heads"], c["num_key_value_heads"], c["num_hidden_layers"]
        qkv_n = (Hq + 2 * Hkv) * hd
        if stream is None:  # default: off until opted in (BD_LLM_STREAM=1) — see DESIGN.md section 7
            stream = os.environ.get("BD_LLM_STREAM", "0") == "1"
        fits = D % 64 == 0 and I % 64 == 0 and qkv_n % 16 == 0 and D <= 6144
        if stream and not fits:
            note(f"decoder dims outside the persistent engine's limits (hidden={D}: multiple of 64 and <= 6144; MLP={I}: "
                 f"multiple of 64): AR blocks run as chained kernels")
        stream = bool(stream) and fits
        n_ctas = ops.stream_num_ctas() if stream else 0
        self._keep = []
        layers = (LlmLayer * L)()
        gen = None
        if state_dict is None:
            assert synthetic_seed is not None
            gen = torch.Generator(device=dev)
            gen.manual_seed(synthetic_seed)

        def rand(shape, std=0.02, one=False):
            t = torch.randn(shape, generator=gen, device=dev, dtype=torch.float32)
            t = (1.0 + 0.1 * t) if one else t * std
            return t.to(torch.bfloat16)

        def get(name, shape, one=False):
            if state_dict is None:
                return rand(shape, one=one)
            t = state_dict[prefix + name]
            assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
            return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()

        for i in range(L):
            p = f"layers.{i}."
            lw = layers[i]
            keep = self._keep
            ln1 = get(p + "input_layernorm.weight", (D,), one=True)
            ln2 = get(p + "post_attention_layernorm.weight", (D,), one=True)
            qn = get(p + "self_attn.q_norm.weight", (hd,), one=True) if qk_norm else None
            kn = get(p + "self_attn.k_norm.weight", (hd,), one=True) if qk_norm else None
            wqkv = torch.cat([get(p + "self_attn.q_proj.weight", (Hq * hd, D)),
                              get(p + "self_attn.k_proj.weight", (Hkv * hd, D)),
                              get(p + "self_attn.v_proj.weight", (Hkv * hd, D))], dim=0).contiguous()
            wo = get(p + "self_attn.o_proj.weight", (D, Hq * hd))
            gate, up = get(p + "mlp.gate_proj.weight", (I, D)), get(p + "mlp.up_proj.weight", (I, D))
            wgu, _ = ops.interleave16(gate, up)
            wd = get(p + "mlp.down_proj.weight", (D, I))
            if n_ctas:
                # stream-major copies for the persistent engine (AR blocks of one 128-row tile)
                sq = ops.stream_pack_weight(wqkv, None, n_ctas=n_ctas)
                so = ops.stream_pack_weight(wo, None, ksplit=ops.stream_ksplit(D, Hq * hd, n_ctas), n_ctas=n_ctas)
                sg = ops.stream_pack_weight(torch.cat([gate, up], dim=0), None, swiglu=True, n_ctas=n_ctas)
                sdn = ops.stream_pack_weight(wd, None, ksplit=ops.stream_ksplit(D, I, n_ctas), n_ctas=n_ctas)
                keep += [sq.data, so.data, sg.data, sdn.data]
                lw.wqkv_s, lw.wo_s, lw.w_gate_up_s, lw.w_down_s = (t.data_ptr() for t in (sq, so, sg, sdn))
            del gate, up
            wqkv, wo, wgu, wd = (ops.pack_weight(t).data for t in (wqkv, wo, wgu, wd))  # tile-major for HBM streaming
            keep += [ln1, ln2, qn, kn, wqkv, wo, wgu, wd]
            lw.ln1_w, lw.ln2_w = ln1.data_ptr(), ln2.data_ptr()
            lw.q_norm_w, lw.k_norm_w = (qn.data_ptr(), kn.data_ptr()) if qk_norm else (None, None)
            lw.wqkv, lw.wo, lw.w_gate_up, lw.w_down = wqkv.data_ptr(), wo.data_ptr(), wgu.data_ptr(), wd.data_ptr()
        fn = get("norm.weight", (D,), one=True)
        self._keep.append(fn)
        torch.cuda.synchronize(dev)
        w = LlmWeights()
        w.D, w.I, w.n_layers, w.Hq, w.Hkv, w.head_dim = D, I, L, Hq, Hkv, hd
        w.eps = c["rms_norm_eps"]
        w.w_tiled = 1
        w.stream_ctas = n_ctas
        w.final_norm_w = fn.data_ptr()
        w.variant = ROPE_PAIRS if rope_pairs is not None else 0
        if n_ctas and qk_norm:
            # per-layer pointer table of the one-launch AR block (bd_llm_weights_t.layer_tab): [L + 1][8]
            rows = [[lw.wqkv_s, lw.wo_s, lw.w_gate_up_s, lw.w_down_s, lw.ln1_w, lw.ln2_w, lw.q_norm_w, lw.k_norm_w]
                    for lw in layers] + [[0] * 8]
            tab = torch.tensor([[int(v or 0) for v in r] for r in rows], dtype=torch.int64).to(dev)
            self._keep.append(tab)
            w.layer_tab = tab.data_ptr()
        if emb_norm is not None:
            en = emb_norm.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
            self._keep.append(en)
            w.emb_norm_w = en.data_ptr()
        self._layers = layers
        w.layers = C.cast(layers, C.POINTER(LlmLayer))
        self.w = w
        self._ws = None
        if rope_pairs is not None:
            cos, sin = rope_pairs
            assert cos.shape == sin.shape and cos.shape[1] == hd // 2
            self.rope_cos = cos.detach().to(device=dev, dtype=torch.float32).contiguous()
            self.rope_sin = sin.detach().to(device=dev, dtype=torch.float32).contiguous()
        else:
            self._build_rope(max_positions)

    def param_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._keep if t is not None)

    def _build_rope(self, n):
        # Qwen3RotaryEmbedding: inv_freq = 1 / theta^(arange(0,d,2)/d); emb = cat(freqs, freqs); fp32 cos/sin
        hd, theta = self.cfg["head_dim"], self.cfg["rope_theta"]
        inv_freq = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).to(device=self.device, dtype=torch.float) / hd))
        pos = torch.arange(n, device=self.device).float()
        freqs = pos[:, None] * inv_freq[None, :]
        emb = torch.cat([freqs, freqs], dim=-1)
        self.rope_cos = emb.cos().contiguous()
        self.rope_sin = emb.sin().contiguous()

    def new_cache(self, R: int, max_tokens: int) -> KVCache:
        c = self.cfg
        max_pages = (max_tokens + PAGE - 1) // PAGE
        assert max_pages * PAGE <= self.rope_cos.shape[0], "raise max_positions"
        pool = torch.zeros((c["num_hidden_layers"], 2, R * max_pages, c["num_key_value_heads"], PAGE, c["head_dim"]),
                           dtype=torch.bfloat16, device=self.device)
        pt = torch.arange(R * max_pages, dtype=torch.int32, device=self.device).view(R, max_pages).contiguous()
        return KVCache(pool, pt, torch.zeros(R, dtype=torch.int32, device=self.device), [0] * R, max_pages * PAGE)

    def plan_splits(self, R, S, max_tokens):
        """Fixed split-KV factor for a cache of up to max_tokens (fixed so that a captured CUDA graph stays valid)."""
        ctas = ((S + 63) // 64) * R * self.cfg["num_attention_heads"]
        tiles = (max_tokens + 63) // 64
        s = max(1, min((2 * 148) // max(ctas, 1), tiles // 4, 16))
        return s

    def forward(self, hidden: torch.Tensor, cache: KVCache, r0: int, R: int, *, causal: bool, out_add=None,
                out_add_mod: int = 0, attn_splits: int | None = None, pdl: bool = True, sk_bound: int | None = None,
                track_host: bool = True) -> torch.Tensor:
        """hidden: [R, S, D] fp32 (AR stream) or bf16 (prefill stream) for sequences r0..r0+R; OVERWRITTEN.
        Returns the final-norm output [R, S, D] in the stream dtype (+ out_add rows when given)."""
        lib = _lib.load()
        assert hidden.is_cuda and hidden.is_contiguous() and hidden.dim() == 3 and hidden.shape[0] == R
        S, D = hidden.shape[1], hidden.shape[2]
        stream_f32 = hidden.dtype == torch.float32
        assert stream_f32 or hidden.dtype == torch.bfloat16
        if sk_bound is None:
            sk_bound = max(cache.host_lens[r0:r0 + R]) + S
        assert sk_bound <= cache.max_tokens, "KV cache capacity exceeded"
        if attn_splits is None:
            attn_splits = self.plan_splits(R, S, sk_bound)
        lib.bd_llm_workspace_bytes.restype = C.c_size_t
        need = lib.bd_llm_workspace_bytes(C.byref(self.w), R, S, attn_splits)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(hidden.shape, dtype=torch.float32 if (stream_f32 or out_add is not None) else torch.bfloat16,
                          device=self.device)
        pool = cache.pool
        if out_add is not None:
            assert out_add.dtype == torch.float32 and out_add.is_contiguous() and out_add.shape[-1] == D
        st = lib.bd_llm_forward(
            C.byref(self.w), ptr(hidden), 1 if stream_f32 else 0, R, S, ptr(cache.seq_lens[r0:r0 + R]), sk_bound,
            1 if causal else 0, ptr(pool), C.c_int64(pool.stride(0)), C.c_int64(pool.stride(1)),
            ptr(cache.page_table[r0:r0 + R]), cache.page_table.shape[1], ptr(self.rope_cos), ptr(self.rope_sin),
            ptr(out), ptr(out_add), out_add_mod, attn_splits, ptr(self._ws), C.c_size_t(self._ws.numel()),
            1 if pdl else 0, stream_ptr())
        check(st, "bd_llm_forward")
        if track_host:
            for r in range(r0, r0 + R):
                cache.host_lens[r] += S
        return out


class LmHead:
    """``Qwen3ForCausalLM.lm_head`` (``nn.Linear(hidden, vocab, bias=False)``; reference call site modeling/mllm.py:845-846)
    on the tcgen05 weight-streaming GEMM: the [vocab, hidden] matrix is prepacked tile-major once (vocab padded to a
    multiple of 128 rows with zeros), a decode step is ONE launch that streams it once (1.56 GB for Qwen3-14B).
    Output is rounded to bf16 — what a bf16 ``nn.Linear`` returns."""

    def __init__(self, weight: torch.Tensor, device="cuda"):
        w = weight.detach().to(device=device, dtype=torch.bfloat16).contiguous()
        assert w.dim() == 2
        self.vocab, self.hidden = w.shape
        pad = (-self.vocab) % 128
        if pad:
            w = torch.cat([w, torch.zeros((pad, self.hidden), dtype=w.dtype, device=w.device)], dim=0)
        self.w = ops.pack_weight(w)

    @torch.no_grad()
    def __call__(self, hidden: torch.Tensor) -> torch.Tensor:
        """hidden [..., hidden] -> logits [..., vocab] bf16"""
        shp = hidden.shape
        a = hidden.reshape(-1, shp[-1]).to(torch.bfloat16).contiguous()
        out = ops.gemm(a, self.w)
        return out[:, :self.vocab].reshape(*shp[:-1], self.vocab)
