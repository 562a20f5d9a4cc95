"""The image-generation engine: prefill -> AR loop (diffusion head -> sign -> projector -> Qwen3 block) -> decode.

Mirrors ``BitDanceT2IPipeline.gen_image`` (modeling/t2i_pipeline.py:158-272) step for step, on the B200-native ops:

  reference, per AR step                          here
  ------------------------------------------      ----------------------------------------------------------------
  h_fused = cat[h_c, h_u] + pos[block]  :244-245  fused into the final RMSNorm of the previous LLM pass (out_add)
  vision_head.sample(...)               :246      ONE C call: bd_head_sample (51 network evaluations + SDE updates)
  curr_tokens = sign(pred)              :248      bd_sign_tokens_ex (+ packed bits, + bf16 copy for both CFG groups)
  embed_vision_mlp(curr_tokens)         :249      2 tcgen05 GEMMs; gelu fused in fc1's epilogue
  model_input = embeds + pos[block]     :253      fused into fc2's epilogue (fp32 residual table, row % pn)
  model(cond), model(uncond)            :261,266  ONE bd_llm_forward over both groups (weights streamed once)
  (last iteration's LLM pass is discarded by the reference and skipped here)
"""
from __future__ import annotations

import ctypes as C
import time

import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr
from .ae import AERunner
from .head import HeadRunner
from .llm import LlmRunner


def sincos_1d(dim: int, max_len: int, device) -> torch.Tensor:
    """_get_1d_sincos_pos_embed (t2i_pipeline.py:85-96): [max_len, dim] = sin(pos*w) || cos(pos*w), w_k = 1e4^(-k/(dim/2))."""
    omega = torch.arange(dim // 2, dtype=torch.float32, device=device)
    omega /= dim / 2.0
    omega = 1.0 / 10000 ** omega
    pos = torch.arange(max_len, dtype=torch.float32, device=device)
    out = torch.einsum("m,d->md", pos, omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1)


def pos_embed_2d(pos_1d: torch.Tensor, h: int, w: int, ps: int) -> torch.Tensor:
    """get_2d_embed (t2i_pipeline.py:98-107): cat[grid_h (column index), grid_v (row index)], patch-raster order."""
    d2 = pos_1d.shape[1]
    grid_v = pos_1d[:h].view(h, 1, d2).expand(h, w, d2)
    grid_h = pos_1d[:w].view(1, w, d2).expand(h, w, d2)
    pe = torch.cat([grid_h, grid_v], dim=-1)  # h w c
    D = pe.shape[-1]
    pe = pe.reshape(h // ps, ps, w // ps, ps, D).permute(0, 2, 1, 3, 4).reshape(h * w, D)
    return pe.contiguous()


class T2IEngine:
    """Owns the three runners + projector weights; ``gen_image`` is the hot path."""

    def __init__(self, llm: LlmRunner, head: HeadRunner, ae: AERunner | None, fc1_w, fc1_b, fc2_w, fc2_b, *,
                 parallel_num: int, vae_patch_size: int, device="cuda", pe_max_len: int = 4096):
        self.llm, self.head, self.ae = llm, head, ae
        self.device = torch.device(device)
        bf = lambda t: t.detach().to(self.device, torch.bfloat16).contiguous()
        self.fc1_w, self.fc1_b = ops.pack_weight(bf(fc1_w)), bf(fc1_b)
        self.fc2_w, self.fc2_b = ops.pack_weight(bf(fc2_w)), bf(fc2_b)
        self.pn = parallel_num
        self.ps = int(parallel_num ** 0.5)
        self.vae_patch_size = vae_patch_size
        self.D = llm.cfg["hidden_size"]
        self.zc = self.fc1_w.K
        self.pos_1d = sincos_1d(self.D // 2, pe_max_len // vae_patch_size, self.device)
        self._pos_cache = {}
        self.timings = {}
        self.use_graph = False
        self._graph_state = {}
        self._kv = {}

    def _get_cache(self, R, max_len):
        """KV pools are reused across calls (sized for the largest request so far); only the lengths are reset."""
        c = self._kv.get(R)
        if c is None or c.max_tokens < max_len:
            # captured graphs are bound to their cache (and pin its multi-GB pool): drop them with it
            self._graph_state = {k: v for k, v in self._graph_state.items() if v.get("R") != R}
            c = None
            c = self.llm.new_cache(R, max_len)
            self._kv[R] = c
        c.seq_lens.zero_()
        c.host_lens = [0] * R
        return c

    def weight_tensors(self):
        """Every prepacked weight tensor the engine streams (decoder, head, projector, tokenizer) — the set a start-up
        ``parallel.broadcast_tensors`` ships from rank 0 instead of N disk reads."""
        # floating-point payload only: llm._keep also holds the per-layer DEVICE-POINTER table of the one-launch AR block
        # (int64 addresses of this process's allocations) — shipping rank 0's addresses would corrupt the receivers
        ts = [t for t in self.llm._keep if t is not None and t.is_floating_point()]
        ts += [t for t in self.head._keep if t is not None and t.is_floating_point()]
        ts += [self.fc1_w.data, self.fc1_b, self.fc2_w.data, self.fc2_b]
        if self.ae is not None:
            for c in self.ae.convs.values():
                ts += [c.w] + ([c.b] if c.b is not None else [])
            for a, b in list(self.ae.norms.values()) + list(self.ae.ada.values()):
                ts += [a, b]
        return ts

    def pos_embed(self, h, w):
        if (h, w) not in self._pos_cache:
            self._pos_cache[(h, w)] = pos_embed_2d(self.pos_1d, h, w, self.ps)
        return self._pos_cache[(h, w)]

    @torch.no_grad()
    def gen_tokens(self, cond_emb, uncond_emb, img_start_emb, *, h: int, w: int, num_images: int, guidance_scale: float,
                   num_sampling_steps: int, num_steps: int | None = None, timers: bool = False,
                   use_graph: bool | None = None):
        """cond_emb [Lc, D], uncond_emb [Lu, D] | None, img_start_emb [pn + 2, D] (bf16 embeddings).
        Returns tokens fp32 [B, h*w, zc] (patch-raster order) and packed bits int32 [B, h*w, zc/32]."""
        lib = _lib.load()
        dev, pn, D, B = self.device, self.pn, self.D, num_images
        use_cfg = guidance_scale > 1.0
        G = 2 if use_cfg else 1
        R = G * B
        total_steps = (h * w) // pn
        steps = total_steps if num_steps is None else min(num_steps, total_steps)
        pos = self.pos_embed(h, w)  # [h*w, D] fp32
        prompts = [cond_emb] + ([uncond_emb] if use_cfg else [])
        max_len = max(p.shape[0] for p in prompts) + 2 + pn + h * w
        cache = self._get_cache(R, max_len)
        t0 = time.perf_counter()
        # ---- prefill: causal over [prompt, <vision_start>, <res_h>], then the first block with an all-ones mask ----
        h_fused = torch.empty((R, pn, D), dtype=torch.float32, device=dev)
        for gi, pe in enumerate(prompts):
            emb = torch.cat([pe, img_start_emb], dim=0).to(torch.bfloat16)
            emb = emb.unsqueeze(0).expand(B, -1, -1).contiguous()
            self.llm.forward(emb[:, :-pn].contiguous(), cache, gi * B, B, causal=True)
            h_fused[gi * B:(gi + 1) * B] = self.llm.forward(emb[:, -pn:].contiguous(), cache, gi * B, B, causal=False,
                                                            out_add=pos[:pn], out_add_mod=pn)
        if timers:
            torch.cuda.synchronize()
            self.timings["prefill_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
        # ---- AR loop ----
        tokens = torch.zeros((B, h * w, self.zc), dtype=torch.float32, device=dev)   # rows beyond a truncated loop: 0
        packed = torch.zeros((B, h * w, self.zc // 32), dtype=torch.int32, device=dev)
        tok_bf = torch.empty((R * pn, self.zc), dtype=torch.bfloat16, device=dev)
        e1 = torch.empty((R * pn, D), dtype=torch.bfloat16, device=dev)
        splits = self.llm.plan_splits(R, pn, max_len)
        if use_graph is None:
            use_graph = self.use_graph
        steps_done = steps
        if use_graph and steps > 2:
            self._ar_loop_graph(cache, h_fused, pos, tokens, packed, B=B, G=G, hw=h * w, steps=steps,
                                total_steps=total_steps, guidance_scale=guidance_scale,
                                num_sampling_steps=num_sampling_steps, splits=splits)
            steps = 0
        for step in range(steps):
            x = self.head.sample(h_fused, guidance_scale, num_sampling_steps)  # [B, pn, zc] fp32
            check(lib.bd_sign_tokens_ex(ptr(x), B, pn, self.zc, ptr(tokens), C.c_longlong(h * w),
                                        C.c_longlong(step * pn), ptr(tok_bf), G, ptr(packed), stream_ptr()),
                  "bd_sign_tokens_ex")
            if step == total_steps - 1:
                break  # the reference runs (and discards) one more LLM pass here (:261-268)
            ops.gemm(tok_bf, self.fc1_w, bias=self.fc1_b, act="gelu_tanh", out=e1, pdl=True)
            hidden = ops.gemm(e1, self.fc2_w, bias=self.fc2_b, res=pos[step * pn:(step + 1) * pn], res_row_mod=pn,
                              out_dtype=torch.float32, pdl=True).view(R, pn, D)
            nxt = pos[(step + 1) * pn:(step + 2) * pn]
            h_fused = self.llm.forward(hidden, cache, 0, R, causal=False, out_add=nxt, out_add_mod=pn,
                                       attn_splits=splits)
        if timers:
            torch.cuda.synchronize()
            self.timings["ar_s"] = time.perf_counter() - t0
            self.timings["ar_steps"] = steps_done
        return tokens, packed

    # ---- CUDA-graph AR loop ------------------------------------------------------------------------------------------
    def _ar_step_body(self, st, cache, guidance_scale, num_sampling_steps, splits, with_llm=True):
        """One AR step on STATIC buffers (graph-capturable): reads st.h_fused / st.pos_cur / st.pos_next, writes
        st.tok_stage / st.packed_stage and the next st.h_fused."""
        lib = _lib.load()
        pn, D = self.pn, self.D
        x = self.head.sample(st["h_fused"], guidance_scale, num_sampling_steps)
        check(lib.bd_sign_tokens_ex(ptr(x), st["B"], pn, self.zc, ptr(st["tok_stage"]), C.c_longlong(pn), C.c_longlong(0),
                                    ptr(st["tok_bf"]), st["G"], ptr(st["packed_stage"]), stream_ptr()),
              "bd_sign_tokens_ex")
        if not with_llm:
            return
        ops.gemm(st["tok_bf"], self.fc1_w, bias=self.fc1_b, act="gelu_tanh", out=st["e1"], pdl=True)
        ops.gemm(st["e1"], self.fc2_w, bias=self.fc2_b, res=st["pos_cur"], res_row_mod=pn, out=st["hidden"], pdl=True)
        R = st["G"] * st["B"]
        out = self.llm.forward(st["hidden"].view(R, pn, D), cache, 0, R, causal=False, out_add=st["pos_next"],
                               out_add_mod=pn, attn_splits=splits, sk_bound=cache.max_tokens, track_host=False)
        st["h_fused"].copy_(out)

    def _scratch_ptrs(self):
        """Addresses of every grow-only scratch buffer a captured AR step bakes in (they live outside the graph pool)."""
        llm_ws = self.llm._ws.data_ptr() if self.llm._ws is not None else 0
        head_ws = tuple(sorted((k, v.data_ptr()) for k, v in self.head._ws.items()))
        gemm_ws = ops.default_workspace(self.device).buf
        return (llm_ws, head_ws, gemm_ws.data_ptr() if gemm_ws is not None else 0)

    def _ar_loop_graph(self, cache, h_fused, pos, tokens, packed, *, B, G, hw, steps, total_steps, guidance_scale,
                       num_sampling_steps, splits):
        """Capture ONE AR step (head sampler + sign + projector + LLM block: ~3 500 kernel launches with their
        programmatic-dependent-launch edges) into a CUDA graph and replay it per step; per-step inputs (pos-embed blocks)
        are staged into fixed buffers, per-step outputs (tokens) copied out. Sequence lengths live on the device."""
        dev, pn, D = self.device, self.pn, self.D
        R = G * B
        key = (B, G, guidance_scale, num_sampling_steps, splits, cache.max_tokens)
        st = self._graph_state.get(key)
        if st is None:
            st = dict(B=B, G=G, R=R,
                      h_fused=torch.empty((R, pn, D), dtype=torch.float32, device=dev),
                      pos_cur=torch.empty((pn, D), dtype=torch.float32, device=dev),
                      pos_next=torch.empty((pn, D), dtype=torch.float32, device=dev),
                      tok_stage=torch.empty((B, pn, self.zc), dtype=torch.float32, device=dev),
                      packed_stage=torch.empty((B, pn, self.zc // 32), dtype=torch.int32, device=dev),
                      tok_bf=torch.empty((R * pn, self.zc), dtype=torch.bfloat16, device=dev),
                      e1=torch.empty((R * pn, D), dtype=torch.bfloat16, device=dev),
                      hidden=torch.empty((R * pn, D), dtype=torch.float32, device=dev), graph=None, cache_id=None)
            self._graph_state[key] = st
        st["h_fused"].copy_(h_fused)

        def stage(step):
            st["pos_cur"].copy_(pos[step * pn:(step + 1) * pn])
            if step + 1 < total_steps:
                st["pos_next"].copy_(pos[(step + 1) * pn:(step + 2) * pn])

        def collect(step):
            tokens[:, step * pn:(step + 1) * pn].copy_(st["tok_stage"])
            packed[:, step * pn:(step + 1) * pn].copy_(st["packed_stage"])

        first = 0
        if st["graph"] is not None and st.get("ws_ptrs") != self._scratch_ptrs():
            # a grow-only scratch buffer (LLM / head / GEMM workspace) was re-allocated since the capture — e.g. by a
            # prefill with a longer prompt: the graph holds the freed address. Capture again.
            st["graph"] = None
        if st["graph"] is None or st["cache_id"] != id(cache.pool):
            # the graph bakes in the KV pool / page-table / seq_lens addresses of this cache: keep the cache with the graph
            stage(0)
            self._ar_step_body(st, cache, guidance_scale, num_sampling_steps, splits)   # eager step 0 = warm-up
            collect(0)
            first = 1
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            stage(1)
            with torch.cuda.graph(g):
                self._ar_step_body(st, cache, guidance_scale, num_sampling_steps, splits)
            # capture does not execute: seq_lens was not bumped, h_fused not advanced -> replay step 1 for real
            st["graph"], st["cache_id"], st["cache"] = g, id(cache.pool), cache
            st["ws_ptrs"] = self._scratch_ptrs()
        else:
            # re-use the captured graph: it is bound to its own cache object -> copy the prefilled state into it
            gc = st["cache"]
            gc.pool.copy_(cache.pool)
            gc.seq_lens.copy_(cache.seq_lens)
            gc.host_lens = list(cache.host_lens)
        for step in range(first, steps):
            last = step == total_steps - 1
            stage(step)
            if last:
                self._ar_step_body(st, st["cache"], guidance_scale, num_sampling_steps, splits, with_llm=False)
            else:
                st["graph"].replay()
            collect(step)


    @torch.no_grad()
    def decode(self, tokens, h, w):
        return self.ae.decode_tokens(tokens, h, w, self.ps)
