// bd_llm.cu — Qwen3 decoder stack over a paged KV cache: one C call per forward (prefill chunk or AR block).
//
// Replaces the third-party arithmetic the reference calls at modeling/t2i_pipeline.py:199,211,224,229,261,266
// (transformers Qwen3Model.forward: RMSNorm -> q/k/v -> q/k RMSNorm(head_dim) -> RoPE -> cache update -> SDPA ->
// o_proj -> +res -> RMSNorm -> SwiGLU -> +res, final RMSNorm). B200-first differences from the reference's execution:
//   * cond and uncond sequences run in ONE pass (the reference makes two model() calls per AR step, streaming the
//     26 GB of weights twice); sequences may have different past lengths (device-side seq_lens);
//   * q/k/v and gate/up weights are fused ([q|k|v] rows; interleave16(gate, up));
//   * KV lives in 64-token pages (= one parallel block of the 64x model) instead of a torch.cat-grown DynamicCache;
//     K is stored as bf16(RoPE(k)) — identical to what SDPA consumes under autocast (oracle/llm.py docstring).
// Rounding policy: oracle/llm.py (stream_f32 = AR steps, else prefill).
#include "bd_host.h"
#include "bd_ptx.cuh"
#include "bd_rowops.cuh"
#include "bd_stream.cuh"

namespace bd {

struct AttnParams;
int attn_run_llm(const __nv_bfloat16* q, const __nv_bfloat16* kpool, const __nv_bfloat16* vpool, const int* page_table,
                 int max_pages, const int* sk_dev, int sk_bound, __nv_bfloat16* out, int R, int S, int Hq, int Hkv,
                 int head_dim, int causal, int splits, void* ws, size_t ws_bytes, bool pdl, cudaStream_t stream,
                 int out_blocked = 0);
size_t attn_llm_workspace_bytes(int R, int S, int Hq, int head_dim, int splits);

__device__ __forceinline__ float block_sum256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// Qwen3RMSNorm. x: [M, D] fp32 (IN_F32) or bf16. weight bf16 [D].
//   fp32 stream: y = float(w) * (x * rstd)                      (fp32; the consumer Linear rounds it to bf16)
//   bf16 stream: y = bf16(w * bf16(x * rstd))
// out: bf16 [M, D] (GEMM operand) or, for the final norm, the stream dtype (OUT_F32) with an optional fp32 row table
// added (add[m % add_mod, :]): h_fused = last_hidden_state + pos_embed (modeling/t2i_pipeline.py:245).
template <bool IN_F32, bool OUT_F32>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const void* __restrict__ x_, const __nv_bfloat16* __restrict__ w,
                                                      void* __restrict__ out_, int D, float eps,
                                                      const float* __restrict__ add, int add_mod) {
  __shared__ float red[32];
  grid_dep_launch();
  grid_dep_wait();
  const long long m = blockIdx.x;
  float ss = 0.f;
  for (int d = threadIdx.x * 4; d < D; d += 256 * 4) {
    float v[4];
    if (IN_F32) {
      const float4 t = *reinterpret_cast<const float4*>(static_cast<const float*>(x_) + m * D + d);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x_) + m * D + d);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&t);
      const float2 a = __bfloat1622float2(p[0]), b = __bfloat1622float2(p[1]);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    ss += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  }
  const float rstd = rsqrtf(block_sum256(ss, red) / static_cast<float>(D) + eps);
  for (int d = threadIdx.x * 4; d < D; d += 256 * 4) {
    float v[4];
    if (IN_F32) {
      const float4 t = *reinterpret_cast<const float4*>(static_cast<const float*>(x_) + m * D + d);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x_) + m * D + d);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&t);
      const float2 a = __bfloat1622float2(p[0]), b = __bfloat1622float2(p[1]);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    float y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float wf = __bfloat162float(w[d + j]);
      if (IN_F32)
        y[j] = wf * (v[j] * rstd);
      else
        y[j] = bf16_round(wf * bf16_round(v[j] * rstd));
      if (add) y[j] += add[static_cast<long long>(m % add_mod) * D + d + j];
    }
    if (OUT_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(out_) + m * D + d) = make_float4(y[0], y[1], y[2], y[3]);
    } else {
      __nv_bfloat162 a = __floats2bfloat162_rn(y[0], y[1]), b = __floats2bfloat162_rn(y[2], y[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out_) + m * D + d) = pk;
    }
  }
}

// q/k RMSNorm over head_dim + RoPE + KV-cache append. One warp per (token, head); head_dim = 32 * VPT.
//   qkv: bf16 [M, (Hq + 2 Hkv) * hd] rows = [q heads | k heads | v heads]
//   q_out: bf16 [M, Hq * hd];  K/V pools: [page][Hkv][64][hd];  position of token (b, s) = seq_lens[b] + s
//   rope tables: fp32 [max_pos, hd] (cos | sin of cat(freqs, freqs), built by torch exactly like Qwen3RotaryEmbedding)
//   ROPE_F32 (AR steps): rot = q*cos + rotate_half(q)*sin in fp32 (unfused), then bf16 (the SDPA autocast cast)
//   else (prefill):      cos/sin rounded to bf16, every product and the sum round to bf16
//   ROPE_PAIRS (ImageNet class-conditional model, imagenet_gen/src/layers_parallel.py:255-290): no q/k RMSNorm (qn_w / kn_w
//   NULL), interleaved (even, odd) pairs rotated by the angle of table entry [pos][pair] (rope_cos / rope_sin are
//   [max_pos, hd/2] here: the 2-D RoPE table of precompute_freqs_cis_2d, half of the pairs turn with x, half with y), fp32
//   arithmetic, one rounding to bf16 (apply_rotary_emb computes in fp32 and casts back, :273-290)
template <int HD, bool ROPE_F32, bool ROPE_PAIRS = false>
__global__ void __launch_bounds__(256) qk_norm_rope_append_kernel(
    const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ qn_w, const __nv_bfloat16* __restrict__ kn_w,
    const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, const int* __restrict__ seq_lens,
    const int* __restrict__ page_table, int max_pages, __nv_bfloat16* __restrict__ q_out,
    __nv_bfloat16* __restrict__ kpool, __nv_bfloat16* __restrict__ vpool, int S, int Hq, int Hkv, float eps, int M) {
  grid_dep_launch();
  grid_dep_wait();
  constexpr int VPT = HD / 32;  // consecutive elements per lane
  const int heads = Hq + 2 * Hkv;
  const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (gw >= static_cast<long long>(M) * heads) return;
  const int lane = threadIdx.x & 31;
  const int m = static_cast<int>(gw / heads), hh = static_cast<int>(gw % heads);
  const int b = m / S, s = m % S;
  const int pos = seq_lens[b] + s;
  if (pos < 0 || pos >= max_pages * 64) {
    // the device-side sequence length ran past the cache (the CUDA-graph path has no host-side length check; the RoPE tables
    // cover max_pages * 64 positions, llm.py::new_cache): fail loudly instead of writing K/V outside the pool
    if (lane == 0 && hh == 0) printf("bd_llm: sequence %d position %d outside the KV cache (%d tokens)\n", b, pos, max_pages * 64);
    __trap();
  }
  const __nv_bfloat16* src = qkv + static_cast<long long>(m) * heads * HD + static_cast<long long>(hh) * HD + lane * VPT;
  float x[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) x[j] = __bfloat162float(src[j]);
  const bool is_q = hh < Hq, is_k = !is_q && hh < Hq + Hkv;
  __nv_bfloat16* dst;
  if (is_q) {
    dst = q_out + static_cast<long long>(m) * Hq * HD + static_cast<long long>(hh) * HD;
  } else {
    const int hk = is_k ? hh - Hq : hh - Hq - Hkv;
    const int page = page_table[b * max_pages + pos / 64];
    dst = (is_k ? kpool : vpool) + ((static_cast<long long>(page) * Hkv + hk) * 64 + (pos % 64)) * HD;
  }
  if (!is_q && !is_k) {  // V: plain copy
#pragma unroll
    for (int j = 0; j < VPT; ++j) dst[lane * VPT + j] = __float2bfloat16_rn(x[j]);
    return;
  }
  if (ROPE_PAIRS) {
    // lane holds VPT consecutive elements = VPT / 2 whole (even, odd) pairs: the rotation is lane-local
    float y[VPT];
#pragma unroll
    for (int j = 0; j < VPT; j += 2) {
      const int pr = (lane * VPT + j) >> 1;
      const float c = rope_cos[static_cast<long long>(pos) * (HD / 2) + pr], sn = rope_sin[static_cast<long long>(pos) * (HD / 2) + pr];
      y[j] = __fsub_rn(__fmul_rn(x[j], c), __fmul_rn(x[j + 1], sn));
      y[j + 1] = __fadd_rn(__fmul_rn(x[j + 1], c), __fmul_rn(x[j], sn));
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) dst[lane * VPT + j] = __float2bfloat16_rn(y[j]);
    return;
  }
  // RMSNorm over the head (bf16 in -> bf16 out): bf16(w * bf16(x * rstd))
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < VPT; ++j) ss += x[j] * x[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / static_cast<float>(HD) + eps);
  const __nv_bfloat16* nw = is_q ? qn_w : kn_w;
#pragma unroll
  for (int j = 0; j < VPT; ++j)
    x[j] = bf16_round(__bfloat162float(nw[lane * VPT + j]) * bf16_round(x[j] * rstd));
  // rotate_half partner: element d pairs with d +- HD/2, i.e. lane +- 16
  float y[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const float other = __shfl_xor_sync(0xffffffffu, x[j], 16);
    const float rot = (lane < 16) ? -other : other;  // cat(-x2, x1)
    const int d = lane * VPT + j;
    const float c = rope_cos[static_cast<long long>(pos) * HD + d], sn = rope_sin[static_cast<long long>(pos) * HD + d];
    if (ROPE_F32)
      y[j] = __fadd_rn(__fmul_rn(x[j], c), __fmul_rn(rot, sn));
    else
      y[j] = bf16_round(bf16_round(x[j] * bf16_round(c)) + bf16_round(rot * bf16_round(sn)));
  }
#pragma unroll
  for (int j = 0; j < VPT; ++j) dst[lane * VPT + j] = __float2bfloat16_rn(y[j]);
}

__global__ void bump_seq_lens_kernel(int* seq_lens, int R, int S) {
  grid_dep_launch();
  grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) seq_lens[i] += S;
}

template <typename... KArgs, typename... Args>
static int launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  LaunchCfg lc(grid, block, smem, st, pdl);
  BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, kern, static_cast<KArgs>(args)...));
  return BD_OK;
}

static size_t al(size_t v) { return (v + 255) & ~size_t(255); }
constexpr int kLlmMaxSplits = 16;

struct LlmWs {
  size_t a, qkv, q, o, g, attn, gemm, total;
  size_t s_a, s_o, s_g, s_part, s_sync;  // persistent path: blocked operands, fp32 k-split partials, barrier counter
};

static LlmWs llm_ws_layout(const bd_llm_weights_t& w, int M, int R, int S, int attn_splits) {
  LlmWs L{};
  size_t off = 0;
  auto take = [&](size_t b) {
    size_t o = off;
    off = al(off + b);
    return o;
  };
  const size_t qkv_n = static_cast<size_t>(w.Hq + 2 * w.Hkv) * w.head_dim;
  L.a = take(static_cast<size_t>(M) * w.D * 2);
  L.qkv = take(M * qkv_n * 2);
  L.q = take(static_cast<size_t>(M) * w.Hq * w.head_dim * 2);
  L.o = take(static_cast<size_t>(M) * w.Hq * w.head_dim * 2);
  L.g = take(static_cast<size_t>(M) * w.I * 2);
  {
    size_t ab = attn_llm_workspace_bytes(R, S, w.Hq, w.head_dim, attn_splits);
    if (w.layer_tab && w.stream_ctas > 0 && M <= 128) {  // in-engine attention: partials for up to kLlmMaxSplits key ranges
      const size_t eb = static_cast<size_t>(kLlmMaxSplits) * R * w.Hq * S * (w.head_dim + 2) * sizeof(float);
      ab = eb > ab ? eb : ab;
    }
    L.attn = take(ab);
  }
  size_t gm = 0;
  auto gw = [&](int n, int k) {
    size_t b = gemm_workspace_bytes(M, n, k, 0, 0);
    if (b > gm) gm = b;
  };
  gw(static_cast<int>(qkv_n), w.D);
  gw(w.D, w.Hq * w.head_dim);
  gw(2 * w.I, w.D);
  gw(w.D, w.I);
  L.gemm = take(gm);
  if (w.stream_ctas > 0 && M <= 128) {
    auto blk = [](size_t K) { return ((K + 63) / 64) * static_cast<size_t>(kSlotBytes); };
    L.s_a = take(blk(w.D));
    L.s_o = take(blk(static_cast<size_t>(w.Hq) * w.head_dim));
    L.s_g = take(blk(w.I));
    L.s_part = take(static_cast<size_t>(4) * 128 * w.D * 4);
    L.s_sync = take(64);
  }
  L.total = off;
  return L;
}

// One persistent launch: o_proj (+res) -> RMSNorm -> gate_up (SwiGLU) -> down (+res) -> next RMSNorm -> next layer's qkv
// (or, after the last layer, the final norm). `first`: only the input RMSNorm + layer 0's qkv.
static int llm_stream_segment(const bd_llm_weights_t& w, int li, bool first, void* hidden, int M, uint8_t* base,
                              const LlmWs& L, __nv_bfloat16* qkv, int qkv_n, void* out, const float* out_add,
                              int out_add_mod, cudaStream_t st) {
  static thread_local StreamProgram prog;
  prog = StreamProgram{};
  prog.family = kStreamFamLlm;
  prog.M = M;
  prog.n_ctas = w.stream_ctas;
  prog.n_iter = 1;
  prog.cfg_mult = 1;
  prog.sync = reinterpret_cast<unsigned int*>(base + L.s_sync);
  const int D = w.D, G = w.stream_ctas, Ko = w.Hq * w.head_dim;
  int n = 0;
  auto gemm_op = [&](const void* W, const void* A, int N, int K, int ksplit, int epi, void* o, long long ld, bool blocked) {
    StreamOp& op = prog.ops[n++];
    op = StreamOp{};
    op.kind = kOpGemm;
    op.sub = epi;
    op.N = N;
    op.K = K;
    op.ksplit = ksplit;
    op.flags = blocked ? kFlagBlocked : 0;
    op.wait_prev = 1;
    op.p0 = W;
    op.p1 = A;
    op.o0 = o;
    op.l0 = ld;
  };
  auto row = [&](int sub) -> StreamOp& {
    StreamOp& op = prog.ops[n++];
    op = StreamOp{};
    op.kind = kOpRow;
    op.sub = sub;
    op.wait_prev = 1;
    op.N = D;
    op.f0 = w.eps;
    op.o1 = hidden;
    return op;
  };
  if (first) {
    StreamOp& op = row(kRowLlmRms);
    op.wait_prev = 0;
    op.p1 = w.layers[0].ln1_w;
    op.o0 = base + L.s_a;
    gemm_op(w.layers[0].wqkv_s, base + L.s_a, qkv_n, D, 1, kEpiBias, qkv, qkv_n, false);
  } else {
    const bd_llm_layer_t& lw = w.layers[li];
    const bool last = li + 1 == w.n_layers;
    const int ks_o = stream_ksplit_for(D, Ko, G), ks_d = stream_ksplit_for(D, w.I, G);
    gemm_op(lw.wo_s, base + L.s_o, D, Ko, ks_o, kEpiPartial, base + L.s_part, 0, false);
    prog.ops[n - 1].wait_prev = 0;  // its operand comes from the attention kernel launched before this one
    {
      StreamOp& op = row(kRowLlmResRms);
      op.p0 = base + L.s_part;
      op.i0 = ks_o;
      op.p1 = lw.ln2_w;
      op.o0 = base + L.s_a;
    }
    gemm_op(lw.w_gate_up_s, base + L.s_a, 2 * w.I, D, 1, kEpiSwiglu8, base + L.s_g, 0, true);
    gemm_op(lw.w_down_s, base + L.s_g, D, w.I, ks_d, kEpiPartial, base + L.s_part, 0, false);
    {
      StreamOp& op = row(kRowLlmResRms);
      op.p0 = base + L.s_part;
      op.i0 = ks_d;
      if (last) {
        op.p1 = w.final_norm_w;
        op.i1 = 1;
        op.o0 = out;
        op.p3 = out_add;
        op.i2 = out_add_mod > 0 ? out_add_mod : 1;
      } else {
        op.p1 = w.layers[li + 1].ln1_w;
        op.o0 = base + L.s_a;
      }
    }
    if (!last) gemm_op(w.layers[li + 1].wqkv_s, base + L.s_a, qkv_n, D, 1, kEpiBias, qkv, qkv_n, false);
  }
  prog.n_pre = 0;
  prog.n_body = n;
  prog.n_post = 0;
  return stream_launch(prog, st);
}

// The whole AR block as ONE persistent launch (bd_stream.cuh): program iterations = decoder layers, per-layer weight
// pointers from the device table w.layer_tab.
//   pre : RMSNorm(ln1[0]) -> blocked a;  qkv GEMM of layer 0
//   body (layer l): q/k norm + RoPE + KV append | paged attention (split-KV partials) | combine -> blocked o |
//         o_proj (k-split partials) | residual + RMSNorm(ln2[l]) | gate/up GEMM (SwiGLU epilogue) | down (k-split partials) |
//         residual + RMSNorm(ln1[l+1]) | qkv GEMM of layer l+1          (the last two skipped in the last layer)
//   post: residual + final RMSNorm (+ pos-embed rows) -> out fp32
static int llm_stream_all(const bd_llm_weights_t& w, void* hidden, int R, int S, const int* seq_lens, int sk_bound,
                          void* kv_pool, int64_t kv_layer_stride, int64_t kv_v_offset, const int32_t* page_table,
                          int max_pages, const float* rope_cos, const float* rope_sin, void* out, const float* out_add,
                          int out_add_mod, uint8_t* base, const LlmWs& L, cudaStream_t st) {
  const int M = R * S, D = w.D, G = w.stream_ctas, hd = w.head_dim, Ko = w.Hq * hd;
  const int qkv_n = (w.Hq + 2 * w.Hkv) * hd;
  BD_REQUIRE(w.n_layers <= kStreamMaxIter && S <= 64 && R * w.Hq <= 4096);
  static thread_local StreamProgram prog;
  prog = StreamProgram{};
  prog.family = kStreamFamLlm;
  prog.M = M;
  prog.n_ctas = G;
  prog.n_iter = w.n_layers;
  prog.cfg_mult = 1;
  prog.sync = reinterpret_cast<unsigned int*>(base + L.s_sync);
  const void* const* tab = static_cast<const void* const*>(w.layer_tab);
  enum { T_WQKV = 1, T_WO, T_WGU, T_WDOWN, T_LN1, T_LN2, T_QN, T_KN };  // slot + 1
  // attention: the (sequence, q head, 32-key half tile) items are dealt evenly over the first att_ctas CTAs whatever the
  // lengths are (bd_stream.cu: llm_attn_stream); a (sequence, head) block is touched by at most att_ctas / Hq + 2 CTAs,
  // one partial slot each (+1 slack; the kernel traps beyond). The bound on att_ctas keeps the slots within the
  // kLlmMaxSplits the workspace is sized for.
  const int att_ctas = std::min(G, (kLlmMaxSplits - 3) * w.Hq);
  const int splits = att_ctas / w.Hq + 3;
  // the kernel's work-list arithmetic is 32-bit: (items) x (CTAs) must fit
  BD_REQUIRE(static_cast<long long>(R) * w.Hq * ((sk_bound + 63) / 64) * att_ctas < (1ll << 31));
  float* part_o = reinterpret_cast<float*>(base + L.attn);
  float* part_ml = part_o + static_cast<size_t>(splits) * R * w.Hq * S * hd;
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(base + L.qkv);
  __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(base + L.q);
  const int ks_o = stream_ksplit_for(D, Ko, G), ks_d = stream_ksplit_for(D, w.I, G);
  int n = 0;
  auto new_op = [&](int kind) -> StreamOp& {
    StreamOp& op = prog.ops[n++];
    op = StreamOp{};
    op.kind = kind;
    op.wait_prev = 1;
    op.tab = tab;
    return op;
  };
  auto gemm_op = [&](int tslot, int toff, const void* A, int N, int K, int ksplit, int epi, void* o, long long ld,
                     bool blocked) -> StreamOp& {
    StreamOp& op = new_op(kOpGemm);
    op.sub = epi;
    op.N = N;
    op.K = K;
    op.ksplit = ksplit;
    op.flags = blocked ? kFlagBlocked : 0;
    op.tab_p0 = tslot;
    op.tab_off = toff;
    op.p1 = A;
    op.o0 = o;
    op.l0 = ld;
    return op;
  };
  auto row = [&](int sub) -> StreamOp& {
    StreamOp& op = new_op(kOpRow);
    op.sub = sub;
    op.N = D;
    op.f0 = w.eps;
    op.o1 = hidden;
    return op;
  };
  // ---- pre ----
  {
    StreamOp& op = row(kRowLlmRms);
    op.wait_prev = 0;
    op.tab_p1 = T_LN1;
    op.o0 = base + L.s_a;
  }
  gemm_op(T_WQKV, 0, base + L.s_a, qkv_n, D, 1, kEpiBias, qkv, qkv_n, false);
  prog.n_pre = n;
  // ---- one layer ----
  {
    StreamOp& op = new_op(kOpLlmRope);
    op.p0 = qkv;
    op.tab_p1 = T_QN;
    op.tab_p2 = T_KN;
    op.p3 = rope_cos;
    op.p4 = rope_sin;
    op.p5 = seq_lens;
    op.p6 = page_table;
    op.o0 = q;
    op.o1 = kv_pool;
    op.l0 = kv_layer_stride;
    op.l1 = kv_v_offset;
    op.sub = R;
    op.i0 = S;
    op.i1 = w.Hq;
    op.i2 = w.Hkv;
    op.N = max_pages;
    op.K = hd;
    op.f0 = w.eps;
  }
  {
    StreamOp& op = new_op(kOpLlmAttn);
    op.p0 = q;
    op.p1 = seq_lens;
    op.p2 = page_table;
    op.p3 = kv_pool;
    op.l0 = kv_layer_stride;
    op.l1 = kv_v_offset;
    op.o0 = part_o;
    op.o1 = part_ml;
    op.sub = R;
    op.ksplit = splits;
    op.act = att_ctas;
    op.i0 = S;
    op.i1 = w.Hq;
    op.i2 = w.Hkv;
    op.N = max_pages;
    op.K = hd;
    op.f0 = (1.0f / sqrtf(static_cast<float>(hd))) * 1.4426950408889634f;
  }
  {
    StreamOp& op = new_op(kOpRow);
    op.sub = kRowLlmAttnCombine;
    op.p0 = part_o;
    op.p1 = part_ml;
    op.p2 = seq_lens;
    op.act = att_ctas;
    op.i0 = splits;
    op.i1 = w.Hq;
    op.i2 = S;
    op.N = Ko;
    op.K = hd;
    op.o0 = base + L.s_o;
  }
  gemm_op(T_WO, 0, base + L.s_o, D, Ko, ks_o, kEpiPartial, base + L.s_part, 0, false);
  {
    StreamOp& op = row(kRowLlmResRms);
    op.p0 = base + L.s_part;
    op.i0 = ks_o;
    op.tab_p1 = T_LN2;
    op.o0 = base + L.s_a;
  }
  gemm_op(T_WGU, 0, base + L.s_a, 2 * w.I, D, 1, kEpiSwiglu8, base + L.s_g, 0, true);
  gemm_op(T_WDOWN, 0, base + L.s_g, D, w.I, ks_d, kEpiPartial, base + L.s_part, 0, false);
  {
    StreamOp& op = row(kRowLlmResRms);
    op.p0 = base + L.s_part;
    op.i0 = ks_d;
    op.tab_p1 = T_LN1;
    op.tab_off = 1;
    op.o0 = base + L.s_a;
    op.flags |= kFlagSkipLast;
  }
  gemm_op(T_WQKV, 1, base + L.s_a, qkv_n, D, 1, kEpiBias, qkv, qkv_n, false).flags |= kFlagSkipLast;
  prog.n_body = n - prog.n_pre;
  // ---- post ----
  {
    StreamOp& op = row(kRowLlmResRms);
    op.tab = nullptr;
    op.p0 = base + L.s_part;
    op.i0 = ks_d;
    op.p1 = w.final_norm_w;
    op.i1 = 1;
    op.o0 = out;
    op.p3 = out_add;
    op.i2 = out_add_mod > 0 ? out_add_mod : 1;
  }
  prog.n_post = 1;
  return stream_launch(prog, st);
}

}  // namespace bd

using namespace bd;

extern "C" {

size_t bd_llm_workspace_bytes(const bd_llm_weights_t* w, int R, int S, int attn_splits) {
  if (!w || R <= 0 || S <= 0) return 0;
  return llm_ws_layout(*w, R * S, R, S, attn_splits).total;
}

#define BD_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != BD_OK) return _rc; \
  } while (0)

int bd_llm_forward(const bd_llm_weights_t* wp, void* hidden, int stream_f32, int R, int S, int* seq_lens,
                   int sk_bound, int causal, void* kv_pool, int64_t kv_layer_stride, int64_t kv_v_offset,
                   const int32_t* page_table, int max_pages, const float* rope_cos, const float* rope_sin, void* out,
                   const float* out_add, int out_add_mod, int attn_splits, void* workspace, size_t workspace_bytes,
                   int flags, bd_stream_t stream_) {
  BD_REQUIRE(wp && hidden && seq_lens && kv_pool && page_table && rope_cos && rope_sin && out && workspace);
  const bd_llm_weights_t& w = *wp;
  BD_REQUIRE(R > 0 && S > 0 && w.n_layers > 0 && w.layers);
  BD_REQUIRE(w.head_dim == 128 || w.head_dim == 64);
  BD_REQUIRE((w.D % 64) == 0 && w.D <= 6144 && (w.I % 64) == 0 && (w.Hq % w.Hkv) == 0);
  BD_REQUIRE(attn_splits >= 1 && sk_bound >= S && max_pages * 64 >= sk_bound);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const bool pdl = (flags & 1) != 0;
  const int M = R * S, D = w.D, hd = w.head_dim;
  const int qkv_n = (w.Hq + 2 * w.Hkv) * hd;
  const bool rope_pairs = (w.variant & BD_LLM_ROPE_PAIRS) != 0;
  for (int li = 0; li < w.n_layers; ++li)  // q/k RMSNorm weights are optional only for the pair-RoPE (ImageNet) variant
    BD_REQUIRE(rope_pairs || (w.layers[li].q_norm_w && w.layers[li].k_norm_w));
  BD_REQUIRE(!rope_pairs || stream_f32);  // that model keeps an fp32 residual stream throughout
  const LlmWs L = llm_ws_layout(w, M, R, S, attn_splits);
  if (workspace_bytes < L.total) return BD_ERR_WORKSPACE;
  BD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  auto bfp = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(base + off); };
  __nv_bfloat16 *a = bfp(L.a), *qkv = bfp(L.qkv), *q = bfp(L.q), *o = bfp(L.o), *g = bfp(L.g);
  void* aws = base + L.attn;
  const size_t aws_bytes = L.gemm - L.attn;
  void* gws = base + L.gemm;
  const size_t gws_bytes = L.total - L.gemm;
  auto bf = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };

  auto norm_to_bf16 = [&](const void* nw) {
    if (stream_f32)
      return launch_k(rmsnorm_kernel<true, false>, dim3(M), dim3(256), 0, st, pdl, (const void*)hidden, bf(nw),
                      (void*)a, D, w.eps, (const float*)nullptr, 1);
    return launch_k(rmsnorm_kernel<false, false>, dim3(M), dim3(256), 0, st, pdl, (const void*)hidden, bf(nw), (void*)a,
                    D, w.eps, (const float*)nullptr, 1);
  };

  // residual add of a projection + the RMSNorm that follows it. When K is split, the reduction of the fp32 partials,
  // the residual add and the norm run in ONE row kernel; otherwise the GEMM epilogue adds the residual and the
  // stand-alone norm kernel follows. next_norm == nullptr: no norm (last layer; the final norm has its own kernel).
  auto proj_res_then_norm = [&](const void* A_, int Kdim, const void* W_, const void* next_norm) -> int {
    GemmEpi e;
    e.res = hidden;
    e.ld_res = D;
    e.res_f32 = stream_f32;
    e.out = hidden;
    e.ld_out = D;
    e.out_f32 = stream_f32;
    int S_used = 1;
    BD_TRY(gemm_bf16(A_, Kdim, W_, Kdim, M, D, Kdim, e, gws, gws_bytes, 0, 0, pdl, st, w.w_tiled != 0, &S_used));
    if (S_used > 1) {
      LlmRowArgs ra;
      ra.partial = static_cast<const float*>(gws);
      ra.splits = S_used;
      ra.M = M;
      ra.D = D;
      ra.hidden = hidden;
      ra.stream_f32 = stream_f32;
      ra.norm_w = bf(next_norm);
      ra.a = a;
      ra.eps = w.eps;
      return launch_k(llm_splitk_row_kernel, dim3(M), dim3(kRowThreads), 0, st, pdl, ra);
    }
    return next_norm ? norm_to_bf16(next_norm) : BD_OK;
  };

  if (w.emb_norm_w) {
    // BitDance.forward_model (imagenet_gen/src/model_parallel.py:343-350): x = emb_norm(x) before the first block, in place
    BD_REQUIRE(stream_f32);
    BD_TRY(launch_k(rmsnorm_kernel<true, true>, dim3(M), dim3(256), 0, st, pdl, (const void*)hidden, bf(w.emb_norm_w), hidden,
                    D, w.eps, (const float*)nullptr, 1));
  }

  // ---- AR block of one 128-row tile with stream-packed weights: the Linears / residual adds / norms of every layer run as
  // persistent bd_stream_kernel segments; RoPE + KV append and the paged attention stay the kernels below ----
  if (w.stream_ctas > 0 && stream_f32 && !causal && !rope_pairs && M <= 128 && w.stream_ctas == num_sms() && w.stream_ctas >= M &&
      (qkv_n % 16) == 0 &&
      D <= 6144 && (w.I % 64) == 0) {
    // blocked operands are read in whole 128-row x 64-column tiles: padding rows / columns must be finite
    BD_CUDA_TRY(cudaMemsetAsync(base + L.s_a, 0, L.s_sync - L.s_a, st));
    if (w.layer_tab && S <= 64 && w.n_layers <= kStreamMaxIter) {
      BD_TRY(llm_stream_all(w, hidden, R, S, seq_lens, sk_bound, kv_pool, kv_layer_stride, kv_v_offset, page_table, max_pages,
                            rope_cos, rope_sin, out, out_add, out_add_mod, base, L, st));
      BD_TRY(launch_k(bump_seq_lens_kernel, dim3(1), dim3(256), 0, st, false, seq_lens, R, S));
      return BD_OK;
    }
    BD_TRY(llm_stream_segment(w, 0, true, hidden, M, base, L, qkv, qkv_n, out, out_add, out_add_mod, st));
    for (int li = 0; li < w.n_layers; ++li) {
      const bd_llm_layer_t& lw = w.layers[li];
      __nv_bfloat16* kpool = reinterpret_cast<__nv_bfloat16*>(kv_pool) + static_cast<long long>(li) * kv_layer_stride;
      __nv_bfloat16* vpool = kpool + kv_v_offset;
      const long long warps = static_cast<long long>(M) * (w.Hq + 2 * w.Hkv);
      const unsigned grid = static_cast<unsigned>((warps * 32 + 255) / 256);
      if (hd == 128)
        BD_TRY(launch_k(qk_norm_rope_append_kernel<128, true>, dim3(grid), dim3(256), 0, st, false,
                        (const __nv_bfloat16*)qkv, bf(lw.q_norm_w), bf(lw.k_norm_w), rope_cos, rope_sin,
                        (const int*)seq_lens, (const int*)page_table, max_pages, q, kpool, vpool, S, w.Hq, w.Hkv, w.eps, M));
      else
        BD_TRY(launch_k(qk_norm_rope_append_kernel<64, true>, dim3(grid), dim3(256), 0, st, false,
                        (const __nv_bfloat16*)qkv, bf(lw.q_norm_w), bf(lw.k_norm_w), rope_cos, rope_sin,
                        (const int*)seq_lens, (const int*)page_table, max_pages, q, kpool, vpool, S, w.Hq, w.Hkv, w.eps, M));
      BD_TRY(attn_run_llm(q, kpool, vpool, page_table, max_pages, seq_lens, sk_bound,
                          reinterpret_cast<__nv_bfloat16*>(base + L.s_o), R, S, w.Hq, w.Hkv, hd, causal, attn_splits, aws,
                          aws_bytes, pdl, st, /*out_blocked=*/1));
      BD_TRY(llm_stream_segment(w, li, false, hidden, M, base, L, qkv, qkv_n, out, out_add, out_add_mod, st));
    }
    BD_TRY(launch_k(bump_seq_lens_kernel, dim3(1), dim3(256), 0, st, false, seq_lens, R, S));
    return BD_OK;
  }

  BD_TRY(norm_to_bf16(w.layers[0].ln1_w));
  for (int li = 0; li < w.n_layers; ++li) {
    const bd_llm_layer_t& lw = w.layers[li];
    __nv_bfloat16* kpool = reinterpret_cast<__nv_bfloat16*>(kv_pool) + static_cast<long long>(li) * kv_layer_stride;
    __nv_bfloat16* vpool = kpool + kv_v_offset;
    {
      GemmEpi e;
      e.out = qkv;
      e.ld_out = qkv_n;
      BD_TRY(gemm_bf16(a, D, lw.wqkv, D, M, qkv_n, D, e, gws, gws_bytes, 0, 0, pdl, st, w.w_tiled != 0));
    }
    {
      const long long warps = static_cast<long long>(M) * (w.Hq + 2 * w.Hkv);
      const unsigned grid = static_cast<unsigned>((warps * 32 + 255) / 256);
      int rc;
#define BD_QK_LAUNCH(HD_, F32_, PAIRS_)                                                                                  \
  launch_k(qk_norm_rope_append_kernel<HD_, F32_, PAIRS_>, dim3(grid), dim3(256), 0, st, pdl, (const __nv_bfloat16*)qkv,  \
           bf(lw.q_norm_w), bf(lw.k_norm_w), rope_cos, rope_sin, (const int*)seq_lens, (const int*)page_table, max_pages, \
           q, kpool, vpool, S, w.Hq, w.Hkv, w.eps, M)
      if (rope_pairs) rc = hd == 128 ? BD_QK_LAUNCH(128, true, true) : BD_QK_LAUNCH(64, true, true);
      else if (hd == 128) rc = stream_f32 ? BD_QK_LAUNCH(128, true, false) : BD_QK_LAUNCH(128, false, false);
      else rc = stream_f32 ? BD_QK_LAUNCH(64, true, false) : BD_QK_LAUNCH(64, false, false);
#undef BD_QK_LAUNCH
      BD_TRY(rc);
    }
    // keys visible to block b: seq_lens[b] (past) + S (this block), read on the device by the attention kernel
    BD_TRY(attn_run_llm(q, kpool, vpool, page_table, max_pages, seq_lens, sk_bound, o, R, S, w.Hq, w.Hkv, hd, causal,
                        attn_splits, aws, aws_bytes, pdl, st));
    BD_TRY(proj_res_then_norm(o, w.Hq * hd, lw.wo, lw.ln2_w));
    {
      GemmEpi e;
      e.swiglu = 1;
      e.out = g;
      e.ld_out = w.I;
      BD_TRY(gemm_bf16(a, D, lw.w_gate_up, D, M, 2 * w.I, D, e, gws, gws_bytes, 0, 0, pdl, st, w.w_tiled != 0));
    }
    BD_TRY(proj_res_then_norm(g, w.I, lw.w_down, li + 1 < w.n_layers ? w.layers[li + 1].ln1_w : nullptr));
  }
  if (stream_f32)
    BD_TRY(launch_k(rmsnorm_kernel<true, true>, dim3(M), dim3(256), 0, st, pdl, (const void*)hidden,
                    bf(w.final_norm_w), out, D, w.eps, out_add, out_add_mod > 0 ? out_add_mod : 1));
  else if (out_add)  // bf16 stream, fp32 output: float(bf16 norm) + pos (first h_fused after prefill, :218,245)
    BD_TRY(launch_k(rmsnorm_kernel<false, true>, dim3(M), dim3(256), 0, st, pdl, (const void*)hidden,
                    bf(w.final_norm_w), out, D, w.eps, out_add, out_add_mod > 0 ? out_add_mod : 1));
  else
    BD_TRY(launch_k(rmsnorm_kernel<false, false>, dim3(M), dim3(256), 0, st, pdl, (const void*)hidden,
                    bf(w.final_norm_w), out, D, w.eps, (const float*)nullptr, 1));
  BD_TRY(launch_k(bump_seq_lens_kernel, dim3(1), dim3(256), 0, st, pdl, seq_lens, R, S));
  return BD_OK;
}

}  // extern "C"
