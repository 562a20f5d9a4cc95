// bd_head.cu — the binary-diffusion vision head: x-prediction transformer + Euler–Maruyama sampler, one C call.
//
// Replaces, per AR step, DiffHead.sample -> euler_maruyama (modeling/vision_head/sampling_x.py:44-97) driving
// TransEncoder.forward (modeling/vision_head/flow_head_parallel_x.py:325-342) S+1 times. Every Linear goes through the
// tcgen05 weight-streaming GEMM (bd_gemm.cuh) with the surrounding elementwise work fused in its epilogue; what is
// left are the small HBM/L2-resident kernels in this file. Exact restructurings (do not change results):
//   * cond_embed(c) is evaluated once per sample() call instead of S+1 times (c is constant across evaluations);
//   * time_embed(t) is evaluated for all S+1 timesteps in one batched GEMM (t is the same for every row);
//   * the 2 adaLN Linears and final_layer.ada_ln_modulation share their input y: one GEMM over concatenated weights;
//   * the rows of cat([x, x]) (cond | uncond halves, sampling_x.py:71) are written once and duplicated.
// Noise is an input (drawn by torch in the reference's call order), so the sampler itself is deterministic.
#include <cstdlib>
#include "bd_host.h"
#include "bd_ptx.cuh"
#include "bd_rowops.cuh"
#include "bd_stream.cuh"

namespace bd {

int attn_run_head(const __nv_bfloat16* qkv, __nv_bfloat16* out, int R, int pn, int D, int head_dim, bool pdl,
                  cudaStream_t stream);  // bd_attn.cu

// ---------------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
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

// timestep_embedding (flow_head_parallel_x.py:12-27): [cos(1000 t f_k) | sin(1000 t f_k)], f_k = exp(-ln(1e4) k/half),
// fp32 math, written as bf16 (it only feeds a Linear under autocast).
struct TVals {
  float t[128];
};
__global__ void timestep_embedding_kernel(TVals tv_, int row0, int dim, __nv_bfloat16* __restrict__ out) {
  grid_dep_launch();
  grid_dep_wait();
  const int i = row0 + blockIdx.x;
  const int half = dim / 2;
  const float tv = 1000.0f * tv_.t[blockIdx.x];
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float f = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
    const float a = tv * f;
    out[static_cast<long long>(i) * dim + k] = __float2bfloat16_rn(cosf(a));
    out[static_cast<long long>(i) * dim + half + k] = __float2bfloat16_rn(sinf(a));
  }
}

// fp32 -> bf16 cast of a [rows, cols] matrix (the autocast input cast of cond_embed / MLPconnector).
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  grid_dep_launch();
  grid_dep_wait();
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(out + i) = pk;
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

// y[m, :] = bf16(silu(bf16(t_emb[:] + c_emb[m, :])))      (TransEncoder.forward :330: y = F.silu(t + c))
__global__ void silu_add_kernel(const __nv_bfloat16* __restrict__ t_emb, const __nv_bfloat16* __restrict__ c_emb,
                                __nv_bfloat16* __restrict__ y, int M, int D) {
  grid_dep_launch();
  grid_dep_wait();
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i >= static_cast<long long>(M) * D) return;
  const int d = static_cast<int>(i % D);
  const uint4 tv = *reinterpret_cast<const uint4*>(t_emb + d);
  const uint4 cv = *reinterpret_cast<const uint4*>(c_emb + i);
  const __nv_bfloat162* t2 = reinterpret_cast<const __nv_bfloat162*>(&tv);
  const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&cv);
  uint4 ov;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = __bfloat1622float2(t2[j]), b = __bfloat1622float2(c2[j]);
    const float s0 = bf16_round(a.x + b.x), s1 = bf16_round(a.y + b.y);
    o2[j] = __floats2bfloat162_rn(siluf_(s0), siluf_(s1));
  }
  *reinterpret_cast<uint4*>(y + i) = ov;
}

// out[m,:] = bf16( LN(x[m,:]) (*w + b) * bf16(1 + scale[m,:]) + shift[m,:] )     (TransBlock.forward :243,246;
// FinalLayer.forward :171). LN statistics in fp32 (autocast runs layer_norm in fp32), eps 1e-6.
// One CTA per row, 256 threads, the row cached in registers (D <= 8192, D % 8 == 0).
constexpr int kLnThreads = 256;
constexpr int kLnMaxVec = 4;
__global__ void __launch_bounds__(kLnThreads) layernorm_mod_kernel(
    const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ w, const float* __restrict__ b,
    const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift, long long ld_mod,
    __nv_bfloat16* __restrict__ out, long long ldo, int D, float eps) {
  __shared__ float red[32];
  grid_dep_launch();
  grid_dep_wait();
  const int m = blockIdx.x;
  const int nvec = D / 8;
  float v[kLnMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
      const uint4 raw = *reinterpret_cast<const uint4*>(x + m * ldx + c * 8);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(p[j]);
        v[i][2 * j] = f.x;
        v[i][2 * j + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
  const float mean = block_sum(sum, red) / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / static_cast<float>(D) + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
      const uint4 sraw = *reinterpret_cast<const uint4*>(scale + m * ld_mod + c * 8);
      const uint4 hraw = *reinterpret_cast<const uint4*>(shift + m * ld_mod + c * 8);
      const __nv_bfloat16* sc = reinterpret_cast<const __nv_bfloat16*>(&sraw);
      const __nv_bfloat16* sh = reinterpret_cast<const __nv_bfloat16*>(&hraw);
      uint4 ov;
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(&ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float h = (v[i][j] - mean) * rstd;
        if (w) h = h * w[c * 8 + j] + b[c * 8 + j];
        const float one_plus = bf16_round(1.0f + __bfloat162float(sc[j]));
        o[j] = __float2bfloat16_rn(h * one_plus + __bfloat162float(sh[j]));
      }
      *reinterpret_cast<uint4*>(out + m * ldo + c * 8) = ov;
    }
  }
}

// FinalLayer.linear + output map on an already LN-modulated row a (bf16 [M, D]); one CTA per token row:
//   o_c = bf16(sum_d a_d W[c,d] + bias_c);  pred_c = out_sigmoid ? bf16(bf16(2 * bf16(sigmoid(o_c))) - 1) : o_c
// 8 warps x 4 channels each; every lane keeps 4 independent 16-byte weight loads in flight per iteration.
__global__ void __launch_bounds__(256) head_out_kernel(const __nv_bfloat16* __restrict__ a, int D,
                                                       const __nv_bfloat16* __restrict__ Wf,
                                                       const __nv_bfloat16* __restrict__ bfin, int C, int out_sigmoid,
                                                       float* __restrict__ pred) {
  extern __shared__ float hrow[];  // [D]
  grid_dep_launch();
  grid_dep_wait();
  const long long m = blockIdx.x;
  for (int c8 = threadIdx.x; c8 < D / 8; c8 += blockDim.x) {
    float v[8];
    load_bf16x8(a + m * D + c8 * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) hrow[c8 * 8 + j] = v[j];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = warp * 4; c0 < C; c0 += 32) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = lane * 8; d < D; d += 256) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        raw[u] = (c0 + u < C) ? *reinterpret_cast<const uint4*>(Wf + static_cast<long long>(c0 + u) * D + d)
                              : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(p2[j]);
          acc[u] = fmaf(hrow[d + 2 * j], f.x, acc[u]);
          acc[u] = fmaf(hrow[d + 2 * j + 1], f.y, acc[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if (lane < 4 && c0 + lane < C) {
      float o = bf16_round(acc[lane] + (bfin ? __bfloat162float(bfin[c0 + lane]) : 0.f));
      if (out_sigmoid) {
        const float sg = bf16_round(1.0f / (1.0f + expf(-o)));
        o = bf16_round(bf16_round(2.0f * sg) - 1.0f);
      }
      pred[m * C + c0 + lane] = o;
    }
  }
}

// One Euler–Maruyama step (sampling_x.py:33-41) or the last deterministic Euler step (:24-30), fp32, written with
// explicit round-to-nearest intrinsics so that nvcc cannot contract a*b+c: each torch op rounds separately.
struct SdeStep {
  float t, dt, denom, var, one_minus_t, noise_scale;
  float cfg;
  int cfg_mult;
  int last;
};
__global__ void sde_step_kernel(float* __restrict__ x, const float* __restrict__ pred, const float* __restrict__ noise,
                                int n, SdeStep s, __nv_bfloat16* __restrict__ xb, int C, int Cp) {
  grid_dep_launch();
  grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv = x[i];
  float v = __fdiv_rn(__fsub_rn(pred[i], xv), s.denom);  // v = (output - combined) / clamp(1 - t, 0.05)
  if (s.cfg_mult == 2) {
    const float vu = __fdiv_rn(__fsub_rn(pred[n + i], xv), s.denom);
    v = __fadd_rn(vu, __fmul_rn(s.cfg, __fsub_rn(v, vu)));  // uncond + cfg * (cond - uncond)
  }
  float xn;
  if (s.last) {
    xn = __fadd_rn(xv, __fmul_rn(v, s.dt));
  } else {
    const float score = __fdiv_rn(__fsub_rn(__fmul_rn(s.t, v), xv), s.var);
    const float drift = __fadd_rn(v, __fmul_rn(s.one_minus_t, score));
    xn = __fadd_rn(__fadd_rn(xv, __fmul_rn(drift, s.dt)), __fmul_rn(s.noise_scale, noise[i]));
  }
  x[i] = xn;
  if (xb) {  // xb rows are padded to Cp columns (zeros beyond C): the input_proj GEMM reads full 128-byte k-blocks
    const __nv_bfloat16 b = __float2bfloat16_rn(xn);
    const int r = i / C, c = i % C, rows = n / C;
    xb[static_cast<long long>(r) * Cp + c] = b;
    if (s.cfg_mult == 2) xb[static_cast<long long>(rows + r) * Cp + c] = b;
  }
}

// x0 = noise[0]; xb = bf16(cat[x0] * mult)
__global__ void sde_init_kernel(float* __restrict__ x, const float* __restrict__ noise0, int n, int cfg_mult,
                                __nv_bfloat16* __restrict__ xb, int C, int Cp) {
  grid_dep_launch();
  grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = noise0[i];
  x[i] = v;
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  const int r = i / C, c = i % C, rows = n / C;
  xb[static_cast<long long>(r) * Cp + c] = b;
  if (cfg_mult == 2) xb[static_cast<long long>(rows + r) * Cp + c] = b;
}

template <typename... KArgs, typename... Args>
static int launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  LaunchCfg lc(grid, block, smem, st, pdl);
  BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, kern, static_cast<KArgs>(args)...));
  return BD_OK;
}

struct HeadWs {
  // offsets (bytes) into the workspace
  size_t xb, h, a, o, qkv, g, y, mod, cemb, tfreq, th, temb, pred, x, condb, tvals, gemm, total;
};

static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

static HeadWs head_ws_layout(const bd_head_weights_t& w, int M, int n_rows_x, int S) {
  HeadWs L{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes);
    return o;
  };
  const size_t D = w.D, C = w.C;
  const int n_mod = w.n_ada * 6 * w.D + 2 * w.D;
  const size_t wide = 3 * D > static_cast<size_t>(2 * w.hidden) ? 3 * D : 2 * w.hidden;
  L.xb = take(static_cast<size_t>(M) * ((C + 63) / 64 * 64) * 2);
  L.h = take(M * D * 2);
  L.a = take(M * D * 2);
  L.o = take(M * D * 2);
  L.qkv = take(M * wide * 2);
  L.g = take(static_cast<size_t>(M) * w.hidden * 2);
  L.y = take(M * D * 2);
  L.mod = take(static_cast<size_t>(M) * n_mod * 2);
  L.cemb = take(M * D * 2);
  L.tfreq = take(static_cast<size_t>(S + 1) * 256 * 2);
  L.th = take(static_cast<size_t>(S + 1) * D * 2);
  L.temb = take(static_cast<size_t>(S + 1) * D * 2);
  L.pred = take(static_cast<size_t>(M) * C * 4);
  L.x = take(static_cast<size_t>(n_rows_x) * C * 4);
  L.condb = take(static_cast<size_t>(M) * w.Dz * 2);
  L.tvals = take(static_cast<size_t>(S + 1) * 4);
  size_t gm = 0;
  auto gw = [&](int m, int n, int k) {
    size_t b = gemm_workspace_bytes(m, n, k, 0, 0);
    if (b > gm) gm = b;
  };
  gw(M, w.D, w.C);
  gw(S + 1, w.D, 256);
  gw(S + 1, w.D, w.D);
  gw(M, w.D, w.Dz);
  gw(M, n_mod, w.D);
  gw(M, 3 * w.D, w.D);
  gw(M, w.D, w.D);
  gw(M, 2 * w.hidden, w.D);
  gw(M, w.hidden, w.D);
  gw(M, w.D, w.hidden);
  L.gemm = take(gm);
  L.total = off;
  return L;
}


// ---------------------------------------------------------------------------------------------------------------------
// persistent path: the whole sampler as one program for bd_stream_kernel (bd_stream.cuh)
// ---------------------------------------------------------------------------------------------------------------------
struct HeadStreamWs {
  size_t xb, h, y, mod, a, qkv, o, g, cemb, condb, tfreq, th, temb, part, pred, x, sync, total;
  size_t mod_bytes, y_bytes;  // one modulation buffer / one y image (fillers: 2 mod buffers, S + 1 y images)
};

static size_t blocked_bytes(int K) { return static_cast<size_t>((K + 63) / 64) * kSlotBytes; }

// Filler policy (bd_head_set_fillers): 0 = the adaLN GEMM runs in line at the start of every evaluation (round-1 program);
// 1 = the adaLN GEMM of evaluation i + 1 runs as filler pieces in the bubbles of evaluation i (bd_stream.cuh, FILLERS).
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
static int g_head_fillers = env_int("BD_HEAD_FILLERS", 1);  // env: debugging aid (tests run both policies explicitly)
static int g_head_fill_row_kb = 22;   // k-blocks (of 128 weight rows x 64) wanted in a slot that hides a row op
static int g_head_fill_gemm_kb = 8;   // ... in a slot that hides a GEMM -> GEMM boundary

static HeadStreamWs head_stream_ws_layout(const bd_head_weights_t& w, int S) {
  HeadStreamWs L{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 1023) & ~size_t(1023);
    return o;
  };
  const size_t D = w.D;
  const size_t n_mod = static_cast<size_t>(w.n_ada) * 6 * D + 2 * D;
  L.mod_bytes = (128 * n_mod * 2 + 1023) & ~size_t(1023);
  L.y_bytes = blocked_bytes(w.D);
  L.xb = take(blocked_bytes(w.C));
  L.h = take(128 * D * 2);
  L.y = take(L.y_bytes * static_cast<size_t>(S + 1));
  L.mod = take(2 * L.mod_bytes);
  L.a = take(blocked_bytes(w.D));
  L.qkv = take(blocked_bytes(3 * w.D));  // blocked: the attention op fetches its Q / K / V tiles by bulk copy
  L.o = take(blocked_bytes(w.D));
  L.g = take(blocked_bytes(w.hidden));
  L.cemb = take(128 * D * 2);
  L.condb = take(blocked_bytes(w.Dz));
  L.tfreq = take(blocked_bytes(256));
  L.th = take(blocked_bytes(w.D));
  L.temb = take(128 * D * 2);
  L.part = take(static_cast<size_t>(4) * 128 * D * 4);
  L.pred = take(static_cast<size_t>(128) * w.C * 4);
  L.x = take(static_cast<size_t>(128) * w.C * 4);
  L.sync = take(64);
  L.total = off;
  return L;
}

// Cut the adaLN GEMM (P passes per CTA x KB k-blocks) into pieces for `n_slots` bubbles whose wanted sizes are `want`
// (k-blocks). Sizes are scaled so that the pieces add up to the whole GEMM; a piece never spans two passes; sizes are
// multiples of the ring step. out[slot] = list of (pass, kb0, kbn).
struct HeadPiece { int slot, pass, kb0, kbn; };
static int plan_head_pieces(int P, int KB, const int* want, int n_slots, HeadPiece* out, int cap) {
  long long total_want = 0;
  for (int i = 0; i < n_slots; ++i) total_want += want[i];
  const long long supply = static_cast<long long>(P) * KB;
  if (total_want <= 0 || supply <= 0) return 0;
  int n = 0, pass = 0, kb = 0;
  long long given = 0, wanted = 0;
  for (int sl = 0; sl < n_slots && pass < P; ++sl) {
    wanted += want[sl];
    // cumulative target keeps rounding errors from piling up; the last slot takes whatever is left
    long long target = sl == n_slots - 1 ? supply : (wanted * supply + total_want / 2) / total_want;
    long long d = target - given;
    d = (d + kKbPerStep - 1) / kKbPerStep * kKbPerStep;
    while (d > 0 && pass < P) {
      int take = static_cast<int>(d < KB - kb ? d : KB - kb);
      if (KB - kb - take < 2 * kKbPerStep) take = KB - kb;  // no crumbs at the end of a pass
      if (n >= cap) return -1;
      out[n++] = HeadPiece{sl, pass, kb, take};
      given += take;
      d -= take;
      kb += take;
      if (kb >= KB) {
        kb = 0;
        ++pass;
      }
    }
  }
  return (pass >= P) ? n : -1;
}

static int head_sample_stream(const bd_head_weights_t& w, const float* cond, const float* noise, const float* sched_host,
                              int B, int pn, int cfg_mult, float cfg, int S, float* x_out, float* trace, void* workspace,
                              size_t workspace_bytes, cudaStream_t st) {
  const int R = B * cfg_mult, M = R * pn, nx = B * pn;
  const int D = w.D, C = w.C, G = w.stream_ctas;
  BD_REQUIRE(M <= 128 && S + 1 <= kStreamMaxIter && S + 1 <= 128 && G > 0 && G <= num_sms());
  // row ops: token row r (timestep row r) is handled by CTA r mod G (G < 128: several engines side by side)
  BD_REQUIRE((D % 64) == 0 && (w.Dz % 8) == 0 && C <= 64 && (w.hidden % 8) == 0 && D <= 6144);
  const int n_mod = w.n_ada * 6 * D + 2 * D;
  const HeadStreamWs L = head_stream_ws_layout(w, S);
  if (workspace_bytes < L.total) return BD_ERR_WORKSPACE;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  // blocked operands are read in whole 64-column k-blocks and 128-row tiles: padding must be finite (zero)
  BD_CUDA_TRY(cudaMemsetAsync(base, 0, L.total, st));

  // ---- filler plan: slots in GEMM order = after ln0, and per block after wo (row op), after w1 (boundary), after w2
  // (row op; not for the last block, whose final-row op uses the A ring as scratch) ----
  constexpr int kMaxPieces = 64;
  HeadPiece pieces[kMaxPieces];
  int n_pieces = 0;
  const int n_slots = 3 * w.n_blocks;  // 1 + 3 * n_blocks - 1
  bool fill = g_head_fillers != 0 && S >= 1;
  if (fill) {
    int P = 0;
    for (int c = 0; c < G; ++c) {
      const StreamPart p = stream_partition(n_mod, D, 1, G, c);
      P = p.npass > P ? p.npass : P;
    }
    int want[3 * BD_HEAD_MAX_BLOCKS];
    want[0] = g_head_fill_row_kb;
    for (int b = 0; b < w.n_blocks; ++b) {
      want[1 + 3 * b] = g_head_fill_row_kb;
      want[2 + 3 * b] = g_head_fill_gemm_kb;
      if (b + 1 < w.n_blocks) want[3 + 3 * b] = g_head_fill_row_kb;
    }
    n_pieces = plan_head_pieces(P, (D + 63) / 64, want, n_slots, pieces, kMaxPieces);
    if (n_pieces <= 0) fill = false;
  }
  BD_REQUIRE(8 + 4 + 7 * w.n_blocks + 1 + n_pieces <= kStreamMaxOps);

  static thread_local StreamProgram prog;
  prog = StreamProgram{};
  prog.family = G < 128 ? kStreamFamHeadSmall : kStreamFamHead;
  prog.M = M;
  prog.n_ctas = G;
  prog.rows_x = nx;
  prog.cfg_mult = cfg_mult;
  prog.cfg = cfg;
  prog.n_iter = S + 1;
  prog.sync = reinterpret_cast<unsigned int*>(base + L.sync);
  for (int i = 0; i <= S; ++i)
    for (int j = 0; j < 6; ++j) prog.sched[i][j] = sched_host[i * 8 + j];
  int n = 0;
  auto gemm_op = [&](const void* W, const void* A, const void* bias, int N, int K, int ksplit, int epi, int act, void* out,
                     long long ld, bool blocked) -> StreamOp& {
    StreamOp& op = prog.ops[n++];
    op = StreamOp{};
    op.kind = kOpGemm;
    op.sub = epi;
    op.N = N;
    op.K = K;
    op.ksplit = ksplit;
    op.act = act;
    op.flags = blocked ? 1 : 0;
    op.wait_prev = 1;
    op.p0 = W;
    op.p1 = A;
    op.p2 = bias;
    op.o0 = out;
    op.l0 = ld;
    return op;
  };
  auto row_op_ = [&](int sub) -> StreamOp& {
    StreamOp& op = prog.ops[n++];
    op = StreamOp{};
    op.kind = kOpRow;
    op.sub = sub;
    op.wait_prev = 1;
    op.f0 = 1e-6f;
    return op;
  };
  // row ops read the modulation of THIS evaluation: buffer it & 1 when the fillers double-buffer it
  auto mod_reader = [&](StreamOp& op) {
    if (fill) {
      op.flags |= kFlagParityIt;
      op.l1 = static_cast<long long>(L.mod_bytes);
    }
  };
  int next_piece = 0;
  auto fill_slot = [&](int slot) {  // the adaLN pieces of evaluation it + 1 that hide the bubble at this point
    while (fill && next_piece < n_pieces && pieces[next_piece].slot == slot) {
      const HeadPiece& pc = pieces[next_piece++];
      StreamOp& op = gemm_op(w.ada_w, base + L.y, w.ada_b, n_mod, D, 1, kEpiBias, kActNone, base + L.mod, n_mod, false);
      op.wait_prev = 0;
      op.flags |= kFlagFiller | kFlagAPerIt | kFlagParityNext | kFlagSkipLast;
      op.i0 = 1;                                        // A = y[it + 1]
      op.l1 = static_cast<long long>(L.mod_bytes);      // out = mod[(it + 1) & 1]
      op.pc_pass = pc.pass;
      op.pc_kb0 = pc.kb0;
      op.pc_kbn = pc.kbn;
      op.pc_flags = (pc.kb0 == 0 ? kPieceFirst : 0) | (pc.kb0 + pc.kbn == (D + 63) / 64 ? kPieceLast : 0);
    }
  };
  __nv_bfloat16* mod = reinterpret_cast<__nv_bfloat16*>(base + L.mod);
  // ---- once per call ----
  {
    StreamOp& op = row_op_(kRowCastCond);
    op.wait_prev = 0;
    op.p0 = cond;
    op.o0 = base + L.condb;
    op.N = w.Dz;
  }
  {
    StreamOp& op = row_op_(kRowTFreq);
    op.wait_prev = 0;
    op.o0 = base + L.tfreq;
    op.N = 256;
  }
  {
    StreamOp& op = row_op_(kRowInit);
    op.wait_prev = 0;
    op.p0 = noise;
    op.o0 = base + L.x;
    op.o1 = base + L.xb;
    op.N = C;
  }
  gemm_op(w.time0_w, base + L.tfreq, w.time0_b, D, 256, 1, kEpiBias, kActSilu, base + L.th, 0, true);
  prog.ops[n - 1].i1 = S + 1;  // one row per timestep
  gemm_op(w.cond_w, base + L.condb, w.cond_b, D, w.Dz, 1, kEpiBias, kActNone, base + L.cemb, D, false);
  gemm_op(w.time2_w, base + L.th, w.time2_b, D, D, 1, kEpiBias, kActNone, base + L.temb, D, false);
  prog.ops[n - 1].i1 = S + 1;
  if (fill) {
    {  // y of EVERY evaluation, then the modulation of evaluation 0 (those of 1.. are filler pieces)
      StreamOp& op = row_op_(kRowSiluAddAll);
      op.p0 = base + L.temb;
      op.p1 = base + L.cemb;
      op.o0 = base + L.y;
      op.l1 = static_cast<long long>(L.y_bytes);
      op.N = D;
    }
    gemm_op(w.ada_w, base + L.y, w.ada_b, n_mod, D, 1, kEpiBias, kActNone, mod, n_mod, false);
  } else {  // y of evaluation 0 (the y of evaluation i+1 is produced by evaluation i's SDE op)
    StreamOp& op = row_op_(kRowSiluAdd);
    op.p0 = base + L.temb;
    op.p1 = base + L.cemb;
    op.o0 = base + L.y;
    op.N = D;
  }
  prog.n_pre = n;
  // ---- one evaluation + SDE step ----
  gemm_op(w.input_proj_w, base + L.xb, w.input_proj_b, D, C, 1, kEpiBias, kActNone, base + L.h, D, false);
  if (!fill) {
    // input_proj (needs xb) and the adaLN GEMM (needs y) both depend only on the previous SDE op: the second one skips the
    // barrier of the first (wait_prev = -1), so the two weight streams run back to back
    gemm_op(w.ada_w, base + L.y, w.ada_b, n_mod, D, 1, kEpiBias, kActNone, mod, n_mod, false);
    prog.ops[n - 1].wait_prev = -1;
  }
  {
    StreamOp& op = row_op_(kRowLnMod);
    op.p0 = base + L.h;
    op.p1 = w.blocks[0].norm1_w;
    op.p2 = w.blocks[0].norm1_b;
    op.p3 = mod;
    op.p4 = mod + D;
    op.l0 = n_mod;
    op.o0 = base + L.a;
    op.N = D;
    mod_reader(op);
  }
  fill_slot(0);
  const int switch_freq = w.n_blocks / w.n_ada > 0 ? w.n_blocks / w.n_ada : 1;
  const int ks_wo = stream_ksplit_for(D, D, G), ks_w2 = stream_ksplit_for(D, w.hidden, G);
  const __nv_bfloat16* mf = mod + static_cast<long long>(w.n_ada) * 6 * D;
  for (int blk = 0; blk < w.n_blocks; ++blk) {
    const bd_head_block_t& bw = w.blocks[blk];
    const __nv_bfloat16* md = mod + static_cast<long long>(blk / switch_freq) * 6 * D;
    gemm_op(bw.wqkv_w, base + L.a, bw.wqkv_b, 3 * D, D, 1, kEpiBias, kActNone, base + L.qkv, 0, true);
    {
      StreamOp& op = prog.ops[n++];
      op = StreamOp{};
      op.kind = kOpAttn;
      op.wait_prev = 1;
      op.p0 = base + L.qkv;
      op.o0 = base + L.o;
      op.N = D;
      op.K = w.head_dim;
      op.i0 = pn;
    }
    gemm_op(bw.wo_w, base + L.o, nullptr, D, D, ks_wo, kEpiPartial, 0, base + L.part, 0, false);
    {
      StreamOp& op = row_op_(kRowSplitkLnMod);
      op.p0 = base + L.part;
      op.i0 = ks_wo;
      op.p5 = bw.wo_b;
      op.p6 = md + 2 * D;
      op.o1 = base + L.h;
      op.p1 = bw.norm2_w;
      op.p2 = bw.norm2_b;
      op.p3 = md + 3 * D;
      op.p4 = md + 4 * D;
      op.l0 = n_mod;
      op.o0 = base + L.a;
      op.N = D;
      mod_reader(op);
    }
    fill_slot(1 + 3 * blk);
    if (w.use_swiglu)
      gemm_op(bw.w1_w, base + L.a, bw.w1_b, 2 * w.hidden, D, 1, kEpiSwiglu8, 0, base + L.g, 0, true);
    else
      gemm_op(bw.w1_w, base + L.a, bw.w1_b, w.hidden, D, 1, kEpiBias, kActSilu, base + L.g, 0, true);
    fill_slot(2 + 3 * blk);
    gemm_op(bw.w2_w, base + L.g, nullptr, D, w.hidden, ks_w2, kEpiPartial, 0, base + L.part, 0, false);
    {
      const bool last = blk + 1 == w.n_blocks;
      StreamOp& op = row_op_(last ? kRowFinal : kRowSplitkLnMod);
      op.p0 = base + L.part;
      op.i0 = ks_w2;
      op.p5 = bw.w2_b;
      op.p6 = md + 5 * D;
      op.o1 = base + L.h;
      op.l0 = n_mod;
      op.N = D;
      if (!last) {
        const bd_head_block_t& nb = w.blocks[blk + 1];
        const __nv_bfloat16* mdn = mod + static_cast<long long>((blk + 1) / switch_freq) * 6 * D;
        op.p1 = nb.norm1_w;
        op.p2 = nb.norm1_b;
        op.p3 = mdn;
        op.p4 = mdn + D;
        op.o0 = base + L.a;
      } else {
        op.p1 = w.final_w;
        op.p2 = w.final_b;
        op.p3 = mf;
        op.p4 = mf + D;
        op.o0 = base + L.pred;
        op.o2 = trace;
        op.i1 = C;
        op.i2 = w.out_sigmoid;
      }
      mod_reader(op);
      if (!last) fill_slot(3 + 3 * blk);
    }
  }
  if (fill && next_piece != n_pieces) return BD_ERR_INVALID;  // every piece placed exactly once
  {
    StreamOp& op = row_op_(kRowSde);
    op.p0 = base + L.pred;
    op.p1 = noise;
    op.o0 = base + L.x;
    op.o1 = base + L.xb;
    op.o2 = x_out;
    op.N = C;
    if (!fill) {
      op.p4 = base + L.temb;  // + y of the next evaluation
      op.p5 = base + L.cemb;
      op.o3 = base + L.y;
      op.i0 = D;
    }
  }
  prog.n_body = n - prog.n_pre;
  prog.n_post = 0;
  return stream_launch(prog, st);
}

}  // namespace bd

using namespace bd;

extern "C" {

// Filler policy of the persistent sampler: mode 0 = adaLN GEMM in line (round-1 program), 1 = as filler pieces of the
// previous evaluation; row_kb / gemm_kb = wanted k-blocks per slot (<= 0: keep).
int bd_head_set_fillers(int mode, int row_kb, int gemm_kb) {
  BD_REQUIRE(mode == 0 || mode == 1);
  g_head_fillers = mode;
  if (row_kb > 0) g_head_fill_row_kb = row_kb;
  if (gemm_kb > 0) g_head_fill_gemm_kb = gemm_kb;
  return BD_OK;
}

// tests: the filler plan for P passes x KB k-blocks over slots of wanted sizes; out = n x {slot, pass, kb0, kbn}
int bd_head_plan_pieces(int P, int KB, const int* want, int n_slots, int* out, int cap) {
  BD_REQUIRE(want && out && P > 0 && KB > 0 && n_slots > 0 && cap > 0);
  HeadPiece tmp[256];
  const int n = plan_head_pieces(P, KB, want, n_slots, tmp, cap < 256 ? cap : 256);
  for (int i = 0; i < n; ++i) {
    out[4 * i] = tmp[i].slot;
    out[4 * i + 1] = tmp[i].pass;
    out[4 * i + 2] = tmp[i].kb0;
    out[4 * i + 3] = tmp[i].kbn;
  }
  return n;
}

size_t bd_head_workspace_bytes(const bd_head_weights_t* w, int B, int pn, int cfg_mult, int S) {
  if (!w || B <= 0 || pn <= 0 || cfg_mult < 1 || cfg_mult > 2 || S < 0) return 0;
  if (w->w_tiled == 2) return head_stream_ws_layout(*w, S).total;
  return head_ws_layout(*w, B * cfg_mult * pn, B * pn, S).total;
}

// debug / tests: byte offsets of the workspace regions (order of the HeadWs / HeadStreamWs fields), returns the count
int bd_head_ws_offsets(const bd_head_weights_t* w, int B, int pn, int cfg_mult, int S, size_t* out, int cap) {
  if (!w || !out) return 0;
  if (w->w_tiled == 2) {
    const HeadStreamWs L = head_stream_ws_layout(*w, S);
    const size_t v[] = {L.xb, L.h, L.y, L.mod, L.a, L.qkv, L.o, L.g, L.cemb, L.condb, L.tfreq, L.th, L.temb, L.part, L.pred,
                        L.x, L.sync, L.total};
    const int n = static_cast<int>(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
  }
  const HeadWs L = head_ws_layout(*w, B * cfg_mult * pn, B * pn, S);
  const size_t v[] = {L.xb, L.h, L.a, L.o, L.qkv, L.g, L.y, L.mod, L.cemb, L.tfreq, L.th, L.temb, L.pred, L.x, L.condb,
                      L.tvals, L.gemm, L.total};
  const int n = static_cast<int>(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
  return n;
}

#define BD_TRY(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != BD_OK) return _rc; \
  } while (0)

int bd_head_sample(const bd_head_weights_t* wp, const float* cond, const float* noise, const float* sched_host, int B,
                   int pn, int cfg_mult, float cfg, int S, float* x_out, float* trace, void* workspace,
                   size_t workspace_bytes, int flags, bd_stream_t stream_) {
  BD_REQUIRE(wp && cond && noise && sched_host && x_out && workspace);
  const bd_head_weights_t& w = *wp;
  BD_REQUIRE(B > 0 && pn > 0 && (cfg_mult == 1 || cfg_mult == 2) && S >= 0);
  BD_REQUIRE(w.D > 0 && (w.D % 64) == 0 && w.D <= 6144 && w.C > 0 && (w.C % 8) == 0 && w.Dz > 0 && (w.Dz % 8) == 0);
  BD_REQUIRE(w.n_blocks > 0 && w.n_blocks <= BD_HEAD_MAX_BLOCKS && w.n_ada > 0 && (w.n_blocks % w.n_ada) == 0);
  BD_REQUIRE(w.head_dim == 64 || w.head_dim == 128);
  BD_REQUIRE(pn <= 64);  // one KV tile per block of parallel tokens
  BD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (w.w_tiled == 2) {
    if (B * cfg_mult * pn > 128) return BD_ERR_UNSUPPORTED;  // stream-packed weights: one 128-row tile (use the tiled path)
    BD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
    return head_sample_stream(w, cond, noise, sched_host, B, pn, cfg_mult, cfg, S, x_out, trace, workspace,
                              workspace_bytes, st);
  }
  const bool pdl = (flags & 1) != 0;
  const int R = B * cfg_mult, M = R * pn, nx = B * pn;
  const int D = w.D, C = w.C;
  const int n_mod = w.n_ada * 6 * D + 2 * D;
  const HeadWs L = head_ws_layout(w, M, nx, S);
  if (workspace_bytes < L.total) return BD_ERR_WORKSPACE;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  auto bfp = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(base + off); };
  __nv_bfloat16 *xb = bfp(L.xb), *h = bfp(L.h), *a = bfp(L.a), *o = bfp(L.o), *qkv = bfp(L.qkv), *g = bfp(L.g),
                *y = bfp(L.y), *mod = bfp(L.mod), *cemb = bfp(L.cemb), *tfreq = bfp(L.tfreq), *th = bfp(L.th),
                *temb = bfp(L.temb), *condb = bfp(L.condb);
  float* pred = reinterpret_cast<float*>(base + L.pred);
  float* x = reinterpret_cast<float*>(base + L.x);
  void* gws = base + L.gemm;
  const size_t gws_bytes = L.total - L.gemm;
  const int switch_freq = w.n_blocks / w.n_ada;

  auto gemm = [&](const void* A, long long lda, const void* W, int m, int n, int k, GemmEpi e) {
    return gemm_bf16(A, lda, W, k, m, n, k, e, gws, gws_bytes, 0, 0, pdl, st, w.w_tiled != 0);
  };
  auto bf = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };

  // ---- once per call: timestep values -> time embeddings for all S+1 evaluations; cond_embed(c) ----
  // sched_host rows: [t, dt, denom, var, one_minus_t, noise_scale, _, _]
  for (int r0 = 0; r0 <= S; r0 += 128) {  // timestep values travel as kernel parameters (graph-capture safe)
    TVals tv{};
    const int cnt = (S + 1 - r0) < 128 ? (S + 1 - r0) : 128;
    for (int i = 0; i < cnt; ++i) tv.t[i] = sched_host[(r0 + i) * 8];
    BD_TRY(launch(timestep_embedding_kernel, dim3(cnt), dim3(128), 0, st, false, tv, r0, 256, tfreq));
  }
  {
    GemmEpi e;
    e.bias = bf(w.time0_b);
    e.act = kActSilu;
    e.out = th;
    e.ld_out = D;
    BD_TRY(gemm(tfreq, 256, w.time0_w, S + 1, D, 256, e));
    GemmEpi e2;
    e2.bias = bf(w.time2_b);
    e2.out = temb;
    e2.ld_out = D;
    BD_TRY(gemm(th, D, w.time2_w, S + 1, D, D, e2));
  }
  {
    const long long n = static_cast<long long>(M) * w.Dz;
    BD_TRY(launch(cast_f32_bf16_kernel, dim3(static_cast<unsigned>((n / 4 + 255) / 256 + 1)), dim3(256), 0, st, pdl,
                  cond, condb, n));
    GemmEpi e;
    e.bias = bf(w.cond_b);
    e.out = cemb;
    e.ld_out = D;
    BD_TRY(gemm(condb, w.Dz, w.cond_w, M, D, w.Dz, e));
  }
  const int nel = nx * C;
  const int Cp = (C + 63) / 64 * 64;
  BD_CUDA_TRY(cudaMemsetAsync(xb, 0, static_cast<size_t>(M) * Cp * 2, st));
  BD_TRY(launch(sde_init_kernel, dim3((nel + 255) / 256), dim3(256), 0, st, false, x, noise, nel, cfg_mult, xb, C, Cp));

  // ---- S stochastic evaluations + 1 deterministic ----
  for (int it = 0; it <= S; ++it) {
    const float* sc = sched_host + it * 8;
    {  // h = input_proj(x)
      GemmEpi e;
      e.bias = bf(w.input_proj_b);
      e.out = h;
      e.ld_out = D;
      // tile-major input_proj weights are zero-padded to 64 columns, so K can be the padded width
      BD_TRY(gemm(xb, Cp, w.input_proj_w, M, D, w.w_tiled ? Cp : C, e));
    }
    {  // y = silu(t_emb + c_emb)
      const long long n = static_cast<long long>(M) * D / 8;
      BD_TRY(launch(silu_add_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, pdl,
                    (const __nv_bfloat16*)(temb + static_cast<long long>(it) * D), (const __nv_bfloat16*)cemb, y, M, D));
    }
    {  // all adaLN modulations in one GEMM: [ada_0 (6D) | ada_1 (6D) | ... | final (2D)]
      GemmEpi e;
      e.bias = bf(w.ada_b);
      e.out = mod;
      e.ld_out = n_mod;
      BD_TRY(gemm(y, D, w.ada_w, M, n_mod, D, e));
    }
    const __nv_bfloat16* mf = mod + static_cast<long long>(w.n_ada) * 6 * D;  // final scale | shift
    // a = norm1_0(h) * (1 + scale1) + shift1 for the first block (h comes straight from input_proj)
    BD_TRY(launch(layernorm_mod_kernel, dim3(M), dim3(kLnThreads), 0, st, pdl, (const __nv_bfloat16*)h, (long long)D,
                  w.blocks[0].norm1_w, w.blocks[0].norm1_b, (const __nv_bfloat16*)mod, (const __nv_bfloat16*)(mod + D),
                  (long long)n_mod, a, (long long)D, D, 1e-6f));
    for (int blk = 0; blk < w.n_blocks; ++blk) {
      const bd_head_block_t& bw = w.blocks[blk];
      const __nv_bfloat16* md = mod + static_cast<long long>(blk / switch_freq) * 6 * D;
      // chunk order: scale1, shift1, gate1, scale2, shift2, gate2
      {
        GemmEpi e;
        e.bias = bf(bw.wqkv_b);
        e.out = qkv;
        e.ld_out = 3 * D;
        BD_TRY(gemm(a, D, bw.wqkv_w, M, 3 * D, D, e));
      }
      BD_TRY(attn_run_head(qkv, o, R, pn, D, w.head_dim, pdl, st));
      // h = h + (wo(o) + b) * gate1 ; a = norm2(h) * (1 + scale2) + shift2
      // When K is split the reduction, this epilogue and the following LayerNorm-modulate run in ONE row kernel.
      auto gated_res_then_norm = [&](const void* A_, long long lda, const void* W_, int Kdim, const void* bias_,
                                     const __nv_bfloat16* gate_, const float* nw, const float* nb,
                                     const __nv_bfloat16* sc, const __nv_bfloat16* sh) -> int {
        GemmEpi e;
        e.bias = bf(bias_);
        e.gate = gate_;
        e.ld_gate = n_mod;
        e.res = h;
        e.ld_res = D;
        e.out = h;
        e.ld_out = D;
        int S_used = 1;
        BD_TRY(gemm_bf16(A_, lda, W_, Kdim, M, D, Kdim, e, gws, gws_bytes, 0, 0, pdl, st, w.w_tiled != 0, &S_used));
        if (S_used > 1) {
          HeadRowArgs ra;
          ra.partial = static_cast<const float*>(gws);
          ra.splits = S_used;
          ra.M = M;
          ra.D = D;
          ra.bias = bf(bias_);
          ra.gate = gate_;
          ra.h = h;
          ra.ln_w = nw;
          ra.ln_b = nb;
          ra.scale = sc;
          ra.shift = sh;
          ra.ld_mod = n_mod;
          ra.a = a;
          ra.eps = 1e-6f;
          return launch(head_splitk_row_kernel, dim3(M), dim3(kRowThreads), 0, st, pdl, ra);
        }
        return launch(layernorm_mod_kernel, dim3(M), dim3(kLnThreads), 0, st, pdl, (const __nv_bfloat16*)h, (long long)D,
                      nw, nb, sc, sh, (long long)n_mod, a, (long long)D, D, 1e-6f);
      };
      BD_TRY(gated_res_then_norm(o, D, bw.wo_w, D, bw.wo_b, md + 2 * D, bw.norm2_w, bw.norm2_b, md + 3 * D, md + 4 * D));
      if (w.use_swiglu) {
        GemmEpi e;
        e.bias = bf(bw.w1_b);
        e.swiglu = 1;
        e.out = g;
        e.ld_out = w.hidden;
        BD_TRY(gemm(a, D, bw.w1_w, M, 2 * w.hidden, D, e));
      } else {
        GemmEpi e;
        e.bias = bf(bw.w1_b);
        e.act = kActSilu;
        e.out = g;
        e.ld_out = w.hidden;
        BD_TRY(gemm(a, D, bw.w1_w, M, w.hidden, D, e));
      }
      // h = h + (w2(g) + b) * gate2 ; a = the NEXT normalisation: next block's norm1-modulate, or the final layer's
      if (blk + 1 < w.n_blocks) {
        const bd_head_block_t& nb = w.blocks[blk + 1];
        const __nv_bfloat16* mdn = mod + static_cast<long long>((blk + 1) / switch_freq) * 6 * D;
        BD_TRY(gated_res_then_norm(g, w.hidden, bw.w2_w, w.hidden, bw.w2_b, md + 5 * D, nb.norm1_w, nb.norm1_b, mdn, mdn + D));
      } else {
        BD_TRY(gated_res_then_norm(g, w.hidden, bw.w2_w, w.hidden, bw.w2_b, md + 5 * D, nullptr, nullptr, mf, mf + D));
      }
    }
    BD_TRY(launch(head_out_kernel, dim3(M), dim3(256), static_cast<size_t>(D) * 4, st, pdl, (const __nv_bfloat16*)a, D,
                  bf(w.final_w), bf(w.final_b), C, w.out_sigmoid, pred));
    if (trace)
      BD_CUDA_TRY(cudaMemcpyAsync(trace + static_cast<long long>(it) * M * C, pred, sizeof(float) * M * C,
                                  cudaMemcpyDeviceToDevice, st));
    SdeStep s;
    s.t = sc[0];
    s.dt = sc[1];
    s.denom = sc[2];
    s.var = sc[3];
    s.one_minus_t = sc[4];
    s.noise_scale = sc[5];
    s.cfg = cfg;
    s.cfg_mult = cfg_mult;
    s.last = (it == S) ? 1 : 0;
    const float* nz = (it < S) ? noise + static_cast<long long>(it + 1) * nel : noise;
    BD_TRY(launch(sde_step_kernel, dim3((nel + 255) / 256), dim3(256), 0, st, pdl, x, (const float*)pred, nz, nel, s,
                  (it < S) ? xb : (__nv_bfloat16*)nullptr, C, Cp));
  }
  BD_CUDA_TRY(cudaMemcpyAsync(x_out, x, sizeof(float) * nel, cudaMemcpyDeviceToDevice, st));
  return BD_OK;
}

}  // extern "C"
