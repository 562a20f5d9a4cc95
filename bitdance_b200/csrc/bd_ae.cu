// bd_ae.cu — the HBM-bound pieces of the binary tokenizer around the tcgen05 convolutions: layout changes,
// GroupNorm (stats + apply + swish), AdaptiveGroupNorm, all on NHWC tensors, vectorised 16-byte accesses.
//
// Reference: modeling/vision_encoder/autoencoder.py — swish :10, ResBlock GroupNorm(32, eps 1e-6) :28-29,
// AdaptiveGroupNorm.forward :260-277 (unbiased variance over the +-1 grid), decode_image rearrange
// modeling/t2i_pipeline.py:280. Rounding policy (autocast): GroupNorm runs and returns fp32, swish on fp32 stays
// fp32, and the consuming convolution rounds its input to bf16 — so "apply" writes bf16(swish(gn(x))) directly.
#include "bd_host.h"
#include "bd_ptx.cuh"

namespace bd {

// ---------------------------------------------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------------------------------------------
// x fp32 NCHW [B,C,H,W] -> bf16 NHWC [B,H,W,Cp] (channels >= C zero): the autocast input cast of conv_in.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           int B, int C, long long HW, int Cp) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  const long long b = i / HW, p = i % HW;
  for (int c = 0; c < Cp; ++c) {
    const float v = c < C ? x[(b * C + c) * HW + p] : 0.f;
    out[i * Cp + c] = __float2bfloat16_rn(v);
  }
}

// tokens fp32 [B, h*w, C] in patch-raster order '(hb wb p1 p2)' -> bf16 NHWC grid [B, h, w, C]
// (decode_image: 'b (h w p1 p2) c -> b c (h p1) (w p2)', t2i_pipeline.py:280). One thread per 8 channels.
__global__ void __launch_bounds__(256) tokens_to_grid_kernel(const float* __restrict__ tok, __nv_bfloat16* __restrict__ out,
                                                             int B, int h, int w, int C, int ps) {
  const long long n = static_cast<long long>(B) * h * w * (C / 8);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cv = static_cast<int>(i % (C / 8));
  long long r = i / (C / 8);
  const int x = static_cast<int>(r % w);
  r /= w;
  const int y = static_cast<int>(r % h);
  const int b = static_cast<int>(r / h);
  const int wb = w / ps;
  const long long t = ((static_cast<long long>(y / ps) * wb + x / ps) * ps + (y % ps)) * ps + (x % ps);
  const float* s = tok + (static_cast<long long>(b) * h * w + t) * C + cv * 8;
  uint4 pk;
  __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int j = 0; j < 4; ++j) p2[j] = __floats2bfloat162_rn(s[2 * j], s[2 * j + 1]);
  *reinterpret_cast<uint4*>(out + i * 8) = pk;
}

// bf16 NHWC [B, HW, C] -> NCHW [B, C, HW] (bf16 or fp32): the API-facing layout of VQModel.encode's output.
template <bool OUT_F32>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, void* __restrict__ out,
                                                           int B, int C, long long HW) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over B*C*HW, hw fastest
  if (i >= B * C * HW) return;
  const long long p = i % HW;
  const long long bc = i / HW;
  const long long c = bc % C, b = bc / C;
  const __nv_bfloat16 v = x[(b * HW + p) * C + c];
  if (OUT_F32)
    static_cast<float*>(out)[i] = __bfloat162float(v);
  else
    static_cast<__nv_bfloat16*>(out)[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm statistics (32 groups) over NHWC, deterministic two-stage reduction
// ---------------------------------------------------------------------------------------------------------------
template <bool IN_F32>
__device__ __forceinline__ void load8(const void* base, long long idx8, float (&v)[8]) {
  if (IN_F32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx8 * 8);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 raw = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + idx8 * 8);
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(p[j]);
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  }
}

// grid (chunks, B). part: [B][chunks][32][2] (sum, sumsq). Requires C = 8 * 2^k, 32 <= C <= 2048.
template <bool IN_F32>
__global__ void __launch_bounds__(256) gn_partial_kernel(const void* __restrict__ x, long long HW, int C,
                                                         long long rows_per_cta, float* __restrict__ part) {
  extern __shared__ float sm[];  // [256][16] per-thread partials, then [C][2]
  grid_dep_launch();
  grid_dep_wait();
  const int nv = C / 8;
  const int v = threadIdx.x % nv, roff = threadIdx.x / nv, rstep = 256 / nv;
  const long long r0 = blockIdx.x * rows_per_cta;
  const long long r1 = min(HW, r0 + rows_per_cta);
  const long long img = static_cast<long long>(blockIdx.y) * HW;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  for (long long r = r0 + roff; r < r1; r += rstep) {
    float val[8];
    load8<IN_F32>(x, (img + r) * nv + v, val);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += val[j];
      q[j] += val[j] * val[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm[threadIdx.x * 16 + j] = s[j];
    sm[threadIdx.x * 16 + 8 + j] = q[j];
  }
  __syncthreads();
  float* chan = sm + 256 * 16;  // [C][2]
  for (int c = threadIdx.x; c < C; c += 256) {
    const int cv = c / 8, j = c % 8;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < rstep; ++k) {  // fixed order
      a += sm[(cv + k * nv) * 16 + j];
      b += sm[(cv + k * nv) * 16 + 8 + j];
    }
    chan[2 * c] = a;
    chan[2 * c + 1] = b;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int cpg = C / 32;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < cpg; ++k) {
      a += chan[2 * (threadIdx.x * cpg + k)];
      b += chan[2 * (threadIdx.x * cpg + k) + 1];
    }
    float* o = part + ((static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * 32 + threadIdx.x) * 2;
    o[0] = a;
    o[1] = b;
  }
}

// stats: [B][32][2] = (mean, rstd)
__global__ void gn_finalize_kernel(const float* __restrict__ part, int chunks, double count, float eps,
                                   float* __restrict__ stats) {
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x, g = threadIdx.x;
  if (g >= 32) return;
  double s = 0.0, q = 0.0;
  for (int c = 0; c < chunks; ++c) {
    const float* p = part + ((static_cast<long long>(b) * chunks + c) * 32 + g) * 2;
    s += p[0];
    q += p[1];
  }
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[(b * 32 + g) * 2] = static_cast<float>(mean);
  stats[(b * 32 + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// mode 0: out bf16 = swish(gn(x) * w[c] + b[c])                       (ResBlock / norm_out, autoencoder.py:44-49,124-125)
// mode 1: out fp32 = gn(x) * gamma[b,c] + beta[b,c]                    (AdaptiveGroupNorm.forward :274-275)
template <bool IN_F32, int MODE>
__global__ void __launch_bounds__(256) gn_apply_kernel(const void* __restrict__ x, const float* __restrict__ stats,
                                                       const float* __restrict__ w, const float* __restrict__ bsh,
                                                       long long HW, int C, long long n8, void* __restrict__ out) {
  grid_dep_launch();
  grid_dep_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int nv = C / 8;
  const int cv = static_cast<int>(i % nv);
  const long long b = i / (nv * HW);
  const int cpg = C / 32;
  float v[8];
  load8<IN_F32>(x, i, v);
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const int g = c / cpg;
    const float mean = stats[(b * 32 + g) * 2], rstd = stats[(b * 32 + g) * 2 + 1];
    const float xh = (v[j] - mean) * rstd;
    if (MODE == 0) {
      const float t = xh * w[c] + bsh[c];
      y[j] = t / (1.0f + __expf(-t));
    } else {
      y[j] = w[b * C + c] * xh + bsh[b * C + c];
    }
  }
  if (MODE == 0) {
    uint4 pk;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) p2[j] = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + i * 8) = pk;
  } else {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(out) + i * 8);
    o[0] = make_float4(y[0], y[1], y[2], y[3]);
    o[1] = make_float4(y[4], y[5], y[6], y[7]);
  }
}

// AdaptiveGroupNorm parameters from the +-1 style grid z (bf16 NHWC [B, hw, Cz]):
//   s_c = sqrt(var_unbiased(z_c) + eps), m_c = mean(z_c); gamma = Linear(s), beta = Linear(m) (bf16 autocast Linears)
// One CTA per image; z is tiny (<= 128 x 128 x 32).
__global__ void __launch_bounds__(256) adagn_params_kernel(const __nv_bfloat16* __restrict__ z, int hw, int Cz,
                                                           const __nv_bfloat16* __restrict__ gw,
                                                           const __nv_bfloat16* __restrict__ gb,
                                                           const __nv_bfloat16* __restrict__ bw,
                                                           const __nv_bfloat16* __restrict__ bb, int C, float eps,
                                                           float* __restrict__ gamma, float* __restrict__ beta) {
  extern __shared__ float sm[];  // s[Cz], m[Cz]
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < Cz; c += 8) {
    double s1 = 0.0, s2 = 0.0;
    for (int p = lane; p < hw; p += 32) {
      const float v = __bfloat162float(z[(static_cast<long long>(b) * hw + p) * Cz + c]);
      s1 += v;
      s2 += static_cast<double>(v) * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      const double mean = s1 / hw;
      const double var = hw > 1 ? (s2 - hw * mean * mean) / (hw - 1) : 0.0;  // torch.var default: unbiased
      sm[c] = bf16_round(sqrtf(static_cast<float>(var) + eps));
      sm[Cz + c] = bf16_round(static_cast<float>(mean));
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ag = 0.f, ab = 0.f;
    for (int k = 0; k < Cz; ++k) {
      ag = fmaf(sm[k], __bfloat162float(gw[c * Cz + k]), ag);
      ab = fmaf(sm[Cz + k], __bfloat162float(bw[c * Cz + k]), ab);
    }
    gamma[b * C + c] = bf16_round(ag + __bfloat162float(gb[c]));
    beta[b * C + c] = bf16_round(ab + __bfloat162float(bb[c]));
  }
}

__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  grid_dep_launch();
  grid_dep_wait();
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 7 < n) {
    const float4 a = *reinterpret_cast<const float4*>(in + i), b = *reinterpret_cast<const float4*>(in + i + 4);
    uint4 pk;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
    p2[0] = __floats2bfloat162_rn(a.x, a.y);
    p2[1] = __floats2bfloat162_rn(a.z, a.w);
    p2[2] = __floats2bfloat162_rn(b.x, b.y);
    p2[3] = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4*>(out + i) = pk;
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

static bool pow2_c(int C) { return C >= 32 && C <= 2048 && (C & (C - 1)) == 0; }

}  // namespace bd

using namespace bd;

extern "C" {

int bd_nchw_to_nhwc_bf16(const float* x, void* out, int B, int C, int H, int W, int C_pad, bd_stream_t stream) {
  BD_REQUIRE(x && out && B > 0 && C > 0 && C_pad >= C && (C_pad % 8) == 0);
  const long long n = static_cast<long long>(B) * H * W;
  nchw_to_nhwc_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(out), B, C, static_cast<long long>(H) * W, C_pad);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

int bd_tokens_to_grid(const float* tokens, void* out, int B, int h, int w, int C, int ps, bd_stream_t stream) {
  BD_REQUIRE(tokens && out && B > 0 && h > 0 && w > 0 && (C % 8) == 0 && ps > 0 && (h % ps) == 0 && (w % ps) == 0);
  const long long n = static_cast<long long>(B) * h * w * (C / 8);
  tokens_to_grid_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      tokens, static_cast<__nv_bfloat16*>(out), B, h, w, C, ps);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

int bd_nhwc_to_nchw(const void* x, void* out, int out_f32, int B, int C, int H, int W, bd_stream_t stream) {
  BD_REQUIRE(x && out && B > 0 && C > 0);
  const long long n = static_cast<long long>(B) * C * H * W;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  if (out_f32)
    nhwc_to_nchw_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), out, B, C, static_cast<long long>(H) * W);
  else
    nhwc_to_nchw_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), out, B, C, static_cast<long long>(H) * W);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

int bd_cast_f32_bf16(const float* x, void* out, long long n, bd_stream_t stream) {
  BD_REQUIRE(x && out && n >= 0);
  BD_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (n == 0) return BD_OK;
  cast_kernel<<<static_cast<unsigned>((n / 8 + 256) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(out), n);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

size_t bd_groupnorm_workspace_bytes(int B, long long HW) {
  const long long chunks = (HW + 1023) / 1024;
  return static_cast<size_t>(B) * chunks * 32 * 2 * sizeof(float) + static_cast<size_t>(B) * 32 * 2 * sizeof(float);
}

int bd_groupnorm_nhwc(const void* x, int x_f32, int B, long long HW, int C, const float* weight, const float* bias,
                      int mode, void* out, void* workspace, size_t workspace_bytes, float eps, int flags,
                      bd_stream_t stream_) {
  BD_REQUIRE(x && out && weight && bias && workspace && B > 0 && HW > 0 && pow2_c(C));
  BD_REQUIRE(mode == 0 || mode == 1);
  if (workspace_bytes < bd_groupnorm_workspace_bytes(B, HW)) return BD_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const bool pdl = (flags & 1) != 0;
  const long long rows = 1024;
  const int chunks = static_cast<int>((HW + rows - 1) / rows);
  float* part = static_cast<float*>(workspace);
  float* stats = part + static_cast<size_t>(B) * chunks * 32 * 2;
  const size_t smem = (256 * 16 + 2 * C) * sizeof(float);
  {
    LaunchCfg lc(dim3(chunks, B), dim3(256), smem, st, pdl);
    if (x_f32)
      BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, gn_partial_kernel<true>, x, HW, C, rows, part));
    else
      BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, gn_partial_kernel<false>, x, HW, C, rows, part));
  }
  {
    LaunchCfg lc(dim3(B), dim3(32), 0, st, pdl);
    BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, gn_finalize_kernel, (const float*)part, chunks,
                                   static_cast<double>(HW) * (C / 32), eps, stats));
  }
  const long long n8 = static_cast<long long>(B) * HW * (C / 8);
  LaunchCfg lc(dim3(static_cast<unsigned>((n8 + 255) / 256)), dim3(256), 0, st, pdl);
  if (mode == 0) {
    if (x_f32)
      BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, gn_apply_kernel<true, 0>, x, (const float*)stats, weight, bias, HW, C, n8, out));
    else
      BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, gn_apply_kernel<false, 0>, x, (const float*)stats, weight, bias, HW, C, n8, out));
  } else {
    if (x_f32)
      BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, gn_apply_kernel<true, 1>, x, (const float*)stats, weight, bias, HW, C, n8, out));
    else
      BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, gn_apply_kernel<false, 1>, x, (const float*)stats, weight, bias, HW, C, n8, out));
  }
  return BD_OK;
}

int bd_adagn_params(const void* z, int B, int hw, int Cz, const void* gamma_w, const void* gamma_b, const void* beta_w,
                    const void* beta_b, int C, float* gamma, float* beta, bd_stream_t stream) {
  BD_REQUIRE(z && gamma_w && gamma_b && beta_w && beta_b && gamma && beta && B > 0 && hw > 0 && Cz > 0 && C > 0);
  adagn_params_kernel<<<B, 256, 2 * Cz * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(z), hw, Cz, static_cast<const __nv_bfloat16*>(gamma_w),
      static_cast<const __nv_bfloat16*>(gamma_b), static_cast<const __nv_bfloat16*>(beta_w),
      static_cast<const __nv_bfloat16*>(beta_b), C, 1e-6f, gamma, beta);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

}  // extern "C"
