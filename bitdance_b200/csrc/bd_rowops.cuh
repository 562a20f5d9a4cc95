// bd_rowops.cuh — row-wise fused kernels that close a split-K GEMM: deterministic reduction of the fp32 partials,
// the Linear's epilogue (bias / gate / residual with the autocast rounding points) AND the normalisation that follows
// it in the network, in one pass over the row (one CTA per token row, the row lives in registers).
//   head:  h = bf16(res + bf16(bf16(acc + bias) * gate));  a = bf16(LN(h)(*w+b) * bf16(1 + scale) + shift)
//          (TransBlock.forward flow_head_parallel_x.py:242-252: the x + h*gate of one sub-block and the
//           norm*(1+scale)+shift of the next; FinalLayer.forward :169-173 when w == nullptr)
//   llm:   hidden = res + bf16(acc) (fp32 stream) | bf16(res + bf16(acc)) (bf16 stream);  a = RMSNorm(hidden) as bf16
//          (Qwen3DecoderLayer: residual add, then the next RMSNorm)
#pragma once
#include "bd_host.h"
#include "bd_ptx.cuh"

namespace bd {

constexpr int kRowThreads = 256;
constexpr int kRowMaxVec = 3;  // D <= 256 * 3 * 8 = 6144

__device__ __forceinline__ float row_block_sum(float v, float* red) {
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

__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* q = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(q[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 pk;
  __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int j = 0; j < 4; ++j) q[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = pk;
}

struct HeadRowArgs {
  const float* partial;  // [S][M][D]
  int splits, M, D;
  const __nv_bfloat16* bias;   // [D]
  const __nv_bfloat16* gate;   // [M, ld_mod]
  __nv_bfloat16* h;            // [M, D] residual stream, in/out
  const float* ln_w;           // [D] or nullptr (no affine)
  const float* ln_b;
  const __nv_bfloat16* scale;  // [M, ld_mod]
  const __nv_bfloat16* shift;
  long long ld_mod;
  __nv_bfloat16* a;            // [M, D] LN-modulated output
  float eps;
};

static __global__ void __launch_bounds__(kRowThreads) head_splitk_row_kernel(HeadRowArgs p) {
  __shared__ float red[32];
  grid_dep_launch();
  grid_dep_wait();
  const long long m = blockIdx.x;
  const int nvec = p.D / 8;
  float v[kRowMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kRowMaxVec; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nvec) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int s = 0; s < p.splits; ++s) {  // fixed order: deterministic
        const float4* q = reinterpret_cast<const float4*>(p.partial + (static_cast<long long>(s) * p.M + m) * p.D + c * 8);
        const float4 x0 = q[0], x1 = q[1];
        acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w;
        acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
      }
      float b[8], g[8], r[8];
      load_bf16x8(p.bias + c * 8, b);
      load_bf16x8(p.gate + m * p.ld_mod + c * 8, g);
      load_bf16x8(p.h + m * p.D + c * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = bf16_round(bf16_round(acc[j] + b[j]) * g[j]);
        v[i][j] = bf16_round(r[j] + y);
        sum += v[i][j];
      }
      store_bf16x8(p.h + m * p.D + c * 8, v[i]);
    }
  }
  const float mean = row_block_sum(sum, red) / static_cast<float>(p.D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kRowMaxVec; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(row_block_sum(sq, red) / static_cast<float>(p.D) + p.eps);
#pragma unroll
  for (int i = 0; i < kRowMaxVec; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nvec) {
      float sc[8], sh[8], o[8];
      load_bf16x8(p.scale + m * p.ld_mod + c * 8, sc);
      load_bf16x8(p.shift + m * p.ld_mod + c * 8, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float hn = (v[i][j] - mean) * rstd;
        if (p.ln_w) hn = hn * p.ln_w[c * 8 + j] + p.ln_b[c * 8 + j];
        o[j] = hn * bf16_round(1.0f + sc[j]) + sh[j];
      }
      store_bf16x8(p.a + m * p.D + c * 8, o);
    }
  }
}

struct LlmRowArgs {
  const float* partial;  // [S][M][D]
  int splits, M, D;
  void* hidden;          // [M, D] residual stream (fp32 or bf16), in/out
  int stream_f32;
  const __nv_bfloat16* norm_w;  // [D] RMSNorm weight of the NEXT norm, or nullptr (no norm: only the residual add)
  __nv_bfloat16* a;             // [M, D] bf16 normalised output (GEMM operand)
  float eps;
};

static __global__ void __launch_bounds__(kRowThreads) llm_splitk_row_kernel(LlmRowArgs p) {
  __shared__ float red[32];
  grid_dep_launch();
  grid_dep_wait();
  const long long m = blockIdx.x;
  const int nvec = p.D / 8;
  float v[kRowMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kRowMaxVec; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nvec) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int s = 0; s < p.splits; ++s) {
        const float4* q = reinterpret_cast<const float4*>(p.partial + (static_cast<long long>(s) * p.M + m) * p.D + c * 8);
        const float4 x0 = q[0], x1 = q[1];
        acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w;
        acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
      }
      if (p.stream_f32) {
        float* hp = static_cast<float*>(p.hidden) + m * p.D + c * 8;
        const float4 r0 = reinterpret_cast<const float4*>(hp)[0], r1 = reinterpret_cast<const float4*>(hp)[1];
        const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = r[j] + bf16_round(acc[j]);
        reinterpret_cast<float4*>(hp)[0] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        reinterpret_cast<float4*>(hp)[1] = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
      } else {
        __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(p.hidden) + m * p.D + c * 8;
        float r[8];
        load_bf16x8(hp, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = bf16_round(r[j] + bf16_round(acc[j]));
        store_bf16x8(hp, v[i]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
    }
  }
  if (!p.norm_w) return;
  const float rstd = rsqrtf(row_block_sum(ss, red) / static_cast<float>(p.D) + p.eps);
#pragma unroll
  for (int i = 0; i < kRowMaxVec; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nvec) {
      float w[8], o[8];
      load_bf16x8(p.norm_w + c * 8, w);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = p.stream_f32 ? w[j] * (v[i][j] * rstd) : bf16_round(w[j] * bf16_round(v[i][j] * rstd));
      store_bf16x8(p.a + m * p.D + c * 8, o);
    }
  }
}

}  // namespace bd
