// bd_quant.cu — binary quantiser, bit packing and GFQ index packing (integer/byte work: HBM-bound, bit-exact).
//
// Reference semantics:
//   VQModel.encode  modeling/vision_encoder/autoencoder.py:385-390   where(h > 0, +1, -1)  (0 -> -1, NaN -> -1)
//   GFQ.forward     imagenet_gen/src/gfq.py:221-239                   same sign rule; idx = sum_i [x_i > 0] * 2^i
//   torch.sign      modeling/t2i_pipeline.py:248                      (0 < x) - (x < 0): sign(0) = 0, sign(NaN) = 0
#include "bd_host.h"
#include "bd_ptx.cuh"

namespace bd {

// h is NCHW: for one (b, hw) the C channels are HW elements apart. A warp owns 32 consecutive hw positions of one
// image: every channel read is a coalesced 128 B (fp32) / 64 B (bf16) row segment; the 32 bits of one token are
// accumulated in a register by the lane that owns the token, so the packed store is one coalesced 128 B line.
template <bool F32>
__global__ void __launch_bounds__(256) sign_pack_nchw_kernel(const void* __restrict__ h_, int C, int HW, void* quant_,
                                                             uint32_t* __restrict__ packed,
                                                             int32_t* __restrict__ indices, int ncb,
                                                             long long BHW) {
  const int b = blockIdx.y;
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= HW) return;
  const long long img = static_cast<long long>(b) * C * HW;
  const int words = C / 32;
  const int cpg = ncb > 0 ? C / ncb : 0;  // channels per codebook group
  int32_t idx = 0;
  for (int w = 0; w * 32 < C; ++w) {
    uint32_t bits = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const int c = w * 32 + i;
      if (c >= C) break;
      const long long off = img + static_cast<long long>(c) * HW + hw;
      float v;
      if (F32)
        v = static_cast<const float*>(h_)[off];
      else
        v = __bfloat162float(static_cast<const __nv_bfloat16*>(h_)[off]);
      const bool pos = v > 0.0f;  // NaN > 0 is false -> -1, exactly like torch.where(h > 0, 1, -1)
      bits |= (pos ? 1u : 0u) << i;
      if (quant_) {
        if (F32)
          static_cast<float*>(quant_)[off] = pos ? 1.0f : -1.0f;
        else
          static_cast<__nv_bfloat16*>(quant_)[off] = __float2bfloat16_rn(pos ? 1.0f : -1.0f);
      }
      if (indices) {
        const int g = c / cpg, k = c % cpg;
        if (k == 0) idx = 0;
        idx |= (pos ? 1 : 0) << k;
        if (k == cpg - 1) indices[static_cast<long long>(g) * BHW + static_cast<long long>(b) * HW + hw] = idx;
      }
    }
    if (packed && words > 0) packed[(static_cast<long long>(b) * HW + hw) * words + w] = bits;
  }
}

// Token-major fp32 [rows, C]: one thread per 32-bit word (32 consecutive channels = 128 contiguous bytes).
__global__ void __launch_bounds__(256) sign_tokens_kernel(const float* __restrict__ x, long long nwords,
                                                          float* __restrict__ tokens, uint32_t* __restrict__ packed) {
  const long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  const float4* src = reinterpret_cast<const float4*>(x + w * 32);
  float4* dst = tokens ? reinterpret_cast<float4*>(tokens + w * 32) : nullptr;
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = src[j];
    const float a[4] = {v.x, v.y, v.z, v.w};
    float s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // torch.sign computes (0 < x) - (x < 0): zero AND NaN map to 0
      s[i] = static_cast<float>((a[i] > 0.f) - (a[i] < 0.f));
      bits |= (a[i] > 0.f ? 1u : 0u) << (4 * j + i);
    }
    if (dst) dst[j] = make_float4(s[0], s[1], s[2], s[3]);
  }
  if (packed) packed[w] = bits;
}

// AR-step variant: x fp32 [B*pn, C] -> tokens fp32 scattered into a [B, rows_per_image, C] grid at row offset row0
// (out_tokens.append + torch.cat, t2i_pipeline.py:250,270), bf16 copies for `dup` sequence groups (the cond | uncond
// rows of curr_tokens feed MLPconnector identically), and packed bits.
__global__ void __launch_bounds__(256) sign_tokens_ex_kernel(const float* __restrict__ x, int B, int pn, int C,
                                                             float* __restrict__ grid, long long rows_per_image,
                                                             long long row0, __nv_bfloat16* __restrict__ tb, int dup,
                                                             uint32_t* __restrict__ packed) {
  const int wpr = C / 32;
  const long long nwords = static_cast<long long>(B) * pn * wpr;
  const long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  const long long r = w / wpr;
  const int cw = static_cast<int>(w % wpr);
  const int b = static_cast<int>(r / pn), s = static_cast<int>(r % pn);
  const float4* src = reinterpret_cast<const float4*>(x + w * 32);
  float4* dst = grid ? reinterpret_cast<float4*>(grid + ((b * rows_per_image + row0 + s) * C + cw * 32)) : nullptr;
  uint32_t bits = 0;
  float sv[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = src[j];
    const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sv[4 * j + i] = static_cast<float>((a[i] > 0.f) - (a[i] < 0.f));
      bits |= (a[i] > 0.f ? 1u : 0u) << (4 * j + i);
    }
    if (dst) dst[j] = make_float4(sv[4 * j], sv[4 * j + 1], sv[4 * j + 2], sv[4 * j + 3]);
  }
  if (packed) packed[(b * rows_per_image + row0 + s) * wpr + cw] = bits;
  if (tb) {
    uint4 pk[4];
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int j = 0; j < 16; ++j) p2[j] = __floats2bfloat162_rn(sv[2 * j], sv[2 * j + 1]);
    for (int d = 0; d < dup; ++d) {
      uint4* o = reinterpret_cast<uint4*>(tb + ((static_cast<long long>(d) * B * pn + r) * C + cw * 32));
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = pk[j];
    }
  }
}

template <bool F32>
__global__ void __launch_bounds__(256) unpack_tokens_kernel(const uint32_t* __restrict__ packed, long long nwords,
                                                            void* __restrict__ out) {
  const long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  const uint32_t bits = packed[w];
  if (F32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out) + w * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] = ((bits >> (4 * j + i)) & 1u) ? 1.0f : -1.0f;
      dst[j] = make_float4(s[0], s[1], s[2], s[3]);
    }
  } else {
    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + w * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t lo = ((bits >> (8 * j + 2 * i)) & 1u) ? 0x3F80u : 0xBF80u;      // +-1.0 in bf16
        const uint32_t hi = ((bits >> (8 * j + 2 * i + 1)) & 1u) ? 0x3F80u : 0xBF80u;
        u[i] = lo | (hi << 16);
      }
      dst[j] = make_uint4(u[0], u[1], u[2], u[3]);
    }
  }
}

}  // namespace bd

using namespace bd;

extern "C" {

int bd_sign_pack_nchw(const void* h, int h_f32, int B, int C, int HW, void* quant, uint32_t* packed,
                      int32_t* indices, int num_codebooks, bd_stream_t stream) {
  BD_REQUIRE(h && B >= 0 && C > 0 && HW >= 0);
  if (B == 0 || HW == 0) return BD_OK;
  BD_REQUIRE(!packed || (C % 32) == 0);
  if (indices) BD_REQUIRE(num_codebooks > 0 && (C % num_codebooks) == 0 && (C / num_codebooks) <= 31);
  dim3 grid((HW + 255) / 256, B);
  const long long BHW = static_cast<long long>(B) * HW;
  if (h_f32)
    sign_pack_nchw_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(h, C, HW, quant, packed, indices,
                                                                                      indices ? num_codebooks : 0, BHW);
  else
    sign_pack_nchw_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(h, C, HW, quant, packed, indices,
                                                                                       indices ? num_codebooks : 0, BHW);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

int bd_sign_tokens(const float* x, long long rows, int C, float* tokens, uint32_t* packed, bd_stream_t stream) {
  BD_REQUIRE(x && rows >= 0 && C > 0 && (C % 32) == 0);
  BD_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(tokens) & 15) == 0);
  const long long nwords = rows * (C / 32);
  if (nwords == 0) return BD_OK;
  sign_tokens_kernel<<<static_cast<unsigned>((nwords + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, nwords, tokens, packed);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

int bd_sign_tokens_ex(const float* x, int B, int pn, int C, float* grid, long long rows_per_image, long long row0,
                      void* tokens_bf16, int dup, uint32_t* packed, bd_stream_t stream) {
  BD_REQUIRE(x && B > 0 && pn > 0 && C > 0 && (C % 32) == 0 && dup >= 0 && row0 >= 0 && rows_per_image >= row0 + pn);
  BD_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(grid) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(tokens_bf16) & 15) == 0);
  const long long nwords = static_cast<long long>(B) * pn * (C / 32);
  sign_tokens_ex_kernel<<<static_cast<unsigned>((nwords + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, B, pn, C, grid, rows_per_image, row0, static_cast<__nv_bfloat16*>(tokens_bf16), dup, packed);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

int bd_unpack_tokens(const uint32_t* packed, long long rows, int C, void* out, int out_f32, bd_stream_t stream) {
  BD_REQUIRE(packed && out && rows >= 0 && C > 0 && (C % 32) == 0);
  BD_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const long long nwords = rows * (C / 32);
  if (nwords == 0) return BD_OK;
  const unsigned grid = static_cast<unsigned>((nwords + 255) / 256);
  if (out_f32)
    unpack_tokens_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, nwords, out);
  else
    unpack_tokens_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, nwords, out);
  BD_LAUNCH_CHECK();
  return BD_OK;
}

}  // extern "C"
