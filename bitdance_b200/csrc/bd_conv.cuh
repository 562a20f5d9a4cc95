// bd_conv.cuh — NHWC bf16 convolution as an implicit GEMM on tcgen05 (no im2col buffer).
//
// The tokenizer's conv blocks (reference modeling/vision_encoder/autoencoder.py:31-38,72-77,94,105,142,170,240):
//   out[b,y,x,:] = sum_taps  A[b, y+dy, x+dx, :] · W[:, tap, :]^T
// One CTA computes 128 output pixels (a TW x TH patch, TW*TH = 128) x BN output channels. For every tap the A operand
// is fetched by ONE 4-D TMA box {64 ch, TW, TH, 1} at the shifted coordinate: out-of-image pixels (the zero padding)
// and channels beyond Cin are zero-filled by the TMA unit, and the box lands in shared memory already in the UMMA
// K-major / 128B-swizzle layout (one pixel = one 128-byte row). Weights are prepacked once as [Cout, taps*Cin_pad]
// (k = tap*Cin_pad + c). Stride-2 convolutions read a 4-phase de-interleaved copy of the input (tap -> phase, shift).
// Epilogues: bias, residual add (bf16/fp32 stream), NHWC bf16/fp32 store, depth-to-space scatter (Upsampler,
// autoencoder.py:198-249), NCHW store for the 3-channel image.
#pragma once
#include "bd_gemm.cuh"

namespace bd {

struct ConvGeom {
  int B, H, W;      // output pixels (also the extent of each input phase plane)
  int Cout;
  int Cin_pad;      // weight K per tap
  int cblocks;      // ceil(Cin / 64)
  int taps;
  int TW, TH;       // tile shape, TW * TH == 128
  int tiles_x, tiles_y;
  signed char dx[9], dy[9], plane[9];  // per tap: input shift and phase plane (plane index p reads image p*B + b)
  int d2s;          // depth-to-space(2) scatter: out is [B, 2H, 2W, Cout/4]
  int nchw_out;     // out is [B, Cout, H, W]
};

constexpr int kConvStages = 3;
constexpr int kConvBN = 128;
constexpr int kConvStageBytes = (128 + kConvBN) * 64 * 2;
constexpr int kConvSmemBytes = kConvStages * kConvStageBytes + 1024 + 256;

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 2)
bd_conv_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, ConvGeom g,
               GemmEpi epi) {
  constexpr int kABytes = 128 * 64 * 2;
  constexpr int kStageBytes = kABytes + BN * 64 * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kConvStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kConvStages;
  uint64_t* acc_bar = empty_bar + kConvStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  int tile = blockIdx.x;
  const int tx_i = tile % g.tiles_x;
  tile /= g.tiles_x;
  const int ty_i = tile % g.tiles_y;
  const int b = tile / g.tiles_y;
  const int x0 = tx_i * g.TW, y0 = ty_i * g.TH;
  const int n0 = blockIdx.y * BN;
  const int nkb = g.taps * g.cblocks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < kConvStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();

  if (warp == 0) {
    if (elect_one()) {
      grid_dep_wait();
      for (int i = 0; i < nkb; ++i) {
        const int s = i % kConvStages;
        const uint32_t ph = static_cast<uint32_t>(i / kConvStages) & 1u;
        if (i >= kConvStages) mbar_wait(&empty_bar[s], ph ^ 1u);
        const int tap = i / g.cblocks, cb = i % g.cblocks;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        tma_load_4d(smem + s * kStageBytes, &tmap_a, &full_bar[s], cb * 64, x0 + g.dx[tap], y0 + g.dy[tap],
                    g.plane[tap] * g.B + b, kEvictNormal);
        tma_load_2d(smem + s * kStageBytes + kABytes, &tmap_w, &full_bar[s], tap * g.Cin_pad + cb * 64, n0,
                    kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % kConvStages;
        const uint32_t ph = static_cast<uint32_t>(i / kConvStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
        const uint32_t w_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(w_addr + k * 32), idesc,
                    (i | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(acc_bar);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + static_cast<int>(lane_id());  // pixel within the tile
    const int x = x0 + r % g.TW, y = y0 + r / g.TW;
    const bool valid = (x < g.W) && (y < g.H);
    const long long m = (static_cast<long long>(b) * g.H + y) * g.W + x;
    grid_dep_wait();
    mbar_wait(acc_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
      const int nc = n0 + c * 32;
      if (!valid || nc >= g.Cout) continue;
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
      if (g.nchw_out) {
        for (int j = 0; j < 32 && nc + j < g.Cout; ++j) {
          float o = acc[j];
          if (epi.bias) o += __bfloat162float(epi.bias[nc + j]);
          o = bf16_round(o);
          const long long idx = ((static_cast<long long>(b) * g.Cout + nc + j) * g.H + y) * g.W + x;
          if (epi.out_f32)
            reinterpret_cast<float*>(epi.out)[idx] = o;
          else
            reinterpret_cast<__nv_bfloat16*>(epi.out)[idx] = __float2bfloat16_rn(o);
        }
      } else if (g.d2s) {
        // conv channel n = (2i + j) * Cq + c'  ->  out[b, 2y+i, 2x+j, c']
        const int Cq = g.Cout >> 2;
        const int ij = nc / Cq, cq = nc % Cq;
        const long long m2 = (static_cast<long long>(b) * (2 * g.H) + 2 * y + (ij >> 1)) * (2 * g.W) + 2 * x + (ij & 1);
        float yv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float o = acc[j];
          if (epi.bias) o += __bfloat162float(epi.bias[nc + j]);
          yv[j] = o;
        }
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(epi.out) + m2 * Cq + cq;
        uint4 pk[4];
        __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
        for (int j = 0; j < 16; ++j) p2[j] = __floats2bfloat162_rn(yv[2 * j], yv[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(o)[j] = pk[j];
      } else {
        epi_apply_store(epi, acc, static_cast<int>(m), nc, g.Cout);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

}  // namespace bd
