// bd_gemm.cuh — weight-streaming bf16 GEMM for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T  (+ fused epilogue)
//
// This is the kernel that bounds the BitDance AR step: at batch 1 every Linear of the diffusion head
// (reference modeling/vision_head/flow_head_parallel_x.py:133-137,189-190,239-240,277-279) and of the
// Qwen3 decoder runs with M = R*parallel_num = 128 rows, so the work is streaming W from HBM exactly once.
//
//   * A (activations, [M,K] row-major bf16) and W (nn.Linear weight, [N,K] row-major bf16) are both K-major,
//     so both are fetched by TMA (128B swizzle, OOB rows/cols zero-filled) straight into the UMMA canonical
//     layout; no prepack of W is needed (SwiGLU pairs are the exception: see bd_interleave16).
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM accumulator;
//   * warp roles: warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer, warps2-5 = epilogue (TMEM -> regs -> HBM);
//   * split-K over blockIdx.z writes fp32 partials that bd_splitk_epilogue_kernel reduces in a fixed order
//     (deterministic), applying the same epilogue;
//   * PDL: weights do not depend on the upstream kernel, so the producer fills the whole smem ring with W tiles
//     BEFORE griddepcontrol.wait, keeping HBM busy across kernel boundaries.
#pragma once
#include <cstdio>
#include "bd_host.h"
#include "bd_ptx.cuh"

namespace bd {

// Apply the epilogue to 32 consecutive accumulator columns [n0, n0+32) of row m.
__device__ __forceinline__ void epi_apply_store(const GemmEpi& e, const float (&acc)[32], int m, int n0, int N) {
  if (e.swiglu) {
    // columns n0..n0+15 = gate features, n0+16..n0+31 = up features of outputs (n0/2 .. n0/2+15)
    const int o0 = n0 >> 1;
    float y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float g = acc[j], u = acc[16 + j];
      if (e.bias) {
        g += __bfloat162float(e.bias[n0 + j]);
        u += __bfloat162float(e.bias[n0 + 16 + j]);
      }
      g = bf16_round(g);
      u = bf16_round(u);
      y[j] = bf16_round(bf16_round(siluf_(g)) * u);
    }
    if (n0 + 32 > N) return;  // N is a multiple of 32 for swiglu weights (checked on the host)
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + static_cast<long long>(m) * e.ld_out + o0;
    uint4 pk[2];
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int j = 0; j < 8; ++j) p2[j] = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
    reinterpret_cast<uint4*>(o)[0] = pk[0];
    reinterpret_cast<uint4*>(o)[1] = pk[1];
    return;
  }
  float y[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float v = acc[j];
    const int n = n0 + j;
    const bool ok = n < N;
    if (e.bias && ok) v += __bfloat162float(e.bias[n]);
    v = bf16_round(v);
    if (e.act == kActSilu) v = bf16_round(siluf_(v));
    if (e.act == kActGeluTanh) v = bf16_round(gelu_tanhf_(v));
    if (e.gate && ok) v = bf16_round(v * __bfloat162float(e.gate[static_cast<long long>(m) * e.ld_gate + n]));
    if (e.res && ok) {
      const long long mr = e.res_mod > 0 ? (m % e.res_mod) : m;
      if (e.res_f32)
        v += reinterpret_cast<const float*>(e.res)[mr * e.ld_res + n];
      else
        v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.res)[mr * e.ld_res + n]);
    }
    y[j] = v;
  }
  const bool full = (n0 + 32 <= N);
  if (e.out_f32) {
    float* o = reinterpret_cast<float*>(e.out) + static_cast<long long>(m) * e.ld_out + n0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) o[j] = y[j];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + static_cast<long long>(m) * e.ld_out + n0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      uint4 pk[4];
      __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
      for (int j = 0; j < 16; ++j) p2[j] = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(o)[j] = pk[j];
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) o[j] = __float2bfloat16_rn(y[j]);
    }
  }
}

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kWBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
bd_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int M, int N,
               int K, int splits, float* __restrict__ partial, GemmEpi epi, int a_hint_last, int w_tiled) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* acc_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * kGemmBM;
  const int n0 = blockIdx.y * BN;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;
  const int kb_begin = static_cast<int>((static_cast<long long>(blockIdx.z) * num_kb) / splits);
  const int kb_end = static_cast<int>((static_cast<long long>(blockIdx.z + 1) * num_kb) / splits);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Let the next kernel in the stream begin its own prologue / weight prefetch as SMs free up.
  grid_dep_launch();

  // W tile of k-block kb into stage memory. Row-major W: one strided box (BN rows x 128 B, rows K*2 bytes apart).
  // Tile-major W (bd_pack_weight_tiles: [n_tile][k_block][128][64], every 128x64 tile = 16 KB CONTIGUOUS in HBM, the
  // k-blocks of one n_tile adjacent): each CTA streams one contiguous region — DRAM-page friendly.
  auto load_w = [&](uint8_t* dst, uint64_t* bar, int kb) {
    if (!w_tiled) {
      tma_load_2d(dst, &tmap_w, bar, kb * kGemmBK, n0, kEvictFirst);
    } else if (BN >= 128) {
#pragma unroll
      for (int hh = 0; hh < BN / 128; ++hh)
        tma_load_2d(dst + hh * (128 * kGemmBK * 2), &tmap_w, bar, 0, ((n0 / 128 + hh) * num_kb + kb) * 128, kEvictFirst);
    } else {
      tma_load_2d(dst, &tmap_w, bar, 0, ((n0 / 128) * num_kb + kb) * 128 + (n0 % 128), kEvictFirst);
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint64_t a_hint = a_hint_last ? kEvictLast : kEvictNormal;
      // Phase 1: weights only (independent of the upstream kernel) for the first ring of stages.
      const int pre = nkb < Cfg::kStages ? nkb : Cfg::kStages;
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&full_bar[i], Cfg::kStageBytes);
        load_w(smem + i * Cfg::kStageBytes + Cfg::kABytes, &full_bar[i], kb_begin + i);
      }
      grid_dep_wait();  // activations are produced by the upstream kernel
      for (int i = 0; i < pre; ++i)
        tma_load_2d(smem + i * Cfg::kStageBytes, &tmap_a, &full_bar[i], (kb_begin + i) * kGemmBK, m0, a_hint);
      // Phase 2: steady state.
      for (int i = pre; i < nkb; ++i) {
        const int s = i % Cfg::kStages;
        const uint32_t ph = static_cast<uint32_t>(i / Cfg::kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
        load_w(smem + s * Cfg::kStageBytes + Cfg::kABytes, &full_bar[s], kb_begin + i);
        tma_load_2d(smem + s * Cfg::kStageBytes, &tmap_a, &full_bar[s], (kb_begin + i) * kGemmBK, m0, a_hint);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::kStages;
        const uint32_t ph = static_cast<uint32_t>(i / Cfg::kStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t w_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k) {
          const uint64_t ad = umma_desc_k_sw128(a_addr + k * 32);
          const uint64_t wd = umma_desc_k_sw128(w_addr + k * 32);
          umma_bf16(tmem_base, ad, wd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
      }
      umma_commit(acc_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====================
    const int q = warp & 3;
    const int m = m0 + q * 32 + static_cast<int>(lane_id());
    grid_dep_wait();  // res / gate come from upstream kernels; out may still be read by them
    if (nkb > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float acc[32];
      if (nkb > 0) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      }
      const int nc = n0 + c * 32;
      if (m < M && nc < N) {
        if (splits > 1) {
          float* p = partial + (static_cast<long long>(blockIdx.z) * M + m) * N + nc;
          if (nc + 32 <= N && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(p)[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (nc + j < N) p[j] = acc[j];
          }
        } else {
          epi_apply_store(epi, acc, m, nc, N);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Deterministic split-K reduction + epilogue: one thread per (row, 32-column chunk).
static __global__ void __launch_bounds__(256) bd_splitk_epilogue_kernel(const float* __restrict__ partial, int M, int N,
                                                                  int splits, GemmEpi epi) {
  grid_dep_launch();
  grid_dep_wait();
  const int chunks = (N + 31) / 32;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(M) * chunks) return;
  // consecutive threads take consecutive rows of the same chunk?  No: consecutive chunks of the same row, so
  // that a warp reads 32 x 128 B = 4 KB contiguous per split.
  const int m = static_cast<int>(idx / chunks);
  const int nc = static_cast<int>(idx % chunks) * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* p = partial + (static_cast<long long>(s) * M + m) * N + nc;
    if (nc + 32 <= N && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = reinterpret_cast<const float4*>(p)[j];
        acc[4 * j] += v.x;
        acc[4 * j + 1] += v.y;
        acc[4 * j + 2] += v.z;
        acc[4 * j + 3] += v.w;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (nc + j < N) acc[j] += p[j];
    }
  }
  epi_apply_store(epi, acc, m, nc, N);
}

}  // namespace bd
