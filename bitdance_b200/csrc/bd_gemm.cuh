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
    if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
      st_global_32B(o, pk[0], pk[1]);
    } else {
      reinterpret_cast<uint4*>(o)[0] = pk[0];
      reinterpret_cast<uint4*>(o)[1] = pk[1];
    }
    return;
  }
  float y[32];
  const long long mr = e.res_mod > 0 ? (m % e.res_mod) : m;
  const bool fullc = (n0 + 32 <= N);
  // vector-load the per-column operands of a full chunk (32 contiguous columns) when 16-byte aligned
  float bv[32], gv[32], rv[32];
  bool have_b = false, have_g = false, have_r = false;
  if (fullc) {
    if (e.bias && ((reinterpret_cast<uintptr_t>(e.bias + n0) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 raw = reinterpret_cast<const uint4*>(e.bias + n0)[q];
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(p2[t]);
          bv[8 * q + 2 * t] = f.x;
          bv[8 * q + 2 * t + 1] = f.y;
        }
      }
      have_b = true;
    }
    if (e.gate) {
      const __nv_bfloat16* gp = e.gate + static_cast<long long>(m) * e.ld_gate + n0;
      if ((reinterpret_cast<uintptr_t>(gp) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 raw = reinterpret_cast<const uint4*>(gp)[q];
          const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(p2[t]);
            gv[8 * q + 2 * t] = f.x;
            gv[8 * q + 2 * t + 1] = f.y;
          }
        }
        have_g = true;
      }
    }
    if (e.res) {
      if (e.res_f32) {
        const float* rp = reinterpret_cast<const float*>(e.res) + mr * e.ld_res + n0;
        if ((reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 f = reinterpret_cast<const float4*>(rp)[q];
            rv[4 * q] = f.x; rv[4 * q + 1] = f.y; rv[4 * q + 2] = f.z; rv[4 * q + 3] = f.w;
          }
          have_r = true;
        }
      } else {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(e.res) + mr * e.ld_res + n0;
        if ((reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 raw = reinterpret_cast<const uint4*>(rp)[q];
            const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(p2[t]);
              rv[8 * q + 2 * t] = f.x;
              rv[8 * q + 2 * t + 1] = f.y;
            }
          }
          have_r = true;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float v = acc[j];
    const int n = n0 + j;
    const bool ok = n < N;
    if (e.bias && ok) v += have_b ? bv[j] : __bfloat162float(e.bias[n]);
    v = bf16_round(v);
    if (e.act == kActSilu) v = bf16_round(siluf_(v));
    if (e.act == kActGeluTanh) v = bf16_round(gelu_tanhf_(v));
    if (e.gate && ok)
      v = bf16_round(v * (have_g ? gv[j] : __bfloat162float(e.gate[static_cast<long long>(m) * e.ld_gate + n])));
    if (e.res && ok) {
      if (have_r)
        v += rv[j];
      else if (e.res_f32)
        v += reinterpret_cast<const float*>(e.res)[mr * e.ld_res + n];
      else
        v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.res)[mr * e.ld_res + n]);
    }
    y[j] = v;
  }
  const bool full = (n0 + 32 <= N);
  if (e.out_f32) {
    float* o = reinterpret_cast<float*>(e.out) + static_cast<long long>(m) * e.ld_out + n0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 31) == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // whole 32-byte sectors (see st_global_32B)
        const uint4 lo = make_uint4(__float_as_uint(y[8 * j]), __float_as_uint(y[8 * j + 1]), __float_as_uint(y[8 * j + 2]),
                                    __float_as_uint(y[8 * j + 3]));
        const uint4 hi = make_uint4(__float_as_uint(y[8 * j + 4]), __float_as_uint(y[8 * j + 5]), __float_as_uint(y[8 * j + 6]),
                                    __float_as_uint(y[8 * j + 7]));
        st_global_32B(o + 8 * j, lo, hi);
      }
    } else if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
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
      if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {  // whole 32-byte sectors (see st_global_32B)
        st_global_32B(o, pk[0], pk[1]);
        st_global_32B(o + 16, pk[2], pk[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(o)[j] = pk[j];
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) o[j] = __float2bfloat16_rn(y[j]);
    }
  }
}

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 224;  // warp0 W producer, warp1 MMA issuer (+TMEM alloc), warp2 A producer, warps3-6 epilogue
constexpr int kConvThreads = 192;  // bd_conv.cuh keeps the single-producer layout

// Shared-memory plan (1 CTA per SM, ~224 KB): the W (weights, from HBM) and A (activations, L2-resident) operands have
// SEPARATE rings with their own producer warps. HBM latency under load is ~2-3 us, so what bounds a weight-streaming CTA
// is the number of W bytes it keeps in flight: the W ring takes 160 KB (10 x 16 KB tiles at BN=128), A gets 4 stages.
template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kWBytes = BN * kGemmBK * 2;
  static constexpr int kAStages = 4;
  static constexpr int kWStages = (160 * 1024) / kWBytes;  // 5 / 10 / 20 for BN = 256 / 128 / 64
  static constexpr int kBarBytes = (2 * kAStages + 2 * kWStages + 1) * 8 + 16;
  static constexpr int kSmemBytes = kAStages * kABytes + kWStages * kWBytes + 1024 /*align slack*/ + kBarBytes;
  static constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
bd_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int M, int N,
               int K, int splits, float* __restrict__ partial, GemmEpi epi, int a_hint_last, int w_tiled) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + Cfg::kWStages * Cfg::kWBytes;
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem_a + Cfg::kAStages * Cfg::kABytes);
  uint64_t* empty_w = full_w + Cfg::kWStages;
  uint64_t* full_a = empty_w + Cfg::kWStages;
  uint64_t* empty_a = full_a + Cfg::kAStages;
  uint64_t* acc_bar = empty_a + Cfg::kAStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * kGemmBM;
  const int n0 = blockIdx.y * BN;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;
  const int kb_begin = static_cast<int>((static_cast<long long>(blockIdx.z) * num_kb) / splits);
  const int kb_end = static_cast<int>((static_cast<long long>(blockIdx.z + 1) * num_kb) / splits);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < Cfg::kWStages; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&empty_w[s], 1);
    }
    for (int s = 0; s < Cfg::kAStages; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&empty_a[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2 && elect_one()) tma_prefetch_desc(&tmap_a);
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Let the next kernel in the stream begin its own prologue / weight prefetch as SMs free up.
  grid_dep_launch();

  if (warp == 0) {
    // ===================== W producer: weights never depend on the upstream kernel -> no griddepcontrol.wait =====
    // Row-major W: one strided box (BN rows x 128 B, rows K*2 bytes apart). Tile-major W (bd_pack_weight_tiles:
    // [n_tile][k_block][128][64], every 128x64 tile = 16 KB CONTIGUOUS in HBM, the k-blocks of one n_tile adjacent):
    // each CTA streams one contiguous region — DRAM-page friendly.
    if (elect_one()) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::kWStages;
        if (i >= Cfg::kWStages) mbar_wait(&empty_w[s], (static_cast<uint32_t>(i / Cfg::kWStages) & 1u) ^ 1u);
        mbar_expect_tx(&full_w[s], Cfg::kWBytes);
        uint8_t* dst = smem_w + s * Cfg::kWBytes;
        const int kb = kb_begin + i;
        if (!w_tiled) {
          tma_load_2d(dst, &tmap_w, &full_w[s], kb * kGemmBK, n0, kEvictFirst);
        } else if (BN >= 128) {
#pragma unroll
          for (int hh = 0; hh < (BN >= 128 ? BN / 128 : 1); ++hh)
            tma_load_2d(dst + hh * (128 * kGemmBK * 2), &tmap_w, &full_w[s], 0, ((n0 / 128 + hh) * num_kb + kb) * 128,
                        kEvictFirst);
        } else {
          tma_load_2d(dst, &tmap_w, &full_w[s], 0, ((n0 / 128) * num_kb + kb) * 128 + (n0 % 128), kEvictFirst);
        }
      }
    }
  } else if (warp == 2) {
    // ===================== A producer: activations come from the upstream kernel =====================
    if (elect_one()) {
      const uint64_t a_hint = a_hint_last ? kEvictLast : kEvictNormal;
      grid_dep_wait();
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::kAStages;
        if (i >= Cfg::kAStages) mbar_wait(&empty_a[s], (static_cast<uint32_t>(i / Cfg::kAStages) & 1u) ^ 1u);
        mbar_expect_tx(&full_a[s], Cfg::kABytes);
        tma_load_2d(smem_a + s * Cfg::kABytes, &tmap_a, &full_a[s], (kb_begin + i) * kGemmBK, m0, a_hint);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int sw = i % Cfg::kWStages, sa = i % Cfg::kAStages;
        mbar_wait(&full_w[sw], static_cast<uint32_t>(i / Cfg::kWStages) & 1u);
        mbar_wait(&full_a[sa], static_cast<uint32_t>(i / Cfg::kAStages) & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + sa * Cfg::kABytes);
        const uint32_t w_addr = smem_u32(smem_w + sw * Cfg::kWBytes);
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k) {
          const uint64_t ad = umma_desc_k_sw128(a_addr + k * 32);
          const uint64_t wd = umma_desc_k_sw128(w_addr + k * 32);
          umma_bf16(tmem_base, ad, wd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_w[sw]);  // frees the stages once these MMAs retire
        umma_commit(&empty_a[sa]);
      }
      umma_commit(acc_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue: warps 3..6, TMEM lane quarter = warp % 4 =====================
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
          if (nc + 32 <= N && ((reinterpret_cast<uintptr_t>(p) & 31) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // whole 32-byte sectors (see st_global_32B)
              const uint4 lo = make_uint4(__float_as_uint(acc[8 * j]), __float_as_uint(acc[8 * j + 1]),
                                          __float_as_uint(acc[8 * j + 2]), __float_as_uint(acc[8 * j + 3]));
              const uint4 hi = make_uint4(__float_as_uint(acc[8 * j + 4]), __float_as_uint(acc[8 * j + 5]),
                                          __float_as_uint(acc[8 * j + 6]), __float_as_uint(acc[8 * j + 7]));
              st_global_32B(p + 8 * j, lo, hi);
            }
          } else if (nc + 32 <= N && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
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

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant for COMPUTE-bound shapes (M >= 256 rows: batches of images, prefill, the ImageNet generator): a cluster
// of two CTAs computes a 256 x 256 tile with tcgen05.mma.cta_group::2 — each CTA stages its own 128 rows of A and HALF of
// the 256 weight rows, so a k-block costs 32 KB of shared-memory ingest per SM instead of 48 KB (the 1-CTA 128 x 256 tile is
// ingest-bound at ~60 % of the tensor peak). Tile-major W only, no split-K; the epilogue is the 1-CTA one on the CTA's own
// 128 accumulator rows.
// ---------------------------------------------------------------------------------------------------------------------
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;       // 16 KB: this CTA's 128 rows
  static constexpr int kWBytes = (BN / 2) * kGemmBK * 2;      // 16 KB: this CTA's half of the weight rows
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kStages = 6;
  static constexpr int kBarBytes = (2 * kStages + 1) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + kBarBytes;
  static constexpr uint32_t kTmemCols = 256;
};

static __global__ void __launch_bounds__(kConvThreads, 1)
bd_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int M, int N,
                int K, GemmEpi epi) {
  using Cfg = Gemm2Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_bar = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const int m0 = (blockIdx.x >> 1) * 256 + static_cast<int>(rank) * kGemmBM;
  const int n0 = blockIdx.y * Cfg::BN;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_a);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers initialised and TMEM allocated before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();

  if (warp == 0) {
    // ===== producer (each CTA): its A rows and its half of the W rows; bytes are counted on the LEADER's full barrier =====
    if (elect_one()) {
      grid_dep_wait();
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % Cfg::kStages;
        if (i >= Cfg::kStages) mbar_wait(&empty[s], (static_cast<uint32_t>(i / Cfg::kStages) & 1u) ^ 1u);
        if (rank == 0) mbar_expect_tx(&full[s], 2u * Cfg::kStageBytes);
        uint8_t* dst = smem + s * Cfg::kStageBytes;
        tma_load_2d_2sm(dst, &tmap_a, &full[s], i * kGemmBK, m0, kEvictNormal);
        // tile-major W [n_tile(128 rows)][k_block][128][64]: this CTA's half = n_tile (n0 / 128 + rank)
        tma_load_2d_2sm(dst + Cfg::kABytes, &tmap_w, &full[s], 0, ((n0 / 128 + static_cast<int>(rank)) * num_kb + i) * 128,
                        kEvictNormal);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA only =====
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, Cfg::BN);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % Cfg::kStages;
        mbar_wait(&full[s], static_cast<uint32_t>(i / Cfg::kStages) & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t w_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k)
          umma_bf16_2sm(tmem_base, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(w_addr + k * 32), idesc,
                        (i | k) != 0 ? 1u : 0u);
        umma_commit_2sm(&empty[s]);  // frees the stage in BOTH CTAs once these MMAs retire
      }
      umma_commit_2sm(acc_bar);
    }
  } else {
    // ===== epilogue: warps 2..5 of each CTA on its own 128 accumulator rows =====
    const int q = warp & 3;
    const int m = m0 + q * 32 + static_cast<int>(lane_id());
    grid_dep_wait();
    mbar_wait(acc_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < Cfg::BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
      const int nc = n0 + c * 32;
      if (m < M && nc < N) epi_apply_store(epi, acc, m, nc, N);
    }
    tc_fence_before();
  }
  cluster_sync_all();  // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
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
