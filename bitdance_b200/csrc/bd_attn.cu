// bd_attn.cu — block-bidirectional / causal attention over a paged (or strided) bf16 KV, flash-style.
//
// Serves both attention sites of the AR step:
//   * diffusion head  Attention.forward flow_head_parallel_x.py:192-220 (non-causal MHA over the parallel_num
//     tokens of one block; flash_attn_func semantics: fp32 scores/softmax, P and O in bf16);
//   * Qwen3Attention (transformers qwen3 modeling; call sites modeling/t2i_pipeline.py:199-266): GQA, the
//     parallel_num new tokens attend to the whole cache and to each other (all-ones mask) or causally (prefill).
//
// Math: tensor cores via mma.sync.m16n8k16 bf16 (legacy HMMA path). The op is <5% of the step's bytes/flops at
// batch 1; a tcgen05 version with the 5 GQA heads stacked on M is the planned upgrade.
// One CTA = 4 warps = 64 query rows of one (batch, q-head); KV walked in tiles of 64 keys; optional split of the
// KV range over blockIdx.z with a fixed-order combine (deterministic).
#include <cfloat>
#include "bd_host.h"
#include "bd_ptx.cuh"

namespace bd {

// byte offset of element (row, col) of a blocked bf16 activation (same rule as bd_stream.cuh::blk_off)
__device__ __forceinline__ long long attn_blk_off(int row, int col) {
  return (static_cast<long long>(col >> 6) << 14) + (row << 7) + ((((col >> 3) & 7) ^ (row & 7)) << 4) + ((col & 7) << 1);
}

struct AttnParams {
  const __nv_bfloat16* q;  // element (b, s, h, d) at q + b*q_sb + s*q_ss + h*q_sh + d
  long long q_sb, q_ss, q_sh;
  // strided KV (paged == 0): element (b, s, hk, d) at k + b*k_sb + s*k_ss + hk*k_sh + d
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  long long k_sb, k_ss, k_sh;
  // paged KV (paged == 1): page p of sequence b is pool page page_table[b*max_pages + p]; a page holds 64 tokens:
  // element (page, hk, t, d) at pool + ((page*Hkv + hk)*64 + t)*D + d
  const int* page_table;
  int max_pages;
  int paged;
  __nv_bfloat16* out;  // (b, s, h, d) at out + b*o_sb + s*o_ss + h*o_sh + d
  long long o_sb, o_ss, o_sh;
  int out_blocked;     // 1: out is the blocked [Hq*hd/64][128][64] swizzled image of rows b*Sq+s (operand of the stream engine)
  float* part_o;  // [splits][B][Hq][Sq][D] fp32 (unnormalised) when splits > 1
  float* part_ml; // [splits][B][Hq][Sq][2] (max, sum)
  int B, Sq, Sk, Hq, Hkv;
  const int* sk_dev;  // optional per-sequence key counts (device): Sk_b = sk_dev[b] + sk_add; Sk = planning bound
  int sk_add;
  int causal;      // key j visible to query i iff j <= i + (Sk - Sq)
  int splits;
  float scale_log2;  // softmax scale * log2(e)
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Shared tile: 64 rows x HD bf16, 16-byte chunks XOR-swizzled by (row & 7) -> conflict-free ldmatrix.
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * (HD * 2) + ((chunk ^ (row & 7)) << 4));
}

// Cooperative 64-row tile load (128 threads), rows >= valid zero-filled. src row r at src + r*row_stride.
template <int HD>
__device__ __forceinline__ void load_tile(uint8_t* smem_tile, const __nv_bfloat16* src, long long row_stride,
                                          int valid_rows) {
  constexpr int kChunks = HD / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < 64 * kChunks; i += 128) {
    const int r = i / kChunks, c = i % kChunks;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < valid_rows) v = *reinterpret_cast<const uint4*>(src + r * row_stride + c * 8);
    *reinterpret_cast<uint4*>(smem_tile + tile_off<HD>(r, c)) = v;
  }
}

template <int HD>
__global__ void __launch_bounds__(128) bd_attn_kernel(AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 64 * HD * 2;
  uint8_t* sV = sK + 64 * HD * 2;
  grid_dep_launch();
  grid_dep_wait();

  const int qt = blockIdx.x;  // query tile
  const int h = blockIdx.y % p.Hq;
  const int b = blockIdx.y / p.Hq;
  const int split = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = qt * 64;
  const int q_valid = min(64, p.Sq - q0);
  const int Sk = p.sk_dev ? p.sk_dev[b] + p.sk_add : p.Sk;

  load_tile<HD>(sQ, p.q + b * p.q_sb + static_cast<long long>(q0) * p.q_ss + h * p.q_sh, p.q_ss, q_valid);

  // KV tile range of this split
  const int n_tiles = (Sk + 63) / 64;
  const int t_begin = static_cast<int>((static_cast<long long>(split) * n_tiles) / p.splits);
  const int t_end = static_cast<int>((static_cast<long long>(split + 1) * n_tiles) / p.splits);

  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-FLT_MAX, -FLT_MAX};
  float l_run[2] = {0.f, 0.f};
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  const int causal_off = Sk - p.Sq;

  for (int kt = t_begin; kt < t_end; ++kt) {
    const int k0 = kt * 64;
    const int k_valid = min(64, Sk - k0);
    __syncthreads();  // previous tile fully consumed (also orders the Q store on the first iteration)
    if (p.paged) {
      const int page = p.page_table[b * p.max_pages + kt];
      const long long base = (static_cast<long long>(page) * p.Hkv + hk) * 64 * HD;
      load_tile<HD>(sK, p.k + base, HD, k_valid);
      load_tile<HD>(sV, p.v + base, HD, k_valid);
    } else {
      load_tile<HD>(sK, p.k + b * p.k_sb + static_cast<long long>(k0) * p.k_ss + hk * p.k_sh, p.k_ss, k_valid);
      load_tile<HD>(sV, p.v + b * p.k_sb + static_cast<long long>(k0) * p.k_ss + hk * p.k_sh, p.k_ss, k_valid);
    }
    __syncthreads();

    // ---- S = Q K^T : per warp 16 x 64 ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ks += 2) {
      uint32_t a0[4], a1[4];
      ldmatrix_x4(a0, smem_u32(sQ + tile_off<HD>(warp * 16 + (lane & 15), 2 * ks + (lane >> 4))));
      ldmatrix_x4(a1, smem_u32(sQ + tile_off<HD>(warp * 16 + (lane & 15), 2 * ks + 2 + (lane >> 4))));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t bk[4];  // keys 8j..8j+7: (b0,b1) for k-step ks, (b0,b1) for k-step ks+1
        ldmatrix_x4(bk, smem_u32(sK + tile_off<HD>(8 * j + (lane & 7), 2 * ks + (lane >> 3))));
        mma_bf16_16816(s[j], a0, bk[0], bk[1]);
        mma_bf16_16816(s[j], a1, bk[2], bk[3]);
      }
    }
    // ---- mask + online softmax (base-2 domain) ----
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int key = k0 + 8 * j + 2 * t + (i & 1);
        const int r = i >> 1;
        bool ok = key < Sk;
        if (p.causal) ok = ok && (key <= qrow[r] + causal_off);
        const float v = ok ? s[j][i] * p.scale_log2 : -FLT_MAX;
        s[j][i] = v;
        m_new[r] = fmaxf(m_new[r], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], l_add[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) corr[r] = exp2f(m_run[r] - m_new[r]);
    uint32_t pa[4][4];  // P as A fragments: 4 k-steps (16 keys each)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float e[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i >> 1;
        // fully-masked rows keep m = -FLT_MAX: force p = 0 there
        e[i] = (s[j][i] == -FLT_MAX) ? 0.f : exp2f(s[j][i] - m_new[r]);
        l_add[r] += e[i];
      }
      const int kk = j >> 1;
      if ((j & 1) == 0) {
        pa[kk][0] = pack_bf16(e[0], e[1]);
        pa[kk][1] = pack_bf16(e[2], e[3]);
      } else {
        pa[kk][2] = pack_bf16(e[0], e[1]);
        pa[kk][3] = pack_bf16(e[2], e[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] = l_run[r] * corr[r] + l_add[r];
      m_run[r] = m_new[r];
    }
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      o_acc[n][0] *= corr[0];
      o_acc[n][1] *= corr[0];
      o_acc[n][2] *= corr[1];
      o_acc[n][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int n = 0; n < HD / 8; n += 2) {
        uint32_t bv[4];  // features chunk n: (b0,b1); chunk n+1: (b0,b1); keys 16kk..16kk+15
        ldmatrix_x4_trans(bv, smem_u32(sV + tile_off<HD>(16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1), n + (lane >> 4))));
        mma_bf16_16816(o_acc[n], pa[kk], bv[0], bv[1]);
        mma_bf16_16816(o_acc[n + 1], pa[kk], bv[2], bv[3]);
      }
    }
  }
  // ---- finalize ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  if (p.splits == 1) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (qrow[r] >= p.Sq) continue;
      const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
      __nv_bfloat16* o = p.out + b * p.o_sb + static_cast<long long>(qrow[r]) * p.o_ss + h * p.o_sh;
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) {
        const uint32_t pk = pack_bf16(o_acc[n][2 * r] * inv, o_acc[n][2 * r + 1] * inv);
        if (p.out_blocked)
          *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p.out) +
                                       attn_blk_off(b * p.Sq + qrow[r], h * HD + 8 * n + 2 * t)) = pk;
        else
          *reinterpret_cast<uint32_t*>(o + 8 * n + 2 * t) = pk;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (qrow[r] >= p.Sq) continue;
      const long long row = ((static_cast<long long>(split) * p.B + b) * p.Hq + h) * p.Sq + qrow[r];
      float* o = p.part_o + row * HD;
#pragma unroll
      for (int n = 0; n < HD / 8; ++n)
        *reinterpret_cast<float2*>(o + 8 * n + 2 * t) = make_float2(o_acc[n][2 * r], o_acc[n][2 * r + 1]);
      if (t == 0) {
        p.part_ml[row * 2] = m_run[r];
        p.part_ml[row * 2 + 1] = l_run[r];
      }
    }
  }
}

// Combine split partials in split order. One warp per (b, h, s) row.
template <int HD>
__global__ void __launch_bounds__(128) bd_attn_combine_kernel(AttnParams p) {
  grid_dep_launch();
  grid_dep_wait();
  const long long rows = static_cast<long long>(p.B) * p.Hq * p.Sq;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int s = static_cast<int>(row % p.Sq);
  const int h = static_cast<int>((row / p.Sq) % p.Hq);
  const int b = static_cast<int>(row / (static_cast<long long>(p.Sq) * p.Hq));
  float m = -FLT_MAX;
  for (int sp = 0; sp < p.splits; ++sp) m = fmaxf(m, p.part_ml[(sp * rows + row) * 2]);
  float l = 0.f;
  float acc[HD / 32];
#pragma unroll
  for (int i = 0; i < HD / 32; ++i) acc[i] = 0.f;
  for (int sp = 0; sp < p.splits; ++sp) {
    const float ms = p.part_ml[(sp * rows + row) * 2];
    const float ls = p.part_ml[(sp * rows + row) * 2 + 1];
    const float w = (ls > 0.f) ? exp2f(ms - m) : 0.f;
    l += ls * w;
    const float* o = p.part_o + (sp * rows + row) * HD;
#pragma unroll
    for (int i = 0; i < HD / 32; ++i) acc[i] += o[lane + 32 * i] * w;
  }
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  __nv_bfloat16* o = p.out + b * p.o_sb + static_cast<long long>(s) * p.o_ss + h * p.o_sh;
#pragma unroll
  for (int i = 0; i < HD / 32; ++i) {
    const __nv_bfloat16 v = __float2bfloat16_rn(acc[i] * inv);
    if (p.out_blocked)
      *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(p.out) + attn_blk_off(b * p.Sq + s, h * HD + lane + 32 * i)) = v;
    else
      o[lane + 32 * i] = v;
  }
}

template <int HD>
static int launch_attn(const AttnParams& p, bool pdl, cudaStream_t stream) {
  const int smem = 3 * 64 * HD * 2;
  static bool set = false;
  if (!set) {
    BD_CUDA_TRY(cudaFuncSetAttribute(bd_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    set = true;
  }
  dim3 grid((p.Sq + 63) / 64, p.B * p.Hq, p.splits);
  LaunchCfg lc(grid, dim3(128), smem, stream, pdl);
  BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, bd_attn_kernel<HD>, p));
  if (p.splits > 1) {
    const long long rows = static_cast<long long>(p.B) * p.Hq * p.Sq;
    LaunchCfg lc2(dim3(static_cast<unsigned>((rows + 3) / 4)), dim3(128), 0, stream, pdl);
    BD_CUDA_TRY(cudaLaunchKernelEx(&lc2.cfg, bd_attn_combine_kernel<HD>, p));
  }
  return BD_OK;
}

int attn_plan_splits(int B, int Hq, int Sq, int Sk) {
  const int ctas = ((Sq + 63) / 64) * B * Hq;
  const int tiles = (Sk + 63) / 64;
  int s = (2 * num_sms()) / (ctas > 0 ? ctas : 1);
  if (s > tiles / 4) s = tiles / 4;  // at least 4 KV tiles per split
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  return s;
}

int attn_run(AttnParams p, int head_dim, bool pdl, cudaStream_t stream) {
  if (head_dim == 128) return launch_attn<128>(p, pdl, stream);
  if (head_dim == 64) return launch_attn<64>(p, pdl, stream);
  return BD_ERR_UNSUPPORTED;
}

// LLM attention over the paged cache: q [R*S, Hq*hd] (token-major), pools [page][Hkv][64][hd], out [R*S, Hq*hd].
// Keys visible to sequence b: sk_dev[b] + S (past + the block just appended).
size_t attn_llm_workspace_bytes(int R, int S, int Hq, int head_dim, int splits) {
  if (splits <= 1) return 0;
  return static_cast<size_t>(splits) * R * Hq * S * (head_dim + 2) * sizeof(float);
}
int attn_run_llm(const __nv_bfloat16* q, const __nv_bfloat16* kpool, const __nv_bfloat16* vpool, const int* page_table,
                 int max_pages, const int* sk_dev, int sk_bound, __nv_bfloat16* out, int R, int S, int Hq, int Hkv,
                 int head_dim, int causal, int splits, void* ws, size_t ws_bytes, bool pdl, cudaStream_t stream,
                 int out_blocked) {
  AttnParams p{};
  p.out_blocked = out_blocked;
  p.q = q;
  p.q_sb = static_cast<long long>(S) * Hq * head_dim;
  p.q_ss = static_cast<long long>(Hq) * head_dim;
  p.q_sh = head_dim;
  p.k = kpool;
  p.v = vpool;
  p.page_table = page_table;
  p.max_pages = max_pages;
  p.paged = 1;
  p.out = out;
  p.o_sb = p.q_sb;
  p.o_ss = p.q_ss;
  p.o_sh = head_dim;
  p.B = R;
  p.Sq = S;
  p.Sk = sk_bound;
  p.sk_dev = sk_dev;
  p.sk_add = S;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.causal = causal;
  p.splits = splits;
  p.scale_log2 = (1.0f / sqrtf(static_cast<float>(head_dim))) * 1.4426950408889634f;
  if (splits > 1) {
    const size_t need = attn_llm_workspace_bytes(R, S, Hq, head_dim, splits);
    if (!ws || ws_bytes < need) return BD_ERR_WORKSPACE;
    p.part_o = static_cast<float*>(ws);
    p.part_ml = p.part_o + static_cast<size_t>(splits) * R * Hq * S * head_dim;
  }
  return attn_run(p, head_dim, pdl, stream);
}

// Head attention: qkv [R*pn, 3D] (q | k | v along the last dim, heads of head_dim inside each), out [R*pn, D].
int attn_run_head(const __nv_bfloat16* qkv, __nv_bfloat16* out, int R, int pn, int D, int head_dim, bool pdl,
                  cudaStream_t stream) {
  AttnParams p{};
  p.q = qkv;
  p.k = qkv + D;
  p.v = qkv + 2 * D;
  p.q_sb = p.k_sb = static_cast<long long>(pn) * 3 * D;
  p.q_ss = p.k_ss = 3ll * D;
  p.q_sh = p.k_sh = head_dim;
  p.out = out;
  p.o_sb = static_cast<long long>(pn) * D;
  p.o_ss = D;
  p.o_sh = head_dim;
  p.B = R;
  p.Sq = p.Sk = pn;
  p.Hq = p.Hkv = D / head_dim;
  p.causal = 0;
  p.splits = 1;
  p.scale_log2 = (1.0f / sqrtf(static_cast<float>(head_dim))) * 1.4426950408889634f;
  return attn_run(p, head_dim, pdl, stream);
}

}  // namespace bd

using namespace bd;

extern "C" {

size_t bd_attention_workspace_bytes(int B, int Hq, int Sq, int Sk, int head_dim, int splits) {
  if (splits == 0) splits = attn_plan_splits(B, Hq, Sq, Sk);
  if (splits <= 1) return 0;
  return static_cast<size_t>(splits) * B * Hq * Sq * (head_dim + 2) * sizeof(float);
}

int bd_attention_bf16(const void* q, int64_t q_sb, int64_t q_ss, int64_t q_sh, const void* k, const void* v,
                      int64_t k_sb, int64_t k_ss, int64_t k_sh, const int32_t* page_table, int max_pages, void* out,
                      int64_t o_sb, int64_t o_ss, int64_t o_sh, int B, int Sq, int Sk, int Hq, int Hkv, int head_dim,
                      int causal, float scale, int splits, void* workspace, size_t workspace_bytes, int flags,
                      bd_stream_t stream) {
  BD_REQUIRE(q && k && v && out && B > 0 && Sq > 0 && Sk > 0 && Hq > 0 && Hkv > 0 && (Hq % Hkv) == 0);
  BD_REQUIRE(head_dim == 64 || head_dim == 128);
  BD_REQUIRE((q_ss % 8) == 0 && (q_sh % 8) == 0 && (q_sb % 8) == 0 && (o_ss % 2) == 0 && (o_sh % 2) == 0);
  BD_REQUIRE(page_table || ((k_ss % 8) == 0 && (k_sh % 8) == 0 && (k_sb % 8) == 0));
  BD_REQUIRE(!page_table || max_pages * 64 >= Sk);
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.q_sb = q_sb; p.q_ss = q_ss; p.q_sh = q_sh;
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.k_sb = k_sb; p.k_ss = k_ss; p.k_sh = k_sh;
  p.page_table = page_table;
  p.max_pages = max_pages;
  p.paged = page_table ? 1 : 0;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.o_sb = o_sb; p.o_ss = o_ss; p.o_sh = o_sh;
  p.B = B; p.Sq = Sq; p.Sk = Sk; p.Hq = Hq; p.Hkv = Hkv;
  p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (splits == 0) splits = attn_plan_splits(B, Hq, Sq, Sk);
  const int tiles = (Sk + 63) / 64;
  if (splits > tiles) splits = tiles;
  p.splits = splits;
  if (splits > 1) {
    const size_t need = static_cast<size_t>(splits) * B * Hq * Sq * (head_dim + 2) * sizeof(float);
    if (!workspace || workspace_bytes < need) return BD_ERR_WORKSPACE;
    p.part_o = static_cast<float*>(workspace);
    p.part_ml = p.part_o + static_cast<size_t>(splits) * B * Hq * Sq * head_dim;
  }
  return attn_run(p, head_dim, (flags & 1) != 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
