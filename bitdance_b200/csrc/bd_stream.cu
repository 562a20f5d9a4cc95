// bd_stream.cu — the persistent weight-streaming kernel (design notes in bd_stream.cuh), the stream-major weight packer
// and the C entry points that expose them.
#include <cfloat>
#include "bd_stream.cuh"

namespace bd {

// ---------------------------------------------------------------------------------------------------------------------
// GEMM-pass epilogues: NC (32 or 16) consecutive accumulator columns of token row m, packed column n
// ---------------------------------------------------------------------------------------------------------------------
// Stores are always whole 32-byte sectors where the tile allows it: 16-byte stores from 128 threads to 128 different rows
// are partial-sector writes the L2 has to read-modify-write (measured: 11.8 us vs 2.1 us for a 128 x 112 bf16 tile).
__device__ __forceinline__ void store_bf16x16(uint8_t* outb, const StreamOp& op, int m, int col, const float (&y0)[8],
                                              const float (&y1)[8]) {
  const uint4 a = f_to_bf16x8(y0), b = f_to_bf16x8(y1);
  if (op.flags & 1) {
    // blocked: chunks 2k and 2k+1 of a row sit at positions {p, p^1} of the same 32-byte sector; swapped on odd rows
    uint8_t* dst = outb + (blk_off(m, col) & ~31ll);
    if (m & 1) st_global_32B(dst, b, a);
    else st_global_32B(dst, a, b);
  } else {
    st_global_32B(outb + (static_cast<long long>(m) * op.l0 + col) * 2, a, b);
  }
}
__device__ __forceinline__ void store_bf16x8(uint8_t* outb, const StreamOp& op, int m, int col, const float (&y)[8]) {
  uint8_t* dst = (op.flags & 1) ? outb + blk_off(m, col) : outb + (static_cast<long long>(m) * op.l0 + col) * 2;
  *reinterpret_cast<uint4*>(dst) = f_to_bf16x8(y);
}

template <int NC>
__device__ __forceinline__ void gemm_epi_chunk(const StreamOp& op, int M, int m, int n, const float (&acc)[NC],
                                               int split, int mode, const float* bias_sm) {
  if (m >= M) return;
  if (mode & 2) {  // experiment: no stores (keep the data dependency alive)
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) t += acc[j];
    if (t == 123456.789f) reinterpret_cast<float*>(op.o0)[0] = t;
    return;
  }
  if (op.sub == kEpiPartial) {
    float* o = reinterpret_cast<float*>(op.o0) + (static_cast<long long>(split) * M + m) * op.N + n;
#pragma unroll
    for (int j = 0; j < NC / 8; ++j) {
      uint4 a, b;
      a.x = __float_as_uint(acc[8 * j]); a.y = __float_as_uint(acc[8 * j + 1]);
      a.z = __float_as_uint(acc[8 * j + 2]); a.w = __float_as_uint(acc[8 * j + 3]);
      b.x = __float_as_uint(acc[8 * j + 4]); b.y = __float_as_uint(acc[8 * j + 5]);
      b.z = __float_as_uint(acc[8 * j + 6]); b.w = __float_as_uint(acc[8 * j + 7]);
      st_global_32B(o + 8 * j, a, b);
    }
    return;
  }
  // the bias slice of this pass sits in shared memory as fp32 (staged before the accumulator was ready)
  float b[NC];
  if (bias_sm && !(mode & 1)) {
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      const float4 t = *reinterpret_cast<const float4*>(bias_sm + 4 * j);
      b[4 * j] = t.x; b[4 * j + 1] = t.y; b[4 * j + 2] = t.z; b[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j) b[j] = 0.f;
  }
  uint8_t* outb = reinterpret_cast<uint8_t*>(op.o0);
  if (op.sub == kEpiSwiglu8) {
    // packed columns: [8 x1 | 8 x2] per 16-column unit -> 8 outputs per unit at column n/2   (x1, x2 = chunk(w1(a)); silu(x1)*x2)
    float y[NC / 16][8];
#pragma unroll
    for (int u = 0; u < NC / 16; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = bf16_round(acc[16 * u + j] + b[16 * u + j]);
        const float up = bf16_round(acc[16 * u + 8 + j] + b[16 * u + 8 + j]);
        // silu with the fast reciprocal (error ~2 fp32 ulps, far below the bf16 rounding that follows)
        y[u][j] = bf16_round(__fdividef(g, 1.0f + __expf(-g))) * up;
      }
    }
    if constexpr (NC == 32) store_bf16x16(outb, op, m, n >> 1, y[0], y[1]);  // n % 32 == 0 here: a whole sector
    else store_bf16x8(outb, op, m, n >> 1, y[0]);
    return;
  }
  float y[NC / 8][8];
  if (op.act == kActNone) {
    // one rounding: bf16(acc + bias) is produced by the pack itself (an explicit bf16_round first would be idempotent)
#pragma unroll
    for (int u = 0; u < NC / 8; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) y[u][j] = acc[8 * u + j] + b[8 * u + j];
  } else {
#pragma unroll
    for (int u = 0; u < NC / 8; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = bf16_round(acc[8 * u + j] + b[8 * u + j]);
        if (op.act == kActSilu) v = bf16_round(siluf_(v));
        if (op.act == kActGeluTanh) v = bf16_round(gelu_tanhf_(v));
        y[u][j] = v;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NC / 16; ++u) store_bf16x16(outb, op, m, n + 16 * u, y[2 * u], y[2 * u + 1]);
}

// ---------------------------------------------------------------------------------------------------------------------
// row ops: executed by the 128 epilogue threads of CTA r for token row r
// split-KV attention of an AR block (llm_attn_stream below): the CTA whose range [c P / Ge, (c + 1) P / Ge) holds item f
__device__ __forceinline__ int llm_attn_cta_of(int f, int Ge, int P) {
  // 32-bit: the host guarantees P * Ge < 2^31 (bd_llm.cu); a 64-bit division is a ~150-instruction subroutine
  return static_cast<int>((static_cast<unsigned>(f + 1) * static_cast<unsigned>(Ge) - 1u) / static_cast<unsigned>(P));
}

// ---------------------------------------------------------------------------------------------------------------------
// A row op is a chain of L2 round trips, not a bandwidth problem (there is no L1 left beside 226 KB of shared memory): the
// round-1 version issued its loads iteration by iteration — every 16-byte store to the row (which may alias in the
// compiler's eyes) fenced the next iteration's loads, the k-split loop had a run-time trip count, the LayerNorm affine
// weights came as 4-byte loads — ~20 dependent round trips, 11 us per op (profiles/r01_rowop_experiments.txt). Here every
// phase issues ALL its loads into registers first (compile-time vector count NV and split count), read-only operands go
// through the non-coherent path (ld.global.nc may be hoisted above stores), and the row itself is kept as packed bf16
// (it is bf16-rounded by construction) so that the operands of the next phase fit in registers while it is reduced.
constexpr int kRowVec = 6;  // D <= 128 * 6 * 8 = 6144

__device__ __forceinline__ uint4 ldg_u4(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ldg_f4(const void* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void compiler_fence() { asm volatile("" ::: "memory"); }

// LN statistics + modulate of a row held as packed bf16; writes `a` blocked (or to smem floats when a_smem != nullptr).
// TPR: threads per row — 128 (the 4 executor warps on one row, block reductions through `red`) or 32 (one warp per row,
// shuffle reductions, `tid` = lane: the small-model mode where a CTA owns several rows of the tile, see kStreamFamHeadSmall)
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* red, int tid) {
  if constexpr (TPR == 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  } else {
    return epi_sum(v, red, tid);
  }
}

template <int NV, int TPR = 128>
__device__ __forceinline__ void row_ln_mod(const StreamOp& op, int r, int tid, const uint4 (&vb)[NV], float sum, float* red,
                                           const float* ln_w, const float* ln_b, float* a_smem) {
  const int D = op.N, nvec = D / 8;
  const __nv_bfloat16* scale = reinterpret_cast<const __nv_bfloat16*>(op.p3) + static_cast<long long>(r) * op.l0;
  const __nv_bfloat16* shift = reinterpret_cast<const __nv_bfloat16*>(op.p4) + static_cast<long long>(r) * op.l0;
  // modulation + affine operands: in flight while the two block reductions run
  uint4 scr[NV], shr[NV];
  float4 lw[NV][2], lb[NV][2];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = tid + i * TPR;
    scr[i] = shr[i] = make_uint4(0, 0, 0, 0);
    lw[i][0] = lw[i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
    lb[i][0] = lb[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nvec) {
      scr[i] = ldcg_u4(scale + c * 8);
      shr[i] = ldcg_u4(shift + c * 8);
      if (ln_w) {
        lw[i][0] = ldg_f4(ln_w + c * 8);
        lw[i][1] = ldg_f4(ln_w + c * 8 + 4);
        lb[i][0] = ldg_f4(ln_b + c * 8);
        lb[i][1] = ldg_f4(ln_b + c * 8 + 4);
      }
    }
  }
  const float mean = row_sum<TPR>(sum, red, tid) / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = tid + i * TPR;
    if (c < nvec) {
      float v[8];
      bf16x8_to_f(vb[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(row_sum<TPR>(sq, red, tid) / static_cast<float>(D) + op.f0);
  uint8_t* a = reinterpret_cast<uint8_t*>(op.o0);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = tid + i * TPR;
    if (c < nvec) {
      float v[8], sc[8], sh[8], o[8];
      bf16x8_to_f(vb[i], v);
      bf16x8_to_f(scr[i], sc);
      bf16x8_to_f(shr[i], sh);
      const float w8[8] = {lw[i][0].x, lw[i][0].y, lw[i][0].z, lw[i][0].w, lw[i][1].x, lw[i][1].y, lw[i][1].z, lw[i][1].w};
      const float b8[8] = {lb[i][0].x, lb[i][0].y, lb[i][0].z, lb[i][0].w, lb[i][1].x, lb[i][1].y, lb[i][1].z, lb[i][1].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float hn = (v[j] - mean) * rstd;
        if (ln_w) hn = hn * w8[j] + b8[j];
        o[j] = hn * bf16_round(1.0f + sc[j]) + sh[j];
      }
      if (a_smem) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a_smem[c * 8 + j] = bf16_round(o[j]);
      } else {
        *reinterpret_cast<uint4*>(a + blk_off(r, c * 8)) = f_to_bf16x8(o);
      }
    }
  }
}

// h = bf16(h + bf16(bf16(sum_s partial_s + bias) * gate)), fixed summation order s = 0, 1, ... (deterministic); returns
// the row as packed bf16 + its sum. Loads: two splits per batch, then [h, gate, bias] in one batch.
template <int NV, int SC, int TPR = 128>  // SC: compile-time split count (0 = run time)
__device__ __forceinline__ float row_splitk(const StreamOp& op, int M, int r, int tid, uint4 (&vb)[NV]) {
  const int D = op.N, nvec = D / 8, S = SC > 0 ? SC : op.i0;
  const float* part = reinterpret_cast<const float*>(op.p0) + static_cast<long long>(r) * D;
  const long long sstride = static_cast<long long>(M) * D;
  const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(op.p5);
  const __nv_bfloat16* gate = reinterpret_cast<const __nv_bfloat16*>(op.p6) + static_cast<long long>(r) * op.l0;
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(op.o1) + static_cast<long long>(r) * D;
  float acc[NV][8];
  auto add_split = [&](int s, bool first) {
    float4 x0[NV], x1[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = tid + i * TPR;
      x0[i] = x1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec) {
        const float* q = part + s * sstride + c * 8;
        x0[i] = ldcg_f4(q);
        x1[i] = ldcg_f4(q + 4);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (first) {
        acc[i][0] = x0[i].x; acc[i][1] = x0[i].y; acc[i][2] = x0[i].z; acc[i][3] = x0[i].w;
        acc[i][4] = x1[i].x; acc[i][5] = x1[i].y; acc[i][6] = x1[i].z; acc[i][7] = x1[i].w;
      } else {
        acc[i][0] += x0[i].x; acc[i][1] += x0[i].y; acc[i][2] += x0[i].z; acc[i][3] += x0[i].w;
        acc[i][4] += x1[i].x; acc[i][5] += x1[i].y; acc[i][6] += x1[i].z; acc[i][7] += x1[i].w;
      }
    }
  };
  // register budget (255 per thread, the op descriptor alone holds ~45): two splits' loads (80) + the sums (40) at a time,
  // then the residual / gate / bias vectors (60) — three L2 round trips instead of ~20
  if constexpr (SC > 0) {
#pragma unroll
    for (int s = 0; s < SC; ++s) {
      if (s > 0 && (s & 1) == 0) compiler_fence();
      add_split(s, s == 0);
    }
  } else {
    add_split(0, true);
    for (int s = 1; s < S; ++s) add_split(s, false);
  }
  compiler_fence();
  uint4 hraw[NV], graw[NV], braw[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = tid + i * TPR;
    hraw[i] = graw[i] = braw[i] = make_uint4(0, 0, 0, 0);
    if (c < nvec) {
      hraw[i] = ldcg_u4(h + c * 8);
      graw[i] = ldcg_u4(gate + c * 8);
      braw[i] = ldg_u4(bias + c * 8);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = tid + i * TPR;
    vb[i] = make_uint4(0, 0, 0, 0);
    if (c < nvec) {
      float b[8], g[8], res[8], v[8];
      bf16x8_to_f(braw[i], b);
      bf16x8_to_f(graw[i], g);
      bf16x8_to_f(hraw[i], res);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = bf16_round(bf16_round(acc[i][j] + b[j]) * g[j]);
        v[j] = bf16_round(res[j] + y);
        sum += v[j];
      }
      vb[i] = f_to_bf16x8(v);
      *reinterpret_cast<uint4*>(h + c * 8) = vb[i];
    }
  }
  return sum;
}

// final layer tail: a (bf16 values, fp32 in shared memory) -> o_c = bf16(sum_d a_d Wf[c,d] + bias_c); optional 2*sigmoid-1
template <int TPR = 128>
__device__ __forceinline__ void row_final_linear(const StreamOp& op, int M, int it, int r, int tid, const float* arow) {
  const int D = op.N, C = op.i1;
  const __nv_bfloat16* Wf = reinterpret_cast<const __nv_bfloat16*>(op.p1);
  const __nv_bfloat16* bfin = reinterpret_cast<const __nv_bfloat16*>(op.p2);
  float* pred = reinterpret_cast<float*>(op.o0);
  float* trace = reinterpret_cast<float*>(op.o2);
  const int warp = TPR == 128 ? (tid >> 5) : 0, lane = tid & 31;
  constexpr int kCStep = TPR == 128 ? 32 : 8;  // warp mode: the one warp takes every group of 8 channels in turn
  for (int c0 = warp * 8; c0 < C; c0 += kCStep) {  // 8 output channels per warp and round: 8 independent loads per step
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
#pragma unroll 2
    for (int d = lane * 8; d < D; d += 256) {
      uint4 raw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        raw[u] = (c0 + u < C) ? ldg_u4(Wf + static_cast<long long>(c0 + u) * D + d) : make_uint4(0, 0, 0, 0);
      const float4 a0 = *reinterpret_cast<const float4*>(arow + d), a1 = *reinterpret_cast<const float4*>(arow + d + 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float wv[8];
        bf16x8_to_f(raw[u], wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[u] = fmaf(av[j], wv[j], acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if (lane < 8 && c0 + lane < C) {
      float sel = acc[0];
#pragma unroll
      for (int u = 1; u < 8; ++u) sel = (lane == u) ? acc[u] : sel;
      float o = bf16_round(sel + (bfin ? __bfloat162float(bfin[c0 + lane]) : 0.f));
      if (op.i2) {
        const float sg = bf16_round(1.0f / (1.0f + expf(-o)));
        o = bf16_round(bf16_round(2.0f * sg) - 1.0f);
      }
      pred[static_cast<long long>(r) * C + c0 + lane] = o;
      if (trace) trace[(static_cast<long long>(it) * M + r) * C + c0 + lane] = o;
    }
  }
}

template <int NV, int TPR = 128>
__device__ __forceinline__ void row_op_ln_family(const StreamProgram& prog, const StreamOp& op, int it, int r, int tid,
                                                 float* red, uint8_t* scratch) {
  const int M = prog.M;
  const float* ln_w = reinterpret_cast<const float*>(op.p1);
  const float* ln_b = reinterpret_cast<const float*>(op.p2);
  uint4 vb[NV];
  if (op.sub == kRowLnMod) {
    const int nvec = op.N / 8;
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(op.p0) + static_cast<long long>(r) * op.N;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = tid + i * TPR;
      vb[i] = c < nvec ? ldcg_u4(h + c * 8) : make_uint4(0, 0, 0, 0);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float v[8];
      bf16x8_to_f(vb[i], v);  // zero beyond nvec
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[j];
    }
    row_ln_mod<NV, TPR>(op, r, tid, vb, sum, red, ln_w, ln_b, nullptr);
    return;
  }
  // h = bf16(h + bf16(bf16(sum_s partial_s + bias) * gate));  then the LayerNorm-modulate that follows in the network
  // (TransBlock.forward flow_head_parallel_x.py:242-252; FinalLayer.forward :169-173 for kRowFinal)
  float sum;
  if (op.i0 == 4) sum = row_splitk<NV, 4, TPR>(op, M, r, tid, vb);
  else if (op.i0 == 2) sum = row_splitk<NV, 2, TPR>(op, M, r, tid, vb);
  else if (op.i0 == 1) sum = row_splitk<NV, 1, TPR>(op, M, r, tid, vb);
  else sum = row_splitk<NV, 0, TPR>(op, M, r, tid, vb);
  if (op.sub == kRowSplitkLnMod) {
    row_ln_mod<NV, TPR>(op, r, tid, vb, sum, red, ln_w, ln_b, nullptr);
    return;
  }
  // final layer: LayerNorm without affine; a stays in shared memory (aliasing the A ring; warp mode: the warp's own
  // slice), then the D -> C Linear
  float* arow = reinterpret_cast<float*>(scratch);
  row_ln_mod<NV, TPR>(op, r, tid, vb, sum, red, nullptr, nullptr, arow);
  if constexpr (TPR == 128) epi_bar();
  else __syncwarp();
  row_final_linear<TPR>(op, M, it, r, tid, arow);
  if constexpr (TPR == 128) epi_bar();  // arow (aliasing the A ring) is dead before anything else may touch it
  else __syncwarp();
}

// Qwen3 residual add + RMSNorm of token row r with the loads of every phase issued together (see the note at the top of
// the row ops). Field meaning: kRowLlmRms / kRowLlmResRms in row_op.
template <int NV>
__device__ __forceinline__ void llm_res_rms(const StreamOp& op, int M, int r, int tid, float* red) {
  const int D = op.N, nvec = D / 8;
  float* hid = reinterpret_cast<float*>(op.o1) + static_cast<long long>(r) * D;
  float v[NV][8];
  if (op.sub == kRowLlmResRms) {
    const float* part = reinterpret_cast<const float*>(op.p0) + static_cast<long long>(r) * D;
    const long long sstride = static_cast<long long>(M) * D;
    float acc[NV][8];
    auto add_split = [&](int s, bool first) {
      float4 x0[NV], x1[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = tid + i * 128;
        x0[i] = x1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nvec) {
          const float* q = part + s * sstride + c * 8;
          x0[i] = ldcg_f4(q);
          x1[i] = ldcg_f4(q + 4);
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (first) {
          acc[i][0] = x0[i].x; acc[i][1] = x0[i].y; acc[i][2] = x0[i].z; acc[i][3] = x0[i].w;
          acc[i][4] = x1[i].x; acc[i][5] = x1[i].y; acc[i][6] = x1[i].z; acc[i][7] = x1[i].w;
        } else {
          acc[i][0] += x0[i].x; acc[i][1] += x0[i].y; acc[i][2] += x0[i].z; acc[i][3] += x0[i].w;
          acc[i][4] += x1[i].x; acc[i][5] += x1[i].y; acc[i][6] += x1[i].z; acc[i][7] += x1[i].w;
        }
      }
    };
    add_split(0, true);  // fixed order: deterministic
    for (int s = 1; s < op.i0; ++s) {
      if ((s & 1) == 0) compiler_fence();
      add_split(s, false);
    }
    compiler_fence();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = tid + i * 128;
      if (c < nvec) {
        const float4 r0 = ldcg_f4(hid + c * 8), r1 = ldcg_f4(hid + c * 8 + 4);
        const float res[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = res[j] + bf16_round(acc[i][j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = tid + i * 128;
      if (c < nvec) {
        *reinterpret_cast<float4*>(hid + c * 8) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        *reinterpret_cast<float4*>(hid + c * 8 + 4) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = tid + i * 128;
      float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
      if (c < nvec) {
        r0 = ldcg_f4(hid + c * 8);
        r1 = ldcg_f4(hid + c * 8 + 4);
      }
      v[i][0] = r0.x; v[i][1] = r0.y; v[i][2] = r0.z; v[i][3] = r0.w;
      v[i][4] = r1.x; v[i][5] = r1.y; v[i][6] = r1.z; v[i][7] = r1.w;
    }
  }
  if (!op.p1) return;
  // norm weights (read-only): in flight during the reduction
  uint4 nwr[NV];
  const __nv_bfloat16* nw = reinterpret_cast<const __nv_bfloat16*>(op.p1);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = tid + i * 128;
    nwr[i] = c < nvec ? ldg_u4(nw + c * 8) : make_uint4(0, 0, 0, 0);
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
  const float rstd = rsqrtf(epi_sum(ss, red, tid) / static_cast<float>(D) + op.f0);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = tid + i * 128;
    if (c < nvec) {
      float wv[8], y[8];
      bf16x8_to_f(nwr[i], wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = wv[j] * (v[i][j] * rstd);
      if (op.i1) {
        float* o = reinterpret_cast<float*>(op.o0) + static_cast<long long>(r) * D + c * 8;
        if (op.p3) {
          const float* add = reinterpret_cast<const float*>(op.p3) + static_cast<long long>(r % op.i2) * D + c * 8;
          const float4 a0 = ldg_f4(add), a1 = ldg_f4(add + 4);
          y[0] += a0.x; y[1] += a0.y; y[2] += a0.z; y[3] += a0.w; y[4] += a1.x; y[5] += a1.y; y[6] += a1.z; y[7] += a1.w;
        }
        *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(op.o0) + blk_off(r, c * 8)) = f_to_bf16x8(y);
      }
    }
  }
}

// FAM: the kernel is compiled once per program family (kStreamFamHead: the diffusion head's row ops + per-(sequence,
// head) attention; kStreamFamLlm: the Qwen3 block's) — each instance carries only its own executors, so that neither
// pays for the other's registers and instruction footprint.
template <int FAM>
__device__ __forceinline__ void row_op(const StreamProgram& prog, const StreamOp& op, int it, int r, int tid, float* red,
                                       uint8_t* scratch) {
  const int M = prog.M;
  if constexpr (FAM != kStreamFamLlm) {
  switch (op.sub) {
    case kRowCastCond: {  // p0 fp32 [M, N] (kernel input) -> o0 blocked bf16
      if (r >= M) return;
      const float* src = reinterpret_cast<const float*>(op.p0) + static_cast<long long>(r) * op.N;
      uint8_t* dst = reinterpret_cast<uint8_t*>(op.o0);
      for (int c = tid; c < op.N / 8; c += 128) {
        const float4 x0 = reinterpret_cast<const float4*>(src + c * 8)[0], x1 = reinterpret_cast<const float4*>(src + c * 8)[1];
        const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        *reinterpret_cast<uint4*>(dst + blk_off(r, c * 8)) = f_to_bf16x8(v);
      }
      return;
    }
    case kRowTFreq: {  // timestep_embedding (flow_head_parallel_x.py:12-27) of t = sched[r][0], dim N -> o0 blocked bf16
      if (r >= prog.n_iter) return;
      const int half = op.N / 2;
      const float tv = 1000.0f * prog.sched[r][0];
      uint8_t* dst = reinterpret_cast<uint8_t*>(op.o0);
      for (int k = tid; k < half; k += 128) {
        const float f = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
        const float a = tv * f;
        *reinterpret_cast<__nv_bfloat16*>(dst + blk_off(r, k)) = __float2bfloat16_rn(cosf(a));
        *reinterpret_cast<__nv_bfloat16*>(dst + blk_off(r, half + k)) = __float2bfloat16_rn(sinf(a));
      }
      return;
    }
    case kRowInit: {  // x = noise[0]; xb = bf16(x) for every CFG group   (sampling_x.py:60,71)
      if (r >= prog.rows_x || tid >= op.N) return;
      const float v = reinterpret_cast<const float*>(op.p0)[static_cast<long long>(r) * op.N + tid];
      reinterpret_cast<float*>(op.o0)[static_cast<long long>(r) * op.N + tid] = v;
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      uint8_t* xb = reinterpret_cast<uint8_t*>(op.o1);
      for (int g = 0; g < prog.cfg_mult; ++g) *reinterpret_cast<__nv_bfloat16*>(xb + blk_off(g * prog.rows_x + r, tid)) = b;
      return;
    }
    case kRowSiluAdd: {  // y = bf16(silu(bf16(temb[it] + cemb[r])))   (TransEncoder.forward :330)
      if (r >= M) return;
      const __nv_bfloat16* te = reinterpret_cast<const __nv_bfloat16*>(op.p0) + static_cast<long long>(it + op.i0) * op.N;
      const __nv_bfloat16* ce = reinterpret_cast<const __nv_bfloat16*>(op.p1) + static_cast<long long>(r) * op.N;
      uint8_t* dst = reinterpret_cast<uint8_t*>(op.o0);
      for (int c = tid; c < op.N / 8; c += 128) {
        float a[8], b[8], o[8];
        bf16x8_to_f(ldcg_u4(te + c * 8), a);
        bf16x8_to_f(ldcg_u4(ce + c * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = siluf_(bf16_round(a[j] + b[j]));
        *reinterpret_cast<uint4*>(dst + blk_off(r, c * 8)) = f_to_bf16x8(o);
      }
      return;
    }
    case kRowSiluAddAll: {
      // y[j] = bf16(silu(bf16(temb[j] + cemb[r]))) for EVERY evaluation j (all known before the first one: the timestep
      // schedule is fixed and c does not change during a sampler call) -> o0 + j * l1, blocked. Lets the adaLN GEMM of
      // evaluation j + 1 run as filler pieces inside evaluation j. The next temb row is fetched before the current one is
      // stored (stores would otherwise fence the loads).
      if (r >= M) return;
      const int nvec = op.N / 8;
      const __nv_bfloat16* te = reinterpret_cast<const __nv_bfloat16*>(op.p0);
      const __nv_bfloat16* ce = reinterpret_cast<const __nv_bfloat16*>(op.p1) + static_cast<long long>(r) * op.N;
      uint8_t* dst = reinterpret_cast<uint8_t*>(op.o0);
      uint4 cer[kRowVec], cur[kRowVec], nxt[kRowVec];
#pragma unroll
      for (int i = 0; i < kRowVec; ++i) {
        const int c = tid + i * 128;
        cer[i] = cur[i] = nxt[i] = make_uint4(0, 0, 0, 0);
        if (c < nvec) {
          cer[i] = ldcg_u4(ce + c * 8);
          cur[i] = ldcg_u4(te + c * 8);
        }
      }
      for (int j = 0; j < prog.n_iter; ++j) {
        if (j + 1 < prog.n_iter) {
#pragma unroll
          for (int i = 0; i < kRowVec; ++i) {
            const int c = tid + i * 128;
            if (c < nvec) nxt[i] = ldcg_u4(te + static_cast<long long>(j + 1) * op.N + c * 8);
          }
        }
#pragma unroll
        for (int i = 0; i < kRowVec; ++i) {
          const int c = tid + i * 128;
          if (c < nvec) {
            float a[8], b[8], o[8];
            bf16x8_to_f(cur[i], a);
            bf16x8_to_f(cer[i], b);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = siluf_(bf16_round(a[k] + b[k]));
            *reinterpret_cast<uint4*>(dst + j * op.l1 + blk_off(r, c * 8)) = f_to_bf16x8(o);
          }
          cur[i] = nxt[i];
        }
      }
      return;
    }
    case kRowLnMod:
    case kRowSplitkLnMod:
    case kRowFinal: {
      if (r >= M) return;
      if (op.N == 5120) row_op_ln_family<5>(prog, op, it, r, tid, red, scratch);  // the 14B head: no predicated tail
      else row_op_ln_family<kRowVec>(prog, op, it, r, tid, red, scratch);
      return;
    }
    case kRowSde: {  // sampling_x.py:24-41 with explicit round-to-nearest ops (torch rounds every op separately)
      // + (p4 != null) y of the NEXT evaluation, y = silu(temb[it + 1] + cemb[r]) -> o3 blocked: it depends on nothing the
      //   sampler computes, and doing it here removes one op (and one grid barrier) from every evaluation
      if (r < prog.rows_x && tid < op.N) {
        const int C = op.N, nx = prog.rows_x;
        const float* pred = reinterpret_cast<const float*>(op.p0);
        float* x = reinterpret_cast<float*>(op.o0);
        const long long i = static_cast<long long>(r) * C + tid;
        const long long n = static_cast<long long>(nx) * C;
        const float t = prog.sched[it][0], dt = prog.sched[it][1], denom = prog.sched[it][2], var = prog.sched[it][3],
                    omt = prog.sched[it][4], nscale = prog.sched[it][5];
        const bool last = (it == prog.n_iter - 1);
        const float xv = x[i];
        float v = __fdiv_rn(__fsub_rn(__ldcg(pred + i), xv), denom);
        if (prog.cfg_mult == 2) {
          const float vu = __fdiv_rn(__fsub_rn(__ldcg(pred + n + i), xv), denom);
          v = __fadd_rn(vu, __fmul_rn(prog.cfg, __fsub_rn(v, vu)));
        }
        float xn;
        if (last) {
          xn = __fadd_rn(xv, __fmul_rn(v, dt));
        } else {
          const float nz = reinterpret_cast<const float*>(op.p1)[static_cast<long long>(it + 1) * n + i];
          const float score = __fdiv_rn(__fsub_rn(__fmul_rn(t, v), xv), var);
          const float drift = __fadd_rn(v, __fmul_rn(omt, score));
          xn = __fadd_rn(__fadd_rn(xv, __fmul_rn(drift, dt)), __fmul_rn(nscale, nz));
        }
        x[i] = xn;
        if (last) {
          reinterpret_cast<float*>(op.o2)[i] = xn;
        } else {
          const __nv_bfloat16 b = __float2bfloat16_rn(xn);
          uint8_t* xb = reinterpret_cast<uint8_t*>(op.o1);
          for (int g = 0; g < prog.cfg_mult; ++g) *reinterpret_cast<__nv_bfloat16*>(xb + blk_off(g * nx + r, tid)) = b;
        }
      }
      if (op.p4 && r < M && it + 1 < prog.n_iter) {
        const int D = op.i0;
        const __nv_bfloat16* te = reinterpret_cast<const __nv_bfloat16*>(op.p4) + static_cast<long long>(it + 1) * D;
        const __nv_bfloat16* ce = reinterpret_cast<const __nv_bfloat16*>(op.p5) + static_cast<long long>(r) * D;
        uint8_t* dst = reinterpret_cast<uint8_t*>(op.o3);
        for (int c = tid; c < D / 8; c += 128) {
          float a[8], b[8], o[8];
          bf16x8_to_f(ldcg_u4(te + c * 8), a);
          bf16x8_to_f(ldcg_u4(ce + c * 8), b);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = siluf_(bf16_round(a[j] + b[j]));
          *reinterpret_cast<uint4*>(dst + blk_off(r, c * 8)) = f_to_bf16x8(o);
        }
      }
      return;
    }
    default: return;
  }
  } else {
  switch (op.sub) {
    case kRowLlmAttnCombine: {
      // token row r = (sequence b, position s): reduce the segments of every (b, q head) block in slot order (see
      // llm_attn_stream: same geometry) -> bf16 -> blocked operand of o_proj. p0 partial O fp32 [R Hq][i0][S][hd], p1
      // partial (max, sum) [R Hq][i0][S][2], p2 seq_lens, i0 = slots per block, i1 = Hq, i2 = S, act = CTA bound of the
      // attention op, K = hd, o0 = out blocked
      if (r >= M) return;
      const int HD = op.K, Hq = op.i1, S = op.i2, maxseg = op.i0, R = M / S;
      const int b = r / S, sq = r % S;
      const float* part_o = reinterpret_cast<const float*>(op.p0);
      const float* part_ml = reinterpret_cast<const float*>(op.p1);
      const int* seq_lens = reinterpret_cast<const int*>(op.p2);
      int P = 0, base = 0, Tb = 0;
      for (int bb = 0; bb < R; ++bb) {
        const int tb = (__ldcg(seq_lens + bb) + S + 63) >> 6;
        if (bb < b) base += Hq * tb;
        if (bb == b) Tb = tb;
        P += Hq * tb;
      }
      const int Ge = min(op.act, P);
      uint8_t* out = reinterpret_cast<uint8_t*>(op.o0);
      const int vph = HD / 8;
      for (int v = tid; v < Hq * vph; v += 128) {
        const int h = v / vph, d0 = (v % vph) * 8;
        const int f0 = base + h * Tb;
        const int nseg = llm_attn_cta_of(f0 + Tb - 1, Ge, P) - llm_attn_cta_of(f0, Ge, P) + 1;
        const long long row0 = (static_cast<long long>(b) * Hq + h) * maxseg * S + sq;  // + seg * S
        constexpr int kMaxS = 16;
        float ms[kMaxS], ls[kMaxS];
#pragma unroll
        for (int sp = 0; sp < kMaxS; ++sp) {
          ms[sp] = -FLT_MAX;
          ls[sp] = 0.f;
          if (sp < nseg) {
            const float2 t = __ldcg(reinterpret_cast<const float2*>(part_ml + (row0 + sp * S) * 2));
            ms[sp] = t.x;
            ls[sp] = t.y;
          }
        }
        float mx = -FLT_MAX;
#pragma unroll
        for (int sp = 0; sp < kMaxS; ++sp) mx = fmaxf(mx, ms[sp]);
        float l = 0.f, acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int s0 = 0; s0 < kMaxS; s0 += 4) {
          if (s0 >= nseg) break;
          float4 x0[4], x1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            x0[u] = x1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + u < nseg) {
              const float* o = part_o + (row0 + (s0 + u) * S) * HD + d0;
              x0[u] = ldcg_f4(o);
              x1[u] = ldcg_f4(o + 4);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {  // slot order: deterministic
            const float wgt = (ls[s0 + u] > 0.f) ? exp2f(ms[s0 + u] - mx) : 0.f;
            l += ls[s0 + u] * wgt;
            acc[0] += x0[u].x * wgt; acc[1] += x0[u].y * wgt; acc[2] += x0[u].z * wgt; acc[3] += x0[u].w * wgt;
            acc[4] += x1[u].x * wgt; acc[5] += x1[u].y * wgt; acc[6] += x1[u].z * wgt; acc[7] += x1[u].w * wgt;
          }
        }
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = acc[j] * inv;
        *reinterpret_cast<uint4*>(out + blk_off(r, h * HD + d0)) = f_to_bf16x8(y);
      }
      return;
    }
    case kRowLlmRms:
    case kRowLlmResRms: {
      // Qwen3DecoderLayer on the fp32 residual stream of an AR block (transformers qwen3; oracle/llm.py stream_f32):
      //   kRowLlmResRms: hidden += bf16(sum_s partial_s)   (the o_proj / down_proj output, rounded to bf16 by the Linear)
      //   then  y = float(w) * (hidden * rstd)             (Qwen3RMSNorm in fp32)
      //   i1 == 0: y -> bf16, blocked (operand of the next GEMM);  i1 == 1: the final norm: out fp32 = y + add[m % i2]
      // p0 partials fp32 [S][M][D] (i0 = S), o1 hidden fp32 [M, D] in/out, p1 norm weight bf16 [D] (nullable for i1 == 0: no
      // norm), p3 add table fp32 [i2][D] (nullable), o0 output
      if (r >= M) return;
      if (op.N == 5120) llm_res_rms<5>(op, M, r, tid, red);
      else llm_res_rms<kRowVec>(op, M, r, tid, red);
      return;
    }
    default: return;
  }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention over the pn <= 64 tokens of one (sequence, head): Attention.forward flow_head_parallel_x.py:192-220
// (flash_attn_func semantics: fp32 scores / softmax, P and O in bf16). 128 threads, mma.sync m16n8k16.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void s_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void s_mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// mma.sync without `volatile`: a pure register operation the compiler may schedule freely between the loads
__device__ __forceinline__ void s_mma_16816_nv(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t s_pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// A Q / K / V / O tile in shared memory: HD/64 half-tiles of [64 rows x 128 B], each the verbatim image of `pn` rows of one
// k-block of a blocked activation (so a tile is filled by HD/64 bulk copies and written back sector by sector). The
// 128-byte swizzle of a blocked buffer uses the row index inside its 128-row block, hence row0 = seq * pn.
__device__ __forceinline__ uint32_t s_tile_off(int row, int chunk, int row0) {
  return static_cast<uint32_t>((chunk >> 3) * 8192 + row * 128 + (((chunk & 7) ^ ((row0 + row) & 7)) << 4));
}

// Attention over the pn <= 64 tokens of one (sequence, head). qkv is the blocked bf16 output of the wqkv GEMM
// ([3D/64][128][64]); the tiles arrive by bulk copy (one mbarrier round trip instead of a chain of 16-byte loads through
// the L1-less load path: measured 12.5 us -> see profiles), the output goes back as whole 32-byte sectors.
template <int HD>
__device__ __forceinline__ void attn_unit(const StreamOp& op, int unit, int tid, uint8_t* smem, uint64_t* aux_bar,
                                          uint32_t aux_parity) {
  const int D = op.N, pn = op.i0, H = D / HD;
  const int seq = unit / H, hd = unit % H;
  const int row0 = seq * pn;
  constexpr int kHalves = HD / 64;
  constexpr int kTile = kHalves * 8192;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;
  uint8_t* sV = sK + kTile;
  const uint8_t* qkv = reinterpret_cast<const uint8_t*>(op.p0);
  if (tid == 0) {
    fence_proxy_async_all();  // the wqkv epilogues' generic-proxy stores (acquired by this thread) -> bulk-copy reads
    mbar_expect_tx(aux_bar, static_cast<uint32_t>(3 * kHalves * pn * 128));
#pragma unroll
    for (int which = 0; which < 3; ++which)
#pragma unroll
      for (int hf = 0; hf < kHalves; ++hf) {
        const int kb = (which * D + hd * HD) / 64 + hf;
        bulk_g2s(smem + which * kTile + hf * 8192, qkv + static_cast<long long>(kb) * kSlotBytes + row0 * 128,
                 static_cast<uint32_t>(pn * 128), aux_bar, kEvictLast);
      }
  }
  if (pn < 64) {  // rows beyond pn: V must be finite (P is exactly 0 there), K / Q rows are masked / never stored
    for (int i = tid; i < (64 - pn) * 8 * kHalves; i += 128) {
      const int hf = i / ((64 - pn) * 8), rem = i % ((64 - pn) * 8);
      *reinterpret_cast<uint4*>(sV + hf * 8192 + (pn + rem / 8) * 128 + (rem % 8) * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  mbar_wait(aux_bar, aux_parity);
  epi_bar();
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float scale_log2 = rsqrtf(static_cast<float>(HD)) * 1.4426950408889634f;
  float s[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ks += 2) {
    uint32_t a0[4], a1[4];
    s_ldmatrix_x4(a0, smem_u32(sQ + s_tile_off(warp * 16 + (lane & 15), 2 * ks + (lane >> 4), row0)));
    s_ldmatrix_x4(a1, smem_u32(sQ + s_tile_off(warp * 16 + (lane & 15), 2 * ks + 2 + (lane >> 4), row0)));
    // the 8 key blocks are 8 independent accumulator chains: all loads first, then the two k-steps chain by chain (the mma
    // is a plain register operation, free to be scheduled; per accumulator the order of the k-steps is unchanged)
    uint32_t bk[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s_ldmatrix_x4(bk[j], smem_u32(sK + s_tile_off(8 * j + (lane & 7), 2 * ks + (lane >> 3), row0)));
#pragma unroll
    for (int j = 0; j < 8; ++j) s_mma_16816_nv(s[j], a0, bk[j][0], bk[j][1]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s_mma_16816_nv(s[j], a1, bk[j][2], bk[j][3]);
  }
  float mx[2] = {-FLT_MAX, -FLT_MAX};
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int key = 8 * j + 2 * t + (i & 1);
      const float v = key < pn ? s[j][i] * scale_log2 : -FLT_MAX;
      s[j][i] = v;
      mx[i >> 1] = fmaxf(mx[i >> 1], v);
    }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
  float l[2] = {0.f, 0.f};
  uint32_t pa[4][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      e[i] = (s[j][i] == -FLT_MAX) ? 0.f : exp2f(s[j][i] - mx[i >> 1]);
      l[i >> 1] += e[i];
    }
    const int kk = j >> 1;
    if ((j & 1) == 0) {
      pa[kk][0] = s_pack_bf16(e[0], e[1]);
      pa[kk][1] = s_pack_bf16(e[2], e[3]);
    } else {
      pa[kk][2] = s_pack_bf16(e[0], e[1]);
      pa[kk][3] = s_pack_bf16(e[2], e[3]);
    }
  }
  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int nh = 0; nh < HD / 8; nh += 8) {
      uint32_t bv[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
        s_ldmatrix_x4_trans(bv[n], smem_u32(sV + s_tile_off(16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1),
                                                            nh + 2 * n + (lane >> 4), row0)));
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        s_mma_16816_nv(o_acc[nh + 2 * n], pa[kk], bv[n][0], bv[n][1]);
        s_mma_16816_nv(o_acc[nh + 2 * n + 1], pa[kk], bv[n][2], bv[n][3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  // O tile -> the Q tile's place (a warp only ever touches its own 16 rows of Q), then back to HBM as whole sectors
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qrow = warp * 16 + g + 8 * r;
    const float inv = l[r] > 0.f ? 1.0f / l[r] : 0.f;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n)
      *reinterpret_cast<uint32_t*>(sQ + s_tile_off(qrow, n, row0) + 4 * t) =
          s_pack_bf16(o_acc[n][2 * r] * inv, o_acc[n][2 * r + 1] * inv);
  }
  __syncwarp();
  uint8_t* out = reinterpret_cast<uint8_t*>(op.o0);
  for (int i = lane; i < 16 * 4 * kHalves; i += 32) {  // 16 rows x (HD / 16) sectors of 32 bytes
    const int hf = i / 64, rr = (i % 64) / 4, sec = i % 4;
    const int qrow = warp * 16 + rr;
    if (qrow >= pn) continue;
    const uint8_t* src = sQ + hf * 8192 + qrow * 128 + sec * 32;
    const uint4 a = *reinterpret_cast<const uint4*>(src), b = *reinterpret_cast<const uint4*>(src + 16);
    st_global_32B(out + static_cast<long long>((hd * HD) / 64 + hf) * kSlotBytes + (row0 + qrow) * 128 + sec * 32, a, b);
  }
  epi_bar();  // tiles (aliasing the A ring) dead
}

// One WARP per (sequence, head) unit when a block is at most 16 tokens and head_dim is 64 (the ImageNet models on a small
// grid, kStreamFamHeadSmall): Q, K and V are 16 rows of 128 bytes each, copied by the warp itself (the rows keep the blocked
// buffer's swizzle), one m16 query tile x 16 keys; no block barrier, so a CTA has 4 units in flight. Same arithmetic as
// attn_unit. wsm: the warp's 6 KB of the A ring.
__device__ __forceinline__ void attn_unit_warp64(const StreamOp& op, int unit, int lane, uint8_t* wsm) {
  constexpr int HD = 64;
  const int D = op.N, pn = op.i0, H = D / HD;
  const int seq = unit / H, hd = unit % H;
  const int row0 = seq * pn;
  uint8_t* sQ = wsm;
  uint8_t* sK = wsm + 2048;
  uint8_t* sV = wsm + 4096;
  const uint8_t* qkv = reinterpret_cast<const uint8_t*>(op.p0);
  __syncwarp();  // the previous unit's reads of these tiles are complete
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    const int kb = (which * D + hd * HD) / 64;
    const uint8_t* src = qkv + static_cast<long long>(kb) * kSlotBytes + row0 * 128;
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane + 32 * k, r = i >> 3;
      v[k] = r < pn ? ldcg_u4(src + r * 128 + (i & 7) * 16) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane + 32 * k;
      *reinterpret_cast<uint4*>(wsm + which * 2048 + (i >> 3) * 128 + (i & 7) * 16) = v[k];
    }
  }
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
  const float scale_log2 = rsqrtf(static_cast<float>(HD)) * 1.4426950408889634f;
  float s[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ks += 2) {
    uint32_t a0[4], a1[4], bk[2][4];
    s_ldmatrix_x4(a0, smem_u32(sQ + s_tile_off(lane & 15, 2 * ks + (lane >> 4), row0)));
    s_ldmatrix_x4(a1, smem_u32(sQ + s_tile_off(lane & 15, 2 * ks + 2 + (lane >> 4), row0)));
#pragma unroll
    for (int j = 0; j < 2; ++j) s_ldmatrix_x4(bk[j], smem_u32(sK + s_tile_off(8 * j + (lane & 7), 2 * ks + (lane >> 3), row0)));
#pragma unroll
    for (int j = 0; j < 2; ++j) s_mma_16816_nv(s[j], a0, bk[j][0], bk[j][1]);
#pragma unroll
    for (int j = 0; j < 2; ++j) s_mma_16816_nv(s[j], a1, bk[j][2], bk[j][3]);
  }
  float mx[2] = {-FLT_MAX, -FLT_MAX};
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int key = 8 * j + 2 * t + (i & 1);
      const float v = key < pn ? s[j][i] * scale_log2 : -FLT_MAX;
      s[j][i] = v;
      mx[i >> 1] = fmaxf(mx[i >> 1], v);
    }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
  float l[2] = {0.f, 0.f};
  uint32_t pa[4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      e[i] = (s[j][i] == -FLT_MAX) ? 0.f : exp2f(s[j][i] - mx[i >> 1]);
      l[i >> 1] += e[i];
    }
    pa[2 * j] = s_pack_bf16(e[0], e[1]);
    pa[2 * j + 1] = s_pack_bf16(e[2], e[3]);
  }
  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  {
    uint32_t bv[HD / 16][4];
#pragma unroll
    for (int n = 0; n < HD / 16; ++n)
      s_ldmatrix_x4_trans(bv[n], smem_u32(sV + s_tile_off((lane & 7) + 8 * ((lane >> 3) & 1), 2 * n + (lane >> 4), row0)));
#pragma unroll
    for (int n = 0; n < HD / 16; ++n) {
      s_mma_16816_nv(o_acc[2 * n], pa, bv[n][0], bv[n][1]);
      s_mma_16816_nv(o_acc[2 * n + 1], pa, bv[n][2], bv[n][3]);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  __syncwarp();  // every lane's ldmatrix of Q is complete: O takes the Q tile's place
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qrow = g + 8 * r;
    const float inv = l[r] > 0.f ? 1.0f / l[r] : 0.f;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n)
      *reinterpret_cast<uint32_t*>(sQ + s_tile_off(qrow, n, row0) + 4 * t) =
          s_pack_bf16(o_acc[n][2 * r] * inv, o_acc[n][2 * r + 1] * inv);
  }
  __syncwarp();
  uint8_t* out = reinterpret_cast<uint8_t*>(op.o0);
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // 16 rows x 4 sectors of 32 bytes
    const int i = lane + 32 * k, rr = i >> 2, sec = i & 3;
    if (rr >= pn) continue;
    const uint8_t* src = sQ + rr * 128 + sec * 32;
    const uint4 a = *reinterpret_cast<const uint4*>(src), b = *reinterpret_cast<const uint4*>(src + 16);
    st_global_32B(out + static_cast<long long>((hd * HD) / 64) * kSlotBytes + (row0 + rr) * 128 + sec * 32, a, b);
  }
}

// The op descriptor is a copy of a kernel parameter: left alone, the compiler re-reads its fields from the constant bank
// with a run-time index wherever they are used (an IMAD + LDC chain in front of every use in the tile loop). pin() makes
// the value opaque, so it stays in a register.
__device__ __forceinline__ int pin(int v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ long long pin(long long v) {
  asm volatile("" : "+l"(v));
  return v;
}
__device__ __forceinline__ float pin(float v) {
  asm volatile("" : "+f"(v));
  return v;
}
template <typename T>
__device__ __forceinline__ T* pin(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}

// ---------------------------------------------------------------------------------------------------------------------
// Qwen3 decoder ops inside the engine (one persistent launch per AR block: csrc/bd_llm.cu::llm_stream_all)
// ---------------------------------------------------------------------------------------------------------------------
// q/k RMSNorm over head_dim + RoPE in fp32 + KV append (the arithmetic of bd_llm.cu::qk_norm_rope_append_kernel<HD, true>;
// transformers qwen3 Qwen3Attention.forward). One warp per (token, head); the 4 x G epilogue warps stride over the tasks.
template <int HD>
__device__ __forceinline__ void llm_rope_append(const StreamOp& op, int it, int c, int G, int tid) {
  constexpr int VPT = HD / 32;
  const int S = pin(op.i0), Hq = pin(op.i1), Hkv = pin(op.i2), heads = pin(Hq + 2 * Hkv), max_pages = pin(op.N);
  const int M = op.i0 * op.sub;  // sub = number of sequences
  const int warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* qkv = pin(reinterpret_cast<const __nv_bfloat16*>(op.p0));
  const __nv_bfloat16* qn_w = reinterpret_cast<const __nv_bfloat16*>(op.p1);
  const __nv_bfloat16* kn_w = reinterpret_cast<const __nv_bfloat16*>(op.p2);
  const float* rope_cos = pin(reinterpret_cast<const float*>(op.p3));
  const float* rope_sin = pin(reinterpret_cast<const float*>(op.p4));
  const int* seq_lens = reinterpret_cast<const int*>(op.p5);
  const int* page_table = pin(reinterpret_cast<const int*>(op.p6));
  __nv_bfloat16* q_out = pin(reinterpret_cast<__nv_bfloat16*>(op.o0));
  __nv_bfloat16* kpool = pin(reinterpret_cast<__nv_bfloat16*>(op.o1) + static_cast<long long>(it) * op.l0);
  __nv_bfloat16* vpool = kpool + op.l1;
  const float eps = pin(op.f0);
  float qw[VPT], kw[VPT];  // the lane's slice of the q / k norm weights
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    qw[j] = __bfloat162float(qn_w[lane * VPT + j]);
    kw[j] = __bfloat162float(kn_w[lane * VPT + j]);
  }
  const int tasks = M * heads;  // (token, head) pairs; 32-bit index arithmetic (a 64-bit division is a ~150-instruction call)
  const int stride = G * 4;
  const int d_m = stride / heads, d_h = stride % heads;  // task + stride = (token + d_m, head + d_h) with one carry
  constexpr int U = 6;  // tasks in flight per warp (12 tasks per warp at the 14B shape): two rounds of one L2 round trip
  // the sequence lengths once per warp (lane b holds sequence b's; more than 32 sequences: read per task)
  const int R = op.sub;
  const int len_lane = (lane < R) ? __ldcg(seq_lens + lane) : 0;
  for (int gw0 = c * 4 + warp; gw0 < tasks; gw0 += stride * U) {
    float x[U][VPT], cs[U][VPT], sn[U][VPT];
    int pos[U], page[U], tm[U], th[U];
    bool live[U];
    int m_u = gw0 / heads, h_u = gw0 % heads;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int gw = gw0 + u * stride;
      live[u] = gw < tasks;
      pos[u] = 0;
      page[u] = 0;
      tm[u] = m_u;
      th[u] = h_u;
      m_u += d_m;
      h_u += d_h;
      if (h_u >= heads) {
        h_u -= heads;
        ++m_u;
      }
#pragma unroll
      for (int j = 0; j < VPT; ++j) x[u][j] = cs[u][j] = sn[u][j] = 0.f;
      if (!live[u]) continue;
      const int m = tm[u], hh = th[u];
      const int b = m / S, sidx = m % S;
      pos[u] = (R <= 32 ? __shfl_sync(0xffffffffu, len_lane, b) : __ldcg(seq_lens + b)) + sidx;
      if (pos[u] < 0 || pos[u] >= max_pages * 64) __trap();  // outside the KV cache (the host checks the bound too)
      const __nv_bfloat16* src = qkv + static_cast<long long>(m) * heads * HD + static_cast<long long>(hh) * HD + lane * VPT;
      if constexpr (VPT == 4) {
        const uint2 raw = __ldcg(reinterpret_cast<const uint2*>(src));
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
        const float2 a = __bfloat1622float2(p2[0]), bb = __bfloat1622float2(p2[1]);
        x[u][0] = a.x; x[u][1] = a.y; x[u][2] = bb.x; x[u][3] = bb.y;
      } else {
        const uint32_t raw = __ldcg(reinterpret_cast<const uint32_t*>(src));
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
        x[u][0] = a.x; x[u][1] = a.y;
      }
      if (hh >= Hq) page[u] = __ldg(page_table + b * max_pages + pos[u] / 64);
      if (hh < Hq + Hkv) {
        const long long ro = static_cast<long long>(pos[u]) * HD + lane * VPT;
        if constexpr (VPT == 4) {
          const float4 cv = __ldg(reinterpret_cast<const float4*>(rope_cos + ro));
          const float4 sv = __ldg(reinterpret_cast<const float4*>(rope_sin + ro));
          cs[u][0] = cv.x; cs[u][1] = cv.y; cs[u][2] = cv.z; cs[u][3] = cv.w;
          sn[u][0] = sv.x; sn[u][1] = sv.y; sn[u][2] = sv.z; sn[u][3] = sv.w;
        } else {
          const float2 cv = __ldg(reinterpret_cast<const float2*>(rope_cos + ro));
          const float2 sv = __ldg(reinterpret_cast<const float2*>(rope_sin + ro));
          cs[u][0] = cv.x; cs[u][1] = cv.y;
          sn[u][0] = sv.x; sn[u][1] = sv.y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;   // warp-uniform
      const int m = tm[u], hh = th[u];
      const bool is_q = hh < Hq, is_k = !is_q && hh < Hq + Hkv;
      __nv_bfloat16* dst;
      if (is_q) {
        dst = q_out + static_cast<long long>(m) * Hq * HD + static_cast<long long>(hh) * HD;
      } else {
        const int hk = is_k ? hh - Hq : hh - Hq - Hkv;
        dst = (is_k ? kpool : vpool) + ((static_cast<long long>(page[u]) * Hkv + hk) * 64 + (pos[u] % 64)) * HD;
      }
      float y[VPT];
      if (!is_q && !is_k) {  // V: plain copy
#pragma unroll
        for (int j = 0; j < VPT; ++j) y[j] = x[u][j];
      } else {
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) ss += x[u][j] * x[u][j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss / static_cast<float>(HD) + eps);
        float xn[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) xn[j] = bf16_round((is_q ? qw[j] : kw[j]) * bf16_round(x[u][j] * rstd));
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const float other = __shfl_xor_sync(0xffffffffu, xn[j], 16);
          const float rot = (lane < 16) ? -other : other;  // rotate_half: cat(-x2, x1)
          y[j] = __fadd_rn(__fmul_rn(xn[j], cs[u][j]), __fmul_rn(rot, sn[u][j]));
        }
      }
      if constexpr (VPT == 4) {
        uint2 pk;
        __nv_bfloat162 a = __floats2bfloat162_rn(y[0], y[1]), bb = __floats2bfloat162_rn(y[2], y[3]);
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&bb);
        *reinterpret_cast<uint2*>(dst + lane * VPT) = pk;
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(y[0], y[1]);
        *reinterpret_cast<uint32_t*>(dst + lane * VPT) = *reinterpret_cast<uint32_t*>(&a);
      }
    }
  }
}

// 16-byte async copies (LDGSTS): the K / V tiles of the NEXT key tile land in shared memory while the current one is used
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t n = valid ? 16u : 0u;  // src-size 0 = zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A 64-row tile of HD bf16 per row in shared memory, 16-byte chunks XOR-swizzled by (row & 7): conflict-free ldmatrix
template <int HD>
__device__ __forceinline__ uint32_t ltile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * (HD * 2) + ((chunk ^ (row & 7)) << 4));
}
template <int HD>
__device__ __forceinline__ void ltile_load_async(uint8_t* tile, const __nv_bfloat16* src, long long row_stride, int valid_rows,
                                                 int tid) {
  constexpr int kChunks = HD / 8;
  for (int i = tid; i < 64 * kChunks; i += 128) {
    const int r = i / kChunks, ch = i % kChunks;
    const bool ok = r < valid_rows;
    cp_async16(tile + ltile_off<HD>(r, ch), src + (ok ? r : 0) * row_stride + ch * 8, ok);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention of an AR block over the paged cache, split "stream-K" style. The work is the flattened list of
// (sequence b, q head h, 64-key tile kt) triples — P = Hq * sum_b ceil(Sk_b / 64) of them, h-major inside a sequence so
// that the q heads of one kv head are neighbours in time and share its pages in L2. CTA c < Ge = min(op.act, P) takes the
// contiguous range [c P / Ge, (c + 1) P / Ge): every CTA gets the same number of tiles (+-1) whatever the lengths are (they
// are device data: no host planning, one captured graph for every step of an image), and a (b, h) block is cut into at
// most op.ksplit segments, one per CTA that touches it; each segment writes its unnormalised partial (O, max, sum) to slot
// (b Hq + h) * ksplit + (c - first CTA of the block), which the combine row op reduces in slot order (the same arithmetic
// on both sides: llm_attn_cta_of).
// The 4 executor warps own 16 query rows each (S <= 64; Q fragments live in registers, loaded straight from global memory
// when a block starts); K / V tiles are double-buffered in the A ring with cp.async (one barrier per tile, the next tile
// and the page index after it in flight); fp32 scores / softmax, P in bf16, like bd_attn.cu.
// The op is bound by instruction issue — one warp per scheduler, so nothing hides a dependent instruction's latency
// (measured: no time waiting for K/V; profiles/r02_llm_timeline.txt) — hence: per-lane shared-memory offsets and the cp.async
// source / destination offsets are computed once per op; key masking only in the last tile of a sequence; one FFMA + ex2 per
// score; the accumulator rescale is skipped when no row maximum of the warp moved; independent mma chains are interleaved.
// ---------------------------------------------------------------------------------------------------------------------
struct LlmAttnCur {
  int b, h, kt, T, Sk, base;  // base: flattened index of (b, 0, 0)
  int hk, hg;                 // kv head of h, position of h in its group
};
__device__ __forceinline__ void llm_attn_advance(LlmAttnCur& k, const int* seq_lens, int R, int S, int Hq, int gqa) {
  if (++k.kt < k.T) return;
  k.kt = 0;
  if (++k.hg == gqa) {
    k.hg = 0;
    ++k.hk;
  }
  if (++k.h < Hq) return;
  k.h = 0;
  k.hk = 0;
  k.base += Hq * k.T;
  if (++k.b < R) {
    k.Sk = __ldcg(seq_lens + k.b) + S;
    k.T = (k.Sk + 63) >> 6;
  }
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_sz(uint32_t smem_dst, const void* gsrc, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
__device__ __forceinline__ void llm_attn_stream(const StreamOp& op, int it, int c, int tid, uint8_t* smem, int dbg_mode,
                                                unsigned long long* dbg_slot) {
  constexpr int kTile = 64 * HD * 2;  // one K (or V) tile
  constexpr int kStage = 2 * kTile;
  constexpr int kChunks = HD / 8;     // 16-byte chunks per row
  constexpr int kCopies = 64 * kChunks / 128;  // cp.async per thread per tile
  constexpr int kRowStep = 128 / kChunks;      // rows between two copies of a thread
  static_assert(kRowStep % 8 == 0, "the swizzle phase of a thread's rows must be constant");
  const int R = pin(op.sub), S = pin(op.i0), Hq = pin(op.i1), Hkv = pin(op.i2), maxseg = op.ksplit, max_pages = pin(op.N);
  const int gqa = pin(Hq / Hkv);
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(op.p0);
  const int* seq_lens = pin(reinterpret_cast<const int*>(op.p1));
  const int* page_table = pin(reinterpret_cast<const int*>(op.p2));
  const __nv_bfloat16* kpool = pin(reinterpret_cast<const __nv_bfloat16*>(op.p3) + static_cast<long long>(it) * op.l0);
  const long long v_off = pin(op.l1);
  float* part_o = reinterpret_cast<float*>(op.o0);
  float* part_ml = reinterpret_cast<float*>(op.o1);
  const float scale_log2 = pin(op.f0);
  int P = 0;
  for (int b = 0; b < R; ++b) P += Hq * ((__ldcg(seq_lens + b) + S + 63) >> 6);
  const int Ge = min(op.act, P);
  if (c >= Ge) return;
  const int lo = static_cast<int>(static_cast<unsigned>(c) * static_cast<unsigned>(P) / static_cast<unsigned>(Ge));
  const int hi = static_cast<int>(static_cast<unsigned>(c + 1) * static_cast<unsigned>(P) / static_cast<unsigned>(Ge));
  LlmAttnCur cu;  // compute cursor
  {
    int base = 0;
    for (int b = 0;; ++b) {
      const int Sk = __ldcg(seq_lens + b) + S, T = (Sk + 63) >> 6;
      if (lo - base < Hq * T || b == R - 1) {
        const int h = (lo - base) / T;
        cu = LlmAttnCur{b, h, (lo - base) % T, T, Sk, base, h / gqa, h % gqa};
        break;
      }
      base += Hq * T;
    }
  }
  LlmAttnCur ld = cu;  // load cursor, one tile ahead
  int f_ld = lo;
  int pg = __ldg(page_table + ld.b * max_pages + ld.kt);
  // measurement switches (bd_stream_set_tuning mode; scripts/llm_timeline.py): 64 no K/V loads, 128 no tensor work / softmax
  const bool no_loads = (dbg_mode & 64) != 0, no_math = (dbg_mode & 128) != 0;
  long long t_wait = 0, t_math = 0;
  // per-thread constants of the tile copies: chunk ch of rows r0 + kRowStep * k
  const int cp_r0 = tid / kChunks, cp_ch = tid % kChunks;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t cp_dst0 = static_cast<uint32_t>(cp_r0 * (HD * 2) + ((cp_ch ^ (cp_r0 & 7)) << 4));
  const int cp_src0 = cp_r0 * HD + cp_ch * 8;
  auto issue = [&](int buf) {
    if (f_ld < hi) {
      const __nv_bfloat16* src = kpool + (static_cast<long long>(pg) * Hkv + ld.hk) * (64 * HD) + cp_src0;
      const uint32_t dst = smem_base + buf * kStage + cp_dst0;
      const int valid = ld.Sk - ld.kt * 64;
      if (!no_loads) {
        if (valid >= 64) {
#pragma unroll
          for (int k = 0; k < kCopies; ++k) {
            cp_async16_s(dst + k * (kRowStep * HD * 2), src + k * (kRowStep * HD));
            cp_async16_s(dst + kTile + k * (kRowStep * HD * 2), src + v_off + k * (kRowStep * HD));
          }
        } else {  // the last tile of a sequence: rows past the end are zero-filled
#pragma unroll
          for (int k = 0; k < kCopies; ++k) {
            const bool ok = cp_r0 + k * kRowStep < valid;
            const __nv_bfloat16* sk = ok ? src + k * (kRowStep * HD) : src - cp_r0 * HD;
            cp_async16_sz(dst + k * (kRowStep * HD * 2), sk, ok);
            cp_async16_sz(dst + kTile + k * (kRowStep * HD * 2), sk + v_off, ok);
          }
        }
      }
      llm_attn_advance(ld, seq_lens, R, S, Hq, gqa);
      if (++f_ld < hi) pg = __ldg(page_table + ld.b * max_pages + ld.kt);  // consumed one tile later
    }
    cp_async_commit();
  };
  issue(0);
  // per-lane ldmatrix offsets inside a tile (the swizzle phase is the lane's row & 7)
  uint32_t k_off[HD / 32], v_offs[HD / 16];
#pragma unroll
  for (int i = 0; i < HD / 32; ++i) k_off[i] = ltile_off<HD>(lane & 7, 4 * i + (lane >> 3));
#pragma unroll
  for (int i = 0; i < HD / 16; ++i) v_offs[i] = ltile_off<HD>((lane & 7) + 8 * ((lane >> 3) & 1), 2 * i + (lane >> 4));
  uint32_t qf[HD / 16][4];
  float o_acc[HD / 8][4];
  float m_run[2], l_run[2];  // running maximum of the RAW scores, running sum of exp2((s - m) * scale_log2) per thread
  bool fresh = true;
  const int n_st = hi - lo;
  for (int i = 0; i < n_st; ++i) {
    if (fresh) {  // a (b, h) block starts (or continues from another CTA): Q fragments, empty accumulators
      fresh = false;
      const __nv_bfloat16* qb = q + (static_cast<long long>(cu.b) * S * Hq + cu.h) * HD;
      const int r0 = warp * 16 + g, r1 = r0 + 8;
      const unsigned int* p0 = reinterpret_cast<const unsigned int*>(qb + static_cast<long long>(r0) * Hq * HD + 2 * t);
      const unsigned int* p1 = reinterpret_cast<const unsigned int*>(qb + static_cast<long long>(r1) * Hq * HD + 2 * t);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        qf[ks][0] = r0 < S ? __ldcg(p0 + 8 * ks) : 0u;
        qf[ks][1] = r1 < S ? __ldcg(p1 + 8 * ks) : 0u;
        qf[ks][2] = r0 < S ? __ldcg(p0 + 8 * ks + 4) : 0u;
        qf[ks][3] = r1 < S ? __ldcg(p1 + 8 * ks + 4) : 0u;
      }
#pragma unroll
      for (int n = 0; n < HD / 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) o_acc[n][j] = 0.f;
      m_run[0] = m_run[1] = -FLT_MAX;
      l_run[0] = l_run[1] = 0.f;
    }
    const long long tw0 = dbg_slot ? clock64() : 0;
    epi_bar();           // every warp is done with tile i - 1: its buffer can be refilled
    issue((i + 1) & 1);
    cp_async_wait<1>();  // tile i landed (tile i + 1 in flight)
    epi_bar();           // ... for every thread's share
    const long long tw1 = dbg_slot ? clock64() : 0;
    t_wait += tw1 - tw0;
    if (no_math) {
      llm_attn_advance(cu, seq_lens, R, S, Hq, gqa);
      continue;
    }
    const uint32_t sK = smem_base + (i & 1) * kStage;
    const uint32_t sV = sK + kTile;
    const int k0 = cu.kt * 64;
    // two halves of 32 keys, each QK^T -> online softmax -> P V: half the score / P registers of a whole-tile pass, which
    // is what lets the compiler keep loads of the next fragments in flight (255 registers: 64 of O, 32 of Q)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float sc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[j][e] = 0.f;
#pragma unroll
      for (int kq = 0; kq < HD / 32; ++kq) {  // 4 key blocks = 4 independent chains, 2 k-steps each
        uint32_t bk[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) s_ldmatrix_x4(bk[j], sK + k_off[kq] + (4 * half + j) * (8 * HD * 2));
#pragma unroll
        for (int j = 0; j < 4; ++j) s_mma_16816_nv(sc[j], qf[2 * kq], bk[j][0], bk[j][1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) s_mma_16816_nv(sc[j], qf[2 * kq + 1], bk[j][2], bk[j][3]);
      }
      if (k0 + 32 * half + 32 > cu.Sk) {  // the last tile of the sequence: keys past the end do not count
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (k0 + 32 * half + 8 * j + 2 * t + (e & 1) >= cu.Sk) sc[j][e] = -FLT_MAX;
      }
      float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) m_new[e >> 1] = fmaxf(m_new[e >> 1], sc[j][e]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
        m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
      }
      if (__any_sync(0xffffffffu, (m_new[0] > m_run[0]) || (m_new[1] > m_run[1]))) {  // a maximum moved: rescale
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          corr[r] = ex2_approx((m_run[r] - m_new[r]) * scale_log2);
          l_run[r] *= corr[r];
          m_run[r] = m_new[r];
        }
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
          o_acc[n][0] *= corr[0];
          o_acc[n][1] *= corr[0];
          o_acc[n][2] *= corr[1];
          o_acc[n][3] *= corr[1];
        }
      }
      const float ms[2] = {m_run[0] * scale_log2, m_run[1] * scale_log2};
      uint32_t pa[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          e[x] = ex2_approx(fmaf(sc[j][x], scale_log2, -ms[x >> 1]));
          l_run[x >> 1] += e[x];
        }
        pa[j >> 1][(j & 1) * 2] = s_pack_bf16(e[0], e[1]);
        pa[j >> 1][(j & 1) * 2 + 1] = s_pack_bf16(e[2], e[3]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
        for (int nh = 0; nh < HD / 16; nh += 4) {  // 8 dim blocks at a time
          uint32_t bv[4][4];
#pragma unroll
          for (int n = 0; n < 4; ++n)
            s_ldmatrix_x4_trans(bv[n], sV + v_offs[nh + n] + (2 * half + kk) * (16 * HD * 2));
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            s_mma_16816_nv(o_acc[2 * (nh + n)], pa[kk], bv[n][0], bv[n][1]);
            s_mma_16816_nv(o_acc[2 * (nh + n) + 1], pa[kk], bv[n][2], bv[n][3]);
          }
        }
      }
    }
    if (cu.kt == cu.T - 1 || i == n_st - 1) {  // the block (or this CTA's share of it) ends: store the segment
      fresh = true;
      const int hb = cu.b * Hq + cu.h;
      const int seg = c - llm_attn_cta_of(cu.base + cu.h * cu.T, Ge, P);
      if (seg < 0 || seg >= maxseg) __trap();
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const int qrow = warp * 16 + g + 8 * r;
        if (qrow >= S) continue;
        const long long row = (static_cast<long long>(hb) * maxseg + seg) * S + qrow;
        float* o = part_o + row * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n)
          *reinterpret_cast<float2*>(o + 8 * n + 2 * t) = make_float2(o_acc[n][2 * r], o_acc[n][2 * r + 1]);
        // the combine works in the exp2 domain of bd_attn.cu: maximum of the SCALED scores
        if (t == 0) *reinterpret_cast<float2*>(part_ml + row * 2) = make_float2(m_run[r] * scale_log2, l);
      }
    }
    llm_attn_advance(cu, seq_lens, R, S, Hq, gqa);
    if (dbg_slot) t_math += clock64() - tw1;
  }
  cp_async_wait<0>();
  epi_bar();
  if (dbg_slot && tid == 0) {  // timeline slots 6 / 7 of this op: cycles waiting for K/V, cycles from landed to next wait
    dbg_slot[6] = static_cast<unsigned long long>(t_wait);
    dbg_slot[7] = static_cast<unsigned long long>(t_math);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kAccBufs = 3;            // TMEM accumulator buffers: 0 / 1 alternate over the passes of the dependent chain,
constexpr uint32_t kTmemCols = 512;    //   2 collects the pieces of a filler pass (128 columns each; 512 allocated)

struct StreamSmem {
  static constexpr int kRing = kStreamSlots * kStepBytes;
  // rings, accumulators, the epilogue warps' own bulk-copy barrier, + the scratch handshake (A ring lent to an executor)
  static constexpr int kBars = (2 * kStreamSlots + 2 * kAccBufs + 1 + 2) * 8;
  static constexpr int kTmemSlotOff = 192, kRedOff = 200;
  static constexpr int kBias = kAccBufs * kAccStride * 4;  // one pass's bias slice (<= 160 values, fp32) per accumulator buffer
  static constexpr int kMisc = 256;      // barriers (<= 168 B), TMEM slot at +192, reduction scratch at +208
  static_assert(kBars <= 192, "barrier area");
  // 227 KB is the most a CTA can have: 896 bytes are left for aligning the ring to 1024 (128B-swizzle atom). The dynamic
  // shared-memory window of a kernel without static shared memory starts 1024-aligned; the kernel checks and traps otherwise.
  static constexpr int kAlignSlack = 896;
  static constexpr int kTotal = kRing + kMisc + kBias + kAlignSlack;
  static_assert(kTotal <= 227 * 1024, "shared memory");
};

struct RingPos {  // position in a ring of n slots: slot index + how many times the ring wrapped
  uint32_t slot = 0, round = 0, n;
  __device__ explicit RingPos(uint32_t n_) : n(n_) {}
  __device__ __forceinline__ void advance() {
    if (++slot == n) {
      slot = 0;
      ++round;
    }
  }
};

__device__ __forceinline__ void op_at(const StreamProgram& prog, int q, int& idx, int& it) {
  if (q < prog.n_pre) {
    idx = q;
    it = 0;
    return;
  }
  const int b = q - prog.n_pre;
  const int nb = prog.n_body * prog.n_iter;
  if (b < nb) {
    it = b / prog.n_body;
    idx = prog.n_pre + b % prog.n_body;
    return;
  }
  idx = prog.n_pre + prog.n_body + (b - nb);
  it = prog.n_iter - 1;
}

// Walks the program in order. mseq = the op's sequence number in the grid-barrier protocol = number of NON-filler ops
// before it (fillers do not arrive, so they must not be counted by those who wait).
struct OpCursor {
  const StreamProgram& prog;
  int total;
  int q = -1, idx = 0, it = 0;
  unsigned int mseq = 0;
  __device__ OpCursor(const StreamProgram& p, int total_) : prog(p), total(total_) {}
  __device__ __forceinline__ bool next() {
    if (q >= 0 && !(prog.ops[idx].flags & kFlagFiller)) ++mseq;
    if (++q >= total) return false;
    op_at(prog, q, idx, it);
    return true;
  }
};

// The part of a GEMM op one CTA works on: passes [p0, p1) of its share, k-blocks [kb_first, kb_first + kbn) of the split.
struct GemmWork {
  StreamPart part;
  int p0, p1, kb_first, kbn;
  bool piece, first, last;
  __device__ __forceinline__ bool any() const { return p1 > p0; }
};
__device__ __forceinline__ GemmWork gemm_work(const StreamOp& op, int G, int c) {
  GemmWork w;
  w.part = stream_partition(op.N, op.K, op.ksplit, G, c);
  w.piece = op.pc_kbn > 0;
  if (w.piece) {
    const bool has = w.part.units > 0 && op.pc_pass < w.part.npass;
    w.p0 = op.pc_pass;
    w.p1 = has ? op.pc_pass + 1 : op.pc_pass;
    w.kb_first = op.pc_kb0;
    w.kbn = op.pc_kbn;
    w.first = (op.pc_flags & kPieceFirst) != 0;
    w.last = (op.pc_flags & kPieceLast) != 0;
  } else {
    w.p0 = 0;
    w.p1 = w.part.units > 0 ? w.part.npass : 0;
    w.kb_first = 0;
    w.kbn = w.part.kbs;
    w.first = w.last = true;
  }
  return w;
}
__device__ __forceinline__ bool op_skipped(const StreamProgram& prog, const StreamOp& op, int it) {
  return (op.flags & kFlagSkipLast) && it == prog.n_iter - 1;
}

__device__ __forceinline__ const void* tab_ptr(const StreamOp& op, int it, int slot1, const void* dflt) {
  if (!op.tab || slot1 == 0) return dflt;
  return op.tab[static_cast<long long>(it + op.tab_off) * kTabSlots + (slot1 - 1)];
}

// The attention executor and the final-row executor borrow the A ring as scratch. Without fillers nothing can be in flight
// there (the A producer is parked at the next GEMM's grid barrier); with fillers — GEMM work without dependencies — the A
// producer could be streaming a piece's operand into the ring at that very moment (always in a CTA that has no share of
// the GEMM before the op, e.g. small models). So the ring is handed over explicitly: the A producer drains it, signals
// scr_free, and resumes only after the executors signal scr_done.
__device__ __forceinline__ bool op_uses_scratch(const StreamProgram& prog, const StreamOp& op, int c) {
  if (op.kind == kOpAttn) return c < (prog.M / op.i0) * (op.N / op.K);
  if (op.kind == kOpLlmAttn) return c < op.act;
  return op.kind == kOpRow && op.sub == kRowFinal && c < prog.M;
}

// The weight stream of one CTA as a flat sequence of steps (op after op, pass after pass, k rotation inside a pass): two
// cursors walk it — the L2 prefetch cursor a fixed distance ahead of the shared-memory load cursor.
struct WStream {
  const StreamProgram& prog;
  int G, c, total;
  int q = -1, i = 0, t = 0;           // op sequence number, pass, step
  int pass_end = 0, nsteps = 0, rot = 0, kbn = 0, kb_first = 0, N = 0, kps = kKbPerStep;
  StreamPart part{};
  const uint8_t* w = nullptr;
  const uint8_t* base = nullptr;      // first slot of the current pass
  uint32_t kb_bytes = 0;
  __device__ WStream(const StreamProgram& p, int G_, int c_, int total_) : prog(p), G(G_), c(c_), total(total_) {}
  __device__ __forceinline__ bool next_op() {
    for (++q; q < total; ++q) {
      int idx, it;
      op_at(prog, q, idx, it);
      const StreamOp& op = prog.ops[idx];
      if (op.kind != kOpGemm) continue;
      if (op_skipped(prog, op, it)) continue;
      const GemmWork gw = gemm_work(op, G, c);
      if (!gw.any()) continue;
      part = gw.part;
      N = op.N;
      w = reinterpret_cast<const uint8_t*>(tab_ptr(op, it, op.tab_p0, op.p0));
      i = gw.p0;
      pass_end = gw.p1;
      kb_first = gw.kb_first;
      kbn = gw.kbn;
      t = 0;
      set_pass();
      return true;
    }
    return false;
  }
  __device__ __forceinline__ void set_pass() {
    const int wd = (stream_pass_u0(part, i + 1) - stream_pass_u0(part, i)) * 16;
    kb_bytes = static_cast<uint32_t>(wd) * 128u;
    base = w + stream_pass_offset(N, part, i) * 2048;
    kps = stream_kps(wd);
    nsteps = stream_steps(kbn, kps);
    rot = stream_k_rot(c, nsteps, prog.dbg_mode);
  }
  // address / size of the next step; false at the end of the program
  __device__ __forceinline__ bool next(const uint8_t*& addr, uint32_t& bytes) {
    if (q < 0 || t >= nsteps) {
      if (q >= 0 && i + 1 < pass_end) {
        ++i;
        t = 0;
        set_pass();
      } else if (!next_op()) {
        return false;
      }
    }
    int st = rot + t;
    if (st >= nsteps) st -= nsteps;
    const int kbl = st * kps;
    const int nkb = min(kps, kbn - kbl);
    bytes = kb_bytes * static_cast<uint32_t>(nkb);  // the k-blocks of a pass are adjacent in HBM
    addr = base + static_cast<long long>(kb_first + kbl) * kb_bytes;
    ++t;
    return true;
  }
};

// wait_prev: 1 = all earlier (non-filler) ops, 0 = none, -k = all earlier ops except the k most recent ones
__device__ __forceinline__ unsigned int wait_target(int wait_prev, unsigned int mseq, int G) {
  const unsigned int n = wait_prev < 0 ? (mseq > static_cast<unsigned int>(-wait_prev) ? mseq + wait_prev : 0u) : mseq;
  return static_cast<unsigned int>(G) * n;
}

#define BD_STAMP(q_, e_)                                                                             \
  do {                                                                                               \
    if (prog.dbg && (q_) < prog.dbg_ops) prog.dbg[(static_cast<long long>(q_) * G + c) * 8 + (e_)] = gtimer(); \
  } while (0)

template <int FAM>
__global__ void __launch_bounds__(kStreamThreads, 1) bd_stream_kernel(const __grid_constant__ StreamProgram prog) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && static_cast<int>(smem - smem_raw) > StreamSmem::kAlignSlack) {
    printf("bd_stream: dynamic shared memory base %p is not aligned enough\n", static_cast<void*>(smem_raw));
    __trap();
  }
  const uint32_t kStreamWSlots = static_cast<uint32_t>(prog.w_slots), kStreamASlots = static_cast<uint32_t>(prog.a_slots);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kStreamWSlots * kStepBytes;
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem + StreamSmem::kRing);
  uint64_t* empty_w = full_w + kStreamWSlots;
  uint64_t* full_a = empty_w + kStreamWSlots;
  uint64_t* empty_a = full_a + kStreamASlots;
  uint64_t* acc_full = empty_a + kStreamASlots;
  uint64_t* acc_empty = acc_full + kAccBufs;
  uint64_t* aux_bar = acc_empty + kAccBufs;
  uint64_t* scr_free = aux_bar + 1;   // A producer -> executors: the A ring is drained and stays untouched
  uint64_t* scr_done = aux_bar + 2;   // executors -> A producer: scratch use finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + StreamSmem::kRing + 192);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);
  float* bias_s = reinterpret_cast<float*>(smem + StreamSmem::kRing + StreamSmem::kMisc);  // 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int c = blockIdx.x;
  const int G = prog.n_ctas;
  const int total = prog.n_pre + prog.n_body * prog.n_iter + prog.n_post;

  if (warp == 0 && elect_one()) {
    for (uint32_t s = 0; s < kStreamWSlots; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&empty_w[s], 1);
    }
    for (uint32_t s = 0; s < kStreamASlots; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < kAccBufs; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_init(aux_bar, 1);
    mbar_init(scr_free, 1);
    mbar_init(scr_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== W producer: runs ahead of every dependency =====================
    if (elect_one()) {
      RingPos wr(kStreamWSlots);
      WStream ld(prog, G, c, total), pf(prog, G, c, total);
      const uint8_t* addr;
      uint32_t bytes;
      // optional L2 prefetch cursor `pf_steps` steps ahead of the load cursor (measured slower: off by default)
      bool pf_live = prog.pf_steps > 0;
      for (int k = 0; pf_live && k < prog.pf_steps; ++k) {
        pf_live = pf.next(addr, bytes);
        if (pf_live) bulk_prefetch_l2(addr, bytes);
      }
      while (ld.next(addr, bytes)) {
        if (pf_live) {
          const uint8_t* pa;
          uint32_t pb;
          pf_live = pf.next(pa, pb);
          if (pf_live) bulk_prefetch_l2(pa, pb);
        }
        const uint32_t s = wr.slot;
        if (wr.round > 0) mbar_wait(&empty_w[s], (wr.round & 1u) ^ 1u);
        mbar_expect_tx(&full_w[s], bytes);
        bulk_g2s(smem_w + s * kStepBytes, addr, bytes, &full_w[s], kEvictFirst);
        wr.advance();
      }
    }
  } else if (warp == 2) {
    // ===================== A producer: waits for the op's grid-wide dependency, then streams the blocked activations
    if (elect_one()) {
      RingPos ar(kStreamASlots);
      OpCursor cur(prog, total);
      uint32_t scr_uses = 0;
      while (cur.next()) {
        const StreamOp& op = prog.ops[cur.idx];
        if (op.kind != kOpGemm) {
          if (op_uses_scratch(prog, op, c)) {
            // drain: every slot's latest fill has been consumed by the MMAs (slots before ar.slot were filled in round
            // ar.round, the others one round earlier)
            for (uint32_t s2 = 0; s2 < kStreamASlots; ++s2) {
              if (s2 < ar.slot) mbar_wait(&empty_a[s2], ar.round & 1u);
              else if (ar.round > 0) mbar_wait(&empty_a[s2], (ar.round & 1u) ^ 1u);
            }
            mbar_arrive(scr_free);
            mbar_wait(scr_done, scr_uses & 1u);
            ++scr_uses;
          }
          continue;
        }
        if (op_skipped(prog, op, cur.it)) continue;
        const GemmWork gw = gemm_work(op, G, c);
        if (!gw.any()) continue;
        if (op.wait_prev) {
          grid_wait(prog.sync, wait_target(op.wait_prev, cur.mseq, G), static_cast<unsigned int>(prog.poll_ns));
          fence_proxy_async_all();  // other CTAs' generic-proxy stores -> this thread's async-proxy (bulk copy) reads
        }
        BD_STAMP(cur.q, 0);
        const uint8_t* abase = reinterpret_cast<const uint8_t*>(op.p1);
        if (op.flags & kFlagAPerIt) abase += static_cast<long long>(cur.it + op.i0) * ((op.K + 63) / 64) * kSlotBytes;
        abase += static_cast<long long>(gw.part.kb0 + gw.kb_first) * kSlotBytes;
        for (int i = gw.p0; i < gw.p1; ++i) {
          const int kps = stream_kps((stream_pass_u0(gw.part, i + 1) - stream_pass_u0(gw.part, i)) * 16);
          const int nsteps = stream_steps(gw.kbn, kps);
          const int rot = stream_k_rot(c, nsteps, prog.dbg_mode);
          for (int t = 0; t < nsteps; ++t) {
            int st = rot + t;
            if (st >= nsteps) st -= nsteps;
            const int kbl = st * kps;
            const uint32_t bytes = static_cast<uint32_t>(min(kps, gw.kbn - kbl)) * kSlotBytes;
            const uint32_t s = ar.slot;
            if (ar.round > 0) mbar_wait(&empty_a[s], (ar.round & 1u) ^ 1u);
            mbar_expect_tx(&full_a[s], bytes);
            bulk_g2s(smem_a + s * kStepBytes, abase + static_cast<long long>(kbl) * kSlotBytes, bytes, &full_a[s],
                     kEvictLast);
            ar.advance();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      RingPos wr(kStreamWSlots), ar(kStreamASlots);
      uint32_t pi = 0, pf = 0;  // passes of the dependent chain (buffers 0 / 1), completed filler accumulations (buffer 2)
      OpCursor cur(prog, total);
      while (cur.next()) {
        const StreamOp& op = prog.ops[cur.idx];
        if (op.kind != kOpGemm) continue;
        if (op_skipped(prog, op, cur.it)) continue;
        const GemmWork gw = gemm_work(op, G, c);
        if (!gw.any()) continue;
        const int q = cur.q;
        for (int i = gw.p0; i < gw.p1; ++i) {
          const int w = (stream_pass_u0(gw.part, i + 1) - stream_pass_u0(gw.part, i)) * 16;
          const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(w));
          const uint32_t buf = gw.piece ? 2u : (pi & 1u);
          if (gw.piece) {
            if (gw.first && pf >= 1) mbar_wait(&acc_empty[2], (pf & 1u) ^ 1u);
          } else if (pi >= 2) {
            mbar_wait(&acc_empty[buf], ((pi >> 1) & 1u) ^ 1u);
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * static_cast<uint32_t>(kAccStride);
          const int kps = stream_kps(w);
          const int nsteps = stream_steps(gw.kbn, kps);
          const int rot = stream_k_rot(c, nsteps, prog.dbg_mode);
          const uint32_t kb_bytes = static_cast<uint32_t>(w) * 128u;
          const uint32_t keep = gw.first ? 0u : 1u;  // later pieces of a pass accumulate onto the earlier ones
          for (int t = 0; t < nsteps; ++t) {
            int st = rot + t;
            if (st >= nsteps) st -= nsteps;
            const int nkb = min(kps, gw.kbn - st * kps);
            const uint32_t sw = wr.slot, sa = ar.slot;
            mbar_wait(&full_w[sw], wr.round & 1u);
            mbar_wait(&full_a[sa], ar.round & 1u);
            tc_fence_after();
            if (i == gw.p0 && t == 0) BD_STAMP(q, 1);
            const uint32_t a_addr = smem_u32(smem_a + sa * kStepBytes);
            const uint32_t w_addr = smem_u32(smem_w + sw * kStepBytes);
            for (int kk = 0; kk < nkb; ++kk) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, umma_desc_k_sw128(a_addr + kk * kSlotBytes + k * 32),
                          umma_desc_k_sw128(w_addr + kk * kb_bytes + k * 32), idesc, (t | kk | k) != 0 ? 1u : keep);
            }
            umma_commit(&empty_w[sw]);
            umma_commit(&empty_a[sa]);
            wr.advance();
            ar.advance();
          }
          if (gw.piece) {
            if (gw.last) {
              umma_commit(&acc_full[2]);
              ++pf;
            }
          } else {
            umma_commit(&acc_full[buf]);
            ++pi;
          }
          if (i == gw.p1 - 1) BD_STAMP(q, 2);
        }
      }
    }
  } else {
    // ===================== epilogue / row / attention warps (3..6); TMEM lane quarter = warp % 4 =====================
    const int tid = threadIdx.x - kStreamEpiWarp0 * 32;
    const int lane = tid & 31;
    const int qd = warp & 3;
    const int m = qd * 32 + lane;
    uint32_t pi = 0, pf = 0, aux_uses = 0, scr_uses = 0;
    OpCursor cur(prog, total);
    while (cur.next()) {
      const int q = cur.q, it = cur.it;
      // by value: the descriptor lives in registers for the whole op. (Reading it through the kernel-parameter bank with a
      // run-time index inside the per-element epilogue loops cost ~3 us per 32-column chunk: measured 11 us -> 1 us.)
      StreamOp op = prog.ops[cur.idx];
      if (op.tab) {
        op.p0 = tab_ptr(op, it, op.tab_p0, op.p0);
        op.p1 = tab_ptr(op, it, op.tab_p1, op.p1);
        op.p2 = tab_ptr(op, it, op.tab_p2, op.p2);
      }
      const bool skipped = op_skipped(prog, op, it);
      const bool filler = (op.flags & kFlagFiller) != 0;
      if ((op.flags & kFlagParityIt) ? (it & 1) : ((op.flags & kFlagParityNext) ? ((it + 1) & 1) : 0)) {
        if (op.kind == kOpGemm) {
          op.o0 = reinterpret_cast<uint8_t*>(op.o0) + op.l1;  // double-buffered GEMM output
        } else {  // double-buffered modulation operands of the row ops
          op.p3 = reinterpret_cast<const uint8_t*>(op.p3) + op.l1;
          op.p4 = reinterpret_cast<const uint8_t*>(op.p4) + op.l1;
          op.p6 = reinterpret_cast<const uint8_t*>(op.p6) + op.l1;
        }
      }
      if (skipped) {
        // nothing to do, but (non-filler ops) the arrival below keeps the cumulative barrier count in step
      } else if (op.kind == kOpGemm) {
        const GemmWork gw = gemm_work(op, G, c);
        const int rows = op.i1 > 0 ? op.i1 : prog.M;  // valid token rows of this op
        if (!gw.any() && op.wait_prev && !filler) {
          // a CTA without work must not run ahead: the arrival counter is cumulative, so every CTA has to pass every
          // dependency (CTAs with work do so through their A producer -> MMA -> accumulator chain)
          if (tid == 0) grid_wait(prog.sync, wait_target(op.wait_prev, cur.mseq, G), static_cast<unsigned int>(prog.poll_ns));
          epi_bar();
        }
        const bool run_epi = !gw.piece || gw.last;  // earlier pieces of a pass only accumulate
        for (int i = gw.p0; run_epi && i < gw.p1; ++i) {
          const int u0 = stream_pass_u0(gw.part, i);
          const int w = (stream_pass_u0(gw.part, i + 1) - u0) * 16;
          const int n0 = (gw.part.unit0 + u0) * 16;
          const uint32_t buf = gw.piece ? 2u : (pi & 1u);
          const uint32_t par = gw.piece ? (pf & 1u) : ((pi >> 1) & 1u);
          // bias slice of this pass -> shared memory while the MMAs are still running (a global load per chunk after the
          // accumulator is ready would put an L2 round trip — there is no L1 left — on the dependency path)
          const float* bias_sm = nullptr;
          if (op.p2 && op.sub != kEpiPartial) {
            if (tid < w / 8) {
              float t8[8];
              bf16x8_to_f(ldg_u4(reinterpret_cast<const __nv_bfloat16*>(op.p2) + n0 + tid * 8), t8);
              float4* dst4 = reinterpret_cast<float4*>(bias_s + buf * kAccStride + tid * 8);
              dst4[0] = make_float4(t8[0], t8[1], t8[2], t8[3]);
              dst4[1] = make_float4(t8[4], t8[5], t8[6], t8[7]);
            }
            epi_bar();
            bias_sm = bias_s + buf * kAccStride;
          }
          mbar_wait(&acc_full[buf], par);
          tc_fence_after();
          if (tid == 0 && i == gw.p1 - 1) BD_STAMP(q, 3);
          const uint32_t tbase = tmem_base + buf * static_cast<uint32_t>(kAccStride) + (static_cast<uint32_t>(qd * 32) << 16);
          // 32-column chunks start at packed columns that are multiples of 32, so that every store is whole 32-byte
          // sectors; a pass that starts / ends mid-way gets a 16-column chunk at that edge
          int col = 0;
          auto chunk16 = [&](int cc) {
            uint32_t v[16];
            tmem_ld_32x32_x16(tbase + static_cast<uint32_t>(cc), v);
            tmem_ld_wait();
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
            gemm_epi_chunk<16>(op, rows, m, n0 + cc, acc, gw.part.split, prog.dbg_mode, bias_sm ? bias_sm + cc : nullptr);
          };
          if ((n0 & 31) != 0) {
            chunk16(0);
            col = 16;
          }
          for (; col + 32 <= w; col += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tbase + static_cast<uint32_t>(col), v);
            tmem_ld_wait();
            if (tid == 0 && i == gw.p1 - 1 && col <= 16) BD_STAMP(q, 6);
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
            gemm_epi_chunk<32>(op, rows, m, n0 + col, acc, gw.part.split, prog.dbg_mode, bias_sm ? bias_sm + col : nullptr);
            if (tid == 0 && i == gw.p1 - 1 && col <= 16) BD_STAMP(q, 7);
          }
          if (col < w) chunk16(col);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
          if (gw.piece) ++pf;
          else ++pi;
        }
      } else {
        if (op.wait_prev) {
          if (tid == 0) grid_wait(prog.sync, wait_target(op.wait_prev, cur.mseq, G), static_cast<unsigned int>(prog.poll_ns));
          epi_bar();
        }
        const bool scratch = op_uses_scratch(prog, op, c);
        if (scratch) mbar_wait(scr_free, scr_uses & 1u);  // the A ring is ours (see op_uses_scratch)
        if (op.kind == kOpRow) {
          // one token row (or schedule row) per CTA; a grid smaller than the 128-row tile (several engines side by side,
          // each on its share of the SMs: the ImageNet sampler) takes its rows round-robin
          bool done = false;
          if constexpr (FAM == kStreamFamHeadSmall) {
            if ((op.sub == kRowLnMod || op.sub == kRowSplitkLnMod || op.sub == kRowFinal) && op.N <= 32 * kRowVec * 8 &&
                !(prog.dbg_mode & 512)) {
              // one warp per row: the CTA's rows c, c + G, ... are dealt to its 4 executor warps (no block barrier inside;
              // the final row's activations live in the warp's own 6 KB of the A ring)
              for (int r = c + (tid >> 5) * G; r < prog.M; r += 4 * G)
                row_op_ln_family<kRowVec, 32>(prog, op, it, r, tid & 31, nullptr, smem_a + 24576 + (tid >> 5) * 6144);
              done = true;
            }
          }
          if (!done) {
            for (int r = c; r < 128; r += G) {
              if (r != c) epi_bar();  // the reduction scratch of the previous row is dead
              row_op<FAM>(prog, op, it, r, tid, red, smem_a);
            }
          }
        } else if constexpr (FAM == kStreamFamLlm) {
          if (op.kind == kOpLlmRope) {
            if (op.K == 128) llm_rope_append<128>(op, it, c, G, tid);
            else llm_rope_append<64>(op, it, c, G, tid);
          } else if (op.kind == kOpLlmAttn && c < op.act) {
            unsigned long long* slot =
                (prog.dbg && q < prog.dbg_ops) ? prog.dbg + (static_cast<long long>(q) * G + c) * 8 : nullptr;
            if (op.K == 128) llm_attn_stream<128>(op, it, c, tid, smem_a, prog.dbg_mode, slot);
            else llm_attn_stream<64>(op, it, c, tid, smem_a, prog.dbg_mode, slot);
          }
        } else if (op.kind == kOpAttn) {
          const int units = (prog.M / op.i0) * (op.N / op.K);
          bool done = false;
          if constexpr (FAM == kStreamFamHeadSmall) {
            if (op.K == 64 && op.i0 <= 16 && !(prog.dbg_mode & 256)) {  // one warp per unit (mode 256: measurement, off)
              for (int u = c + (tid >> 5) * G; u < units; u += 4 * G)
                attn_unit_warp64(op, u, tid & 31, smem_a + (tid >> 5) * 6144);
              done = true;
            }
          }
          for (int u = c; u < units && !done; u += G) {
            if (op.K == 128) attn_unit<128>(op, u, tid, smem_a, aux_bar, aux_uses & 1u);
            else attn_unit<64>(op, u, tid, smem_a, aux_bar, aux_uses & 1u);
            ++aux_uses;
          }
        }
        // generic-proxy writes to the A ring (attention tiles / final row) before later async-proxy (bulk copy) writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (scratch) {
          epi_bar();
          if (tid == 0) mbar_arrive(scr_done);
          ++scr_uses;
        }
      }
      if (filler) continue;  // outside the barrier protocol: its consumers sit behind a later full barrier
      // ---- this CTA's part of op q is complete: publish ----
      if (tid == 0) BD_STAMP(q, 4);
      epi_bar();
      if (tid == 0) {
        red_release_gpu_add(prog.sync, 1u);
        BD_STAMP(q, 5);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// stream-major weight packing: W [N, K] row-major bf16 -> for CTA c, pass i, k-block kb: a [w x 64] slot in the
// 128B-swizzled K-major image (contiguous per CTA). perm 1 = SwiGLU-8: packed unit u = [8 rows of x1 | 8 rows of x2].
// ---------------------------------------------------------------------------------------------------------------------
__device__ __host__ inline int stream_src_row(int p, int perm, int hidden) {
  if (perm == 0) return p;
  const int u = p >> 4, wi = p & 15;
  return wi < 8 ? u * 8 + wi : hidden + u * 8 + (wi - 8);
}

__global__ void __launch_bounds__(256) stream_pack_kernel(const __nv_bfloat16* __restrict__ w, long long ldw, int N, int K,
                                                          int S, int G, int perm, int hidden,
                                                          const __nv_bfloat16* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ out,
                                                          __nv_bfloat16* __restrict__ bias_out) {
  const int c = blockIdx.x;
  const StreamPart part = stream_partition(N, K, S, G, c);
  if (part.units == 0) return;
  const int kbl = blockIdx.y;  // local k-block
  if (kbl >= part.kbs) return;
  const int kb = part.kb0 + kbl;
  for (int i = 0; i < part.npass; ++i) {
    const int u0 = stream_pass_u0(part, i);
    const int wd = (stream_pass_u0(part, i + 1) - u0) * 16;
    const int n0 = (part.unit0 + u0) * 16;
    uint8_t* slot = reinterpret_cast<uint8_t*>(out) + stream_pass_offset(N, part, i) * 2048 +
                    static_cast<long long>(kbl) * wd * 128;
    for (int t = threadIdx.x; t < wd * 8; t += blockDim.x) {
      const int r = t >> 3, j = t & 7;
      const int src = stream_src_row(n0 + r, perm, hidden);
      const int k = kb * 64 + j * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (k + 8 <= K && ((ldw & 7) == 0)) {
        v = *reinterpret_cast<const uint4*>(w + src * ldw + k);
      } else {
        __nv_bfloat16 tmp[8];
        for (int e = 0; e < 8; ++e) tmp[e] = (k + e < K) ? w[src * ldw + k + e] : __float2bfloat16_rn(0.f);
        v = *reinterpret_cast<uint4*>(tmp);
      }
      *reinterpret_cast<uint4*>(slot + r * 128 + ((j ^ (r & 7)) << 4)) = v;
    }
    if (bias_out && kbl == 0 && part.split == 0) {
      for (int r = threadIdx.x; r < wd; r += blockDim.x)
        bias_out[n0 + r] = bias ? bias[stream_src_row(n0 + r, perm, hidden)] : __float2bfloat16_rn(0.f);
    }
  }
}

static unsigned long long* g_stream_dbg = nullptr;
static int g_stream_dbg_ops = 0;
static int g_stream_dbg_mode = 0;
static int g_stream_pf_steps = 0;
static int g_stream_poll_ns = 32;
static int g_stream_w_slots = kStreamWSlotsDefault, g_stream_a_slots = kStreamASlotsDefault;

int stream_tuning_mode() { return g_stream_dbg_mode; }

int stream_launch(const StreamProgram& prog_in, cudaStream_t stream) {
  static thread_local StreamProgram prog;
  prog = prog_in;
  prog.dbg = g_stream_dbg;
  prog.dbg_ops = g_stream_dbg_ops;
  prog.dbg_mode = g_stream_dbg_mode;
  prog.pf_steps = g_stream_pf_steps;
  prog.poll_ns = g_stream_poll_ns;
  if (prog.w_slots <= 0 || prog.a_slots <= 0 || prog.w_slots + prog.a_slots > kStreamSlots) {
    prog.w_slots = g_stream_w_slots;
    prog.a_slots = g_stream_a_slots;
  }
  {  // the attribute is per device (a process may drive several GPUs)
    static bool attr_set[64] = {};
    int dev = 0;
    BD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      BD_CUDA_TRY(cudaFuncSetAttribute(bd_stream_kernel<kStreamFamHead>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       StreamSmem::kTotal));
      BD_CUDA_TRY(cudaFuncSetAttribute(bd_stream_kernel<kStreamFamLlm>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       StreamSmem::kTotal));
      BD_CUDA_TRY(cudaFuncSetAttribute(bd_stream_kernel<kStreamFamHeadSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       StreamSmem::kTotal));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  BD_CUDA_TRY(cudaMemsetAsync(prog.sync, 0, sizeof(unsigned int), stream));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(prog.n_ctas);
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = StreamSmem::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ++g_launch_count;
  BD_REQUIRE(prog.family == kStreamFamHead || prog.family == kStreamFamLlm || prog.family == kStreamFamHeadSmall);
  if (prog.family == kStreamFamLlm) BD_CUDA_TRY(cudaLaunchKernelEx(&cfg, bd_stream_kernel<kStreamFamLlm>, prog));
  else if (prog.family == kStreamFamHeadSmall) BD_CUDA_TRY(cudaLaunchKernelEx(&cfg, bd_stream_kernel<kStreamFamHeadSmall>, prog));
  else BD_CUDA_TRY(cudaLaunchKernelEx(&cfg, bd_stream_kernel<kStreamFamHead>, prog));
  return BD_OK;
}

}  // namespace bd

using namespace bd;

extern "C" {

int bd_stream_num_ctas(void) { return num_sms(); }

int bd_stream_set_debug(void* buf, int max_ops) {
  g_stream_dbg = static_cast<unsigned long long*>(buf);
  g_stream_dbg_ops = buf ? max_ops : 0;
  return BD_OK;
}

int bd_stream_set_ksplit(int ksplit) {  // before packing any weights: the packer and the program builder must agree
  BD_REQUIRE(ksplit == 1 || ksplit == 2 || ksplit == 4);
  stream_ksplit_small() = ksplit;
  return BD_OK;
}

int bd_stream_set_poll_ns(int ns) {
  BD_REQUIRE(ns >= 0 && ns <= 100000);
  g_stream_poll_ns = ns;
  return BD_OK;
}

int bd_stream_set_prefetch(int steps) {
  BD_REQUIRE(steps >= 0 && steps <= 256);
  g_stream_pf_steps = steps;
  return BD_OK;
}

int bd_stream_set_tuning(int w_slots, int a_slots, int mode) {
  BD_REQUIRE(w_slots >= 2 && a_slots >= 2 && w_slots + a_slots <= kStreamSlots);  // A ring also hosts the attention tiles
  g_stream_w_slots = w_slots;
  g_stream_a_slots = a_slots;
  g_stream_dbg_mode = mode;
  return BD_OK;
}

size_t bd_stream_packed_elems(int N, int K) {
  if (N <= 0 || K <= 0 || (N % 16) != 0) return 0;
  return static_cast<size_t>(N) * (static_cast<size_t>((K + 63) / 64) * 64);
}

int bd_stream_ksplit(int N, int K, int n_ctas) { return stream_ksplit_for(N, K, n_ctas); }

// tests: the work split of CTA c for a GEMM op. out[0..5] = split, unit0, units, kb0, kbs, npass; then per pass i
// (up to 8): out[6 + 3 i ..] = first unit (relative to unit0), width in rows, offset of the pass's first slot in 2 KB units.
int bd_stream_partition_info(int N, int K, int ksplit, int n_ctas, int c, long long* out, int cap) {
  BD_REQUIRE(out && cap >= 6 && N > 0 && (N % 16) == 0 && K > 0 && ksplit >= 1 && n_ctas >= ksplit && c >= 0 && c < n_ctas);
  BD_REQUIRE(((K + 63) / 64) % ksplit == 0);
  const StreamPart p = stream_partition(N, K, ksplit, n_ctas, c);
  out[0] = p.split; out[1] = p.unit0; out[2] = p.units; out[3] = p.kb0; out[4] = p.kbs; out[5] = p.npass;
  for (int i = 0; i < p.npass && 6 + 3 * i + 2 < cap; ++i) {
    out[6 + 3 * i] = stream_pass_u0(p, i);
    out[6 + 3 * i + 1] = (stream_pass_u0(p, i + 1) - stream_pass_u0(p, i)) * 16;
    out[6 + 3 * i + 2] = stream_pass_offset(N, p, i);
  }
  return BD_OK;
}

int bd_stream_pack_weight(const void* W, int64_t ldw, int N, int K, int ksplit, int n_ctas, int perm, int hidden,
                          const void* bias, void* out, void* bias_out, bd_stream_t stream) {
  BD_REQUIRE(W && out && N > 0 && K > 0 && (N % 16) == 0 && ldw >= K && n_ctas > 0 && ksplit >= 1);
  BD_REQUIRE(((K + 63) / 64) % ksplit == 0 && n_ctas / ksplit >= 1);
  BD_REQUIRE(perm == 0 || (perm == 1 && hidden > 0 && N == 2 * hidden && (hidden % 8) == 0));
  const int kbs = ((K + 63) / 64) / ksplit;
  dim3 grid(n_ctas, kbs);
  stream_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(W), ldw, N, K, ksplit, n_ctas, perm, hidden,
      static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(bias_out));
  BD_LAUNCH_CHECK();
  return BD_OK;
}

// n_w GEMM ops x repeat through the persistent kernel (tests / micro-benchmarks): A blocked bf16 [ceil(K/64)][128][64],
// W stream-packed, out per `epi` (0 bias(+act) -> bf16 row-major or blocked; 1 SwiGLU-8; 2 fp32 partials [ksplit][M][N]).
int bd_stream_gemm(const void* A_blocked, const void* W_packed, int64_t w_stride_bytes, int n_w, const void* bias_packed,
                   void* out, int64_t ld_out, int M, int N, int K, int ksplit, int epi, int act, int out_blocked,
                   int n_ctas, int repeat, void* sync, bd_stream_t stream) {
  BD_REQUIRE(A_blocked && W_packed && out && sync && M > 0 && M <= 128 && N > 0 && (N % 16) == 0 && K > 0);
  BD_REQUIRE(n_ctas > 0 && ksplit >= 1 && ((K + 63) / 64) % ksplit == 0 && repeat >= 1 && repeat <= kStreamMaxIter);
  BD_REQUIRE(n_w >= 1 && n_w <= kStreamMaxOps);
  static StreamProgram prog;  // large (kernel parameter image); built per call
  prog = StreamProgram{};
  prog.n_pre = 0;
  prog.n_body = n_w;
  prog.n_iter = repeat;
  prog.n_post = 0;
  prog.M = M;
  prog.n_ctas = n_ctas;
  prog.rows_x = 0;
  prog.cfg_mult = 1;
  prog.sync = static_cast<unsigned int*>(sync);
  for (int j = 0; j < n_w; ++j) {
    StreamOp& op = prog.ops[j];
    op.kind = kOpGemm;
    op.sub = epi;
    op.N = N;
    op.K = K;
    op.ksplit = ksplit;
    op.act = act;
    op.flags = out_blocked ? 1 : 0;
    op.wait_prev = 1;
    op.p0 = static_cast<const uint8_t*>(W_packed) + static_cast<long long>(j) * w_stride_bytes;
    op.p1 = A_blocked;
    op.p2 = bias_packed;
    op.o0 = out;
    op.l0 = ld_out;
  }
  return stream_launch(prog, static_cast<cudaStream_t>(stream));
}

int bd_stream_gemm_filler(const void* A_blocked, const void* W_main, const void* W_fill, const void* bias_fill,
                          void* out_main, void* out_fill, int M, int N, int K, int n_ctas, int n_slices, int repeat,
                          void* sync, bd_stream_t stream) {
  BD_REQUIRE(A_blocked && W_main && W_fill && out_main && out_fill && sync && M > 0 && M <= 128 && N > 0 && (N % 16) == 0);
  BD_REQUIRE(K > 0 && n_ctas > 0 && n_slices >= 1 && repeat >= 1 && repeat <= kStreamMaxIter);
  const int KB = (K + 63) / 64, nsteps = stream_steps(KB);
  BD_REQUIRE(n_slices <= nsteps);
  int P = 0;
  for (int c = 0; c < n_ctas; ++c) {
    const StreamPart p = stream_partition(N, K, 1, n_ctas, c);
    P = p.npass > P ? p.npass : P;
  }
  BD_REQUIRE(1 + P * n_slices <= kStreamMaxOps);
  static StreamProgram prog;
  prog = StreamProgram{};
  prog.n_pre = 0;
  prog.n_iter = repeat;
  prog.n_post = 0;
  prog.M = M;
  prog.n_ctas = n_ctas;
  prog.cfg_mult = 1;
  prog.sync = static_cast<unsigned int*>(sync);
  int n = 0;
  auto base_op = [&](const void* W, const void* bias, void* out) -> StreamOp& {
    StreamOp& op = prog.ops[n++];
    op = StreamOp{};
    op.kind = kOpGemm;
    op.sub = kEpiBias;
    op.N = N;
    op.K = K;
    op.ksplit = 1;
    op.wait_prev = 1;
    op.p0 = W;
    op.p1 = A_blocked;
    op.p2 = bias;
    op.o0 = out;
    op.l0 = N;
    return op;
  };
  base_op(W_main, nullptr, out_main);
  for (int pass = 0; pass < P; ++pass) {
    for (int j = 0; j < n_slices; ++j) {
      const int s0 = (j * nsteps) / n_slices, s1 = ((j + 1) * nsteps) / n_slices;
      StreamOp& op = base_op(W_fill, bias_fill, out_fill);
      op.wait_prev = 0;
      op.flags |= kFlagFiller;
      op.pc_pass = pass;
      op.pc_kb0 = s0 * kKbPerStep;
      op.pc_kbn = (s1 * kKbPerStep < KB ? s1 * kKbPerStep : KB) - op.pc_kb0;
      op.pc_flags = (j == 0 ? kPieceFirst : 0) | (j == n_slices - 1 ? kPieceLast : 0);
    }
  }
  prog.n_body = n;
  return stream_launch(prog, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
