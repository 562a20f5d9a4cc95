// bd_stream.cuh — the persistent weight-streaming engine (sm_100a).
//
// Measured on B200 (profiles/r01_*): the tcgen05 GEMM reaches ~5.8 TB/s in steady state but a 52-157 MB weight matrix is
// only 8-24 us of HBM time, and every kernel boundary costs 6-9 us of ramp + drain during which HBM idles. One diffusion
// head evaluation (reference modeling/vision_head/flow_head_parallel_x.py:325-342) is 27 such GEMMs, 51 evaluations per
// AR step (sampling_x.py:44-97). So the whole sampler runs as ONE cooperative kernel, one CTA per SM, interpreting a small
// op program (GEMM / row / attention ops):
//   * warp 0  — W producer. Weights never depend on activations, so it walks the program AHEAD of everybody else and
//               keeps the shared-memory ring full across op boundaries: HBM never idles while the other warps wait on a
//               grid-wide dependency. Weights are pre-packed "stream-major": the bytes CTA c needs for an op are ONE
//               contiguous range, already in the 128B-swizzled K-major image tcgen05 wants, fetched with 1-D bulk copies
//               (no tensor map);
//   * warp 2  — A producer: activations live in HBM/L2 in the same blocked image ([K/64][128 rows][64] bf16, swizzled),
//               one 16 KB bulk copy per k-block, issued after the grid-wide dependency of the op is satisfied;
//   * warp 1  — MMA issuer: tcgen05.mma M=128 (token rows) x N=pass width (16..128 weight rows) x K=16, fp32 accumulators
//               in TMEM, two accumulator buffers so the epilogue of one pass overlaps the MMAs of the next;
//   * warps 3-6 — epilogue of the GEMM passes (TMEM -> registers -> HBM) and the executors of row ops (LayerNorm-modulate,
//               split-K reduction + gated residual, SiLU, the final Linear + SDE update) and of the 64x64 attention.
// Ops are separated by a grid barrier (one monotonic counter; release/acquire at gpu scope). Work split of a GEMM op:
// N is cut in units of 16 weight rows dealt evenly to the CTAs (and K in `ksplit` ranges when N is small), so all 148
// SMs stream equal shares of every matrix.
#pragma once
#include "bd_host.h"
#include "bd_ptx.cuh"

namespace bd {

constexpr int kStreamThreads = 224;
constexpr int kStreamEpiThreads = 128;   // warps 3..6
constexpr int kStreamEpiWarp0 = 3;
constexpr int kSlotBytes = 16384;        // 128 rows x 64 bf16
constexpr int kKbPerStep = 2;            // k-blocks per ring slot: one mbarrier round trip / MMA batch per 128 k
constexpr int kStepBytes = kKbPerStep * kSlotBytes;
constexpr int kStreamSlots = 7;          // shared-memory ring slots (32 KB each) in total (W ring + A ring)
constexpr int kStreamWSlotsDefault = 5;
constexpr int kStreamASlotsDefault = 2;
constexpr int kStreamMaxOps = 128;  // the program travels as a kernel parameter (< 32 KB)
constexpr int kStreamMaxIter = 104;

enum : int {
  kOpGemm = 0, kOpRow = 1, kOpAttn = 2,
  kOpLlmRope = 3,   // Qwen3: q/k RMSNorm over head_dim + RoPE (fp32) + paged KV append, one warp per (token, head)
  kOpLlmAttn = 4,   // Qwen3: flash attention of the block's queries over the paged cache, split-KV partials
};
enum : int {
  kFlagBlocked = 1, kFlagParityIt = 2, kFlagParityNext = 4, kFlagSkipLast = 8,
  kFlagFiller = 16,   // GEMM op outside the grid-barrier sequence (see "fillers" below)
  kFlagAPerIt = 32,   // GEMM op: the A operand of iteration `it` is p1 + (it + i0) * blocked bytes of K
};
enum : int { kPieceFirst = 1, kPieceLast = 2 };
enum : int { kEpiBias = 0, kEpiSwiglu8 = 1, kEpiPartial = 2 };
enum : int {
  kRowCastCond = 0,   // fp32 [M, K] -> blocked bf16
  kRowTFreq = 1,      // timestep embedding rows -> blocked bf16
  kRowInit = 2,       // x = noise[0]; xb = bf16(x) duplicated per CFG group
  kRowSiluAdd = 3,    // y = silu(temb[it] + cemb[r]) -> blocked
  kRowLnMod = 4,      // a = LN(h) (*w + b) * (1 + scale) + shift -> blocked
  kRowSplitkLnMod = 5,// h += gate * (sum(partials) + bias); then kRowLnMod
  kRowFinal = 6,      // kRowSplitkLnMod without affine, then final Linear (+ 2 sigmoid - 1) -> pred
  kRowSde = 7,        // Euler–Maruyama / last Euler step on x; xb for the next evaluation
  kRowLlmRms = 8,     // Qwen3 RMSNorm of an fp32 residual row -> blocked bf16
  kRowLlmResRms = 9,  // residual += bf16(sum partials); then RMSNorm -> blocked bf16, or the final norm (+ pos table) -> fp32
  kRowSiluAddAll = 10,// kRowSiluAdd for EVERY iteration at once: y[it] = silu(temb[it] + cemb[r]) -> o0 + it * l1 (blocked)
  kRowLlmAttnCombine = 11,  // fixed-order combine of the split-KV attention partials of token row r -> blocked bf16 operand
};

// One op. Field meaning per kind:
//  GEMM: p0 = W stream-packed, p1 = A blocked, p2 = bias (packed order), o0 = out, l0 = ld_out (elements), N, K, ksplit,
//        sub = epilogue kind, act, flags bit0 = out is blocked bf16 (else row-major), i1 = valid rows (0: prog.M)
//  flags bit1 / bit2: double-buffered operand — add l1 bytes to o0 (GEMM) or to p3, p4, p6 (ROW) when the iteration `it`
//        (bit1) or `it + 1` (bit2) is odd; flags bit3: the op is skipped in the last iteration (it prepares the next one).
//  wait_prev = 0 on a GEMM makes it a FILLER: its operands were complete long before, so its weight stream and MMAs
//        overlap the epilogue / row op / barrier of the ops around it (the head's adaLN GEMM of the next evaluation).
//  FILLERS (kFlagFiller, wait_prev = 0): GEMM work that depends on nothing the surrounding ops compute (the head's adaLN
//        modulation of the NEXT evaluation). A filler does not take part in the grid barrier (no arrival, not counted in the
//        sequence numbers the other ops wait for); its result is consumed only after a later full barrier, which every CTA
//        passes after its own fillers (ops are processed in program order). A filler may be a PIECE of a GEMM
//        (pc_kbn > 0): pass `pc_pass` of the CTA's share, k-blocks [pc_kb0, pc_kb0 + pc_kbn) only, accumulated in a third
//        TMEM buffer across the pieces of that pass (pc_flags: kPieceFirst zero-initialises, kPieceLast runs the epilogue).
//        Pieces are sized to the bubbles of the dependent chain (a row op, an op boundary): the MMA warp and the weight
//        stream stay busy while the epilogue warps run the row op / wait for the grid barrier. Because the A ring doubles
//        as scratch of the attention / final-row executors, a filler must never sit between a GEMM and such an op in GEMM
//        order (the host program builder guarantees it).
//  PER-ITERATION POINTERS (tab != nullptr): a body op of a program whose iterations are the LAYERS of a decoder takes its
//        weight pointers from a device table tab[(it + tab_off) * kTabSlots + slot]; tab_p0 / tab_p1 / tab_p2 = slot + 1 of
//        the table entry that replaces p0 / p1 / p2 (0: keep the field).
//  LLM_ROPE: p0 = qkv row-major bf16 [M, (Hq + 2 Hkv) hd], p1 / p2 = q_norm / k_norm weights bf16 [hd], p3 / p4 = RoPE cos /
//        sin fp32 [pos, hd], p5 = seq_lens int32 [R], p6 = page_table int32 [R, N], o0 = q out row-major bf16 [M, Hq hd],
//        o1 = K pool of layer 0 (layer `it` at + it * l0 elements, V pool at + l1 elements; pools [page][Hkv][64][hd]),
//        i0 = S tokens per sequence, i1 = Hq, i2 = Hkv, N = max_pages, K = hd, f0 = eps
//  LLM_ATTN: p0 = q (as above), p1 = seq_lens, p2 = page_table, p3 = K pool of layer 0 (l0, l1 as above), o0 = partial O fp32
//        [ksplit][R][Hq][S][hd] (unnormalised), o1 = partial (max, sum) fp32 [ksplit][R][Hq][S][2], sub = R, ksplit = splits of
//        the key range, i0 = S, i1 = Hq, i2 = Hkv, N = max_pages, K = hd, f0 = softmax scale * log2(e)
//  ROW:  sub = row kind; pointers documented at each row function
//  ATTN: p0 = qkv row-major bf16 [M, 3D], o0 = out blocked bf16, N = D, K = head_dim
struct StreamOp {
  int kind, sub, N, K;
  int ksplit, act, flags, wait_prev;
  const void* p0;
  const void* p1;
  const void* p2;
  const void* p3;
  const void* p4;
  const void* p5;
  const void* p6;
  void* o0;
  void* o1;
  void* o2;
  void* o3;
  long long l0, l1;
  float f0;
  int i0, i1, i2;
  int pc_pass, pc_kb0, pc_kbn, pc_flags;  // piece of a GEMM (pc_kbn > 0), see FILLERS
  const void* const* tab;                 // per-iteration pointer table (device), see PER-ITERATION POINTERS
  int tab_p0, tab_p1, tab_p2, tab_off;
};
constexpr int kTabSlots = 8;

// which instance of the kernel runs the program (its executors). kStreamFamHeadSmall = the head's program on a grid smaller
// than the 128-row tile (several engines side by side, small models): a CTA owns several rows, so the LayerNorm row ops run
// one WARP per row (4 rows in flight per CTA) instead of one row per CTA at a time.
enum : int { kStreamFamHead = 0, kStreamFamLlm = 1, kStreamFamHeadSmall = 2 };

struct StreamProgram {
  int family;
  int n_pre, n_body, n_iter, n_post;
  int M;          // token rows (<= 128)
  int n_ctas;
  int rows_x;     // B * pn rows of the sampler state
  int cfg_mult;
  float cfg;
  unsigned int* sync;  // [1] zero before launch
  unsigned long long* dbg;  // optional timeline [dbg_ops][n_ctas][8] of %globaltimer stamps (bd_stream_set_debug)
  int dbg_ops;
  int dbg_mode;   // experiment switches (bd_stream_set_debug): bit0 no bias loads, bit1 no stores, bit2 32-byte stores
  int w_slots, a_slots;  // ring split, w_slots + a_slots <= kStreamSlots (0: defaults)
  int pf_steps;          // L2 prefetch distance of the weight stream, in ring steps (0: off)
  int poll_ns;           // sleep between two polls of the grid barrier counter
  float sched[kStreamMaxIter][6];  // per iteration: t, dt, denom, var, 1-t, noise_scale   (sampling_x.py:62-68)
  StreamOp ops[kStreamMaxOps];
};

// ---------------------------------------------------------------------------------------------------------------------
// work partition of a GEMM op (shared by the packer, the three pipeline roles and the host)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxPassUnits = 10;       // widest pass: 160 weight rows (TMEM: 3 accumulator buffers x 160 columns <= 512)
constexpr int kAccStride = kMaxPassUnits * 16;

struct StreamPart {
  int split;    // k-range index
  int unit0;    // first 16-row unit of this CTA
  int units;    // number of 16-row units (0: no work)
  int kb0;      // first k-block of the split
  int kbs;      // k-blocks per split
  int npass;    // passes of <= 8 units (128 weight rows)
};

// k-split of a Linear whose N is small next to the CTA count (its output then goes through fp32 partials + a row op)
inline int& stream_ksplit_small() {
  static int v = 4;
  return v;
}
inline int stream_ksplit_for(int N, int K, int G) {
  const int U = N / 16, KB = (K + 63) / 64, S = stream_ksplit_small();
  if (U < 4 * G && (KB % S) == 0 && KB >= 16) return S;
  return 1;
}

__host__ __device__ inline StreamPart stream_partition(int N, int K, int S, int G, int c) {
  StreamPart p;
  const int U = N / 16, KB = (K + 63) / 64;
  const int Gs = G / S;
  p.kbs = KB / S;
  if (c >= Gs * S) {
    p.split = 0; p.unit0 = 0; p.units = 0; p.kb0 = 0; p.npass = 0;
    return p;
  }
  p.split = c / Gs;
  const int j = c % Gs;
  p.unit0 = static_cast<int>((static_cast<long long>(j) * U) / Gs);
  p.units = static_cast<int>((static_cast<long long>(j + 1) * U) / Gs) - p.unit0;
  p.kb0 = p.split * p.kbs;
  // passes of <= 8 units (128 weight rows = one 16 KB ring half-slot per k-block) — except that a share of 9 or 10 units
  // stays ONE pass of 144 / 160 rows: a second pass would re-read the whole activation operand for one or two units
  // (measured on the k-split Linears wo / w2, 8.65 units per CTA: the MMA phase ran at 4.6 TB/s, bound by the A traffic)
  p.npass = p.units <= kMaxPassUnits ? (p.units > 0 ? 1 : 0) : (p.units + 7) / 8;
  return p;
}
// pass i of a CTA covers units [pass_u0, pass_u1) relative to unit0
__host__ __device__ inline int stream_pass_u0(const StreamPart& p, int i) { return (i * p.units) / p.npass; }
// offset (in 2 KB = 16 rows x 64 k x 2 B units) of the first slot of CTA c / pass i
__host__ __device__ inline long long stream_pass_offset(int N, const StreamPart& p, int i) {
  const long long U = N / 16;
  return (static_cast<long long>(p.split) * U + p.unit0 + stream_pass_u0(p, i)) * p.kbs;
}
// the K loop advances in steps of kKbPerStep k-blocks (the last step may be short); rotation: CTAs start at different
// steps so that they do not all hit the same L2 lines of A at once
// k-blocks per ring step: 2 for passes of <= 128 rows (2 x 16 KB = one 32 KB slot), 1 for the wide single passes
__host__ __device__ inline int stream_kps(int pass_rows) { return pass_rows <= 128 ? kKbPerStep : 1; }
__host__ __device__ inline int stream_steps(int kbs, int kps = kKbPerStep) { return (kbs + kps - 1) / kps; }
// mode (bd_stream_set_tuning, measurement): bit 3 = no rotation (every CTA walks K in the same order: the L2 sees ~148
// requests for the same lines within a short window); bit 4 = CTAs rotate in groups of 4 (4 requesters per line)
__host__ __device__ inline int stream_k_rot(int c, int nsteps, int mode = 0) {
  if (mode & 8) return 0;
  if (mode & 16) return ((c >> 2) * 3) % nsteps;
  return (c * 3) % nsteps;
}

// byte offset of element (row, col) inside a blocked bf16 activation [K/64][128][64] with the 128-byte swizzle
__host__ __device__ inline long long blk_off(int row, int col) {
  return (static_cast<long long>(col >> 6) << 14) + (row << 7) + ((((col >> 3) & 7) ^ (row & 7)) << 4) + ((col & 7) << 1);
}

struct StreamProgram;
int stream_launch(const StreamProgram& prog, cudaStream_t stream);  // bd_stream.cu
int stream_tuning_mode();                                           // bd_stream_set_tuning's mode bits

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "l"(hint)
      : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Wait until every CTA has completed all ops before sequence number `seq` (counter >= G * seq). Bounded: a protocol bug
// traps instead of hanging the box.
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target, unsigned int poll_ns = 32) {
  unsigned int spins = 0;
  while (ld_acquire_gpu(ctr) < target) {
    __nanosleep(poll_ns);  // (tight polling is slower: 148-296 pollers hammer the line the arrivals have to update)
    if (++spins > (1u << 24)) {
      printf("bd_stream: grid barrier timeout block=%d thread=%d target=%u have=%u\n", blockIdx.x, threadIdx.x, target,
             ld_acquire_gpu(ctr));
      __trap();
    }
  }
}

__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// L2-coherent loads of activations written by other CTAs of this kernel (never through a stale L1 line)
__device__ __forceinline__ uint4 ldcg_u4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ldcg_f4(const void* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void bf16x8_to_f(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* q = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(q[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 f_to_bf16x8(const float (&v)[8]) {
  uint4 pk;
  __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int j = 0; j < 4; ++j) q[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  return pk;
}
// sum over the 128 epilogue threads (tid = 0..127), red = 8 floats of shared memory
__device__ __forceinline__ float epi_sum(float v, float* red, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  epi_bar();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  epi_bar();
  return (red[0] + red[1]) + (red[2] + red[3]);
}
#endif  // __CUDACC__

}  // namespace bd
