// bd_host.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cuda_bf16.h>
#include "../../include/bitdance_b200.h"

namespace bd {

extern thread_local int g_last_cuda_error;
extern unsigned long long g_launch_count;  // kernels launched by this library (bench.py's gpu_launches)

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return BD_ERR_CUDA;
}
#define BD_CUDA_TRY(expr)                                \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return ::bd::cuda_fail(_e);   \
  } while (0)
#define BD_LAUNCH_CHECK()           \
  do {                              \
    ++::bd::g_launch_count;         \
    BD_CUDA_TRY(cudaGetLastError()); \
  } while (0)
#define BD_REQUIRE(cond)              \
  do {                                \
    if (!(cond)) return BD_ERR_INVALID; \
  } while (0)

// 2-D bf16 row-major tensor map: dims {inner, outer}, box {box_inner, box_outer}, 128B swizzle, zero OOB fill.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer);
// 4-D bf16 NHWC activation map for implicit-GEMM convolution: dims {C, W, H, B}.
int make_tmap_4d_bf16(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B,
                      uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_b);

enum : int { kActNone = 0, kActSilu = 1, kActGeluTanh = 2 };

// Generic GEMM epilogue. Rounding points mirror torch autocast(bf16): every Linear output is rounded to bf16,
// every bf16 elementwise op rounds again.
//   y = bf16(acc + bias[n])
//   act:     y = bf16(act(y))
//   swiglu:  (W rows interleaved in groups of 16: 16 gate rows then 16 up rows) y = bf16(bf16(silu(g)) * u)
//   gate:    y = bf16(y * gate[m, n])
//   res:     y = res[m (% res_mod), n] + y   (rounded to bf16 when the output is bf16)
struct GemmEpi {
  const __nv_bfloat16* bias = nullptr;  // [N]
  const __nv_bfloat16* gate = nullptr;  // [M, ld_gate]
  const void* res = nullptr;            // [M, ld_res] bf16 or fp32
  void* out = nullptr;                  // [M, ld_out] bf16 or fp32
  long long ld_gate = 0, ld_res = 0, ld_out = 0;
  int act = 0;
  int swiglu = 0;
  int res_f32 = 0;
  int out_f32 = 0;
  int res_mod = 0;  // > 0: residual row index is (m % res_mod) (a [pn, N] table broadcast over sequences)
};

// Internal C++ entry points shared by the composite ops (bd_head.cu, bd_llm.cu, ...).
size_t gemm_workspace_bytes(int M, int N, int K, int bn, int splits);
// w_tiled: W is in the tile-major layout of bd_pack_weight_tiles (ldw ignored).
// partial_splits != nullptr: when the plan splits K, ONLY the partial-sum GEMM is launched (fp32 partials
// [S][M][N] at the start of `workspace`), *partial_splits = S and the caller runs its own fused reduction; when the
// plan does not split, the epilogue runs inside the GEMM as usual and *partial_splits = 1.
int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const GemmEpi& epi,
              void* workspace, size_t workspace_bytes, int bn, int splits, bool pdl, cudaStream_t stream,
              bool w_tiled = false, int* partial_splits = nullptr);

int num_sms();

struct LaunchCfg {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ++g_launch_count;
  }
};

}  // namespace bd
