// bd_host.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/bitdance_b200.h"

namespace bd {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return BD_ERR_CUDA;
}
#define BD_CUDA_TRY(expr)                                \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return ::bd::cuda_fail(_e);   \
  } while (0)
#define BD_LAUNCH_CHECK() BD_CUDA_TRY(cudaGetLastError())
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
  }
};

}  // namespace bd
