// bd_probe.cu — measurement-only kernels (never on the product path): the HBM->smem streaming ceiling of the
// producer/consumer ring the weight-streaming GEMM is built on, with and without the L2-resident activation re-reads.
// Used by scripts/stream_probe.py to separate "HBM stream" from "L2->SM fabric" limits.
#include "bd_host.h"
#include "bd_ptx.cuh"

namespace bd {

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                             uint64_t hint) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "l"(hint)
      : "memory");
}

struct ProbeArgs {
  const uint8_t* w;        // HBM stream: CTA c reads [c * w_per_cta, (c+1) * w_per_cta)
  long long w_per_cta;     // bytes, multiple of w_chunk
  int w_chunk;             // bytes per bulk copy (multiple of 16)
  int w_stages;
  const uint8_t* x;        // L2-resident buffer re-read by every CTA (x_bytes total, cycled)
  int x_bytes;
  int x_chunk;             // bytes per W chunk re-read from x (0 = none)
  int x_stages;
};

__global__ void __launch_bounds__(96, 1) probe_stream_kernel(ProbeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sw = smem;
  uint8_t* sx = sw + static_cast<size_t>(a.w_stages) * a.w_chunk;
  uint64_t* full_w = reinterpret_cast<uint64_t*>(sx + static_cast<size_t>(a.x_stages) * a.x_chunk);
  uint64_t* empty_w = full_w + a.w_stages;
  uint64_t* full_x = empty_w + a.w_stages;
  uint64_t* empty_x = full_x + a.x_stages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.w_stages; ++s) { mbar_init(&full_w[s], 1); mbar_init(&empty_w[s], 1); }
    for (int s = 0; s < a.x_stages; ++s) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int n = static_cast<int>(a.w_per_cta / a.w_chunk);
  if (warp == 0) {
    if (elect_one()) {
      const uint8_t* src = a.w + static_cast<long long>(blockIdx.x) * a.w_per_cta;
      for (int i = 0; i < n; ++i) {
        const int s = i % a.w_stages;
        if (i >= a.w_stages) mbar_wait(&empty_w[s], ((i / a.w_stages) & 1u) ^ 1u);
        mbar_expect_tx(&full_w[s], a.w_chunk);
        bulk_load_1d(sw + static_cast<size_t>(s) * a.w_chunk, src + static_cast<long long>(i) * a.w_chunk, a.w_chunk,
                     &full_w[s], kEvictFirst);
      }
    }
  } else if (warp == 2 && a.x_chunk > 0) {
    if (elect_one()) {
      const int per = a.x_bytes / a.x_chunk;
      for (int i = 0; i < n; ++i) {
        const int s = i % a.x_stages;
        if (i >= a.x_stages) mbar_wait(&empty_x[s], ((i / a.x_stages) & 1u) ^ 1u);
        mbar_expect_tx(&full_x[s], a.x_chunk);
        bulk_load_1d(sx + static_cast<size_t>(s) * a.x_chunk, a.x + static_cast<long long>(i % per) * a.x_chunk,
                     a.x_chunk, &full_x[s], kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < n; ++i) {
        const int s = i % a.w_stages;
        mbar_wait(&full_w[s], (i / a.w_stages) & 1u);
        mbar_arrive(&empty_w[s]);
        if (a.x_chunk > 0) {
          const int sxs = i % a.x_stages;
          mbar_wait(&full_x[sxs], (i / a.x_stages) & 1u);
          mbar_arrive(&empty_x[sxs]);
        }
      }
    }
  }
}

}  // namespace bd

using namespace bd;

extern "C" int bd_probe_stream(const void* w, long long w_per_cta, int w_chunk, int w_stages, const void* x, int x_bytes,
                               int x_chunk, int x_stages, int n_ctas, bd_stream_t stream) {
  BD_REQUIRE(w && w_per_cta > 0 && w_chunk > 0 && (w_chunk % 16) == 0 && (w_per_cta % w_chunk) == 0 && w_stages > 0);
  BD_REQUIRE(x_chunk == 0 || (x && x_bytes >= x_chunk && (x_chunk % 16) == 0 && x_stages > 0));
  BD_REQUIRE(n_ctas > 0);
  const size_t smem = static_cast<size_t>(w_stages) * w_chunk + static_cast<size_t>(x_stages) * x_chunk + 1024 +
                      (2 * w_stages + 2 * x_stages) * 8 + 64;
  BD_REQUIRE(smem <= 227 * 1024);
  BD_CUDA_TRY(cudaFuncSetAttribute(probe_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ProbeArgs a{static_cast<const uint8_t*>(w), w_per_cta, w_chunk, w_stages, static_cast<const uint8_t*>(x), x_bytes,
              x_chunk, x_stages};
  probe_stream_kernel<<<n_ctas, 96, smem, static_cast<cudaStream_t>(stream)>>>(a);
  BD_LAUNCH_CHECK();
  return BD_OK;
}
