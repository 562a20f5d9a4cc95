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

// Issue rate of the legacy tensor path (mma.sync.m16n8k16 bf16 -> HMMA) with few warps per SM sub-partition: CH independent
// accumulator chains per warp, `iters` rounds; per-warp cycles -> out[cta * warps + warp]. (The in-engine attention runs on
// 4 executor warps: is it bound by HMMA issue, and would more warps or more chains help?)
template <int CH>
__global__ void probe_hmma_kernel(int iters, long long* out, float* sink) {
  float acc[CH][4];
  uint32_t a[4], b[2];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = 0x3c003c00u + threadIdx.x + j;
  b[0] = 0x38003800u + threadIdx.x;
  b[1] = 0x34003400u;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  if (s == 123.456f) sink[0] = s;
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = t1 - t0;
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

extern "C" int bd_probe_hmma(int warps, int chains, int iters, int n_ctas, long long* out_cycles, float* sink,
                             bd_stream_t stream) {
  BD_REQUIRE(warps >= 1 && warps <= 32 && iters > 0 && n_ctas > 0 && out_cycles && sink);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (chains) {
    case 1: probe_hmma_kernel<1><<<n_ctas, warps * 32, 0, st>>>(iters, out_cycles, sink); break;
    case 2: probe_hmma_kernel<2><<<n_ctas, warps * 32, 0, st>>>(iters, out_cycles, sink); break;
    case 4: probe_hmma_kernel<4><<<n_ctas, warps * 32, 0, st>>>(iters, out_cycles, sink); break;
    case 8: probe_hmma_kernel<8><<<n_ctas, warps * 32, 0, st>>>(iters, out_cycles, sink); break;
    case 16: probe_hmma_kernel<16><<<n_ctas, warps * 32, 0, st>>>(iters, out_cycles, sink); break;
    default: return BD_ERR_INVALID;
  }
  BD_LAUNCH_CHECK();
  return BD_OK;
}
