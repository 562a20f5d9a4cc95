// bd_gemm.cu — C-ABI for the weight-streaming tcgen05 GEMM (kernel in bd_gemm.cuh).
#include <cstdlib>
#include "bd_gemm.cuh"
#include "bd_host.h"

namespace bd {

struct GemmPlan {
  int bn;
  int splits;
};

static GemmPlan plan_gemm(int M, int N, int K, int bn, int splits) {
  const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;
  const int sms = num_sms();
  if (bn == 0) {
    if (N <= 64)
      bn = 64;
    else
      bn = (m_tiles * ((N + 127) / 128) > sms) ? 256 : 128;
  }
  const int tiles = m_tiles * ((N + bn - 1) / bn);
  if (splits == 0) {
    splits = sms / tiles;
    if (splits > num_kb / 4) splits = num_kb / 4;
    if (splits > 8) splits = 8;
    if (splits < 1) splits = 1;
  }
  if (splits > num_kb) splits = num_kb;
  return {bn, splits};
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, int M, int N, int K, int splits, float* partial,
                       const GemmEpi& epi, bool pdl, cudaStream_t stream, int w_tiled) {
  using Cfg = GemmCfg<BN>;
  {  // per device, not per process
    static bool attr_set[64] = {};
    int dev = 0;
    BD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      BD_CUDA_TRY(cudaFuncSetAttribute(bd_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  dim3 grid((M + kGemmBM - 1) / kGemmBM, (N + BN - 1) / BN, splits);
  LaunchCfg lc(grid, dim3(kGemmThreads), Cfg::kSmemBytes, stream, pdl);
  // Re-reading A across N tiles should hit L2: ask TMA to keep it when it is small next to W.
  const int a_hint_last = (static_cast<long long>(M) * K * 2 <= (32ll << 20)) ? 1 : 0;
  BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, bd_gemm_kernel<BN>, ta, tw, M, N, K, splits, partial, epi, a_hint_last, w_tiled));
  return BD_OK;
}

// CTA-pair kernel (bd_gemm2_kernel): clusters of 2 along M. BD_GEMM2=0 switches it off (A/B measurements).
static int gemm2_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BD_GEMM2");
    v = (e && *e) ? atoi(e) : 1;
  }
  return v;
}
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tw, int M, int N, int K, const GemmEpi& epi, bool pdl,
                        cudaStream_t stream) {
  {
    static bool attr_set[64] = {};
    int dev = 0;
    BD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      BD_CUDA_TRY(cudaFuncSetAttribute(bd_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg::kSmemBytes));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * ((M + 255) / 256), (N + 255) / 256, 1);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = Gemm2Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  ++g_launch_count;
  BD_CUDA_TRY(cudaLaunchKernelEx(&cfg, bd_gemm2_kernel, ta, tw, M, N, K, epi));
  return BD_OK;
}

__global__ void bd_interleave16_kernel(const __nv_bfloat16* __restrict__ gate, const __nv_bfloat16* __restrict__ up,
                                       __nv_bfloat16* __restrict__ out, int F, int K,
                                       const __nv_bfloat16* __restrict__ bg, const __nv_bfloat16* __restrict__ bu,
                                       __nv_bfloat16* __restrict__ bo) {
  const int r = blockIdx.x;  // output row in [0, 2F)
  const int blk = r >> 5, w = r & 31;
  const int src = blk * 16 + (w & 15);
  const __nv_bfloat16* s = (w < 16 ? gate : up) + static_cast<long long>(src) * K;
  __nv_bfloat16* d = out + static_cast<long long>(r) * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) d[k] = s[k];
  if (threadIdx.x == 0 && bo) bo[r] = (w < 16 ? bg : bu)[src];
}

// W [N, K] row-major (ld) -> tile-major [ceil(N/128)][ceil(K/64)][128][64], zero padded. One CTA per tile row.
__global__ void __launch_bounds__(256) pack_weight_tiles_kernel(const __nv_bfloat16* __restrict__ w, long long ldw, int N,
                                                                int K, __nv_bfloat16* __restrict__ out) {
  const int KB = (K + 63) / 64;
  const long long tile = blockIdx.x;           // nt * KB + kb
  const int nt = static_cast<int>(tile / KB), kb = static_cast<int>(tile % KB);
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {  // 16-byte chunks of the tile
    const int r = i >> 3, c8 = (i & 7) * 8;
    const int n = nt * 128 + r, k = kb * 64 + c8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < N) {
      if (k + 8 <= K && ((ldw & 7) == 0)) {
        v = *reinterpret_cast<const uint4*>(w + n * ldw + k);
      } else {
        __nv_bfloat16 t[8];
        for (int j = 0; j < 8; ++j) t[j] = (k + j < K) ? w[n * ldw + k + j] : __float2bfloat16_rn(0.f);
        v = *reinterpret_cast<uint4*>(t);
      }
    }
    *reinterpret_cast<uint4*>(out + (tile * 128 + r) * 64 + c8) = v;
  }
}

size_t gemm_workspace_bytes(int M, int N, int K, int bn, int splits) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  GemmPlan p = plan_gemm(M, N, K, bn, splits);
  return p.splits > 1 ? static_cast<size_t>(p.splits) * M * N * sizeof(float) : 0;
}

int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const GemmEpi& epi,
              void* workspace, size_t workspace_bytes, int bn, int splits, bool pdl, cudaStream_t stream,
              bool w_tiled, int* partial_splits) {
  BD_REQUIRE(A && W && epi.out);
  if (w_tiled) ldw = ((K + 63) / 64) * 64;
  BD_REQUIRE(M > 0 && N > 0 && K > 0);
  BD_REQUIRE(bn == 0 || bn == 64 || bn == 128 || bn == 256);
  BD_REQUIRE(splits >= 0);
  BD_REQUIRE(lda >= K && ldw >= K && (lda % 8) == 0 && (ldw % 8) == 0);
  BD_REQUIRE(!epi.swiglu || ((N % 32) == 0 && !epi.gate && !epi.res && !epi.out_f32 && epi.act == 0));
  GemmPlan p = plan_gemm(M, N, K, bn, splits);
  // compute-bound shapes: the CTA-pair kernel (256 x 256 tiles, cta_group::2) on tile-major weights
  if (w_tiled && M >= 256 && (N % 256) == 0 && bn == 0 && splits == 0 && !partial_splits && gemm2_enabled()) {
    CUtensorMap ta2, tw2;
    int rc2 = make_tmap_2d_bf16(&ta2, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda),
                                kGemmBK, kGemmBM);
    if (rc2 != BD_OK) return rc2;
    const uint64_t rows2 = static_cast<uint64_t>((N + 127) / 128) * ((K + 63) / 64) * 128;
    rc2 = make_tmap_2d_bf16(&tw2, W, 64, rows2, 64, kGemmBK, 128);
    if (rc2 != BD_OK) return rc2;
    return launch_gemm2(ta2, tw2, M, N, K, epi, pdl, stream);
  }
  float* partial = nullptr;
  if (p.splits > 1) {
    const size_t need = static_cast<size_t>(p.splits) * M * N * sizeof(float);
    if (!workspace || workspace_bytes < need) return BD_ERR_WORKSPACE;
    BD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
    partial = static_cast<float*>(workspace);
  }
  CUtensorMap ta, tw;
  int rc = make_tmap_2d_bf16(&ta, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda),
                             kGemmBK, kGemmBM);
  if (rc != BD_OK) return rc;
  if (w_tiled) {
    const uint64_t rows = static_cast<uint64_t>((N + 127) / 128) * ((K + 63) / 64) * 128;
    rc = make_tmap_2d_bf16(&tw, W, 64, rows, 64, kGemmBK, static_cast<uint32_t>(p.bn < 128 ? p.bn : 128));
  } else {
    rc = make_tmap_2d_bf16(&tw, W, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(ldw),
                           kGemmBK, static_cast<uint32_t>(p.bn));
  }
  if (rc != BD_OK) return rc;
  const int wt = w_tiled ? 1 : 0;
  switch (p.bn) {
    case 64: rc = launch_gemm<64>(ta, tw, M, N, K, p.splits, partial, epi, pdl, stream, wt); break;
    case 128: rc = launch_gemm<128>(ta, tw, M, N, K, p.splits, partial, epi, pdl, stream, wt); break;
    default: rc = launch_gemm<256>(ta, tw, M, N, K, p.splits, partial, epi, pdl, stream, wt); break;
  }
  if (rc != BD_OK) return rc;
  if (partial_splits) {
    *partial_splits = p.splits;
    return BD_OK;
  }
  if (p.splits > 1) {
    const long long work = static_cast<long long>(M) * ((N + 31) / 32);
    dim3 grid(static_cast<unsigned>((work + 255) / 256));
    LaunchCfg lc(grid, dim3(256), 0, stream, pdl);
    BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, bd_splitk_epilogue_kernel, static_cast<const float*>(partial), M, N,
                                   p.splits, epi));
  }
  return BD_OK;
}

}  // namespace bd

using namespace bd;

extern "C" {

size_t bd_gemm_workspace_bytes(int M, int N, int K, int bn, int splits) {
  return gemm_workspace_bytes(M, N, K, bn, splits);
}

int bd_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                 const bd_gemm_epilogue_t* e, void* workspace, size_t workspace_bytes, int bn, int splits, int flags,
                 bd_stream_t stream_) {
  BD_REQUIRE(A && W && e && e->out);
  GemmEpi epi;
  epi.bias = static_cast<const __nv_bfloat16*>(e->bias);
  epi.gate = static_cast<const __nv_bfloat16*>(e->gate);
  epi.res = e->res;
  epi.out = e->out;
  epi.ld_gate = e->ld_gate;
  epi.ld_res = e->ld_res;
  epi.ld_out = e->ld_out;
  epi.act = e->act;
  epi.swiglu = e->swiglu;
  epi.res_f32 = e->res_f32;
  epi.out_f32 = e->out_f32;
  epi.res_mod = e->res_row_mod;
  return gemm_bf16(A, lda, W, ldw, M, N, K, epi, workspace, workspace_bytes, bn, splits, (flags & 1) != 0,
                   static_cast<cudaStream_t>(stream_), (flags & 2) != 0);
}

size_t bd_packed_weight_elems(int N, int K) {
  if (N <= 0 || K <= 0) return 0;
  return static_cast<size_t>((N + 127) / 128) * ((K + 63) / 64) * 128 * 64;
}

int bd_pack_weight_tiles(const void* W, int64_t ldw, int N, int K, void* out, bd_stream_t stream) {
  BD_REQUIRE(W && out && N > 0 && K > 0 && ldw >= K);
  BD_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const long long tiles = static_cast<long long>((N + 127) / 128) * ((K + 63) / 64);
  pack_weight_tiles_kernel<<<static_cast<unsigned>(tiles), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(W), ldw, N, K, static_cast<__nv_bfloat16*>(out));
  BD_LAUNCH_CHECK();
  return BD_OK;
}

int bd_interleave16(const void* gate, const void* up, void* out, int F, int K, const void* bias_gate,
                    const void* bias_up, void* bias_out, bd_stream_t stream) {
  BD_REQUIRE(gate && up && out && F > 0 && K > 0 && (F % 16) == 0);
  BD_REQUIRE(!bias_out || (bias_gate && bias_up));
  bd_interleave16_kernel<<<2 * F, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(gate), static_cast<const __nv_bfloat16*>(up), static_cast<__nv_bfloat16*>(out),
      F, K, static_cast<const __nv_bfloat16*>(bias_gate), static_cast<const __nv_bfloat16*>(bias_up),
      static_cast<__nv_bfloat16*>(bias_out));
  BD_LAUNCH_CHECK();
  return BD_OK;
}

}  // extern "C"
