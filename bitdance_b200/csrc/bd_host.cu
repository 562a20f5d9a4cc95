// bd_host.cu — status strings, device checks, TMA tensor-map construction (driver entry point, no libcuda link).
#include "bd_host.h"

#include <mutex>

namespace bd {

thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return BD_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld_elems * 2) & 15)) return BD_ERR_INVALID;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = static_cast<int>(r);
    return BD_ERR_CUDA;
  }
  return BD_OK;
}

int make_tmap_4d_bf16(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B,
                      uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_b) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return BD_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((C * 2) & 15)) return BD_ERR_INVALID;
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, box_b};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = static_cast<int>(r);
    return BD_ERR_CUDA;
  }
  return BD_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace bd

extern "C" {

const char* bd_strerror(int status) {
  switch (status) {
    case BD_OK: return "ok";
    case BD_ERR_INVALID: return "invalid argument (shape, null pointer or alignment)";
    case BD_ERR_WORKSPACE: return "workspace too small";
    case BD_ERR_CUDA: return "CUDA call failed (see bd_last_cuda_error)";
    case BD_ERR_NO_DEVICE: return "no CUDA device";
    case BD_ERR_ARCH: return "device is not sm_100 (B200)";
    case BD_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

int bd_last_cuda_error(void) { return bd::g_last_cuda_error; }
int bd_abi_version(void) { return 1; }
unsigned long long bd_launch_count(void) { return bd::g_launch_count; }

int bd_device_check(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return BD_ERR_NO_DEVICE;
  int dev = 0, major = 0;
  BD_CUDA_TRY(cudaGetDevice(&dev));
  BD_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? BD_OK : BD_ERR_ARCH;
}

}  // extern "C"
