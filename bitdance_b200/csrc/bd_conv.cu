// bd_conv.cu — C-ABI for the implicit-GEMM convolution (kernel in bd_conv.cuh) and its layout helpers.
#include "bd_conv.cuh"

namespace bd {

// stride-2 input de-interleave: x [B, 2H, 2W, C] -> phases [4, B, H, W, C], phase p = 2a + b holds x[:, 2y+a, 2x+b, :]
__global__ void __launch_bounds__(256) phase_split_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B,
                                                          int H, int W, int Cv /* C/8 */) {
  grid_dep_launch();
  grid_dep_wait();
  const long long n = static_cast<long long>(B) * (2 * H) * (2 * W) * Cv;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % Cv);
  long long r = i / Cv;
  const int xx = static_cast<int>(r % (2 * W));
  r /= (2 * W);
  const int yy = static_cast<int>(r % (2 * H));
  const int b = static_cast<int>(r / (2 * H));
  const int p = ((yy & 1) << 1) | (xx & 1);
  out[(((static_cast<long long>(p) * B + b) * H + (yy >> 1)) * W + (xx >> 1)) * Cv + c] = x[i];
}

static void pick_tile(int H, int W, int* TW, int* TH) {
  long long best = -1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = 128 / tw;
    const long long padded = static_cast<long long>((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th;
    if (best < 0 || padded < best) {
      best = padded;
      *TW = tw;
      *TH = th;
    }
  }
}

}  // namespace bd

using namespace bd;

extern "C" {

int bd_phase_split_nhwc(const void* x, void* out, int B, int H_out, int W_out, int C, bd_stream_t stream) {
  BD_REQUIRE(x && out && B > 0 && H_out > 0 && W_out > 0 && C > 0 && (C % 8) == 0);
  const long long n = static_cast<long long>(B) * (2 * H_out) * (2 * W_out) * (C / 8);
  LaunchCfg lc(dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), false);
  BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, phase_split_kernel, static_cast<const uint4*>(x), static_cast<uint4*>(out), B,
                                 H_out, W_out, C / 8));
  return BD_OK;
}

int bd_conv2d_nhwc(const void* x, const void* w_packed, const void* bias, const void* res, int res_f32, void* out,
                   int out_f32, int out_mode, int B, int H_out, int W_out, int Cin, int Cout, int ksize, int stride,
                   int flags, bd_stream_t stream_) {
  BD_REQUIRE(x && w_packed && out && B > 0 && H_out > 0 && W_out > 0 && Cin > 0 && Cout > 0);
  BD_REQUIRE((ksize == 3 && (stride == 1 || stride == 2)) || (ksize == 1 && stride == 1));
  BD_REQUIRE((Cin % 8) == 0);
  BD_REQUIRE(out_mode >= 0 && out_mode <= 2);
  BD_REQUIRE(out_mode != 1 || ((Cout % 128) == 0 && !res && !out_f32));
  BD_REQUIRE(out_mode != 2 || !res);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ConvGeom g{};
  g.B = B;
  g.H = H_out;
  g.W = W_out;
  g.Cout = Cout;
  g.Cin_pad = Cin;
  g.cblocks = (Cin + 63) / 64;
  g.taps = ksize * ksize;
  pick_tile(H_out, W_out, &g.TW, &g.TH);
  g.tiles_x = (W_out + g.TW - 1) / g.TW;
  g.tiles_y = (H_out + g.TH - 1) / g.TH;
  for (int t = 0; t < g.taps; ++t) {
    const int ky = t / ksize, kx = t % ksize;
    if (ksize == 1) {
      g.dx[t] = g.dy[t] = g.plane[t] = 0;
    } else if (stride == 1) {
      g.dy[t] = static_cast<signed char>(ky - 1);
      g.dx[t] = static_cast<signed char>(kx - 1);
      g.plane[t] = 0;
    } else {  // input row 2y + ky - 1: ky=0 -> odd phase of row y-1; ky=1 -> even phase; ky=2 -> odd phase of row y
      const int a = (ky == 1) ? 0 : 1, bb = (kx == 1) ? 0 : 1;
      g.dy[t] = static_cast<signed char>(ky == 0 ? -1 : 0);
      g.dx[t] = static_cast<signed char>(kx == 0 ? -1 : 0);
      g.plane[t] = static_cast<signed char>(2 * a + bb);
    }
  }
  g.d2s = out_mode == 1;
  g.nchw_out = out_mode == 2;
  const int planes = (stride == 2) ? 4 : 1;
  CUtensorMap ta, tw;
  int rc = make_tmap_4d_bf16(&ta, x, static_cast<uint64_t>(Cin), static_cast<uint64_t>(W_out),
                             static_cast<uint64_t>(H_out), static_cast<uint64_t>(planes) * B, 64, g.TW, g.TH, 1);
  if (rc != BD_OK) return rc;
  const uint64_t Ktot = static_cast<uint64_t>(g.taps) * Cin;
  const int bn = (Cout <= 64) ? 64 : 128;
  rc = make_tmap_2d_bf16(&tw, w_packed, Ktot, static_cast<uint64_t>(Cout), Ktot, 64, bn);
  if (rc != BD_OK) return rc;
  GemmEpi epi;
  epi.bias = static_cast<const __nv_bfloat16*>(bias);
  epi.res = res;
  epi.res_f32 = res_f32;
  epi.ld_res = Cout;
  epi.out = out;
  epi.ld_out = Cout;
  epi.out_f32 = out_f32;
  dim3 grid(static_cast<unsigned>(static_cast<long long>(B) * g.tiles_y * g.tiles_x), (Cout + bn - 1) / bn);
  const bool pdl = (flags & 1) != 0;
  if (bn == 128) {
    static bool set = false;
    const int smem = kConvStages * (128 + 128) * 128 + 1024 + 256;
    if (!set) {
      BD_CUDA_TRY(cudaFuncSetAttribute(bd_conv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = true;
    }
    LaunchCfg lc(grid, dim3(kConvThreads), smem, stream, pdl);
    BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, bd_conv_kernel<128>, ta, tw, g, epi));
  } else {
    static bool set = false;
    const int smem = kConvStages * (128 + 64) * 128 + 1024 + 256;
    if (!set) {
      BD_CUDA_TRY(cudaFuncSetAttribute(bd_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = true;
    }
    LaunchCfg lc(grid, dim3(kConvThreads), smem, stream, pdl);
    BD_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, bd_conv_kernel<64>, ta, tw, g, epi));
  }
  return BD_OK;
}

}  // extern "C"
