/* bitdance_b200.h — C ABI of the B200-native BitDance image-generation hot path.
 *
 * The reference (shallowdream204/BitDance) is pure Python: it has no FFI layer, so nothing on its side dictates
 * these signatures. Each entry point below names the reference function whose arithmetic it
 * replaces (paths relative to the reference repository root). Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host; torch (or any allocator) owns the memory;
 *   - no entry point allocates, frees, synchronises or throws; work is enqueued on `stream`;
 *   - return value: 0 = BD_OK, negative = bd_status_t error (see bd_strerror);
 *   - bf16 = IEEE bfloat16 stored as uint16; "row-major [M,K] with ld" means element (m,k) at m*ld + k;
 *   - rounding points follow torch.autocast(bfloat16) semantics of the reference and are listed per function.
 *   - there is no CPU fallback: a missing GPU / wrong architecture returns BD_ERR_NO_DEVICE / BD_ERR_ARCH.
 */
#ifndef BITDANCE_B200_H_
#define BITDANCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* bd_stream_t; /* cudaStream_t */

typedef enum {
  BD_OK = 0,
  BD_ERR_INVALID = -1,   /* bad shape / null pointer / misaligned */
  BD_ERR_WORKSPACE = -2, /* workspace too small */
  BD_ERR_CUDA = -3,      /* a CUDA runtime/driver call failed (bd_last_cuda_error) */
  BD_ERR_NO_DEVICE = -4,
  BD_ERR_ARCH = -5,      /* device is not sm_100 */
  BD_ERR_UNSUPPORTED = -6
} bd_status_t;

const char* bd_strerror(int status);
int bd_last_cuda_error(void);   /* cudaError_t of the most recent BD_ERR_CUDA on this thread */
int bd_abi_version(void);       /* bumped on any signature change */
int bd_device_check(void);      /* BD_OK when the current device is sm_100 */

/* ------------------------------------------------------------------------------------------------
 * Binary quantiser + bit packing
 * ---------------------------------------------------------------------------------------------- */

/* VQModel.encode: quant = where(h > 0, +1, -1)  (modeling/vision_encoder/autoencoder.py:385-390; 0 -> -1, NaN -> -1)
 * fused with GFQ.forward's index packing idx = sum_i [x_i > 0] << i per codebook group
 * (imagenet_gen/src/gfq.py:221-239; channel 0 of a group is the LSB).
 * h: [B, C, HW] (NCHW, fp32 when h_f32 else bf16). Outputs (either may be NULL):
 *   quant  : same shape/dtype as h, values +-1
 *   packed : uint32 [B, HW, C/32] — bit c%32 of word c/32 is (h[b,c,hw] > 0)            (C % 32 == 0)
 *   indices: int32 [num_codebooks, B*HW] GFQ indices, group g = channels [g*C/ncb, (g+1)*C/ncb), C/ncb <= 31 */
int bd_sign_pack_nchw(const void* h, int h_f32, int B, int C, int HW, void* quant, uint32_t* packed,
                      int32_t* indices, int num_codebooks, bd_stream_t stream);

/* torch.sign on the AR path (modeling/t2i_pipeline.py:248): computed as (0 < x) - (x < 0), so sign(0) = 0 and sign(NaN) = 0.
 * x: fp32 [rows, C] token-major. tokens: fp32 [rows, C] (may alias x; may be NULL);
 * packed: uint32 [rows, C/32], bit set iff x > 0 (may be NULL). */
int bd_sign_tokens(const float* x, long long rows, int C, float* tokens, uint32_t* packed, bd_stream_t stream);

/* Inverse of the packing: packed uint32 [rows, C/32] -> +-1 in fp32 or bf16, token-major [rows, C]. */
int bd_unpack_tokens(const uint32_t* packed, long long rows, int C, void* out, int out_f32, bd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Weight-streaming GEMM (every nn.Linear on the path)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
  const void* bias; /* bf16 [N] or NULL */
  const void* gate; /* bf16 [M, ld_gate] or NULL: y *= gate */
  const void* res;  /* [M, ld_res] bf16 (res_f32=0) or fp32 (res_f32=1), or NULL: y += res */
  void* out;        /* [M, ld_out] bf16 (out_f32=0) or fp32 */
  int64_t ld_gate, ld_res, ld_out;
  int act;     /* 0 none, 1 SiLU, 2 GELU(tanh) */
  int swiglu;  /* 1: W rows are interleave16(gate, up) (bd_interleave16); out has N/2 columns */
  int res_f32;
  int out_f32;
} bd_gemm_epilogue_t;

/* out = epilogue(A[M,K] · W[N,K]^T): F.linear under autocast(bf16) (+ the elementwise ops that follow it in
 * TransBlock.forward flow_head_parallel_x.py:242-252, Qwen3MLP / Qwen3DecoderLayer residuals, MLPconnector
 * modeling/utils.py:16-20). A, W bf16 row-major, lda/ldw multiples of 8 elements, 16-byte aligned bases.
 * fp32 accumulation on tcgen05 tensor cores; y = bf16(acc + bias) then the epilogue stages, each rounding to bf16.
 * bn: 0 = auto, else 64/128/256. splits: 0 = auto, else split-K factor (needs workspace >= splits*M*N*4 bytes).
 * flags: bit0 = launch with programmatic dependent launch. */
int bd_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                 const bd_gemm_epilogue_t* epi, void* workspace, size_t workspace_bytes, int bn, int splits,
                 int flags, bd_stream_t stream);
size_t bd_gemm_workspace_bytes(int M, int N, int K, int bn, int splits);

/* One-time weight re-layout for SwiGLU pairs: out rows [32j, 32j+16) = gate rows [16j, 16j+16),
 * out rows [32j+16, 32j+32) = up rows [16j, 16j+16). gate/up: bf16 [F, K] (F % 16 == 0); out: bf16 [2F, K].
 * bias_* (bf16 [F]) / bias_out (bf16 [2F]) may be NULL. */
int bd_interleave16(const void* gate, const void* up, void* out, int F, int K, const void* bias_gate,
                    const void* bias_up, void* bias_out, bd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BITDANCE_B200_H_ */
