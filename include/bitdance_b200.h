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
unsigned long long bd_launch_count(void); /* kernels launched by this library since load (all threads) */

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

/* AR-step form of the same op: x fp32 [B*pn, C] (the sampler output of one block). Writes (each optional):
 *   grid        fp32 [B, rows_per_image, C] tokens at rows row0..row0+pn (out_tokens.append / torch.cat, :250,270)
 *   tokens_bf16 bf16 [dup*B*pn, C]: the +-1/0 tokens repeated for `dup` sequence groups (cond | uncond rows of
 *               curr_tokens, the MLPconnector input)
 *   packed      uint32 [B, rows_per_image, C/32] bits (x > 0) at the same rows */
int bd_sign_tokens_ex(const float* x, int B, int pn, int C, float* grid, long long rows_per_image, long long row0,
                      void* tokens_bf16, int dup, uint32_t* packed, bd_stream_t stream);

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
  int res_row_mod; /* > 0: res is a [res_row_mod, N] table, row m uses res[m % res_row_mod] (pos-embed broadcast) */
} bd_gemm_epilogue_t;

/* out = epilogue(A[M,K] · W[N,K]^T): F.linear under autocast(bf16) (+ the elementwise ops that follow it in
 * TransBlock.forward flow_head_parallel_x.py:242-252, Qwen3MLP / Qwen3DecoderLayer residuals, MLPconnector
 * modeling/utils.py:16-20). A, W bf16 row-major, lda/ldw multiples of 8 elements, 16-byte aligned bases.
 * fp32 accumulation on tcgen05 tensor cores; y = bf16(acc + bias) then the epilogue stages, each rounding to bf16.
 * bn: 0 = auto, else 64/128/256. splits: 0 = auto, else split-K factor (needs workspace >= splits*M*N*4 bytes).
 * flags: bit0 = launch with programmatic dependent launch; bit1 = W is tile-major (bd_pack_weight_tiles; ldw ignored). */
int bd_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                 const bd_gemm_epilogue_t* epi, void* workspace, size_t workspace_bytes, int bn, int splits,
                 int flags, bd_stream_t stream);
size_t bd_gemm_workspace_bytes(int M, int N, int K, int bn, int splits);

/* One-time weight re-layout for HBM streaming: W bf16 [N, K] row-major -> tile-major
 * [ceil(N/128)][ceil(K/64)][128][64] (zero padded; bd_packed_weight_elems elements). A 128x64 tile is 16 KB CONTIGUOUS and
 * the k-blocks of one 128-row slab are adjacent, so a CTA of the GEMM streams one contiguous HBM region instead of
 * 128-byte pieces K*2 bytes apart (DRAM-page locality). */
int bd_pack_weight_tiles(const void* W, int64_t ldw, int N, int K, void* out, bd_stream_t stream);
size_t bd_packed_weight_elems(int N, int K);

/* One-time weight re-layout for SwiGLU pairs: out rows [32j, 32j+16) = gate rows [16j, 16j+16),
 * out rows [32j+16, 32j+32) = up rows [16j, 16j+16). gate/up: bf16 [F, K] (F % 16 == 0); out: bf16 [2F, K].
 * bias_* (bf16 [F]) / bias_out (bf16 [2F]) may be NULL. */
int bd_interleave16(const void* gate, const void* up, void* out, int F, int K, const void* bias_gate,
                    const void* bias_up, void* bias_out, bd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Attention (diffusion-head MHA and Qwen3 GQA over a paged KV cache)
 * ---------------------------------------------------------------------------------------------- */

/* softmax(Q K^T * scale [+ causal mask]) V with fp32 scores/softmax, bf16 P and output (flash-attention semantics):
 * Attention.forward flow_head_parallel_x.py:192-220 and Qwen3Attention (transformers, called from
 * modeling/t2i_pipeline.py:199-266; an all-ones mask = causal 0). Strides in ELEMENTS: (b, s, h, d) at
 * base + b*sb + s*ss + h*sh + d. KV either strided (page_table NULL) or paged: page_table int32 [B, max_pages],
 * pools k/v laid out [page][Hkv][64 tokens][head_dim]. causal: key j visible to query i iff j <= i + (Sk - Sq).
 * splits: 0 = auto split of the KV range (deterministic fixed-order combine), needs bd_attention_workspace_bytes. */
int bd_attention_bf16(const void* q, int64_t q_sb, int64_t q_ss, int64_t q_sh, const void* k, const void* v,
                      int64_t k_sb, int64_t k_ss, int64_t k_sh, const int32_t* page_table, int max_pages, void* out,
                      int64_t o_sb, int64_t o_ss, int64_t o_sh, int B, int Sq, int Sk, int Hq, int Hkv, int head_dim,
                      int causal, float scale, int splits, void* workspace, size_t workspace_bytes, int flags,
                      bd_stream_t stream);
size_t bd_attention_workspace_bytes(int B, int Hq, int Sq, int Sk, int head_dim, int splits);

/* ------------------------------------------------------------------------------------------------
 * Binary-diffusion vision head: DiffHead.sample in one call
 * ---------------------------------------------------------------------------------------------- */

#define BD_HEAD_MAX_BLOCKS 16

typedef struct {
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b; /* fp32 [D] */
  /* bf16 nn.Linear weights [out, in] and bf16 biases. w1: interleave16(h1 rows, h2 rows) when use_swiglu
   * (w1.weight.chunk(2) = rows [0,hidden) and [hidden,2*hidden)), else mlp.0; w2: w2 / mlp.2. */
  const void *wqkv_w, *wqkv_b, *wo_w, *wo_b, *w1_w, *w1_b, *w2_w, *w2_b;
} bd_head_block_t;

typedef struct {
  int D;        /* ch_latent */
  int Dz;       /* ch_cond */
  int C;        /* ch_target (bits per token) */
  int hidden;   /* int(1.5 * D) */
  int n_blocks; /* depth_latent */
  int n_ada;    /* depth_adanln */
  int head_dim; /* 128 (modeling/) or 64 (imagenet_gen/) */
  int use_swiglu;
  int w_tiled;     /* 1: every Linear weight below (except final_w) is in bd_pack_weight_tiles layout;
                    * 2: stream-major (bd_stream_pack_weight, n_ctas = stream_ctas; w1 with perm 1 + its packed bias when
                    *    use_swiglu; wo / w2 packed with ksplit = bd_stream_ksplit(D, K, stream_ctas), all others ksplit 1):
                    *    bd_head_sample then runs the whole sampler as ONE persistent kernel when B*cfg_mult*pn <= 128 */
  int out_sigmoid; /* 1: 2*sigmoid(.)-1 (flow_head_parallel_x.py:341-342); 0: imagenet_gen diff_head_parallel.py:310 */
  int stream_ctas; /* w_tiled == 2: the CTA count the weights were packed for (bd_stream_num_ctas()) */
  int reserved_;
  const void *input_proj_w, *input_proj_b;
  const void *time0_w, *time0_b, *time2_w, *time2_b; /* time_embed.mlp.0 / .2 */
  const void *cond_w, *cond_b;
  const void *ada_w, *ada_b;     /* rows: ada_ln_blocks[0] (6D) | ... | ada_ln_blocks[n_ada-1] | final_layer.ada_ln_modulation (2D) */
  const void *final_w, *final_b; /* final_layer.linear [C, D] */
  bd_head_block_t blocks[BD_HEAD_MAX_BLOCKS];
} bd_head_weights_t;

/* DiffHead.sample(z, cfg, num_sampling_steps) = euler_maruyama (modeling/vision_head/sampling_x.py:44-97) over
 * TransEncoder.forward (flow_head_parallel_x.py:325-342).
 *   cond       fp32 [B*cfg_mult*pn, Dz]: z, conditional rows first then unconditional rows (t2i_pipeline.py:244)
 *   noise      fp32 [(S+1), B*pn, C]: noise[0] = the torch.randn of :60, noise[1+i] = the randn_like of step i (:40)
 *   sched_host HOST fp32 [(S+1), 8]: per evaluation {t, dt, clamp(1-t,0.05), var, 1-t, sqrt(2(1-t)dt), 0, 0} computed
 *              with the reference's fp32 scalar arithmetic (t is the running sum of dt); row S: t=0.95.., dt=last step
 *   x_out      fp32 [B*pn, C] final continuous sample (the reference returns cat([x]*cfg_mult); sign() is the caller's)
 *   trace      optional fp32 [(S+1), B*cfg_mult*pn, C]: the network output of every evaluation (tests)
 * Rounding: autocast(bf16) policy (oracle/head.py docstring); sampler state fp32 with unfused mul/add. */
int bd_head_sample(const bd_head_weights_t* w, const float* cond, const float* noise, const float* sched_host, int B,
                   int pn, int cfg_mult, float cfg, int S, float* x_out, float* trace, void* workspace,
                   size_t workspace_bytes, int flags, bd_stream_t stream);
size_t bd_head_workspace_bytes(const bd_head_weights_t* w, int B, int pn, int cfg_mult, int S);
/* tests / debugging: byte offsets of the named regions inside the workspace (xb, h, ... in the order documented in
 * csrc/bd_head.cu for the weights' layout kind); returns how many were written (<= cap). */
int bd_head_ws_offsets(const bd_head_weights_t* w, int B, int pn, int cfg_mult, int S, size_t* out, int cap);
/* Program policy of the persistent sampler. mode 1 (default): the adaLN-modulation GEMM of evaluation i + 1 (it depends
 * on t and c only, flow_head_parallel_x.py:330-336) is cut into k-sliced pieces that run as FILLERS inside evaluation i,
 * in the bubbles of the dependent chain (row ops, op boundaries) — see csrc/bd_stream.cuh; mode 0: in line at the start
 * of every evaluation (the round-1 program). row_kb / gemm_kb: wanted piece size (k-blocks of 128 rows x 64) in a slot
 * that hides a row op / a GEMM -> GEMM boundary; <= 0 keeps the current value. Results are identical in both modes up to
 * the fp32 accumulation order of that one GEMM. */
int bd_head_set_fillers(int mode, int row_kb, int gemm_kb);
/* tests: the piece plan (host arithmetic only) for a GEMM of P passes x KB k-blocks per CTA and n_slots bubbles of wanted
 * sizes want[]: out = n x {slot, pass, kb0, kbn}; returns n (< 0: does not fit `cap`). */
int bd_head_plan_pieces(int P, int KB, const int* want, int n_slots, int* out, int cap);

/* ------------------------------------------------------------------------------------------------
 * Qwen3 decoder stack with a paged KV cache
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
  const void *ln1_w, *ln2_w;       /* bf16 [D]: input_layernorm / post_attention_layernorm */
  const void *q_norm_w, *k_norm_w; /* bf16 [head_dim] */
  const void* wqkv;                /* bf16 [(Hq + 2 Hkv) * head_dim, D]: rows q_proj | k_proj | v_proj */
  const void* wo;                  /* bf16 [D, Hq * head_dim] */
  const void* w_gate_up;           /* bf16 [2 I, D]: bd_interleave16(gate_proj, up_proj) */
  const void* w_down;              /* bf16 [D, I] */
  /* optional stream-major copies (bd_stream_pack_weight, n_ctas = stream_ctas) for AR blocks of one 128-row tile:
   * wqkv_s [q|k|v] rows (ksplit 1), wo_s (ksplit bd_stream_ksplit(D, Hq*hd)), w_gate_up_s = cat(gate_proj, up_proj) with
   * perm 1 (ksplit 1), w_down_s (ksplit bd_stream_ksplit(D, I)); NULL when stream_ctas == 0 */
  const void *wqkv_s, *wo_s, *w_gate_up_s, *w_down_s;
} bd_llm_layer_t;

typedef struct {
  int D, I, n_layers, Hq, Hkv, head_dim;
  float eps;
  int w_tiled;                  /* 1: wqkv / wo / w_gate_up / w_down are in bd_pack_weight_tiles layout */
  int stream_ctas;              /* > 0: the *_s weights exist, packed for this many CTAs: fp32-stream, non-causal passes with
                                 * R*S <= 128 run every layer's four Linears, residual adds and RMSNorms as persistent
                                 * bd_stream_kernel segments (RoPE / KV append / paged attention stay separate kernels) */
  int variant;                  /* 0: Qwen3 (q/k RMSNorm over head_dim, rotate_half RoPE, tables fp32 [pos, head_dim]);
                                 * BD_LLM_ROPE_PAIRS: the ImageNet class-conditional decoder (imagenet_gen/src/
                                 * layers_parallel.py): no q/k norm (q_norm_w / k_norm_w NULL), interleaved-pair RoPE in fp32
                                 * with tables [pos, head_dim/2] (the 2-D table of precompute_freqs_cis_2d), fp32 stream only */
  const void* final_norm_w;     /* bf16 [D] */
  const bd_llm_layer_t* layers; /* HOST array [n_layers] */
  const void* emb_norm_w;       /* bf16 [D] or NULL: RMSNorm applied to the input embeddings in place before the first
                                 * layer (BitDance.forward_model, imagenet_gen/src/model_parallel.py:343-345) */
  const void* layer_tab;        /* DEVICE table of pointers [n_layers + 1][8] (the last row unused): {wqkv_s, wo_s,
                                 * w_gate_up_s, w_down_s, ln1_w, ln2_w, q_norm_w, k_norm_w} of every layer, or NULL. With it
                                 * (and stream_ctas > 0) an fp32-stream, non-causal pass of R*S <= 128 rows runs as ONE
                                 * persistent launch: all layers' Linears, norms, RoPE + KV append, paged attention and its
                                 * split-KV combine are ops of bd_stream_kernel (csrc/bd_llm.cu::llm_stream_all) */
} bd_llm_weights_t;
#define BD_LLM_ROPE_PAIRS 1

/* Qwen3Model.forward(inputs_embeds, past_key_values, attention_mask) as the reference calls it at
 * modeling/t2i_pipeline.py:199,211,224,229 (prefill) and :261,266 (AR block) — third-party transformers==4.57.0.
 *   hidden      [R*S, D] residual stream, fp32 (stream_f32=1: AR blocks, inputs_embeds = bf16 + fp32 pos-embed) or bf16
 *               (prefill); holds inputs_embeds on entry and is overwritten.
 *   seq_lens    DEVICE int32 [R] (R <= 256): tokens already cached per sequence; position of token (b,s) is
 *               seq_lens[b]+s; incremented by S on the device at the end (graph-replay safe). sk_bound = host upper
 *               bound of seq_lens[b]+S (planning only).
 *   causal      1: first prefill call; 0: all-ones mask = the S new tokens see the whole cache and each other
 *   kv_pool     bf16; layer l keys at kv_pool + l*kv_layer_stride, values at + kv_v_offset (elements); each pool is
 *               [n_pages][Hkv][64][head_dim]; page_table DEVICE int32 [R, max_pages]
 *   rope_cos/sin fp32 [>= sk_bound, head_dim] tables (cos/sin of cat(freqs, freqs), Qwen3RotaryEmbedding)
 *   out         final-RMSNorm output [R*S, D] in the stream dtype, or fp32 when out_add is given: out_add (fp32
 *               [out_add_mod, D]) is added row-wise with m % out_add_mod: h_fused = last_hidden_state + pos_embed
 *               (t2i_pipeline.py:245)
 * Rounding policy: oracle/llm.py. */
int bd_llm_forward(const bd_llm_weights_t* w, void* hidden, int stream_f32, int R, int S, int* seq_lens, int sk_bound,
                   int causal, void* kv_pool, int64_t kv_layer_stride, int64_t kv_v_offset, const int32_t* page_table,
                   int max_pages, const float* rope_cos, const float* rope_sin, void* out, const float* out_add,
                   int out_add_mod, int attn_splits, void* workspace, size_t workspace_bytes, int flags,
                   bd_stream_t stream);
size_t bd_llm_workspace_bytes(const bd_llm_weights_t* w, int R, int S, int attn_splits);

/* ------------------------------------------------------------------------------------------------
 * Binary tokenizer (conv encoder / decoder) building blocks — all activations NHWC
 * ---------------------------------------------------------------------------------------------- */

/* nn.Conv2d (3x3 pad 1 stride 1|2, or 1x1) as an implicit GEMM on tcgen05 (autoencoder.py:31-38,72-77,94,105,142,170,240).
 *   x        bf16 NHWC [B, H_out, W_out, Cin] (stride 1) or the 4-phase split [4, B, H_out, W_out, Cin] (stride 2,
 *            bd_phase_split_nhwc); Cin % 8 == 0 (pad the 3-channel image to 8)
 *   w_packed bf16 [Cout, k*k*Cin]: element (co, tap, ci) of weight.permute(0,2,3,1)
 *   bias     bf16 [Cout] or NULL;  res [B,H,W,Cout] bf16/fp32 or NULL (ResBlock: x + residual, :57)
 *   out_mode 0: NHWC [B,H,W,Cout] bf16/fp32 (out_f32) = res + bf16(acc + bias)
 *            1: depth_to_space(2) scatter (Upsampler, :198-249): bf16 [B, 2H, 2W, Cout/4]
 *            2: NCHW [B, Cout, H, W] (decoder conv_out -> image) */
int bd_conv2d_nhwc(const void* x, const void* w_packed, const void* bias, const void* res, int res_f32, void* out,
                   int out_f32, int out_mode, int B, int H_out, int W_out, int Cin, int Cout, int ksize, int stride,
                   int flags, bd_stream_t stream);
/* x bf16 [B, 2H, 2W, C] -> [4, B, H, W, C], phase 2a+b = x[:, 2y+a, 2x+b, :]. */
int bd_phase_split_nhwc(const void* x, void* out, int B, int H_out, int W_out, int C, bd_stream_t stream);

/* fp32 NCHW image -> bf16 NHWC with channels zero-padded to C_pad (the autocast input cast of Encoder.conv_in). */
int bd_nchw_to_nhwc_bf16(const float* x, void* out, int B, int C, int H, int W, int C_pad, bd_stream_t stream);
/* tokens fp32 [B, h*w, C] in patch-raster order -> bf16 NHWC [B, h, w, C]
 * ('b (h w p1 p2) c -> b c (h p1) (w p2)', modeling/t2i_pipeline.py:280). */
int bd_tokens_to_grid(const float* tokens, void* out, int B, int h, int w, int C, int ps, bd_stream_t stream);
/* fp32 -> bf16 (round to nearest even): the autocast input cast of a conv / Linear fed from an fp32 tensor. */
int bd_cast_f32_bf16(const float* x, void* out, long long n, bd_stream_t stream);
int bd_nhwc_to_nchw(const void* x_bf16, void* out, int out_f32, int B, int C, int H, int W, bd_stream_t stream);

/* nn.GroupNorm(32, C, eps) over NHWC x [B, HW, C] (bf16 or fp32), statistics in fp32 (two-stage, deterministic).
 *   mode 0: out bf16 = swish(gn(x) * weight[c] + bias[c])   (ResBlock.forward :44-49, norm_out + swish :124-125,191-192)
 *   mode 1: out fp32 = gn(x) * weight[b,c] + bias[b,c]       (AdaptiveGroupNorm :274-275; weight/bias from bd_adagn_params)
 * C must be 8 * 2^k in [32, 2048]. */
int bd_groupnorm_nhwc(const void* x, int x_f32, int B, long long HW, int C, const float* weight, const float* bias,
                      int mode, void* out, void* workspace, size_t workspace_bytes, float eps, int flags,
                      bd_stream_t stream);
size_t bd_groupnorm_workspace_bytes(int B, long long HW);
/* AdaptiveGroupNorm scale/bias (autoencoder.py:263-272): gamma = Linear(sqrt(var_unbiased(z) + 1e-6)),
 * beta = Linear(mean(z)) over the +-1 grid z (bf16 NHWC [B, hw, Cz]); outputs fp32 [B, C] holding bf16 values. */
int bd_adagn_params(const void* z, int B, int hw, int Cz, const void* gamma_w, const void* gamma_b, const void* beta_w,
                    const void* beta_b, int C, float* gamma, float* beta, bd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Persistent weight-streaming engine (csrc/bd_stream.cuh): one cooperative kernel, one CTA per SM, interprets a program
 * of GEMM / row / attention ops with the weight stream running ahead of every dependency. No reference counterpart (the
 * reference launches ~100 eager kernels per head evaluation, flow_head_parallel_x.py:325-342); used by bd_head_sample.
 * ---------------------------------------------------------------------------------------------- */
int bd_stream_num_ctas(void); /* CTAs of the persistent kernel = SMs of the current device */
/* profiling aid: when buf != NULL every later persistent launch records, for its first max_ops ops, 8 %globaltimer stamps
 * per (op, CTA) into buf (uint64 [max_ops][n_ctas][8]): 0 A loads start, 1 first MMA, 2 last MMA issued, 3 accumulator
 * ready, 4 epilogue/row work done, 5 arrival published. NULL switches it off. */
int bd_stream_set_debug(void* buf, int max_ops);
/* tuning / experiments: split of the 7 shared-memory ring slots (32 KB = two 64-wide k-blocks each) between weights and
 * activations (default 5 + 2; the activation ring needs >= 2 slots, it also hosts the attention tiles), and epilogue
 * experiment switches (0 = product behaviour; bit0 no bias loads, bit1 no stores — measurement only, results are wrong) */
int bd_stream_set_tuning(int w_slots, int a_slots, int mode);
/* L2 prefetch distance of the weight stream in ring steps of ~28 KB per CTA (0 = off): the producer warp issues
 * cp.async.bulk.prefetch.L2 that far ahead of its shared-memory loads, so HBM keeps streaming while the ring is full. */
int bd_stream_set_prefetch(int steps);
/* sleep (ns) between two polls of the grid-barrier counter (default 32; tight polling is slower) */
int bd_stream_set_poll_ns(int ns);
/* k-split (1, 2 or 4; default 4) bd_stream_ksplit returns for small-N Linears. Set BEFORE packing weights. */
int bd_stream_set_ksplit(int ksplit);
/* k-split the engine uses for a Linear whose output goes through fp32 partials (N small next to the SM count) */
int bd_stream_ksplit(int N, int K, int n_ctas);
size_t bd_stream_packed_elems(int N, int K); /* bf16 elements of a stream-packed [N, K] weight (K padded to 64) */
/* tests: CTA c's share of a GEMM op: out[0..5] = k-split index, first 16-row unit, units, first k-block, k-blocks, passes;
 * then per pass (first unit relative to unit0, width in rows <= 128, offset of its first slot in the packed weight in
 * 2 KB units). Host arithmetic only (the same inline functions the kernel and the packer use). */
int bd_stream_partition_info(int N, int K, int ksplit, int n_ctas, int c, long long* out, int cap);
/* W [N, K] row-major bf16 (ldw) -> stream-major: the weight bytes CTA c needs are one contiguous range, already in the
 * 128B-swizzled K-major shared-memory image. perm 0: rows as they are; perm 1 (SwiGLU, N = 2*hidden): packed unit u =
 * [rows 8u..8u+7 of W[:hidden] | rows 8u..8u+7 of W[hidden:]]. bias (nullable) -> bias_out [N] in packed row order. */
int bd_stream_pack_weight(const void* W, int64_t ldw, int N, int K, int ksplit, int n_ctas, int perm, int hidden,
                          const void* bias, void* out, void* bias_out, bd_stream_t stream);
/* n_w GEMMs (weights W_packed + j * w_stride_bytes) x `repeat` through the persistent kernel — tests / micro-benchmarks.
 * A: blocked bf16 [ceil(K/64)][128][64] (128B-swizzled rows; see tests/test_stream_gpu.py); epi 0: bf16(act(acc + bias)),
 * row-major (ld_out) or blocked; epi 1: SwiGLU-8 -> [M, N/2]; epi 2: fp32 partials [ksplit][M][N]. sync: 4 bytes. */
int bd_stream_gemm(const void* A_blocked, const void* W_packed, int64_t w_stride_bytes, int n_w, const void* bias_packed,
                   void* out, int64_t ld_out, int M, int N, int K, int ksplit, int epi, int act, int out_blocked,
                   int n_ctas, int repeat, void* sync, bd_stream_t stream);

/* tests: `repeat` x { one main GEMM op (W_main, bf16 row-major out_main [M, N]); then the GEMM (W_fill, bias_fill) ->
 * out_fill [M, N] executed as FILLER pieces: every pass of every CTA cut into n_slices k-ranges accumulated in the third
 * TMEM buffer, outside the grid-barrier protocol }. Both weights [N, K], ksplit 1, same A. */
int bd_stream_gemm_filler(const void* A_blocked, const void* W_main, const void* W_fill, const void* bias_fill,
                          void* out_main, void* out_fill, int M, int N, int K, int n_ctas, int n_slices, int repeat,
                          void* sync, bd_stream_t stream);

/* ---- measurement only (scripts/stream_probe.py; not on the product path, replaces nothing in the reference) ----
 * Streams n_ctas * w_per_cta bytes of `w` HBM -> shared memory through the same bulk-copy/mbarrier ring the GEMM uses
 * (w_stages x w_chunk bytes per CTA) and, when x_chunk > 0, re-reads x_chunk bytes of the L2-resident buffer `x` per W
 * chunk in every CTA (the activation re-read of a weight-streaming GEMM). Gives the ceiling the GEMM is judged against. */
int bd_probe_stream(const void* w, long long w_per_cta, int w_chunk, int w_stages, const void* x, int x_bytes,
                    int x_chunk, int x_stages, int n_ctas, bd_stream_t stream);

/* measurement only (scripts/hmma_probe.py): per-warp cycles of `iters` rounds of `chains` independent
 * mma.sync.m16n8k16 (bf16, fp32 accumulate) with `warps` warps per CTA -> out_cycles[n_ctas * warps]. */
int bd_probe_hmma(int warps, int chains, int iters, int n_ctas, long long* out_cycles, float* sink, bd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BITDANCE_B200_H_ */
