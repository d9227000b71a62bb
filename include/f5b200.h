/* f5b200 — C ABI of the B200-native F5-TTS hot path (libf5b200.so).
 *
 * The reference (hungkq-1724/EraXviF5TTS) has no FFI: its boundary is the Python class API
 * F5TTSWrapper / CFM.sample / DiT.forward / MelSpec / vocoder.decode.  The Python mirror of
 * those classes (package eraxvif5tts_b200) owns torch tensors for device memory and calls THIS library through
 * ctypes with raw device pointers and a CUDA stream.  Every entry point below names the reference code it replaces
 * (paths relative to /root/reference/src/f5_tts/).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types;
 *   - return 0 on success, a negative code on failure; f5b_last_error() returns the message (thread local);
 *   - the caller owns every buffer; the library never allocates device memory;
 *   - every call is stream-ordered, asynchronous and CUDA-graph capturable; `stream` is a cudaStream_t;
 *   - activations are token-major: row = b * rows_per_batch + position;
 *   - "bf16" buffers hold __nv_bfloat16, weights are nn.Linear layout [out_features, in_features] in bf16;
 *   - sm_100a only.  There is no CPU or generic-GPU fallback.
 */
#ifndef F5B200_H_
#define F5B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F5B_ABI_VERSION 1

typedef void* f5b_stream_t; /* cudaStream_t */

/* ---- epilogues of the tcgen05 GEMM engine ------------------------------------------------------------------ */
enum {
  F5B_EPI_BF16 = 0,       /* out bf16[M,ldc]  = act(acc + bias)                                               */
  F5B_EPI_F32 = 1,        /* out f32 [M,ldc]  = act(acc + bias) (+ addsrc[row,:])  ; optional bf16 copy in out2 */
  F5B_EPI_QKV_ROPE = 2,   /* out bf16[M,ldc] = acc + bias with rotary embedding on the first rope_heads heads of the q and k sections */
  F5B_EPI_GATE_RESID = 3, /* out f32[M,ldc] += gate[b,:] * (acc + bias), rows with pos >= lens[b] untouched    */
  F5B_EPI_BF16_DUAL = 4   /* training forward (FeedForward, model/modules.py:348-353 under autograd): out bf16[M,ldc] = acc + bias (the
                             pre-activation the backward needs) AND out2 bf16[M,ldc2] = act(out) from one accumulator tile; act = GELU_TANH */
};
enum { F5B_ACT_NONE = 0, F5B_ACT_GELU_TANH = 1, F5B_ACT_GELU_ERF = 2, F5B_ACT_SILU = 3, F5B_ACT_MISH = 4 /* f5b_act_fwd/bwd only */ };

typedef struct F5bGemmArgs {
  int32_t M, N, K;
  int32_t epi, act;
  const float* bias;         /* [N] or NULL */
  void* out;                 /* see epilogue */
  int32_t ldc;
  void* out2;                /* F32: optional bf16 copy; BF16_DUAL: the activated output */
  int32_t ldc2;
  void* out3;                /* unused (kept for ABI stability) */
  const float* addsrc;       /* F32: optional f32 [M,ld_add] added to the result */
  int32_t ld_add;
  int32_t rows_per_batch;    /* QKV_ROPE / GATE_RESID: n (positions per batch row; position = row % n) */
  const float* gate;         /* GATE_RESID: f32 gate, element (b, col) at gate[b*gate_bstride + col]; NULL -> 1 */
  int64_t gate_bstride;
  const int32_t* lens;       /* GATE_RESID: int32 [B] valid length per batch row, or NULL (all rows valid) */
  int32_t batch_mod;         /* gate / lens are indexed by b % batch_mod (a CFG-fused batch shares them); 0 -> b */
  const float* rope;         /* QKV_ROPE: f32 [n, 32, 2] (cos, sin) */
  int32_t rope_heads;        /* QKV_ROPE: heads that get the rotary embedding (pe_attn_head; H for all) */
  int32_t heads;             /* QKV_ROPE: H (N must be 3*H*64) */
  int32_t tf32;              /* operand mode: 0 = A, W bf16 (kind::f16); 1 = A, W fp32 words rounded to tf32 (kind::tf32) — the
                                BF16 / QKV_ROPE epilogues then write fp32 rounded to tf32, F32's out2 is an fp32 tf32 copy */
} F5bGemmArgs;

const char* f5b_last_error(void);
int f5b_abi_version(void);

/* C = epilogue(A[M,K] (bf16, row pitch lda) x W[N,K]^T (bf16, row pitch ldw)).
 * Replaces every nn.Linear on the path: to_q/to_k/to_v (model/modules.py:452-454, fused to one N=3D GEMM with the
 * rotary embedding of :470-480 in the epilogue), to_out + masked_fill + gated residual (:495-501, :635),
 * FeedForward (:348-353; GELU-tanh :625; gated residual :639), InputEmbedding.proj (model/backbones/dit.py:95),
 * AdaLayerNorm linears (:311, :332), TimestepEmbedding MLP (:727-731), proj_out (dit.py:231), ConvNeXtV2 / Vocos
 * point-wise linears (:263-267). */
int f5b_gemm(const void* A, int lda, const void* W, int ldw, const F5bGemmArgs* args, f5b_stream_t stream);

/* Transposed-operand GEMM (backward of nn.Linear; groundwork of the training step, trainer.py:1280):
 *   C[m,n] (+)= sum_k A(m,k) B(n,k),  A(m,k) = A[m*lda + k] (a_mn_major 0) or A[k*lda + m] (a_mn_major 1), same for B.
 *   dgrad dX = dY W:      f5b_gemm_tn(dY, N, 0, W, K, 1, dX, ...)       wgrad dW = dY^T X: f5b_gemm_tn(dY, N, 1, X, K, 1, dW, ..., splits)
 * out: bf16 [M, ldc] overwritten (out_f32_accumulate 0) or f32 [M, ldc] accumulated with TMA reduce-add (1; required for
 * splits > 1, which cuts the reduction into `splits` ranges processed by different CTAs).  A, B bf16. */
int f5b_gemm_tn(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* out, int ldc,
                int out_f32_accumulate, int M, int N, int K, int splits, f5b_stream_t stream);

/* out bf16[rows,D] = LayerNorm(x f32[rows,D], eps, no affine) * (1 + scale[b,:]) + shift[b,:]
 * (AdaLayerNorm.forward model/modules.py:310-315, DiTBlock ff norm :637, AdaLayerNorm_Final :331-336).
 * scale/shift element (b, c) at ptr[(b % batch_mod)*mod_bstride + c] (batch_mod 0: b); NULL/NULL = plain LayerNorm. */
int f5b_ln_modulate(const float* x, const float* scale, const float* shift, int64_t mod_bstride, int batch_mod,
                    void* out_bf16, int rows, int rows_per_batch, int D, float eps, f5b_stream_t stream);
/* LayerNorm(D, affine w/b, eps) with f32 and/or bf16 outputs (either may be NULL; out_f32 may alias x)
 * (Vocos backbone.norm / final_layer_norm). */
int f5b_ln_affine(const float* x, const float* w, const float* b, float* out_f32, void* out_bf16, int rows, int D, float eps,
                  f5b_stream_t stream);

/* Non-causal softmax(QK^T/sqrt(64))V with a per-batch key length (AttnProcessor, model/modules.py:457-493, dropout_p = 0).
 * q, k, v: bf16 token-major matrices [B*n, ld] whose columns [h*64, h*64+64) belong to head h — typically the three column
 * sections of the fused QKV GEMM output (q = base, k = base + D, v = base + 2D, ld = 3D); heads are gathered by strided TMA
 * boxes, v is consumed as an MN-major tcgen05 operand.  out bf16 [B*n, H*64] token-major.
 * lens int32 [lens_mod] (kv length of batch b = lens[b % lens_mod]) or NULL (= n).  Query rows >= len are written as zeros
 * (the reference zeroes them after to_out, :499-501). */
int f5b_attn_fwd(const void* q, const void* k, const void* v, int ld, void* out, const int32_t* lens, int lens_mod, int B, int H,
                 int n, float scale, f5b_stream_t stream);
/* Training forward: as f5b_attn_fwd, and also writes lse f32 [B, H, n] = log2(sum_j exp2(s_ij * scale * log2 e)) per query row
 * (+inf for query rows >= len, whose output is zero and which therefore take no gradient). */
int f5b_attn_fwd_lse(const void* q, const void* k, const void* v, int ld, void* out, float* lse, const int32_t* lens, int lens_mod,
                     int B, int H, int n, float scale, f5b_stream_t stream);
/* Attention backward (autograd of AttnProcessor's SDPA + RoPE under CFM.forward, model/cfm.py:210-283 / model/modules.py:470-493;
 * dropout_p = 0).  q, k, v, ld as in the forward (post-RoPE values); out / dout bf16 [B*n, ld_o] (forward output, its gradient);
 * lse from f5b_attn_fwd_lse; delta_ws f32 [B*H*n] and dq_ws f32 [B*n, H*64] are workspaces.  Writes dqkv bf16 [B*n, ld_d] =
 * (dq | dk | dv) at columns 0, H*64, 2*H*64, i.e. the gradient of the fused QKV projection's output: the transpose of the RoPE
 * rotation is applied to dq and dk of the first rope_heads heads (rope = f5b_rope_table).  Masked keys get zero gradient. */
int f5b_attn_bwd(const void* q, const void* k, const void* v, int ld, const void* out, const void* dout, int ld_o, const float* lse,
                 float* delta_ws, float* dq_ws, void* dqkv, int ld_d, const int32_t* lens, int lens_mod, int B, int H, int n,
                 float scale, const float* rope, int rope_heads, f5b_stream_t stream);

/* ConvPositionEmbedding conv layer (model/modules.py:171-176,183-185): grouped Conv1d(k, groups, pad k/2) + Mish.
 * x bf16 [B*n, D] token-major; wpk = weights packed by f5b_pack_convpos_weight; bias f32 [D].
 * mode 0: out_bf16[B*n, D] = mish(conv(x)+bias);  mode 1: resid_f32[B*n, D] += mish(conv(x)+bias);
 * mode 2: out_bf16 = conv(x)+bias without the activation (bias may be NULL) — the training forward keeps the pre-activation,
 * and the backward's input gradient is the same conv with weights packed by f5b_pack_convpos_weight_t. */
int f5b_convpos(const void* x_bf16, const void* wpk, const float* bias, void* out_bf16, float* resid_f32, int B, int n,
                int D, int groups, int ksize, int mode, f5b_stream_t stream);
/* w f32 [D, D/groups, ksize] (nn.Conv1d layout) -> bf16 [groups, ksize, NP, 64], NP = roundup(D/groups, 16). */
int f5b_pack_convpos_weight(const float* w, void* wpk, int D, int groups, int ksize, f5b_stream_t stream);
/* same packing of the TRANSPOSED conv (in/out channels of each group swapped, taps reversed): conv(dy, W^T) = d loss / d x */
int f5b_pack_convpos_weight_t(const float* w, void* wpk, int D, int groups, int ksize, f5b_stream_t stream);
size_t f5b_convpos_packed_elems(int D, int groups, int ksize);

/* Depth-wise Conv1d(k=7, pad 3, groups=C) + bias + LayerNorm(C, eps, affine) -> bf16
 * (ConvNeXtV2Block model/modules.py:259-262; Vocos ConvNeXtBlock).  x f32 [B*n, C] token-major. */
int f5b_dwconv7_ln(const float* x, const float* w /*[C,7]*/, const float* b, const float* ln_w, const float* ln_b,
                   void* out_bf16, int B, int n, int C, float eps, f5b_stream_t stream);

/* GRN (model/modules.py:225-234) over h bf16 [B*n, C]: Gx = ||h||_2 over the n positions of each batch row,
 * Nx = Gx / (mean_c Gx + 1e-6), out = gamma*(h*Nx) + beta + h -> bf16.  ws f32 [B*C] scratch. */
int f5b_grn(const void* h_bf16, const float* gamma, const float* beta, void* out_bf16, float* ws, int B, int n, int C,
            f5b_stream_t stream);

/* TextEmbedding front (model/backbones/dit.py:49-72): ids int64 [B, nt] (-1 padded) -> +1, truncate / pad with 0 to n,
 * drop_text -> all 0, embedding lookup (table f32 [V+1, C]) + freqs_cis[pos] (pos f32 [4096, C]) -> f32 [B*n, C].
 * mask_out (optional, uint8 [B*n]) = (token == 0) before drop_text, for text_mask_padding.
 * vocab_rows = rows of `table`: a shifted id outside [0, vocab_rows) traps the kernel (-> CUDA error on the host), like the
 * device-side assert of the reference's nn.Embedding; it is never read out of bounds. */
int f5b_text_lookup(const int64_t* ids, int nt, const float* table, const float* pos, float* out, uint8_t* mask_out,
                    int B, int n, int C, int vocab_rows, int drop_text, int add_pos, f5b_stream_t stream);
/* rows with mask != 0 are set to 0 (masked_fill, dit.py:74-75) */
int f5b_mask_rows_f32(float* x, const uint8_t* mask, int rows, int C, f5b_stream_t stream);

/* SinusPositionEmbedding(256) (model/modules.py:149-161): t f32 [M] -> bf16 [M,256] = cat(sin, cos)(1000 t f_k). */
int f5b_time_sinus(const float* t, void* out_bf16, int M, f5b_stream_t stream);
/* silu(x f32 [n]) -> bf16 */
int f5b_silu_bf16(const float* x, void* out_bf16, int64_t n, f5b_stream_t stream);
/* f32 [rows, cols] (pitch ld_in) -> bf16 columns [0, width) of a [rows, ld_out] matrix; columns >= cols are zero-filled
 * (x may be NULL when cols == 0). */
int f5b_pack_bf16(const float* x, int ld_in, void* out_bf16, int ld_out, int rows, int cols, int width, f5b_stream_t stream);

/* CFG combine + Euler update (fn closure model/cfm.py:159-173 + torchdiffeq euler):
 * v = pc + (pc - pu) * cfg;  y += dt * v;  y_bf16[rows, ld_bf] = bf16(y) (zero padded).  All f32 [rows, C].
 * pu may be NULL (cfg_strength < 1e-5 -> v = pc).  vel_out (optional) receives v. */
int f5b_cfg_euler(float* y, const float* pc, const float* pu, float cfg, float dt, void* y_bf16, int ld_bf, float* vel_out,
                  int rows, int C, f5b_stream_t stream);

/* same, with (cfg, dt) read from device memory params_dev[0..1]: lets one captured CUDA graph serve every ODE step */
int f5b_cfg_euler_dev(float* y, const float* pc, const float* pu, const float* params_dev, void* y_bf16, int ld_bf, float* vel_out,
                      int rows, int C, f5b_stream_t stream);

/* Flow-matching training inputs (CFM.forward, model/cfm.py:255-266): phi = (1-t) x0 + t x1, flow = x1 - x0,
 * cond = where(span_mask, 0, x1).  x1, x0 f32 [B, n, C]; time f32 [B]; span_mask uint8 [B*n]. */
int f5b_fm_prepare(const float* x1, const float* x0, const float* time, const uint8_t* span_mask, float* phi, float* flow, float* cond,
                   int B, int n, int C, f5b_stream_t stream);
/* Flow-matching loss (cfm.py:280-283): out2[0] = mean over rows with mask != 0 of (pred - flow)^2, out2[1] = element count.
 * Deterministic two-pass reduction; ws f32 [2048]. */
int f5b_masked_mse(const float* pred, const float* flow, const uint8_t* mask, float* ws, float* out2, int rows, int C,
                   f5b_stream_t stream);

/* MelSpec "vocos" (model/modules.py:83-101): wav f32 [B, L] -> log-mel f32 [B, T, n_mels] (token-major, T = 1 + L/256):
 * reflect-pad 512, periodic Hann(1024), |rFFT1024|, fb f32 [513, n_mels], log(clamp 1e-5). */
int f5b_melspec(const float* wav, const float* fb, const int32_t* ranges /* [n_mels,2] nonzero rows [f0,f1) of each fb column */,
                float* out, int B, int L, int n_mels, f5b_stream_t stream);

/* Vocos ISTFTHead tail: head f32 [B*T, 1026] (mag logits | phase) -> exp/clip 1e2, cos/sin, irFFT1024, Hann window,
 * overlap-add (hop 256), divide by the window envelope, trim 512 each side -> wav f32 [B, 256*(T-1)].
 * frames_ws f32 [B*T, 1024] scratch. */
int f5b_istft_head(const float* head, int ld, float* frames_ws, float* wav, int B, int T, f5b_stream_t stream);

/* im2col for the Vocos embed Conv1d(n_mels -> C, k=7, pad 3): mel f32 [B, T, n_mels] -> bf16 [B*T, ld] (7*n_mels <= ld) */
int f5b_im2col7(const float* mel, void* out_bf16, int B, int T, int n_mels, int ld, f5b_stream_t stream);

/* ---- model-level drivers ----------------------------------------------------------------------------------- */

/* All pointers are device pointers owned by the caller and must outlive the handle.
 * "stack" pointers hold one tensor per DiT block, contiguous: [depth, ...]. */
/* ---- tf32 operand mode (F5bGemmArgs.tf32 / F5bDitDesc.precision 1): the fp32-tolerance path ------------------------------
 * Tensor-core operands are fp32 words rounded to tf32 (cvt.rna, 10-bit mantissa), accumulation and all element-wise math fp32.
 * It replaces the same reference call sites as the bf16 entry points named in each comment; it exists because bf16 operands
 * cannot hold the 1e-3 relative velocity tolerance against the fp32 reference (measured 3.4e-3; tf32 4e-4). */
/* f5b_ln_modulate with an fp32 (tf32-rounded) output */
int f5b_ln_modulate_tf32(const float* x, const float* scale, const float* shift, int64_t mod_bstride, int batch_mod, float* out,
                         int rows, int rows_per_batch, int D, float eps, f5b_stream_t stream);
/* f5b_attn_fwd with q, k, v, out fp32 (tf32-rounded values in, tf32-rounded out); ld in elements.  kind::tf32 takes K-major
 * operands only, so V is transposed into vt_ws (f5b_attn_tf32_ws_floats(B, H, n) floats) first. */
int f5b_attn_fwd_tf32(const float* q, const float* k, const float* v, int ld, float* out, float* vt_ws, const int32_t* lens,
                      int lens_mod, int B, int H, int n, float scale, f5b_stream_t stream);
size_t f5b_attn_tf32_ws_floats(int B, int H, int n);
/* f5b_convpos with x fp32 [B*n, D] (tf32-rounded), wpk from f5b_pack_convpos_weight_tf32; mode 0: out fp32 = tf32(mish(conv + b)),
 * mode 1: resid += mish(conv + b) */
int f5b_convpos_tf32(const float* x, const float* wpk, const float* bias, float* out, float* resid, int B, int n, int D, int groups,
                     int ksize, int mode, f5b_stream_t stream);
int f5b_pack_convpos_weight_tf32(const float* w, float* wpk, int D, int groups, int ksize, f5b_stream_t stream);
size_t f5b_convpos_packed_elems_tf32(int D, int groups, int ksize);
/* f5b_time_sinus / f5b_pack_bf16 with fp32 (tf32-rounded) outputs */
int f5b_time_sinus_tf32(const float* t, float* out, int M, f5b_stream_t stream);
int f5b_pack_tf32(const float* x, int ld_in, float* out, int ld_out, int rows, int cols, int width, f5b_stream_t stream);

typedef struct F5bDitDesc {
  int32_t dim, depth, heads, dim_head, ff_mult, mel_dim, text_dim, conv_layers, rope_heads, text_mask_padding;
  int32_t convpos_kernel, convpos_groups, vocab_rows;
  /* operand mode of the inference drivers (f5b_dit_modulation / text_embed / input_const / forward):
   *   0  bf16 tensor-core operands (kind::f16), the throughput mode; every "bf16" weight below is bf16.
   *   1  tf32 operands (kind::tf32): every "bf16" weight below is instead an fp32 array of the same shape whose values were
   *      rounded to tf32 (round-to-nearest), cp_w1 / cp_w2 come from f5b_pack_convpos_weight_tf32, the activations between the
   *      kernels are fp32, and f5b_dit_forward's `x` is fp32 [Bx*n, 128].  Holds 1e-3 of the fp32 reference. */
  int32_t precision;
  /* TimestepEmbedding (model/modules.py:721-731) */
  const void* time_w0; const float* time_b0;   /* bf16 [D,256] */
  const void* time_w2; const float* time_b2;   /* bf16 [D,D]   */
  /* all AdaLayerNorm linears stacked: rows [i*6D, (i+1)*6D) = block i attn_norm.linear, last 2D rows = norm_out.linear */
  const void* mod_w; const float* mod_b;       /* bf16 [depth*6D + 2D, D] */
  /* TextEmbedding (model/backbones/dit.py:32-79) */
  const float* text_table;                     /* f32 [vocab_rows, text_dim] */
  const float* text_pos;                       /* f32 [4096, text_dim] */
  const float* tb_dw_w; const float* tb_dw_b;  /* f32 [L, text_dim, 7], [L, text_dim] */
  const float* tb_ln_w; const float* tb_ln_b;  /* f32 [L, text_dim] */
  const void* tb_pw1_w; const float* tb_pw1_b; /* bf16 [L, 2T, T], f32 [L, 2T] */
  const float* tb_grn_g; const float* tb_grn_b;/* f32 [L, 2T] */
  const void* tb_pw2_w; const float* tb_pw2_b; /* bf16 [L, T, 2T], f32 [L, T] */
  /* InputEmbedding (dit.py:85-97): proj weight split by source, K padded */
  const void* in_wx;                           /* bf16 [D, 128]  (x part, mel_dim cols used) */
  const void* in_wct;                          /* bf16 [D, 128 + T] (cond part padded to 128 | text part) */
  const float* in_b;                           /* f32 [D] */
  const void* cp_w1; const float* cp_b1;       /* packed grouped-conv weights (f5b_pack_convpos_weight) */
  const void* cp_w2; const float* cp_b2;
  /* DiT blocks */
  const void* qkv_w; const float* qkv_b;       /* bf16 [depth, 3D, D], f32 [depth, 3D] (q | k | v) */
  const void* out_w; const float* out_b;       /* bf16 [depth, D, D] */
  const void* ff1_w; const float* ff1_b;       /* bf16 [depth, F, D], F = ff_mult*D */
  const void* ff2_w; const float* ff2_b;       /* bf16 [depth, D, F] */
  const void* proj_w; const float* proj_b;     /* bf16 [mel_dim, D], f32 [mel_dim] */
} F5bDitDesc;

typedef struct F5bDit F5bDit;

int f5b_dit_create(const F5bDitDesc* desc, F5bDit** out);
void f5b_dit_destroy(F5bDit* h);
/* bytes of f32/bf16 scratch needed for a forward over `rows` = B*n token rows */
size_t f5b_dit_workspace_bytes(const F5bDit* h, int B, int n);

/* time MLP + every AdaLN modulation for M time values at once (t identical for all batch rows in CFM.sample, so the
 * whole schedule is one GEMM): t f32 [M] -> mod f32 [M, depth*6D + 2D].  ws: f5b_dit_modulation_ws_bytes. */
int f5b_dit_modulation(const F5bDit* h, const float* t, int M, float* mod, void* ws, f5b_stream_t stream);
size_t f5b_dit_modulation_ws_bytes(const F5bDit* h, int M);

/* TextEmbedding.forward -> f32 [B*n, text_dim].  ws: f5b_dit_text_ws_bytes. */
int f5b_dit_text_embed(const F5bDit* h, const int64_t* ids, int nt, int B, int n, int drop_text, float* out, void* ws,
                       f5b_stream_t stream);
size_t f5b_dit_text_ws_bytes(const F5bDit* h, int B, int n);

/* Step-invariant part of InputEmbedding.proj: c0 f32 [B*n, D] = [cond | text_embed] W_ct^T + bias
 * (cond f32 [B*n, mel_dim] or NULL for drop_audio_cond).  ws: B*n*(128+T)*2 bytes. */
int f5b_dit_input_const(const F5bDit* h, const float* cond, const float* text_embed, int B, int n, float* c0, void* ws,
                        f5b_stream_t stream);

/* One DiT.forward (dit.py:185-233) over Bf fused batch rows (e.g. cond rows then uncond rows of a CFG pair).
 *   x_bf16   bf16 [Bx*n, 128]   ODE state (zero padded); fused row b reads state row b % Bx
 *   c0       f32  [Bf*n, D]     f5b_dit_input_const output for each fused row
 *   mod      f32  [rows, depth*6D+2D] modulation; fused row b uses mod + (b % Bx) * mod_bstride (0 = shared)
 *   lens     int32 [Bx] or NULL key-padding lengths (mask), shared by the fused halves
 *   rope     f32  [n, 32, 2]
 *   pred     f32  [Bf*n, mel_dim]
 */
int f5b_dit_forward(const F5bDit* h, const void* x_bf16, int Bx, const float* c0, int Bf, int n, const float* mod,
                    int64_t mod_bstride, const int32_t* lens, const float* rope, float* pred, void* ws, size_t ws_bytes,
                    f5b_stream_t stream);

/* The reference's attention dropout at INFERENCE: /root/reference/src/f5_tts/model/modules.py:490 passes dropout_p = 0.1 to
 * F.scaled_dot_product_attention unconditionally, so the reference's own sampling is stochastic in attention even under eval()
 * (SURVEY.md 9.1).  Parity is defined at p = 0 (the default here); this switch reproduces the reference's behaviour statistically
 * (same distribution, the product's own Philox stream — see f5b_train_set_attn_dropout) for the following f5b_dit_forward calls,
 * bf16 operand mode only.  `seed` selects the stream; `seed_dev` (device uint32, may be NULL) is read by the kernels at run time and
 * mixed into it, so that a captured CUDA graph replays with a fresh mask whenever the word changes.  p = 0 switches it off. */
int f5b_dit_set_attn_dropout(float p, uint64_t seed, const uint32_t* seed_dev);

/* ---- training step: DiT.forward with saved activations + hand-written backward (CFM.forward / accelerator.backward,
 * model/cfm.py:210-283, model/trainer.py:1271-1287; dropout 0) ------------------------------------------------------------------
 * F5bDitGrads: f32 gradient buffers, one per F5bDitDesc tensor and of the same shape, except that the conv_pos_embed weight
 * gradients use nn.Conv1d's [D, D/groups, k] layout and the bf16 weights get f32 gradients.  NULL members are skipped. */
typedef struct F5bDitGrads {
  float *time_w0, *time_b0, *time_w2, *time_b2, *mod_w, *mod_b;
  float *text_table, *tb_dw_w, *tb_dw_b, *tb_ln_w, *tb_ln_b, *tb_pw1_w, *tb_pw1_b, *tb_grn_g, *tb_grn_b, *tb_pw2_w, *tb_pw2_b;
  float *in_wx, *in_wct, *in_b, *cp_w1, *cp_b1, *cp_w2, *cp_b2;
  float *qkv_w, *qkv_b, *out_w, *out_b, *ff1_w, *ff1_b, *ff2_w, *ff2_b, *proj_w, *proj_b;
} F5bDitGrads;
size_t f5b_dit_train_ws_bytes(const F5bDit* h, int B, int n);
/* x f32 [B*n, mel] (phi_t), cond f32 [B*n, mel] or NULL (drop_audio_cond), text_embed f32 [B*n, T], time f32 [B] (per sample),
 * lens int32 [B] or NULL, rope f32 [n, 32, 2] -> pred f32 [B*n, mel]; every activation the backward needs stays in ws. */
int f5b_dit_train_forward(const F5bDit* h, const float* x, const float* cond, const float* text_embed, const float* time, int B, int n,
                          const int32_t* lens, const float* rope, float* pred, void* ws, size_t ws_bytes, f5b_stream_t stream);
/* dpred bf16 [B*n, 128] (f5b_mse_grad).  Gradients are ACCUMULATED into g (zero the buffers first).  cp_w1_t / cp_w2_t: the
 * conv_pos_embed weights packed by f5b_pack_convpos_weight_t.  dtext_bf16 (optional): d loss / d text_embed, bf16 [B*n, T].
 * ws must be the workspace the forward ran in. */
int f5b_dit_train_backward(const F5bDit* h, const void* dpred_bf16, const void* cp_w1_t, const void* cp_w2_t, const F5bDitGrads* g,
                           void* dtext_bf16, int B, int n, const int32_t* lens, const float* rope, void* ws, size_t ws_bytes,
                           f5b_stream_t stream);

/* Train-mode dropout of the DiT blocks (reference: DiT(dropout=0.1), model/backbones/dit.py:132; FeedForward's Dropout after GELU,
 * model/modules.py:342-353, and the Dropout behind attention's to_out, :436-440).  p = 0 (the default) switches it off.  The keep
 * decision is a counter-based hash of (seed, block, site, element) -- the backward regenerates the forward's mask -- so call this
 * once per micro-step with a fresh seed BEFORE f5b_dit_train_forward and leave it untouched until the matching backward returns.
 * The dropout inside scaled_dot_product_attention (:490) has its own switch: f5b_train_set_attn_dropout. */
int f5b_train_set_dropout(float p, uint64_t seed);
/* the dropout inside F.scaled_dot_product_attention (model/modules.py:490, dropout_p = 0.1 in the fork): applied to the normalised
 * attention probabilities of every block by the f5b_dit_train_* drivers (forward and backward regenerate the same mask from the seed
 * of f5b_train_set_dropout); p = 0 (default) switches it off.  Inference never applies it.  The mask stream is the product's own
 * (one Philox-2x32-7 block per 8 consecutive keys of a query row, p quantised to 1/128 with the kept values scaled by the exact
 * reciprocal of the quantised keep probability; eraxvif5tts_b200/csrc/dropout.cuh, restated in oracle.attention_dropout_multipliers):
 * torch's own Philox offsets are an implementation detail, only the distribution is the reference's. */
int f5b_train_set_attn_dropout(float p);
/* Activation checkpointing of the DiT blocks (the reference's `checkpoint_activations`, /root/reference/src/f5_tts/model/backbones/
 * dit.py:121,158,221-223): on, every block keeps only its fp32 input and the backward re-runs the block's forward before
 * differentiating it (workspace 4.6 GB instead of 29.6 GB at 32 x 1200 frames, F5TTS_Base; one extra forward per step).  Set it
 * before f5b_dit_train_ws_bytes / f5b_dit_train_forward and keep it until the matching backward. */
int f5b_train_set_checkpoint(int on);

/* The same backward in pieces (parts bit 0 = head: proj_out + final AdaLN; bit 1 = blocks [blk_lo, blk_hi) in descending order;
 * bit 2 = tail: input embedding + modulation / time MLP), issued head -> blocks from depth down to 0 -> tail, so the host can start
 * the gradient all-reduce of finished blocks (DDP's bucketed overlap, trainer.py:1280) while earlier blocks are differentiated. */
int f5b_dit_train_backward_part(const F5bDit* h, const void* dpred_bf16, const void* cp_w1_t, const void* cp_w2_t, const F5bDitGrads* g,
                                void* dtext_bf16, int B, int n, const int32_t* lens, const float* rope, void* ws, size_t ws_bytes,
                                int parts, int blk_lo, int blk_hi, f5b_stream_t stream);

/* TextEmbedding.forward in training form (every ConvNeXtV2Block input kept in ws) and its backward: dtext_bf16 [B*n, T] is the
 * gradient f5b_dit_train_backward returns; fills g->text_table and g->tb_*.  text_mask_padding is not supported here. */
size_t f5b_dit_text_train_ws_bytes(const F5bDit* h, int B, int n);
int f5b_dit_text_embed_train(const F5bDit* h, const int64_t* ids, int nt, int B, int n, int drop_text, float* out, void* ws,
                             size_t ws_bytes, f5b_stream_t stream);
int f5b_dit_text_embed_backward(const F5bDit* h, const int64_t* ids, int nt, int B, int n, int drop_text, const void* dtext_bf16,
                                const F5bDitGrads* g, void* ws, size_t ws_bytes, f5b_stream_t stream);
/* building blocks of the above */
int f5b_ln_affine_bwd(const void* dy_bf16, const float* x, const float* w, float* dx, int accumulate, float* dw, float* db, int B,
                      int n, int D, float eps, f5b_stream_t stream);
/* GRN (model/modules.py:225-234) + the GELU(erf) in front of it, backward: dp1 = d(t3)/d(p1) applied to dt3 (may alias), with
 * dgamma / dbeta / dbias1 accumulated; stats_ws f32 [B, 3, C] */
int f5b_grn_gelu_bwd(const void* dt3_bf16, const void* t2_bf16, const void* p1_bf16, const float* gamma, void* dp1_bf16, float* dgamma,
                     float* dbeta, float* dbias1, float* stats_ws, int B, int n, int C, f5b_stream_t stream);
int f5b_dwconv7_bwd(const float* dy, const float* x, const float* w, float* dx_accum, float* dw, float* db, int B, int n, int C,
                    f5b_stream_t stream);
/* dtable f32 [vocab_rows, C]; out-of-range ids trap (see f5b_text_lookup) instead of corrupting neighbouring gradients */
int f5b_text_lookup_bwd(const float* dh, const int64_t* ids, int nt, float* dtable, int B, int n, int C, int vocab_rows, int drop_text,
                        f5b_stream_t stream);

/* rope table for n positions, dim_head 64: f32 [n, 32, 2] = (cos, sin)(pos * 10000^(-2j/64))
 * (x_transformers RotaryEmbedding.forward_from_seq_len, call site dit.py:215) */
int f5b_rope_table(float* out, int n, f5b_stream_t stream);

/* Vocos (third-party `vocos`, charactr/vocos-mel-24khz; call sites infer/f5tts_wrapper.py:524, infer/utils_infer.py:488) */
typedef struct F5bVocosDesc {
  int32_t n_mels, dim, intermediate, num_layers, n_fft, hop;
  const void* embed_w; const float* embed_b;     /* bf16 [dim, ld_embed] im2col layout (k-major: col = k*n_mels + c) */
  int32_t ld_embed;
  const float* norm_w; const float* norm_b;      /* f32 [dim] */
  const float* dw_w; const float* dw_b;          /* f32 [L, dim, 7], [L, dim] */
  const float* ln_w; const float* ln_b;          /* f32 [L, dim] */
  const void* pw1_w; const float* pw1_b;         /* bf16 [L, I, dim], f32 [L, I] */
  const void* pw2_w; const float* pw2_b;         /* bf16 [L, dim, I], f32 [L, dim] */
  const float* gamma;                            /* f32 [L, dim] */
  const float* fln_w; const float* fln_b;        /* f32 [dim] */
  const void* head_w; const float* head_b;       /* bf16 [n_fft+2, dim], f32 [n_fft+2] */
} F5bVocosDesc;
typedef struct F5bVocos F5bVocos;
int f5b_vocos_create(const F5bVocosDesc* desc, F5bVocos** out);
void f5b_vocos_destroy(F5bVocos* h);
size_t f5b_vocos_workspace_bytes(const F5bVocos* h, int B, int T);
/* mel f32 [B, T, n_mels] (token-major) -> wav f32 [B, hop*(T-1)] */
int f5b_vocos_decode(const F5bVocos* h, const float* mel, int B, int T, float* wav, void* ws, size_t ws_bytes,
                     f5b_stream_t stream);

/* ---- backward-pass building blocks (autograd of CFM.forward -> DiT, model/cfm.py:210-283; DiTBlock model/modules.py:627-641) ----
 * Un-fused training form of the gated residual: out = x + gate[b] * z (rows >= lens[b] keep x, :499-501); z bf16 [B*n, C]. */
int f5b_gate_add(const float* x, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens, float* out, int B,
                 int n, int C, f5b_stream_t stream);
/* the same fused with the LayerNorm-modulate that consumes x_out (one read of the residual stream instead of two):
 * x_out = x + gate[b] * z;  out_bf16 = LN(x_out) * (1 + scale[b]) + shift[b].  D <= 1024. */
int f5b_gate_add_ln_modulate(const float* x, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens,
                             float* x_out, const float* scale, const float* shift, int64_t mod_bstride, void* out_bf16, int B, int n,
                             int D, float eps, f5b_stream_t stream);
/* ... and its backward: dz = bf16(gate[b] * dx) (0 on rows >= lens[b]); dgate[b] += sum_r dx * z; dbias += sum_r dz (NULL = skip) */
int f5b_gate_bwd(const float* dx, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens, void* dz_bf16,
                 float* dgate, float* dbias, int B, int n, int C, f5b_stream_t stream);
/* out = act(h) on `count` bf16 elements;  dh = bf16(du * act'(h)) over [rows, C] (pitch ld) with dbias[C] += column sums of dh.
 * h NULL: dh = du (dh may be NULL too: plain column sum = bias gradient of a Linear). */
int f5b_act_fwd(const void* h_bf16, void* out_bf16, int64_t count, int act, f5b_stream_t stream);
int f5b_act_bwd(const void* du_bf16, const void* h_bf16, void* dh_bf16, float* dbias, int64_t rows, int C, int ld, int act,
                f5b_stream_t stream);
/* Backward of f5b_ln_modulate: dx (+)= LN-backward(dy * (1 + scale[b])), dscale[b] += sum_r dy * xhat, dshift[b] += sum_r dy.
 * x is the saved fp32 input (statistics are recomputed); scale NULL = plain LayerNorm; D <= 1024. */
int f5b_ln_modulate_bwd(const void* dy_bf16, const float* x, const float* scale, int64_t mod_bstride, float* dx, int accumulate,
                        float* dscale, float* dshift, int B, int n, int D, float eps, f5b_stream_t stream);
/* d loss / d pred of f5b_masked_mse: 2 (pred - flow) / loss2[1] on masked rows -> bf16 [rows, ld] (columns >= C zero) */
int f5b_mse_grad(const float* pred, const float* flow, const uint8_t* mask, const float* loss2, void* out_bf16, int64_t rows, int C,
                 int ld, f5b_stream_t stream);

/* Distillation losses, train/distil_reload.py:1066-1093 (teacher prediction T detached; m = rand_span_mask; cnt = max(sum m, 1)
 * FRAMES -- the channel axis is summed, unlike CFM.forward's mean): student = sum m (p - flow)^2 / cnt; distill = sum m (p - T)^2 / cnt
 * (l1 != 0: sum m |p - T| / cnt); spec_l1 = sum m |p - T| / cnt when spec_l1_weight > 0, else 0;
 * total = (1 - alpha) student + alpha distill + spec_l1_weight spec_l1.  out5 = (total, student, distill, spec_l1, cnt).
 * ws: 4 * 1024 floats.  f5b_distill_grad writes d total / d pred as bf16 [rows, ld] (columns >= C zero), the input of
 * f5b_dit_train_backward. */
int f5b_distill_loss(const float* pred, const float* flow, const float* teacher, const uint8_t* mask, float* ws, float* out5, int rows,
                     int C, int l1, float alpha, float spec_l1_weight, f5b_stream_t stream);
int f5b_distill_grad(const float* pred, const float* flow, const float* teacher, const uint8_t* mask, const float* out5, void* out_bf16,
                     int64_t rows, int C, int ld, int l1, float alpha, float spec_l1_weight, f5b_stream_t stream);

/* ---- optimizer step (trainer.py:1280-1287, 1321): clip_grad_norm_ + torch.optim.AdamW + EMA lerp, one fused pass ------------
 * f5b_grad_sumsq: out[0] = sum(g^2) over a flat f32 gradient buffer (deterministic two-pass; ws f32 [1024]).
 * f5b_adamw_ema_step over flat f32 buffers p, g, m, v (+ optional ema, + optional bf16 copy of the new p):
 *   g' = g * grad_scale * min(1, max_norm / (sqrt(grad_sumsq[0]) * grad_scale + 1e-6))      (grad_sumsq NULL or max_norm <= 0: no clip)
 *   p *= 1 - lr*wd;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps)
 *   ema += (1 - ema_decay) * (p - ema)   when ema != NULL and ema_decay >= 0. */
int f5b_grad_sumsq(const float* g, int64_t n, float* ws, float* out, f5b_stream_t stream);
int f5b_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, const float* grad_sumsq, float max_norm, float grad_scale,
                       float ema_decay, f5b_stream_t stream);

/* ---- launch accounting / per-kernel-class timing (used by bench.py for `gpu_launches` and the roofline leg) ------------
 * kernel classes: 0 gemm, 1 attention, 2 convpos, 3 norm (LN / dwconv+LN), 4 elementwise, 5 spectral.
 * f5b_prof_reset(enable): zero the counters; enable != 0 additionally brackets every launch with CUDA events on its stream.
 * f5b_prof_read(out, 6): out[k*4+{0,1,2,3}] = launches, device ms, algorithmic FLOPs, algorithmic bytes. Synchronises. */
#define F5B_NUM_KERNEL_KINDS 6
void f5b_prof_reset(int enable);
int f5b_prof_read(double* out, int n_kinds);
int f5b_prof_enabled(void);
/* add per-kind (launches, unused, flops, bytes) — used when a captured CUDA graph is replayed */
void f5b_prof_add(const double* delta, int n_kinds);

/* Programmatic dependent launch for the kernels of the DiT forward path (GEMM engine, attention, conv_pos_embed, LN + modulate,
 * CFG / Euler): with on != 0 they are launched with the programmatic-stream-serialization attribute and run their prologue (barrier
 * init, TMEM allocation, tensor-map prefetch) under the previous kernel's tail, waiting (griddepcontrol.wait) before they touch
 * global memory.  Measured on B200 (A/B on one box, twice each): B = 1 utterance 77.4 -> 75.3 ms; no measurable change for the
 * training step (105.3 / 106.4 ms off, 106.4 / 105.9 ms on); 22.4 k -> 22.0 k frames/s for the GPU-bound cfg-2 batch.  So the
 * Python host switches it on for fused batches of <= 16384 rows only.  Process-global; default off. */
int f5b_set_dependent_launch(int on);

/* ---- on-device chunk cross-fade + PCM packing ---- */
/* One fold step of the cross-fade of generated chunks (infer/f5tts_wrapper.py:556-572, infer/utils_infer.py:519-543): blends
 * acc[acc_len - cfs, acc_len) with next[0, cfs) using numpy's linspace(1, 0, cfs) / linspace(0, 1, cfs) in fp64 and appends
 * next[cfs, next_len); the accumulated length becomes acc_len - cfs + next_len (acc must have room).  cfs = 0: plain append. */
int f5b_crossfade_append(float* acc, int64_t acc_len, const float* next, int64_t next_len, int cfs, f5b_stream_t stream);
/* np.int16(x * 32767) (socket_server.py:54): fp32 product truncated toward zero, saturated to the int16 range. */
int f5b_pcm16(const float* x, int16_t* out, int64_t n, f5b_stream_t stream);

/* ---- monotonic alignment search + duration predictor (the duration side of train/distil_reload.py) ---- */

/* viterbi_vectorized_alignment, model/alignment_utils.py:154-212.  sim fp32 [B, nt, T] (token x frame similarity) -> align fp32
 * [B, nt, T] of 0 / 1 (bit-exact: the recurrence path[n,t] = sim[n,t] + max(path[n-1,t], path[n,t-1]) and the backtracking rule
 * "segment of token n starts at the last j < curr with path[n,j+1] - path[n,j] > 0" use the reference's fp32 operations).
 * path_ws: fp32 [B, nt, T] scratch that holds path_prob on return.  durations (optional): int32 [B, nt] = align.sum(-1)
 * (get_durations_from_alignment, :123-133). */
int f5b_align_viterbi(const float* sim, float* path_ws, float* align, int32_t* durations, int B, int nt, int T, f5b_stream_t stream);
/* windowed_monotonic_alignment, model/alignment_utils.py:214-257; window = max(2, int(T * window_size)) computed by the caller.
 * err int32 [B]: set to 1 for a batch item whose search window came out empty (the reference's torch.argmax raises there). */
int f5b_align_window(const float* sim, float* align, int32_t* durations, int32_t* err, int B, int nt, int T, int window,
                     f5b_stream_t stream);
/* DurationPredictor.forward / .phoneme_forward in eval mode, model/duration_predictor.py:27-66 (g = None): embedding of
 * ids + id_shift (1 for text tokens padded with -1, 0 for phoneme indices) -> * mask -> Conv1d(Cin, F, k) -> ReLU -> GroupNorm(1, F)
 * -> * mask -> Conv1d(F, F, k) -> ReLU -> GroupNorm(1, F) -> * mask -> Conv1d(F, 1, 1) -> * mask.  All fp32, nn.Conv1d weight
 * layouts [out, in, k]; ids int64 [B, nt]; mask fp32 [B, nt]; h1_ws fp32 [B, nt, F]; out fp32 [B, nt] (= the reference's [B, 1, nt]). */
int f5b_duration_predictor(const int64_t* ids, int id_shift, const float* mask, const float* table, int vocab_rows, const float* conv1_w,
                           const float* conv1_b, const float* norm1_w, const float* norm1_b, const float* conv2_w, const float* conv2_b,
                           const float* norm2_w, const float* norm2_b, const float* proj_w, const float* proj_b, float* h1_ws, float* out,
                           int B, int nt, int Cin, int F, int ksize, f5b_stream_t stream);

/* DurationPredictor in TRAIN mode (nn.Dropout(p_dropout) after each GroupNorm, duration_predictor.py:16, 36-41) and its backward,
 * for the duration loss of train/distil_reload.py:1096-1124 (logw vs log(attn.sum(2) + 1e-6)).  The same pointer struct carries the
 * parameters and (non-const in effect) their gradient buffers, which the backward ACCUMULATES into (zero them first).
 * The dropout masks are the counter-based generator of f5b_train_set_dropout (site 0 / 1 = first / second Dropout, element index in
 * [B, nt, F] order), regenerated by the backward from (p_dropout, seed).  h1 / a / c / dpre1 workspaces: fp32 [B, nt, F]; stats [B, 4].
 * dlogw fp32 [B, nt] = d loss / d out.  B * nt * F must be a multiple of 4; F <= 64. */
typedef struct F5bDurPredParams {
  const float *table, *conv1_w, *conv1_b, *norm1_w, *norm1_b, *conv2_w, *conv2_b, *norm2_w, *norm2_b, *proj_w, *proj_b;
} F5bDurPredParams;
int f5b_duration_predictor_train_forward(const int64_t* ids, int id_shift, const float* mask, const F5bDurPredParams* p, float p_dropout,
                                         uint64_t seed, float* h1_ws, float* a_ws, float* c_ws, float* stats_ws, float* out, int B, int nt,
                                         int Cin, int F, int ksize, f5b_stream_t stream);
int f5b_duration_predictor_backward(const float* dlogw, const int64_t* ids, int id_shift, const float* mask, const F5bDurPredParams* p,
                                    const F5bDurPredParams* grads, float p_dropout, uint64_t seed, const float* h1_ws, const float* a_ws,
                                    const float* c_ws, const float* stats_ws, float* dpre1_ws, int B, int nt, int Cin, int F, int ksize,
                                    f5b_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* F5B200_H_ */
