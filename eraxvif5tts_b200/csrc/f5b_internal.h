// Internal C++ interfaces between the translation units of libf5b200.so (the public C ABI is include/f5b200.h).
#pragma once
#include <cuda_runtime.h>

#include "../../include/f5b200.h"

namespace f5b {

int gemm(const void* A, int lda, const void* W, int ldw, const F5bGemmArgs& g, cudaStream_t stream);

// convenience wrappers used by the model drivers
int linear_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M, int N, int K,
                int act, cudaStream_t s, bool tf32 = false);
// out = bf16(A W^T + bias) and out_act = bf16(act(out)) from one accumulator tile (F5B_EPI_BF16_DUAL; act = GELU_TANH)
int linear_bf16_dual(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, void* out_act, int ldc_act,
                     int M, int N, int K, int act, cudaStream_t s);
int linear_f32(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldc, int M, int N, int K,
               int act, const float* addsrc, int ld_add, void* out_bf16, int ld_bf16, cudaStream_t s, bool tf32 = false);
int linear_gate_resid(const void* A, int lda, const void* W, int ldw, const float* bias, float* x, int ldc, int M, int N,
                      int K, int rows_per_batch, const float* gate, int64_t gate_bstride, const int32_t* lens,
                      int batch_mod, cudaStream_t s, bool tf32 = false);

// (tf32 = true in the sweeps below: the activation output is fp32 rounded to tf32 instead of bf16 — the tf32 operand mode)
int ln_modulate(const float* x, const float* scale, const float* shift, int64_t mod_bstride, int batch_mod, void* out_bf16,
                int rows, int rows_per_batch, int D, float eps, cudaStream_t s, bool tf32 = false);
int ln_affine(const float* x, const float* w, const float* b, float* out_f32, void* out_bf16, int rows, int D, float eps,
              cudaStream_t s);
int dwconv7_ln(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b, void* out_bf16, int B, int n,
               int C, float eps, cudaStream_t s, float* y_out = nullptr, bool tf32 = false);
int grn(const void* h_bf16, const float* gamma, const float* beta, void* out_bf16, float* ws, int B, int n, int C, cudaStream_t s,
        bool tf32 = false);
// gemm_grad.cu: weight gradient of one 64-channel group of the grouped position-embedding conv, im2col by TMA coordinates
int conv_wgrad_implicit(const void* dy_g, const void* x_g, int ld, float* out, int ldc, int B, int n, int cpg, int ks, int splits,
                        cudaStream_t stream);
bool train_dropout_on();  // train_kernels.cu: f5b_train_set_dropout's p > 0 (sites 0 / 1)
struct AttnDrop;  // dropout.cuh: mask stream of one layer's SDPA dropout; nullptr / addc == 0 = no dropout
int attn_fwd(const void* q, const void* k, const void* v, int ld, void* out, float* lse, const int32_t* lens, int lens_mod, int B, int H,
             int n, float scale, cudaStream_t stream, const AttnDrop* drop = nullptr, const uint32_t* drop_seed_dev = nullptr);
int attn_bwd(const void* q, const void* k, const void* v, int ld, const void* out, const void* dout, int ld_o, const float* lse,
             float* delta, float* dq_ws, void* dqkv, int ld_d, const int32_t* lens, int lens_mod, int B, int H, int n, float scale,
             const float* rope, int rope_heads, cudaStream_t stream, const AttnDrop* drop = nullptr);
int attn_fwd_tf32(const float* q, const float* k, const float* v, int ld, float* out, float* vt_ws, const int32_t* lens, int lens_mod,
                  int B, int H, int n, float scale, cudaStream_t stream);
size_t attn_tf32_ws_floats(int B, int H, int n);
int convpos_tf32(const float* x, const float* wpk, const float* bias, float* out, float* resid, int B, int n, int D, int groups, int ksize,
                 int mode, cudaStream_t stream);
int convpos(const void* x, const void* wpk, const float* bias, void* out, float* resid, int B, int n, int D, int groups,
            int ksize, int mode, cudaStream_t stream);

}  // namespace f5b

// The training drivers' dropout-aware forms of four sweeps (train_kernels.cu).  site 0 = FeedForward's Dropout after GELU
// (model/modules.py:342-353), site 1 = the Dropout behind attention's to_out (:436-440); the mask is a function of
// (f5b_train_set_dropout's seed, layer, site, element).  With p = 0 they are the public f5b_* sweeps.
extern "C" {
int f5b_act_fwd_site(const void* h_bf16, void* out_bf16, int64_t count, int act, int layer, int site, f5b_stream_t stream);
int f5b_act_bwd_site(const void* du_bf16, const void* h_bf16, void* dh_bf16, float* dbias, int64_t rows, int C, int ld, int act, int layer,
                     int site, f5b_stream_t stream);
int f5b_gate_add_ln_modulate_site(const float* x, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens,
                                  float* x_out, const float* scale, const float* shift, int64_t mod_bstride, void* out_bf16, int B,
                                  int n, int D, float eps, int layer, int site, f5b_stream_t stream);
int f5b_gate_bwd_site(const float* dx, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens, void* dz_bf16,
                      float* dgate, float* dbias, int B, int n, int C, int layer, int site, f5b_stream_t stream);
}
