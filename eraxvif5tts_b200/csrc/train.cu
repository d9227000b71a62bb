// Training-step drivers: DiT.forward in training form (activations kept) and its hand-written backward — the autograd graph the
// reference builds under CFM.forward / accelerator.backward (/root/reference/src/f5_tts/model/cfm.py:210-283,
// /root/reference/src/f5_tts/model/trainer.py:1271-1287) as explicit launch sequences.  Host-side only; every launch goes to the
// caller's stream and nothing is allocated.  Dropout is 0 (DESIGN.md "oracle adjustments").
#include "common.cuh"
#include "f5b_internal.h"
#include "dropout.cuh"

struct F5bDit {
  F5bDitDesc d;
};

namespace f5b {

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<uint8_t*>(p)) {}
  template <class T>
  T* take(size_t count) {
    T* r = reinterpret_cast<T*>(base + off);
    off = align_up(off + count * sizeof(T));
    return r;
  }
};

typedef __nv_bfloat16 bf;

struct BlockSave {
  float *x_in, *x_mid, *lse;
  bf *a, *qkv, *o, *z1, *f, *h1, *u, *z2;
};

constexpr int MAX_DEPTH = 64;
// Activation checkpointing (the reference's `checkpoint_activations`, model/backbones/dit.py:121,158,221-223: one
// torch.utils.checkpoint per DiT block).  Off: every block keeps its ten intermediate tensors (1.26 GB per block at 32 x 1200 frames,
// Base).  On: every block keeps only its fp32 input; all blocks share ONE set of intermediates, and the backward re-runs a block's
// forward (same kernels, same dropout masks: they are a pure function of the seed) right before differentiating it.
static int g_ckpt = 0;

struct TrainWs {
  // time / modulation chain (M = B rows)
  bf *sin_bf, *a1, *h1, *a2, *h2, *dmod_bf, *tb1, *tb2;
  float *mod, *dmod, *dh2;
  // input embedding
  bf *a_x, *a_ct, *hb0, *u1, *c1, *u2;
  float* h0;
  BlockSave blk[MAX_DEPTH];
  float* x_fin;
  bf* hbF;
  // backward scratch
  float *dx, *delta, *dq_ws;
  bf *t1, *t2, *t3, *tF, *xcol, *dact;
  float* cpw_tmp;
  size_t bytes;
};

static TrainWs carve_train(const F5bDitDesc& d, int B, int n, void* ws) {
  const size_t rows = (size_t)B * n;
  const int D = d.dim, F = d.ff_mult * d.dim, H = d.heads, T = d.text_dim;
  const size_t mod_dim = (size_t)d.depth * 6 * D + 2 * D;
  Carver c(ws);
  TrainWs w;
  w.sin_bf = c.take<bf>((size_t)B * 256);
  w.a1 = c.take<bf>((size_t)B * D);
  w.h1 = c.take<bf>((size_t)B * D);
  w.a2 = c.take<bf>((size_t)B * D);
  w.h2 = c.take<bf>((size_t)B * D);
  w.tb1 = c.take<bf>((size_t)B * D);
  w.tb2 = c.take<bf>((size_t)B * D);
  w.dh2 = c.take<float>((size_t)B * D);
  w.mod = c.take<float>(B * mod_dim);
  w.dmod = c.take<float>(B * mod_dim);
  w.dmod_bf = c.take<bf>(B * mod_dim);
  w.a_x = c.take<bf>(rows * 128);
  w.a_ct = c.take<bf>(rows * (128 + T));
  w.h0 = c.take<float>(rows * D);
  w.hb0 = c.take<bf>(rows * D);
  w.u1 = c.take<bf>(rows * D);
  w.c1 = c.take<bf>(rows * D);
  w.u2 = c.take<bf>(rows * D);
  float* x_next = c.take<float>(rows * D);
  for (int i = 0; i < d.depth && i < MAX_DEPTH; ++i) {
    BlockSave& s = w.blk[i];
    s.x_in = x_next;
    if (g_ckpt && i > 0) {  // checkpointing: the intermediates of every block alias block 0's
      const float* keep = s.x_in;
      s = w.blk[0];
      s.x_in = const_cast<float*>(keep);
    } else {
      s.x_mid = c.take<float>(rows * D);
      s.lse = c.take<float>((size_t)B * H * n);
      s.a = c.take<bf>(rows * D);
      s.qkv = c.take<bf>(rows * 3 * D);
      s.o = c.take<bf>(rows * D);
      s.z1 = c.take<bf>(rows * D);
      s.f = c.take<bf>(rows * D);
      s.h1 = c.take<bf>(rows * F);
      s.u = c.take<bf>(rows * F);
      s.z2 = c.take<bf>(rows * D);
    }
    x_next = c.take<float>(rows * D);
  }
  w.x_fin = x_next;
  w.hbF = c.take<bf>(rows * D);
  w.dx = c.take<float>(rows * D);
  w.delta = c.take<float>((size_t)B * H * n);
  w.dq_ws = c.take<float>(rows * D);
  w.t1 = c.take<bf>(rows * D);
  w.t2 = c.take<bf>(rows * D);
  w.t3 = c.take<bf>(rows * 3 * D);
  w.tF = c.take<bf>(rows * F);
  w.xcol = c.take<bf>(rows * (size_t)(D / d.convpos_groups) * d.convpos_kernel);
  w.dact = c.take<bf>(rows * (128 + T));
  w.cpw_tmp = c.take<float>((size_t)D * (D / d.convpos_groups) * d.convpos_kernel);
  w.bytes = c.off;
  return w;
}

// Xcol[r, k*cpg + ci] = x[b, t + k - pad, g*cpg + ci]  (zero outside the utterance; 16-byte chunks): the conv weight gradient of group
// g is then the plain product dY_g^T Xcol in [co][k][ci] order; conv_wgrad_fold_kernel adds it into nn.Conv1d's [co][ci][k] layout
__global__ void im2col_convpos_kernel(const bf* __restrict__ x, bf* __restrict__ out, int n, int D, int g, int cpg, int ks) {
  const size_t row = blockIdx.x;
  const int b = (int)(row / n), t = (int)(row - (size_t)b * n);
  const int c8n = cpg >> 3, pad = ks / 2;
  for (int j = threadIdx.x; j < ks * c8n; j += blockDim.x) {
    const int k = j / c8n, c8 = j - k * c8n;
    const int p = t + k - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p >= 0 && p < n) v = *reinterpret_cast<const uint4*>(x + ((size_t)b * n + p) * D + g * cpg + c8 * 8);
    *reinterpret_cast<uint4*>(out + row * (size_t)(cpg * ks) + k * cpg + c8 * 8) = v;
  }
}
__global__ void conv_wgrad_fold_kernel(const float* __restrict__ tmp, float* __restrict__ dw, int D, int cpg, int ks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // index into dw [D][cpg][ks]
  if (i >= D * cpg * ks) return;
  const int k = i % ks;
  const int t = i / ks;
  const int ci = t % cpg, co = t / cpg;
  dw[i] += tmp[((size_t)co * ks + k) * cpg + ci];
}

struct TextSave {
  float *h_in, *y;
  bf *tb, *p1, *t2, *t3;
};
struct TextWs {
  TextSave L[16];
  float *gx, *stats, *dh, *dy;
  bf *dz, *d2, *dtb;
  size_t bytes;
};
static TextWs carve_text(const F5bDitDesc& d, int B, int n, void* ws) {
  const size_t rows = (size_t)B * n;
  const int T = d.text_dim, T2 = 2 * d.text_dim;
  Carver c(ws);
  TextWs w;
  for (int j = 0; j < d.conv_layers && j < 16; ++j) {
    TextSave& s = w.L[j];
    s.h_in = c.take<float>(rows * T);
    s.y = c.take<float>(rows * T);
    s.tb = c.take<bf>(rows * T);
    s.p1 = c.take<bf>(rows * T2);
    s.t2 = c.take<bf>(rows * T2);
    s.t3 = c.take<bf>(rows * T2);
  }
  w.gx = c.take<float>((size_t)B * T2);
  w.stats = c.take<float>((size_t)B * 3 * T2);
  w.dh = c.take<float>(rows * T);
  w.dy = c.take<float>(rows * T);
  w.dz = c.take<bf>(rows * T);
  w.d2 = c.take<bf>(rows * T2);
  w.dtb = c.take<bf>(rows * T);
  w.bytes = c.off;
  return w;
}

int g_conv_wgrad_implicit = [] { const char* e = getenv("F5B_CONV_WGRAD_IM2COL"); return (e && atoi(e)) ? 0 : 1; }();  // A/B switch

static int pick_splits(int M, int N, int K) {
  const int bn = N >= 256 ? 256 : 128;  // f5b_gemm_tn's tile width
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
  const int kblocks = (K + 63) / 64;
  int s = (4 * sm_count() + tiles - 1) / tiles;
  if (s > kblocks) s = kblocks;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

}  // namespace f5b

using namespace f5b;
#define ST(s) static_cast<cudaStream_t>(s)
#define F5B_TRY(call)           \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != 0) return rc__; \
  } while (0)

// dX[rows, K_in] (bf16) = dY[rows, N_out] W[N_out, K_in]
static int dgrad(const void* dY, int ldy, const void* W, int k_in, void* dX, int ldx, int rows, int n_out, f5b_stream_t s) {
  return f5b_gemm_tn(dY, ldy, 0, W, k_in, 1, dX, ldx, 0, rows, k_in, n_out, 1, s);
}
// dW[N_out, K_in] (f32, accumulated) += dY[rows, N_out]^T X[rows, K_in]
static int wgrad(const void* dY, int ldy, const void* X, int ldx, float* dW, int ldw, int rows, int n_out, int k_in, f5b_stream_t s) {
  if (dW == nullptr) return 0;
  return f5b_gemm_tn(dY, ldy, 1, X, ldx, 1, dW, ldw, 1, n_out, k_in, rows, pick_splits(n_out, k_in, rows), s);
}

// Forward of DiT block i up to z2 (everything the block's backward reads): a = AdaLN(x_in) is expected in b.a unless `with_ln`.
static int train_block_forward(const F5bDitDesc& d, const TrainWs& w, int i, int B, int n, const int32_t* lens, const float* rope, bool with_ln,
                               cudaStream_t s) {
  const int D = d.dim, F = d.ff_mult * d.dim, H = d.heads;
  const int rows = B * n;
  const int64_t mod_dim = (int64_t)d.depth * 6 * D + 2 * D;
  f5b_stream_t stream = reinterpret_cast<f5b_stream_t>(s);
  const BlockSave& b = w.blk[i];
  const float* m = w.mod + (size_t)i * 6 * D;  // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp (modules.py:312)
  const bf* qkv_w = reinterpret_cast<const bf*>(d.qkv_w);
  const bf* out_w = reinterpret_cast<const bf*>(d.out_w);
  const bf* ff1_w = reinterpret_cast<const bf*>(d.ff1_w);
  const bf* ff2_w = reinterpret_cast<const bf*>(d.ff2_w);
  if (with_ln) F5B_TRY(ln_modulate(b.x_in, m + D, m, mod_dim, 0, b.a, rows, n, D, 1e-6f, s));
  F5bGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = rows; g.N = 3 * D; g.K = D; g.epi = F5B_EPI_QKV_ROPE; g.act = F5B_ACT_NONE;
  g.bias = d.qkv_b + (size_t)i * 3 * D;
  g.out = b.qkv; g.ldc = 3 * D;
  g.rows_per_batch = n; g.rope = rope; g.rope_heads = d.rope_heads; g.heads = H;
  F5B_TRY(gemm(b.a, D, qkv_w + (size_t)i * 3 * D * D, D, g, s));
  const AttnDrop adrop = attn_drop_for_layer(i);  // SDPA dropout (f5b_train_set_attn_dropout); addc == 0 when off
  F5B_TRY(attn_fwd(b.qkv, b.qkv + D, b.qkv + 2 * D, 3 * D, b.o, b.lse, lens, 0, B, H, n, 0.125f, s, &adrop));
  F5B_TRY(linear_bf16(b.o, D, out_w + (size_t)i * D * D, D, d.out_b + (size_t)i * D, b.z1, D, rows, D, D, F5B_ACT_NONE, s));
  // x_mid = x_in + gate_msa * mask(z1);  f = LN(x_mid) (1 + scale_mlp) + shift_mlp
  // (train mode: to_out's Dropout sits between z1 and the gate -- site 1; FeedForward's follows the GELU -- site 0)
  F5B_TRY(f5b_gate_add_ln_modulate_site(b.x_in, b.z1, m + 2 * D, mod_dim, lens, b.x_mid, m + 4 * D, m + 3 * D, mod_dim, b.f, B, n, D,
                                        1e-6f, i, 1, stream));
  if (!train_dropout_on()) {
    // pre-GELU tensor (kept for the backward) and GELU output from ONE accumulator tile: no sweep that re-reads h1
    F5B_TRY(linear_bf16_dual(b.f, D, ff1_w + (size_t)i * F * D, D, d.ff1_b + (size_t)i * F, b.h1, F, b.u, F, rows, F, D, F5B_ACT_GELU_TANH, s));
  } else {  // FeedForward's Dropout sits between GELU and the second Linear: the masked sweep
    F5B_TRY(linear_bf16(b.f, D, ff1_w + (size_t)i * F * D, D, d.ff1_b + (size_t)i * F, b.h1, F, rows, F, D, F5B_ACT_NONE, s));
    F5B_TRY(f5b_act_fwd_site(b.h1, b.u, (int64_t)rows * F, F5B_ACT_GELU_TANH, i, 0, stream));
  }
  F5B_TRY(linear_bf16(b.u, F, ff2_w + (size_t)i * D * F, F, d.ff2_b + (size_t)i * D, b.z2, D, rows, D, F, F5B_ACT_NONE, s));
  return 0;
}

extern "C" {

/* activation checkpointing for the f5b_dit_train_* drivers (see g_ckpt above): set it BEFORE f5b_dit_train_ws_bytes / _forward and
 * leave it unchanged until the matching backward.  Gradients are bit-identical with and without it. */
int f5b_train_set_checkpoint(int on) {
  g_ckpt = on ? 1 : 0;
  return 0;
}

size_t f5b_dit_train_ws_bytes(const F5bDit* h, int B, int n) {
  if (!h || B <= 0 || n <= 0 || h->d.depth > MAX_DEPTH) return 0;
  return carve_train(h->d, B, n, nullptr).bytes;
}

// DiT.forward (model/backbones/dit.py:185-233) in training form: per-sample time values, activations saved for the backward.
//   x f32 [B*n, mel] (phi_t), cond f32 [B*n, mel] or NULL (drop_audio_cond), text_embed f32 [B*n, T] (TextEmbedding output),
//   time f32 [B], lens int32 [B] or NULL, rope f32 [n, 32, 2]  ->  pred f32 [B*n, mel]
int f5b_dit_train_forward(const F5bDit* h, const float* x, const float* cond, const float* text_embed, const float* time, int B, int n,
                          const int32_t* lens, const float* rope, float* pred, void* ws, size_t ws_bytes, f5b_stream_t stream) {
  F5B_CHECK(h && x && text_embed && time && rope && pred && ws && B > 0 && n > 0, "f5b_dit_train_forward: bad argument");
  const F5bDitDesc& d = h->d;
  F5B_CHECK(d.depth <= MAX_DEPTH && d.dim <= 1024, "f5b_dit_train_forward: depth <= %d and dim <= 1024 supported", MAX_DEPTH);
  F5B_CHECK(d.precision == 0, "f5b_dit_train_forward: the training drivers run bf16 operands (F5bDitDesc.precision 0) only");
  const int D = d.dim, F = d.ff_mult * d.dim, H = d.heads, T = d.text_dim, mel = d.mel_dim;
  const int rows = B * n;
  const int64_t mod_dim = (int64_t)d.depth * 6 * D + 2 * D;
  TrainWs w = carve_train(d, B, n, ws);
  F5B_CHECK(w.bytes <= ws_bytes, "f5b_dit_train_forward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  cudaStream_t s = ST(stream);

  // TimestepEmbedding (model/modules.py:721-731) + all AdaLN linears (:311, :332); pre-activations kept
  F5B_TRY(f5b_time_sinus(time, w.sin_bf, B, stream));
  F5B_TRY(linear_bf16(w.sin_bf, 256, d.time_w0, 256, d.time_b0, w.a1, D, B, D, 256, F5B_ACT_NONE, s));
  F5B_TRY(f5b_act_fwd(w.a1, w.h1, (int64_t)B * D, F5B_ACT_SILU, stream));
  F5B_TRY(linear_bf16(w.h1, D, d.time_w2, D, d.time_b2, w.a2, D, B, D, D, F5B_ACT_NONE, s));
  F5B_TRY(f5b_act_fwd(w.a2, w.h2, (int64_t)B * D, F5B_ACT_SILU, stream));
  F5B_TRY(linear_f32(w.h2, D, d.mod_w, D, d.mod_b, w.mod, (int)mod_dim, B, (int)mod_dim, D, F5B_ACT_NONE, nullptr, 0, nullptr, 0, s));

  // InputEmbedding (dit.py:91-97)
  const int KC = 128 + T;
  F5B_TRY(f5b_pack_bf16(x, mel, w.a_x, 128, rows, mel, 128, stream));
  F5B_TRY(f5b_pack_bf16(cond, mel, w.a_ct, KC, rows, cond ? mel : 0, 128, stream));
  F5B_TRY(f5b_pack_bf16(text_embed, T, w.a_ct + 128, KC, rows, T, T, stream));
  F5B_TRY(linear_f32(w.a_ct, KC, d.in_wct, KC, d.in_b, w.h0, D, rows, D, KC, F5B_ACT_NONE, nullptr, 0, nullptr, 0, s));
  F5B_TRY(linear_f32(w.a_x, 128, d.in_wx, 128, nullptr, w.h0, D, rows, D, 128, F5B_ACT_NONE, w.h0, D, w.hb0, D, s));
  F5B_TRY(convpos(w.hb0, d.cp_w1, d.cp_b1, w.u1, nullptr, B, n, D, d.convpos_groups, d.convpos_kernel, 2, s));
  F5B_TRY(f5b_act_fwd(w.u1, w.c1, (int64_t)rows * D, F5B_ACT_MISH, stream));
  F5B_TRY(convpos(w.c1, d.cp_w2, d.cp_b2, w.u2, nullptr, B, n, D, d.convpos_groups, d.convpos_kernel, 2, s));
  F5B_TRY(f5b_act_fwd(w.u2, w.t1, (int64_t)rows * D, F5B_ACT_MISH, stream));
  F5B_TRY(f5b_gate_add(w.h0, w.t1, nullptr, 0, nullptr, w.blk[0].x_in, B, n, D, stream));

  // the gated residual of each branch is fused with the AdaLN of the next one (f5b_gate_add_ln_modulate): the first AdaLN of block
  // 0 runs alone, the last gated residual feeds AdaLayerNorm_Final
  {
    const float* m0 = w.mod;
    F5B_TRY(ln_modulate(w.blk[0].x_in, m0 + D, m0, mod_dim, 0, w.blk[0].a, rows, n, D, 1e-6f, s));
  }
  for (int i = 0; i < d.depth; ++i) {
    const BlockSave& b = w.blk[i];
    const bool last = i + 1 == d.depth;
    float* x_out = last ? w.x_fin : w.blk[i + 1].x_in;
    const float* m = w.mod + (size_t)i * 6 * D;
    F5B_TRY(train_block_forward(d, w, i, B, n, lens, rope, false, s));
    // x_out = x_mid + gate_mlp * z2;  next AdaLN: block i+1's (shift_msa, scale_msa) or AdaLayerNorm_Final's (scale, shift; :333)
    const float* mn = w.mod + (size_t)(i + 1) * 6 * D;
    const float* nscale = last ? mn : mn + D;
    const float* nshift = last ? mn + D : mn;
    bf* nout = last ? w.hbF : w.blk[i + 1].a;
    F5B_TRY(f5b_gate_add_ln_modulate(b.x_mid, b.z2, m + 5 * D, mod_dim, nullptr, x_out, nscale, nshift, mod_dim, nout, B, n, D, 1e-6f, stream));
  }
  F5B_TRY(linear_f32(w.hbF, D, d.proj_w, D, d.proj_b, pred, mel, rows, mel, D, F5B_ACT_NONE, nullptr, 0, nullptr, 0, s));
  return 0;
}

// Backward of f5b_dit_train_forward.  dpred bf16 [B*n, 128] (columns >= mel zero; f5b_mse_grad).  Gradients are ACCUMULATED into
// the f32 buffers of `g` (zero them first; NULL members are skipped).  cp_w1_t / cp_w2_t: conv_pos_embed weights packed by
// f5b_pack_convpos_weight_t.  dtext_bf16 (optional) receives d loss / d text_embed as bf16 [B*n, T].
static int train_backward_impl(const F5bDit* h, const void* dpred_bf16, const void* cp_w1_t, const void* cp_w2_t, const F5bDitGrads* gr,
                               void* dtext_bf16, int B, int n, const int32_t* lens, const float* rope, void* ws, size_t ws_bytes,
                               int parts, int blk_lo, int blk_hi, f5b_stream_t stream) {
  F5B_CHECK(h && dpred_bf16 && cp_w1_t && cp_w2_t && gr && rope && ws && B > 0 && n > 0, "f5b_dit_train_backward: bad argument");
  const F5bDitDesc& d = h->d;
  F5B_CHECK(d.depth <= MAX_DEPTH && d.dim <= 1024, "f5b_dit_train_backward: depth <= %d and dim <= 1024 supported", MAX_DEPTH);
  const int D = d.dim, F = d.ff_mult * d.dim, H = d.heads, T = d.text_dim, mel = d.mel_dim;
  const int rows = B * n;
  const int64_t mod_dim = (int64_t)d.depth * 6 * D + 2 * D;
  TrainWs w = carve_train(d, B, n, ws);
  F5B_CHECK(w.bytes <= ws_bytes, "f5b_dit_train_backward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  cudaStream_t s = ST(stream);
  const F5bDitGrads& g = *gr;
  const bf* qkv_w = reinterpret_cast<const bf*>(d.qkv_w);
  const bf* out_w = reinterpret_cast<const bf*>(d.out_w);
  const bf* ff1_w = reinterpret_cast<const bf*>(d.ff1_w);
  const bf* ff2_w = reinterpret_cast<const bf*>(d.ff2_w);
  auto off = [](float* p, size_t o) { return p ? p + o : nullptr; };

  F5B_CHECK(blk_lo >= 0 && blk_hi <= d.depth && blk_lo <= blk_hi, "f5b_dit_train_backward: bad block range [%d, %d)", blk_lo, blk_hi);
  if (parts & 1) {
    F5B_CUDA(cudaMemsetAsync(w.dmod, 0, sizeof(float) * B * mod_dim, s));
    // proj_out + AdaLayerNorm_Final
    F5B_TRY(f5b_act_bwd(dpred_bf16, nullptr, nullptr, g.proj_b, rows, mel, 128, F5B_ACT_NONE, stream));
    F5B_TRY(wgrad(dpred_bf16, 128, w.hbF, D, g.proj_w, D, rows, mel, D, stream));
    F5B_TRY(dgrad(dpred_bf16, 128, d.proj_w, D, w.t1, D, rows, mel, stream));
    float* dmf = w.dmod + (size_t)d.depth * 6 * D;
    F5B_TRY(f5b_ln_modulate_bwd(w.t1, w.x_fin, w.mod + (size_t)d.depth * 6 * D, mod_dim, w.dx, 0, dmf, dmf + D, B, n, D, 1e-6f, stream));
  }

  for (int i = ((parts & 2) ? blk_hi : 0) - 1; i >= ((parts & 2) ? blk_lo : 0); --i) {
    const BlockSave& b = w.blk[i];
    const float* m = w.mod + (size_t)i * 6 * D;
    float* dm = w.dmod + (size_t)i * 6 * D;
    if (g_ckpt) F5B_TRY(train_block_forward(d, w, i, B, n, lens, rope, true, s));  // re-materialise the block's intermediates
    // ---- x_out = x_mid + gate_mlp * (W2 gelu(W1 f + b1) + b2),  f = LN(x_mid) (1 + scale_mlp) + shift_mlp
    F5B_TRY(f5b_gate_bwd(w.dx, b.z2, m + 5 * D, mod_dim, nullptr, w.t1, dm + 5 * D, off(g.ff2_b, (size_t)i * D), B, n, D, stream));
    F5B_TRY(wgrad(w.t1, D, b.u, F, off(g.ff2_w, (size_t)i * D * F), F, rows, D, F, stream));
    F5B_TRY(dgrad(w.t1, D, ff2_w + (size_t)i * D * F, F, w.tF, F, rows, D, stream));
    F5B_TRY(f5b_act_bwd_site(w.tF, b.h1, w.tF, off(g.ff1_b, (size_t)i * F), rows, F, F, F5B_ACT_GELU_TANH, i, 0, stream));
    F5B_TRY(wgrad(w.tF, F, b.f, D, off(g.ff1_w, (size_t)i * F * D), D, rows, F, D, stream));
    F5B_TRY(dgrad(w.tF, F, ff1_w + (size_t)i * F * D, D, w.t1, D, rows, F, stream));
    F5B_TRY(f5b_ln_modulate_bwd(w.t1, b.x_mid, m + 4 * D, mod_dim, w.dx, 1, dm + 4 * D, dm + 3 * D, B, n, D, 1e-6f, stream));
    // ---- x_mid = x_in + gate_msa * mask(Wo attn(rope(Wqkv a + b)) + bo),  a = LN(x_in) (1 + scale_msa) + shift_msa
    F5B_TRY(f5b_gate_bwd_site(w.dx, b.z1, m + 2 * D, mod_dim, lens, w.t1, dm + 2 * D, off(g.out_b, (size_t)i * D), B, n, D, i, 1, stream));
    F5B_TRY(wgrad(w.t1, D, b.o, D, off(g.out_w, (size_t)i * D * D), D, rows, D, D, stream));
    F5B_TRY(dgrad(w.t1, D, out_w + (size_t)i * D * D, D, w.t2, D, rows, D, stream));
    const AttnDrop adrop = attn_drop_for_layer(i);
    F5B_TRY(attn_bwd(b.qkv, b.qkv + D, b.qkv + 2 * D, 3 * D, b.o, w.t2, D, b.lse, w.delta, w.dq_ws, w.t3, 3 * D, lens, 0, B, H, n, 0.125f,
                     rope, d.rope_heads, s, &adrop));
    F5B_TRY(f5b_act_bwd(w.t3, nullptr, nullptr, off(g.qkv_b, (size_t)i * 3 * D), rows, 3 * D, 3 * D, F5B_ACT_NONE, stream));
    F5B_TRY(wgrad(w.t3, 3 * D, b.a, D, off(g.qkv_w, (size_t)i * 3 * D * D), D, rows, 3 * D, D, stream));
    F5B_TRY(dgrad(w.t3, 3 * D, qkv_w + (size_t)i * 3 * D * D, D, w.t1, D, rows, 3 * D, stream));
    F5B_TRY(f5b_ln_modulate_bwd(w.t1, b.x_in, m + D, mod_dim, w.dx, 1, dm + D, dm, B, n, D, 1e-6f, stream));
  }

  if (!(parts & 4)) return 0;
  // ---- InputEmbedding: x0 = h0 + mish(conv2(mish(conv1(h0)))),  h0 = [x | cond | text] W^T + b
  const int G = d.convpos_groups, cpg = D / G, ks = d.convpos_kernel, ccols = cpg * ks;
  F5B_CHECK((cpg & 7) == 0, "f5b_dit_train_backward: channels per conv group must be a multiple of 8");
  // grouped-conv weight gradient: per group, im2col of the layer input (k-major) and one split-K wgrad GEMM into a [co][k][ci] buffer
  auto conv_wgrad = [&](const bf* dy, const bf* xin, float* dw) -> int {
    if (dw == nullptr) return 0;
    F5B_CUDA(cudaMemsetAsync(w.cpw_tmp, 0, sizeof(float) * (size_t)D * ccols, s));
    for (int gi = 0; gi < G; ++gi) {
      if (cpg == 64 && g_conv_wgrad_implicit) {  // no im2col buffer: the taps are TMA position offsets (gemm_grad.cu)
        F5B_TRY(conv_wgrad_implicit(dy + gi * cpg, xin + gi * cpg, D, w.cpw_tmp + (size_t)gi * cpg * ccols, ccols, B, n, cpg, ks,
                                    pick_splits(cpg, ccols, rows), s));
        continue;
      }
      {
        LaunchScope scope(K_ELEMENTWISE, s, 0, 4.0 * rows * ccols);
        im2col_convpos_kernel<<<rows, 256, 0, s>>>(xin, w.xcol, n, D, gi, cpg, ks);
        F5B_CUDA(cudaGetLastError());
      }
      F5B_TRY(wgrad(dy + gi * cpg, D, w.xcol, ccols, w.cpw_tmp + (size_t)gi * cpg * ccols, ccols, rows, cpg, ccols, stream));
    }
    LaunchScope scope(K_ELEMENTWISE, s, 0, 12.0 * D * ccols);
    conv_wgrad_fold_kernel<<<(D * ccols + 255) / 256, 256, 0, s>>>(w.cpw_tmp, dw, D, cpg, ks);
    F5B_CUDA(cudaGetLastError());
    return 0;
  };
  F5B_TRY(f5b_gate_bwd(w.dx, nullptr, nullptr, 0, nullptr, w.t1, nullptr, nullptr, B, n, D, stream));       // bf16(dx)
  F5B_TRY(f5b_act_bwd(w.t1, w.u2, w.t2, g.cp_b2, rows, D, D, F5B_ACT_MISH, stream));                         // d u2
  F5B_TRY(conv_wgrad(w.t2, w.c1, g.cp_w2));
  F5B_TRY(convpos(w.t2, cp_w2_t, nullptr, w.t1, nullptr, B, n, D, G, ks, 2, s));                             // d c1
  F5B_TRY(f5b_act_bwd(w.t1, w.u1, w.t2, g.cp_b1, rows, D, D, F5B_ACT_MISH, stream));                         // d u1
  F5B_TRY(conv_wgrad(w.t2, w.hb0, g.cp_w1));
  F5B_TRY(convpos(w.t2, cp_w1_t, nullptr, w.t1, nullptr, B, n, D, G, ks, 2, s));                             // conv path of d h0
  F5B_TRY(f5b_gate_add(w.dx, w.t1, nullptr, 0, nullptr, w.dx, B, n, D, stream));                             // d h0 = dx + conv path
  F5B_TRY(f5b_gate_bwd(w.dx, nullptr, nullptr, 0, nullptr, w.t1, nullptr, g.in_b, B, n, D, stream));         // bf16(d h0), d bias
  const int KC = 128 + T;
  F5B_TRY(wgrad(w.t1, D, w.a_x, 128, g.in_wx, 128, rows, D, 128, stream));
  F5B_TRY(wgrad(w.t1, D, w.a_ct, KC, g.in_wct, KC, rows, D, KC, stream));
  if (dtext_bf16 != nullptr) {
    // d text_embed = d h0 W_ct[:, 128:]  (the text columns of the input projection)
    F5B_TRY(f5b_gemm_tn(w.t1, D, 0, reinterpret_cast<const bf*>(d.in_wct) + 128, KC, 1, dtext_bf16, T, 0, rows, T, D, 1, stream));
  }

  // ---- modulation + time MLP: mod = silu(a2) Wm^T + bm, a2 = silu(a1) W2^T + b2, a1 = sinus(t) W0^T + b0
  F5B_TRY(f5b_pack_bf16(w.dmod, (int)mod_dim, w.dmod_bf, (int)mod_dim, B, (int)mod_dim, (int)mod_dim, stream));
  F5B_TRY(f5b_act_bwd(w.dmod_bf, nullptr, nullptr, g.mod_b, B, (int)mod_dim, (int)mod_dim, F5B_ACT_NONE, stream));
  F5B_TRY(wgrad(w.dmod_bf, (int)mod_dim, w.h2, D, g.mod_w, D, B, (int)mod_dim, D, stream));
  F5B_CUDA(cudaMemsetAsync(w.dh2, 0, sizeof(float) * B * D, s));
  {
    const int kb = ((int)mod_dim + 63) / 64;
    F5B_TRY(f5b_gemm_tn(w.dmod_bf, (int)mod_dim, 0, d.mod_w, D, 1, w.dh2, D, 1, B, D, (int)mod_dim, kb < 32 ? kb : 32, stream));
  }
  F5B_TRY(f5b_pack_bf16(w.dh2, D, w.tb1, D, B, D, D, stream));
  F5B_TRY(f5b_act_bwd(w.tb1, w.a2, w.tb1, g.time_b2, B, D, D, F5B_ACT_SILU, stream));                        // d a2
  F5B_TRY(wgrad(w.tb1, D, w.h1, D, g.time_w2, D, B, D, D, stream));
  F5B_TRY(dgrad(w.tb1, D, d.time_w2, D, w.tb2, D, B, D, stream));
  F5B_TRY(f5b_act_bwd(w.tb2, w.a1, w.tb2, g.time_b0, B, D, D, F5B_ACT_SILU, stream));                        // d a1
  F5B_TRY(wgrad(w.tb2, D, w.sin_bf, 256, g.time_w0, 256, B, D, 256, stream));
  return 0;
}

int f5b_dit_train_backward(const F5bDit* h, const void* dpred_bf16, const void* cp_w1_t, const void* cp_w2_t, const F5bDitGrads* gr,
                           void* dtext_bf16, int B, int n, const int32_t* lens, const float* rope, void* ws, size_t ws_bytes,
                           f5b_stream_t stream) {
  return train_backward_impl(h, dpred_bf16, cp_w1_t, cp_w2_t, gr, dtext_bf16, B, n, lens, rope, ws, ws_bytes, 7, 0, h ? h->d.depth : 0, stream);
}

// The same backward in pieces, so that the host can start the gradient all-reduce of finished blocks while earlier blocks are
// still being differentiated: parts bit 0 = head (proj_out, final AdaLN), bit 1 = blocks [blk_lo, blk_hi) in descending order,
// bit 2 = tail (input embedding, modulation / time MLP).  Pieces must be issued head, blocks from depth down to 0, tail.
int f5b_dit_train_backward_part(const F5bDit* h, const void* dpred_bf16, const void* cp_w1_t, const void* cp_w2_t, const F5bDitGrads* gr,
                                void* dtext_bf16, int B, int n, const int32_t* lens, const float* rope, void* ws, size_t ws_bytes,
                                int parts, int blk_lo, int blk_hi, f5b_stream_t stream) {
  return train_backward_impl(h, dpred_bf16, cp_w1_t, cp_w2_t, gr, dtext_bf16, B, n, lens, rope, ws, ws_bytes, parts, blk_lo, blk_hi, stream);
}

size_t f5b_dit_text_train_ws_bytes(const F5bDit* h, int B, int n) {
  if (!h || B <= 0 || n <= 0 || h->d.conv_layers > 16) return 0;
  return carve_text(h->d, B, n, nullptr).bytes;
}

// TextEmbedding.forward (model/backbones/dit.py:49-79) in training form: un-fused GELU, every ConvNeXtV2Block input kept
int f5b_dit_text_embed_train(const F5bDit* h, const int64_t* ids, int nt, int B, int n, int drop_text, float* out, void* ws,
                             size_t ws_bytes, f5b_stream_t stream) {
  F5B_CHECK(h && ids && out && ws && B > 0 && n > 0 && nt > 0, "f5b_dit_text_embed_train: bad argument");
  const F5bDitDesc& d = h->d;
  F5B_CHECK(d.conv_layers <= 16 && d.text_dim <= 1024, "f5b_dit_text_embed_train: conv_layers <= 16, text_dim <= 1024");
  F5B_CHECK(!(d.text_mask_padding && d.conv_layers > 0), "f5b_dit_text_embed_train: text_mask_padding is not built for training");
  const int T = d.text_dim, T2 = 2 * d.text_dim;
  const size_t rows = (size_t)B * n;
  TextWs w = carve_text(d, B, n, ws);
  F5B_CHECK(w.bytes <= ws_bytes, "f5b_dit_text_embed_train: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  cudaStream_t s = ST(stream);
  float* first = d.conv_layers > 0 ? w.L[0].h_in : out;
  F5B_TRY(f5b_text_lookup(ids, nt, d.text_table, d.text_pos, first, nullptr, B, n, T, d.vocab_rows, drop_text, d.conv_layers > 0, stream));
  for (int j = 0; j < d.conv_layers; ++j) {
    const TextSave& L = w.L[j];
    float* h_out = (j + 1 < d.conv_layers) ? w.L[j + 1].h_in : out;
    F5B_TRY(dwconv7_ln(L.h_in, d.tb_dw_w + (size_t)j * T * 7, d.tb_dw_b + (size_t)j * T, d.tb_ln_w + (size_t)j * T, d.tb_ln_b + (size_t)j * T,
                       L.tb, B, n, T, 1e-6f, s, L.y));
    F5B_TRY(linear_bf16(L.tb, T, reinterpret_cast<const bf*>(d.tb_pw1_w) + (size_t)j * T2 * T, T, d.tb_pw1_b + (size_t)j * T2, L.p1, T2,
                        (int)rows, T2, T, F5B_ACT_NONE, s));
    F5B_TRY(f5b_act_fwd(L.p1, L.t2, (int64_t)rows * T2, F5B_ACT_GELU_ERF, stream));
    F5B_TRY(grn(L.t2, d.tb_grn_g + (size_t)j * T2, d.tb_grn_b + (size_t)j * T2, L.t3, w.gx, B, n, T2, s));
    F5B_CUDA(cudaMemcpyAsync(h_out, L.h_in, rows * T * sizeof(float), cudaMemcpyDeviceToDevice, s));
    F5B_TRY(linear_gate_resid(L.t3, T2, reinterpret_cast<const bf*>(d.tb_pw2_w) + (size_t)j * T * T2, T2, d.tb_pw2_b + (size_t)j * T, h_out,
                              T, (int)rows, T, T2, n, nullptr, 0, nullptr, 0, s));
  }
  return 0;
}

// ... and its backward: dtext bf16 [B*n, T] (f5b_dit_train_backward) -> gradients of the table and the ConvNeXtV2 blocks
int f5b_dit_text_embed_backward(const F5bDit* h, const int64_t* ids, int nt, int B, int n, int drop_text, const void* dtext_bf16,
                                const F5bDitGrads* gr, void* ws, size_t ws_bytes, f5b_stream_t stream) {
  F5B_CHECK(h && ids && dtext_bf16 && gr && ws && B > 0 && n > 0 && nt > 0, "f5b_dit_text_embed_backward: bad argument");
  const F5bDitDesc& d = h->d;
  F5B_CHECK(d.conv_layers <= 16 && d.text_dim <= 1024, "f5b_dit_text_embed_backward: conv_layers <= 16, text_dim <= 1024");
  const int T = d.text_dim, T2 = 2 * d.text_dim;
  const int rows = B * n;
  TextWs w = carve_text(d, B, n, ws);
  F5B_CHECK(w.bytes <= ws_bytes, "f5b_dit_text_embed_backward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  cudaStream_t s = ST(stream);
  const F5bDitGrads& g = *gr;
  auto off = [](float* p, size_t o) { return p ? p + o : nullptr; };
  F5B_CUDA(cudaMemsetAsync(w.dh, 0, sizeof(float) * (size_t)rows * T, s));
  F5B_TRY(f5b_gate_add(w.dh, dtext_bf16, nullptr, 0, nullptr, w.dh, B, n, T, stream));  // fp32 running gradient
  for (int j = d.conv_layers - 1; j >= 0; --j) {
    const TextSave& L = w.L[j];
    const bf* w1 = reinterpret_cast<const bf*>(d.tb_pw1_w) + (size_t)j * T2 * T;
    const bf* w2 = reinterpret_cast<const bf*>(d.tb_pw2_w) + (size_t)j * T * T2;
    F5B_TRY(f5b_gate_bwd(w.dh, nullptr, nullptr, 0, nullptr, w.dz, nullptr, off(g.tb_pw2_b, (size_t)j * T), B, n, T, stream));
    F5B_TRY(wgrad(w.dz, T, L.t3, T2, off(g.tb_pw2_w, (size_t)j * T * T2), T2, rows, T, T2, stream));
    F5B_TRY(dgrad(w.dz, T, w2, T2, w.d2, T2, rows, T, stream));
    F5B_TRY(f5b_grn_gelu_bwd(w.d2, L.t2, L.p1, d.tb_grn_g + (size_t)j * T2, w.d2, off(g.tb_grn_g, (size_t)j * T2),
                             off(g.tb_grn_b, (size_t)j * T2), off(g.tb_pw1_b, (size_t)j * T2), w.stats, B, n, T2, stream));
    F5B_TRY(wgrad(w.d2, T2, L.tb, T, off(g.tb_pw1_w, (size_t)j * T2 * T), T, rows, T2, T, stream));
    F5B_TRY(dgrad(w.d2, T2, w1, T, w.dtb, T, rows, T2, stream));
    F5B_TRY(f5b_ln_affine_bwd(w.dtb, L.y, d.tb_ln_w + (size_t)j * T, w.dy, 0, off(g.tb_ln_w, (size_t)j * T), off(g.tb_ln_b, (size_t)j * T), B,
                              n, T, 1e-6f, stream));
    F5B_TRY(f5b_dwconv7_bwd(w.dy, L.h_in, d.tb_dw_w + (size_t)j * T * 7, w.dh, off(g.tb_dw_w, (size_t)j * T * 7),
                            off(g.tb_dw_b, (size_t)j * T), B, n, T, stream));
  }
  if (g.text_table) F5B_TRY(f5b_text_lookup_bwd(w.dh, ids, nt, g.text_table, B, n, T, d.vocab_rows, drop_text, stream));
  return 0;
}

}  // extern "C"

extern "C" void f5b_debug_conv_wgrad_implicit(int on) { f5b::g_conv_wgrad_implicit = on; }
