// Model-level drivers: the launch sequences of DiT.forward, TextEmbedding, the per-sample modulation table, and Vocos.decode.
// Host-side only (no kernels here); every launch goes to the caller's stream, nothing is allocated, so a whole ODE step is
// CUDA-graph capturable.  Reference citations are relative to /root/reference/src/f5_tts/.
#include <new>

#include "common.cuh"
#include "f5b_internal.h"
#include "dropout.cuh"

struct F5bDit {
  F5bDitDesc d;
};
struct F5bVocos {
  F5bVocosDesc d;
};

namespace f5b {

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int round8(int x) { return (x + 7) / 8 * 8; }

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

// activation buffers between the kernels: bf16, or fp32 words (rounded to tf32) in the tf32 operand mode (F5bDitDesc.precision 1)
struct DitWs {
  float* x;
  uint8_t *hb, *qkv, *ab, *fb;
  float* vt;
  size_t bytes;
};

static inline bool is_tf32(const F5bDitDesc& d) { return d.precision == 1; }
static inline size_t act_bytes(const F5bDitDesc& d) { return is_tf32(d) ? 4 : 2; }
// element `elems` of a weight array stored in the mode's operand type
static inline const void* wat(const F5bDitDesc& d, const void* base, size_t elems) {
  return reinterpret_cast<const uint8_t*>(base) + elems * act_bytes(d);
}
static inline uint8_t* aat(const F5bDitDesc& d, uint8_t* base, size_t elems) { return base + elems * act_bytes(d); }

static DitWs carve_dit(const F5bDitDesc& d, int B, int n, void* ws) {
  const size_t rows = (size_t)B * n;
  const int D = d.dim, F = d.ff_mult * d.dim, H = d.heads;
  const size_t es = act_bytes(d);
  Carver c(ws);
  DitWs w;
  w.x = c.take<float>(rows * D);
  w.hb = c.take<uint8_t>(rows * D * es);
  w.qkv = c.take<uint8_t>(rows * (size_t)H * 64 * 3 * es);  // token-major [rows, 3D]: q | k | v
  w.ab = c.take<uint8_t>(rows * D * es);
  w.fb = c.take<uint8_t>(rows * F * es);
  w.vt = is_tf32(d) ? c.take<float>(attn_tf32_ws_floats(B, H, n)) : nullptr;  // V^T of the tf32 attention
  w.bytes = c.off;
  return w;
}

// f5b_dit_set_attn_dropout: the reference's SDPA dropout at INFERENCE (model/modules.py:490 passes dropout_p = 0.1 in eval() too)
static float g_infer_attn_p = 0.f;
static uint64_t g_infer_attn_seed = 0;
static const uint32_t* g_infer_attn_seed_dev = nullptr;

}  // namespace f5b

using namespace f5b;
#define ST(s) static_cast<cudaStream_t>(s)
#define F5B_TRY(call)        \
  do {                       \
    int rc__ = (call);       \
    if (rc__ != 0) return rc__; \
  } while (0)

extern "C" {

const char* f5b_last_error(void) { return f5b::last_error(); }
int f5b_abi_version(void) { return F5B_ABI_VERSION; }

int f5b_dit_set_attn_dropout(float p, uint64_t seed, const uint32_t* seed_dev) {
  F5B_CHECK(p >= 0.f && p < 1.f, "f5b_dit_set_attn_dropout: p must be in [0, 1)");
  g_infer_attn_p = p;
  g_infer_attn_seed = seed;
  g_infer_attn_seed_dev = seed_dev;
  return 0;
}

int f5b_dit_create(const F5bDitDesc* desc, F5bDit** out) {
  F5B_CHECK(desc && out, "f5b_dit_create: null argument");
  const F5bDitDesc& d = *desc;
  F5B_CHECK(d.dim > 0 && d.depth > 0 && d.heads > 0, "f5b_dit_create: bad dims");
  F5B_CHECK(d.dim_head == 64, "f5b_dit_create: dim_head must be 64 (got %d)", d.dim_head);
  F5B_CHECK(d.heads * d.dim_head == d.dim, "f5b_dit_create: heads*dim_head (%d) must equal dim (%d)", d.heads * d.dim_head, d.dim);
  F5B_CHECK(d.dim % 8 == 0 && d.text_dim % 8 == 0, "f5b_dit_create: dim and text_dim must be multiples of 8");
  F5B_CHECK(d.precision == 0 || d.precision == 1, "f5b_dit_create: precision must be 0 (bf16 operands) or 1 (tf32 operands)");
  F5B_CHECK(d.mel_dim > 0 && d.mel_dim <= 128, "f5b_dit_create: mel_dim must be <= 128");
  F5B_CHECK(d.convpos_groups > 0 && d.dim % d.convpos_groups == 0 && d.dim / d.convpos_groups <= 64,
            "f5b_dit_create: conv_pos_embed needs <= 64 channels per group");
  F5B_CHECK(d.rope_heads >= 0 && d.rope_heads <= d.heads, "f5b_dit_create: rope_heads");
  F5B_CHECK(d.time_w0 && d.time_w2 && d.mod_w && d.text_table && d.in_wx && d.in_wct && d.cp_w1 && d.cp_w2 && d.qkv_w && d.out_w &&
                d.ff1_w && d.ff2_w && d.proj_w,
            "f5b_dit_create: null weight pointer");
  F5bDit* h = new (std::nothrow) F5bDit;
  F5B_CHECK(h != nullptr, "f5b_dit_create: out of host memory");
  h->d = d;
  *out = h;
  return 0;
}

void f5b_dit_destroy(F5bDit* h) { delete h; }

size_t f5b_dit_workspace_bytes(const F5bDit* h, int B, int n) {
  if (!h || B <= 0 || n <= 0) return 0;
  return carve_dit(h->d, B, n, nullptr).bytes;
}

// TimestepEmbedding (model/modules.py:721-731) + every AdaLayerNorm / AdaLayerNorm_Final linear (:311, :332) for M time values
int f5b_dit_modulation(const F5bDit* h, const float* t, int M, float* mod, void* ws, f5b_stream_t stream) {
  F5B_CHECK(h && t && mod && ws && M > 0, "f5b_dit_modulation: bad argument");
  const F5bDitDesc& d = h->d;
  const int D = d.dim;
  const int mod_dim = d.depth * 6 * D + 2 * D;
  const bool tp = is_tf32(d);
  const size_t es = act_bytes(d);
  Carver c(ws);
  uint8_t* sin_bf = c.take<uint8_t>((size_t)M * 256 * es);
  uint8_t* h1 = c.take<uint8_t>((size_t)M * D * es);
  uint8_t* h2 = c.take<uint8_t>((size_t)M * D * es);
  cudaStream_t s = ST(stream);
  if (tp) F5B_TRY(f5b_time_sinus_tf32(t, reinterpret_cast<float*>(sin_bf), M, stream));
  else F5B_TRY(f5b_time_sinus(t, sin_bf, M, stream));
  F5B_TRY(linear_bf16(sin_bf, 256, d.time_w0, 256, d.time_b0, h1, D, M, D, 256, F5B_ACT_SILU, s, tp));
  // every consumer of the time embedding applies SiLU first (AdaLayerNorm.silu), so it is fused here
  F5B_TRY(linear_bf16(h1, D, d.time_w2, D, d.time_b2, h2, D, M, D, D, F5B_ACT_SILU, s, tp));
  F5B_TRY(linear_f32(h2, D, d.mod_w, D, d.mod_b, mod, mod_dim, M, mod_dim, D, F5B_ACT_NONE, nullptr, 0, nullptr, 0, s, tp));
  return 0;
}

size_t f5b_dit_modulation_ws_bytes(const F5bDit* h, int M) {
  if (!h || M <= 0) return 0;
  const size_t es = act_bytes(h->d);
  return align_up((size_t)M * 256 * es) + 2 * align_up((size_t)M * h->d.dim * es);
}

size_t f5b_dit_text_ws_bytes(const F5bDit* h, int B, int n) {
  if (!h || B <= 0 || n <= 0) return 0;
  const size_t rows = (size_t)B * n;
  const int T = h->d.text_dim;
  const size_t es = act_bytes(h->d);
  return align_up(rows * T * es) + 2 * align_up(rows * 2 * T * es) + align_up((size_t)B * 2 * T * 4) + align_up(rows);
}

// TextEmbedding.forward, model/backbones/dit.py:49-79 with ConvNeXtV2Block model/modules.py:241-269
int f5b_dit_text_embed(const F5bDit* h, const int64_t* ids, int nt, int B, int n, int drop_text, float* out, void* ws,
                       f5b_stream_t stream) {
  F5B_CHECK(h && ids && out && ws && B > 0 && n > 0 && nt > 0, "f5b_dit_text_embed: bad argument");
  const F5bDitDesc& d = h->d;
  const int T = d.text_dim, T2 = 2 * d.text_dim;
  const size_t rows = (size_t)B * n;
  const bool tp = is_tf32(d);
  const size_t es = act_bytes(d);
  Carver c(ws);
  uint8_t* tb = c.take<uint8_t>(rows * T * es);
  uint8_t* t2 = c.take<uint8_t>(rows * T2 * es);
  uint8_t* t3 = c.take<uint8_t>(rows * T2 * es);
  float* gx = c.take<float>((size_t)B * T2);
  uint8_t* mask = c.take<uint8_t>(rows);
  cudaStream_t s = ST(stream);
  const bool mp = d.text_mask_padding != 0 && d.conv_layers > 0;
  F5B_TRY(f5b_text_lookup(ids, nt, d.text_table, d.text_pos, out, mp ? mask : nullptr, B, n, T, d.vocab_rows, drop_text, d.conv_layers > 0,
                          stream));
  if (mp) F5B_TRY(f5b_mask_rows_f32(out, mask, (int)rows, T, stream));
  for (int j = 0; j < d.conv_layers; ++j) {
    F5B_TRY(dwconv7_ln(out, d.tb_dw_w + (size_t)j * T * 7, d.tb_dw_b + (size_t)j * T, d.tb_ln_w + (size_t)j * T,
                       d.tb_ln_b + (size_t)j * T, tb, B, n, T, 1e-6f, s, nullptr, tp));
    F5B_TRY(linear_bf16(tb, T, wat(d, d.tb_pw1_w, (size_t)j * T2 * T), T, d.tb_pw1_b + (size_t)j * T2, t2, T2, (int)rows, T2, T,
                        F5B_ACT_GELU_ERF, s, tp));
    F5B_TRY(grn(t2, d.tb_grn_g + (size_t)j * T2, d.tb_grn_b + (size_t)j * T2, t3, gx, B, n, T2, s, tp));
    F5B_TRY(linear_gate_resid(t3, T2, wat(d, d.tb_pw2_w, (size_t)j * T * T2), T2, d.tb_pw2_b + (size_t)j * T, out, T, (int)rows, T, T2,
                              n, nullptr, 0, nullptr, 0, s, tp));
    if (mp) F5B_TRY(f5b_mask_rows_f32(out, mask, (int)rows, T, stream));
  }
  return 0;
}

// Step-invariant part of InputEmbedding.proj (model/backbones/dit.py:91-95): [cond | text_embed] W_ct^T + bias
int f5b_dit_input_const(const F5bDit* h, const float* cond, const float* text_embed, int B, int n, float* c0, void* ws,
                        f5b_stream_t stream) {
  F5B_CHECK(h && text_embed && c0 && ws && B > 0 && n > 0, "f5b_dit_input_const: bad argument");
  const F5bDitDesc& d = h->d;
  const int T = d.text_dim, K = 128 + T;
  const int rows = B * n;
  if (is_tf32(d)) {  // ws: rows * (128 + T) fp32
    float* a = reinterpret_cast<float*>(ws);
    F5B_TRY(f5b_pack_tf32(cond, d.mel_dim, a, K, rows, cond ? d.mel_dim : 0, 128, stream));
    F5B_TRY(f5b_pack_tf32(text_embed, T, a + 128, K, rows, T, T, stream));
    F5B_TRY(linear_f32(a, K, d.in_wct, K, d.in_b, c0, d.dim, rows, d.dim, K, F5B_ACT_NONE, nullptr, 0, nullptr, 0, ST(stream), true));
    return 0;
  }
  __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(ws);
  F5B_TRY(f5b_pack_bf16(cond, d.mel_dim, a, K, rows, cond ? d.mel_dim : 0, 128, stream));
  F5B_TRY(f5b_pack_bf16(text_embed, T, a + 128, K, rows, T, T, stream));
  F5B_TRY(linear_f32(a, K, d.in_wct, K, d.in_b, c0, d.dim, rows, d.dim, K, F5B_ACT_NONE, nullptr, 0, nullptr, 0, ST(stream)));
  return 0;
}

// DiT.forward, model/backbones/dit.py:185-233 (time/text embeddings hoisted out: `mod`, `c0`)
int f5b_dit_forward(const F5bDit* h, const void* x_bf16, int Bx, const float* c0, int Bf, int n, const float* mod,
                    int64_t mod_bstride, const int32_t* lens, const float* rope, float* pred, void* ws, size_t ws_bytes,
                    f5b_stream_t stream) {
  F5B_CHECK(h && x_bf16 && c0 && mod && rope && pred && ws, "f5b_dit_forward: null argument");
  F5B_CHECK(Bx > 0 && Bf > 0 && Bf % Bx == 0 && n > 0, "f5b_dit_forward: Bf (%d) must be a multiple of Bx (%d)", Bf, Bx);
  const F5bDitDesc& d = h->d;
  const int D = d.dim, F = d.ff_mult * d.dim, H = d.heads;
  const int rows = Bf * n;
  DitWs w = carve_dit(d, Bf, n, ws);
  F5B_CHECK(w.bytes <= ws_bytes, "f5b_dit_forward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  cudaStream_t s = ST(stream);
  const int batch_mod = Bx;

  const bool tp = is_tf32(d);
  // InputEmbedding (dit.py:91-97): x W_x^T + c0, then conv_pos_embed(h) + h (no mask)
  const int hrows = Bx * n;
  for (int half = 0; half < Bf / Bx; ++half) {
    const size_t off = (size_t)half * hrows * D;
    F5B_TRY(linear_f32(x_bf16, 128, d.in_wx, 128, nullptr, w.x + off, D, hrows, D, 128, F5B_ACT_NONE, c0 + off, D, aat(d, w.hb, off), D, s, tp));
  }
  if (tp) {
    F5B_TRY(convpos_tf32(reinterpret_cast<const float*>(w.hb), reinterpret_cast<const float*>(d.cp_w1), d.cp_b1,
                         reinterpret_cast<float*>(w.ab), nullptr, Bf, n, D, d.convpos_groups, d.convpos_kernel, 0, s));
    F5B_TRY(convpos_tf32(reinterpret_cast<const float*>(w.ab), reinterpret_cast<const float*>(d.cp_w2), d.cp_b2, nullptr, w.x, Bf, n, D,
                         d.convpos_groups, d.convpos_kernel, 1, s));
  } else {
    F5B_TRY(convpos(w.hb, d.cp_w1, d.cp_b1, w.ab, nullptr, Bf, n, D, d.convpos_groups, d.convpos_kernel, 0, s));
    F5B_TRY(convpos(w.ab, d.cp_w2, d.cp_b2, nullptr, w.x, Bf, n, D, d.convpos_groups, d.convpos_kernel, 1, s));
  }

  for (int i = 0; i < d.depth; ++i) {
    // AdaLayerNorm chunk order (model/modules.py:312): shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const float* m = mod + (size_t)i * 6 * D;
    F5B_TRY(ln_modulate(w.x, m + D, m, mod_bstride, batch_mod, w.hb, rows, n, D, 1e-6f, s, tp));
    F5bGemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = rows; g.N = 3 * D; g.K = D; g.epi = F5B_EPI_QKV_ROPE; g.act = F5B_ACT_NONE;
    g.bias = d.qkv_b + (size_t)i * 3 * D;
    g.out = w.qkv; g.ldc = 3 * D;
    g.rows_per_batch = n; g.rope = rope; g.rope_heads = d.rope_heads; g.heads = H;
    g.tf32 = tp;
    F5B_TRY(gemm(w.hb, D, wat(d, d.qkv_w, (size_t)i * 3 * D * D), D, g, s));
    if (tp) {
      const float* q = reinterpret_cast<const float*>(w.qkv);
      F5B_CHECK(!(g_infer_attn_p > 0.f), "f5b_dit_forward: attention dropout is built for the bf16 operand mode only");
      F5B_TRY(attn_fwd_tf32(q, q + D, q + 2 * D, 3 * D, reinterpret_cast<float*>(w.ab), w.vt, lens, batch_mod, Bf, H, n, 0.125f, s));
    } else {
      const AttnDrop adrop = make_attn_drop(g_infer_attn_p, g_infer_attn_seed, i);  // off unless f5b_dit_set_attn_dropout asked for it
      F5B_TRY(attn_fwd(w.qkv, aat(d, w.qkv, D), aat(d, w.qkv, 2 * D), 3 * D, w.ab, nullptr, lens, batch_mod, Bf, H, n, 0.125f, s, &adrop,
                       g_infer_attn_seed_dev));
    }
    F5B_TRY(linear_gate_resid(w.ab, D, wat(d, d.out_w, (size_t)i * D * D), D, d.out_b + (size_t)i * D, w.x, D, rows, D, D, n, m + 2 * D,
                              mod_bstride, lens, batch_mod, s, tp));
    F5B_TRY(ln_modulate(w.x, m + 4 * D, m + 3 * D, mod_bstride, batch_mod, w.hb, rows, n, D, 1e-6f, s, tp));
    F5B_TRY(linear_bf16(w.hb, D, wat(d, d.ff1_w, (size_t)i * F * D), D, d.ff1_b + (size_t)i * F, w.fb, F, rows, F, D, F5B_ACT_GELU_TANH, s,
                        tp));
    F5B_TRY(linear_gate_resid(w.fb, F, wat(d, d.ff2_w, (size_t)i * D * F), F, d.ff2_b + (size_t)i * D, w.x, D, rows, D, F, n, m + 5 * D,
                              mod_bstride, nullptr, batch_mod, s, tp));
  }
  // AdaLayerNorm_Final chunk order (model/modules.py:333): scale, shift; then proj_out (dit.py:231)
  const float* mf = mod + (size_t)d.depth * 6 * D;
  F5B_TRY(ln_modulate(w.x, mf, mf + D, mod_bstride, batch_mod, w.hb, rows, n, D, 1e-6f, s, tp));
  F5B_TRY(linear_f32(w.hb, D, d.proj_w, D, d.proj_b, pred, d.mel_dim, rows, d.mel_dim, D, F5B_ACT_NONE, nullptr, 0, nullptr, 0, s, tp));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Vocos.decode (third-party `vocos`)
// ---------------------------------------------------------------------------------------------------------------
int f5b_vocos_create(const F5bVocosDesc* desc, F5bVocos** out) {
  F5B_CHECK(desc && out, "f5b_vocos_create: null argument");
  const F5bVocosDesc& d = *desc;
  F5B_CHECK(d.n_fft == 1024 && d.hop == 256, "f5b_vocos_create: only n_fft 1024 / hop 256 (vocos-mel-24khz) is built");
  F5B_CHECK(d.dim > 0 && d.dim % 8 == 0 && d.dim <= 1024 && d.intermediate % 8 == 0 && d.ld_embed >= 7 * d.n_mels && d.ld_embed % 8 == 0,
            "f5b_vocos_create: bad dims");
  F5B_CHECK(d.embed_w && d.dw_w && d.pw1_w && d.pw2_w && d.gamma && d.head_w, "f5b_vocos_create: null weight pointer");
  F5bVocos* h = new (std::nothrow) F5bVocos;
  F5B_CHECK(h != nullptr, "f5b_vocos_create: out of host memory");
  h->d = d;
  *out = h;
  return 0;
}
void f5b_vocos_destroy(F5bVocos* h) { delete h; }

struct VocosWs {
  float* x;
  __nv_bfloat16 *a, *hb, *ib;
  float *head, *frames;
  size_t bytes;
};
static VocosWs carve_vocos(const F5bVocosDesc& d, int B, int T, void* ws) {
  const size_t rows = (size_t)B * T;
  Carver c(ws);
  VocosWs w;
  w.x = c.take<float>(rows * d.dim);
  w.a = c.take<__nv_bfloat16>(rows * d.ld_embed);
  w.hb = c.take<__nv_bfloat16>(rows * d.dim);
  w.ib = c.take<__nv_bfloat16>(rows * d.intermediate);
  w.head = c.take<float>(rows * (d.n_fft + 2));
  w.frames = c.take<float>(rows * d.n_fft);
  w.bytes = c.off;
  return w;
}
size_t f5b_vocos_workspace_bytes(const F5bVocos* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return carve_vocos(h->d, B, T, nullptr).bytes;
}

int f5b_vocos_decode(const F5bVocos* h, const float* mel, int B, int T, float* wav, void* ws, size_t ws_bytes,
                     f5b_stream_t stream) {
  F5B_CHECK(h && mel && wav && ws && B > 0 && T > 1, "f5b_vocos_decode: bad argument");
  const F5bVocosDesc& d = h->d;
  VocosWs w = carve_vocos(d, B, T, ws);
  F5B_CHECK(w.bytes <= ws_bytes, "f5b_vocos_decode: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  cudaStream_t s = ST(stream);
  KindOverride as_vocos(K_VOCOS);
  const int rows = B * T, C = d.dim, I = d.intermediate;
  // backbone.embed Conv1d(n_mels, C, k=7, pad 3) as im2col + GEMM, then backbone.norm
  F5B_TRY(f5b_im2col7(mel, w.a, B, T, d.n_mels, d.ld_embed, stream));
  F5B_TRY(linear_f32(w.a, d.ld_embed, d.embed_w, d.ld_embed, d.embed_b, w.x, C, rows, C, 7 * d.n_mels, F5B_ACT_NONE, nullptr, 0,
                     nullptr, 0, s));
  F5B_TRY(ln_affine(w.x, d.norm_w, d.norm_b, w.x, nullptr, rows, C, 1e-6f, s));
  const __nv_bfloat16* pw1 = reinterpret_cast<const __nv_bfloat16*>(d.pw1_w);
  const __nv_bfloat16* pw2 = reinterpret_cast<const __nv_bfloat16*>(d.pw2_w);
  for (int i = 0; i < d.num_layers; ++i) {
    F5B_TRY(dwconv7_ln(w.x, d.dw_w + (size_t)i * C * 7, d.dw_b + (size_t)i * C, d.ln_w + (size_t)i * C, d.ln_b + (size_t)i * C, w.hb,
                       B, T, C, 1e-6f, s));
    F5B_TRY(linear_bf16(w.hb, C, pw1 + (size_t)i * I * C, C, d.pw1_b + (size_t)i * I, w.ib, I, rows, I, C, F5B_ACT_GELU_ERF, s));
    // x += gamma * (pwconv2(.) + b): the gate-residual epilogue with a batch-independent gate
    F5B_TRY(linear_gate_resid(w.ib, I, pw2 + (size_t)i * C * I, I, d.pw2_b + (size_t)i * C, w.x, C, rows, C, I, T,
                              d.gamma + (size_t)i * C, 0, nullptr, 0, s));
  }
  F5B_TRY(ln_affine(w.x, d.fln_w, d.fln_b, nullptr, w.hb, rows, C, 1e-6f, s));
  const int NO = d.n_fft + 2;
  F5B_TRY(linear_f32(w.hb, C, d.head_w, C, d.head_b, w.head, NO, rows, NO, C, F5B_ACT_NONE, nullptr, 0, nullptr, 0, s));
  F5B_TRY(f5b_istft_head(w.head, NO, w.frames, wav, B, T, stream));
  return 0;
}

}  // extern "C"
