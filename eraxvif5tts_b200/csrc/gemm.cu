// C[M,N] = epilogue(A[M,K] * W[N,K]^T): the one linear-layer engine of the DiT / text / vocoder stacks, a Problem
// policy on top of tile_engine.cuh (TMA -> 128B-swizzled smem ring -> tcgen05.mma into TMEM -> fused epilogue).
//
// Replaces, per reference call site (paths relative to /root/reference/src/f5_tts/): nn.Linear in to_q/to_k/to_v
// (model/modules.py:452-454, fused to one N=3D GEMM with the rotary embedding of :470-480 and the head split of
// :457-465 in the epilogue), to_out + masked_fill + gated residual (:495-501, :635), FeedForward (:348-353, GELU-tanh
// fused; gated residual :639), InputEmbedding.proj (model/backbones/dit.py:95), AdaLayerNorm linears (:311, :332),
// TimestepEmbedding MLP (:727-731), proj_out (dit.py:231), ConvNeXtV2 / Vocos point-wise linears.
#include "f5b_internal.h"
#include "tile_engine.cuh"

namespace f5b {

template <int ACT, bool PRECISE = false>
__device__ __forceinline__ float activate(float x) {
  if constexpr (ACT == F5B_ACT_GELU_TANH && PRECISE) {
    // the tf32 operand mode targets 1e-3 of the fp32 reference: libm tanh instead of the 2^-11 SFU approximation
    const float u = 0.7978845608028654f * x * fmaf(0.044715f * x, x, 1.0f);
    return 0.5f * x * (1.0f + tanhf(u));
  } else if constexpr (ACT == F5B_ACT_SILU && PRECISE) {
    return x / (1.0f + expf(-x));
  } else if constexpr (ACT == F5B_ACT_GELU_TANH) {
    // nn.GELU(approximate="tanh"), model/modules.py:625; tanh on the SFU (tanh.approx, rel. error ~2^-11 << bf16 output)
    const float u = 0.7978845608028654f * x * fmaf(0.044715f * x, x, 1.0f);
    return 0.5f * x * (1.0f + tanh_approx(u));
  } else if constexpr (ACT == F5B_ACT_GELU_ERF) {
    return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
  } else if constexpr (ACT == F5B_ACT_SILU) {
    return x / (1.0f + __expf(-x));
  } else {
    return x;
  }
}

// TF32_: the tf32 operand mode (F5bGemmArgs.tf32): A and W are fp32 words (rounded to tf32 by their producers), 32 per 128-byte
// stage row, kind::tf32 MMAs; the BF16 / QKV_ROPE epilogues then write fp32 (rounded to tf32: they feed the next tensor-core
// operand) through STORE_F32, and the F32 epilogue's optional second output is an fp32 tf32-rounded copy.
template <int BN_, int EPI, int ACT, bool TF32_ = false>
struct LinearProblem {
  static constexpr int BN = BN_;
  static constexpr bool TF32 = TF32_;
  static constexpr int BKE = TF32_ ? 32 : 64;  // elements per k-block (128 bytes)
  static constexpr int STORE = (EPI == F5B_EPI_BF16 || EPI == F5B_EPI_QKV_ROPE || EPI == F5B_EPI_BF16_DUAL)
                                   ? (TF32_ ? STORE_F32 : STORE_BF16)
                                   : (EPI == F5B_EPI_GATE_RESID ? STORE_F32ADD : STORE_DIRECT);
  // BF16_DUAL (training forward): out = bf16(acc + bias), out2 = bf16(act(out)) from one accumulator tile (tile_engine.cuh, P::DUAL)
  static constexpr bool DUAL = EPI == F5B_EPI_BF16_DUAL;
  static constexpr bool F32ADD_DIRECT = F5B_F32ADD_DIRECT != 0 && EPI == F5B_EPI_GATE_RESID && !TF32_;
  static_assert(!(DUAL && TF32_), "the dual-output epilogue exists in the bf16 operand mode only");
  static constexpr int CLUSTER = 2;  // CTA pairs on vertically adjacent tiles share the weight tile through TMA multicast
  F5bGemmArgs g;
  int n_tiles, m_tiles, kblocks;

  struct RowCtx {
    int row, b, pos, n_base;
    bool valid;
    const float* gate;
  };

  // work unit = (pair of vertically adjacent m-blocks, n-block); CTA `rank` of the cluster takes m-block 2*mp + rank
  __device__ __forceinline__ int num_units() const { return n_tiles * ((m_tiles + 1) / 2); }
  __device__ __forceinline__ int unit_tile(int unit, uint32_t rank) const {
    const int mp = unit / n_tiles, n_blk = unit - mp * n_tiles;
    return (2 * mp + (int)rank) * n_tiles + n_blk;  // may address an m-block past the end: loads zero-fill, stores clip
  }
  __device__ __forceinline__ int num_kblocks() const { return kblocks; }
  __device__ __forceinline__ uint32_t umma_n() const { return BN; }
  __device__ __forceinline__ uint32_t idesc() const { return TF32 ? idesc_tf32(BM, umma_n(), 0, 0) : idesc_bf16(BM, umma_n(), 0, 0); }
  __device__ __forceinline__ uint32_t idesc2() const {  // CTA pair: 256 rows
    return TF32 ? idesc_tf32(2 * BM, umma_n(), 0, 0) : idesc_bf16(2 * BM, umma_n(), 0, 0);
  }
  __device__ __forceinline__ uint64_t a_desc(uint32_t addr, int k) const { return desc_kmajor(addr, k); }
  __device__ __forceinline__ uint64_t b_desc(uint32_t addr, int k) const { return desc_kmajor(addr, k); }
  __device__ __forceinline__ uint32_t b_tx_bytes() const { return EngCfg<BN>::B_BYTES; }
  __device__ __forceinline__ int tile_cols(int tile) const {
    const int left = g.N - (tile % n_tiles) * BN;
    return left < BN ? left : BN;
  }
  __device__ __forceinline__ int out_col0(int tile) const { return (tile % n_tiles) * BN; }
  __device__ __forceinline__ int out_row0(int tile) const { return (tile / n_tiles) * BM; }
  __device__ __forceinline__ void load(int tile, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar, const CUtensorMap* tmA,
                                       const CUtensorMap* tmB, uint32_t rank) const {
    const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
    tma_load_2d(sA, tmA, bar, kb * BKE, m_blk * BM);
    // this CTA fetches rows [rank*BN/2, +BN/2) of the weight tile and multicasts them to both CTAs of the pair
    tma_load_2d_mcast(sB + rank * (EngCfg<BN>::B_BYTES / 2), tmB, bar, kb * BKE, n_blk * BN + (int)rank * (BN / 2), (uint16_t)0x3);
  }
  // CTA-pair mode: this CTA stages its own A rows and rows [rank*BN/2, +BN/2) of the weight tile; both complete on the leader's barrier
  __device__ __forceinline__ void load2(int tile, int kb, uint8_t* sA, uint8_t* sB, uint32_t leader_bar, const CUtensorMap* tmA,
                                        const CUtensorMap* tmB, uint32_t rank) const {
    const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
    tma_load_2d_2sm(sA, tmA, leader_bar, kb * BKE, m_blk * BM);
    tma_load_2d_2sm(sB, tmB, leader_bar, kb * BKE, n_blk * BN + (int)rank * (BN / 2));
  }
  __device__ __forceinline__ RowCtx row_ctx(int tile, int r) const {
    RowCtx c;
    const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
    c.row = m_blk * BM + r;
    c.n_base = n_blk * BN;
    c.valid = c.row < g.M;
    c.b = 0;
    c.pos = c.row;
    c.gate = nullptr;
    if constexpr (EPI == F5B_EPI_QKV_ROPE || EPI == F5B_EPI_GATE_RESID) {
      c.b = c.row / g.rows_per_batch;
      c.pos = c.row - c.b * g.rows_per_batch;
    }
    if constexpr (EPI == F5B_EPI_GATE_RESID) {
      const int bb = g.batch_mod > 0 ? c.b % g.batch_mod : c.b;
      if (c.valid && g.lens != nullptr) c.valid = c.pos < __ldg(g.lens + bb);
      if (g.gate != nullptr) c.gate = g.gate + (size_t)bb * g.gate_bstride;
    }
    return c;
  }

  // STORE_BF16 / STORE_F32ADD: final values of 32 consecutive columns (the engine stages + TMA-stores them)
  __device__ __forceinline__ void compute(const RowCtx& c, int c0, const uint32_t (&r)[32], float (&v)[32]) const {
    const int n0 = c.n_base + c0;
    const int left = g.N - n0;  // > 0
    float b[32];
    load_bias32(g.bias, n0, left, b);
    if constexpr (EPI == F5B_EPI_GATE_RESID) {
      // x += gate[b] * (acc + bias); rows at or beyond lens[b] contribute 0 (masked_fill, model/modules.py:499-501)
      if (!c.valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        return;
      }
      if (c.gate != nullptr) {
        float gt[32];
        load_bias32(c.gate, n0, left, gt);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gt[i] * (__uint_as_float(r[i]) + b[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + b[i];
      }
    } else if constexpr (DUAL) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + b[i];  // the activation is applied by second()
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = activate<ACT, TF32>(__uint_as_float(r[i]) + b[i]);
      if constexpr (EPI == F5B_EPI_QKV_ROPE) {
        // x_transformers.apply_rotary_pos_emb on the first rope_heads heads of q and k (model/modules.py:470-480): interleaved
        // pairs (2i, 2i+1), angle pos * 10000^(-2i/64), fp32 math.  Output stays token-major [rows, 3D] (q | k | v); the
        // attention kernel addresses heads through strided TMA boxes, so no head-split scatter exists.
        const int Dm = g.heads * 64;
        const int sec = n0 / Dm;
        const int within = n0 - sec * Dm;
        if (sec < 2 && (within >> 6) < g.rope_heads) {
          const float4* cs = reinterpret_cast<const float4*>(g.rope) + (size_t)c.pos * 16 + ((within & 63) >> 2);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(cs + i);  // (cos, sin) of two consecutive frequency pairs
            const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
            v[4 * i] = x0 * t.x - x1 * t.y;
            v[4 * i + 1] = x1 * t.x + x0 * t.y;
            v[4 * i + 2] = x2 * t.z - x3 * t.w;
            v[4 * i + 3] = x3 * t.z + x2 * t.w;
          }
        }
      }
      if constexpr (TF32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
      }
    }
  }

  // F32ADD_DIRECT: address of 32 consecutive residual-stream columns of this thread's row (nullptr: nothing to add)
  __device__ __forceinline__ float* f32add_row(const RowCtx& c, int, int c0) const {
    return c.valid ? reinterpret_cast<float*>(g.out) + (size_t)c.row * g.ldc + c.n_base + c0 : nullptr;
  }
  // DUAL: the second output as a function of the (bf16-rounded) first one
  __device__ __forceinline__ float second(float x) const { return activate<ACT, false>(x); }

  // STORE_DIRECT epilogues
  __device__ __forceinline__ void epilogue(const RowCtx& c, int c0, const uint32_t (&r)[32]) const {
    if (!c.valid) return;
    const int n0 = c.n_base + c0;
    const int left = g.N - n0;  // > 0
    float v[32];
    {
      float b[32];
      load_bias32(g.bias, n0, left, b);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = activate<ACT, TF32>(__uint_as_float(r[i]) + b[i]);
    }
    if constexpr (EPI == F5B_EPI_F32) {
      if (g.addsrc != nullptr) {
        const float* a = g.addsrc + (size_t)c.row * g.ld_add + n0;
        if (left >= 32 && (g.ld_add & 3) == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(a) + q);
            v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < left) v[i] += __ldg(a + i);
        }
      }
      float* o = reinterpret_cast<float*>(g.out) + (size_t)c.row * g.ldc + n0;
      store_row32_f32(o, v, left, (g.ldc & 3) == 0);
      if (g.out2 != nullptr) {
        if constexpr (TF32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
          store_row32_f32(reinterpret_cast<float*>(g.out2) + (size_t)c.row * g.ldc2 + n0, v, left, (g.ldc2 & 3) == 0);
        } else {
          __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(g.out2) + (size_t)c.row * g.ldc2 + n0;
          store_row32_bf16(o2, v, left, (g.ldc2 & 7) == 0);
        }
      }
    }
  }
};

int g_gemm_pair_mode = 1;  // 1 (default): 256-wide tiles run as tcgen05 CTA pairs (cta_group::2); 0: two cta_group::1 tiles + weight multicast

template <int BN, int EPI, int ACT, bool TF32>
static int launch_linear(const CUtensorMap& tmA, const CUtensorMap& tmB, const F5bGemmArgs& g, cudaStream_t stream) {
  using P = LinearProblem<BN, EPI, ACT, TF32>;
  P p;
  p.g = g;
  p.m_tiles = (g.M + BM - 1) / BM;
  p.n_tiles = (g.N + BN - 1) / BN;
  p.kblocks = (g.K + P::BKE - 1) / P::BKE;
  CUtensorMap tmC = tmA;
  if constexpr (P::STORE == STORE_BF16) {
    F5B_CHECK((g.ldc & 7) == 0, "f5b_gemm: bf16 output pitch %d must be a multiple of 8", g.ldc);
    if (make_tmap_2d(&tmC, g.out, 2, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc * 2, 64, 32, true)) return -1;
  } else if constexpr (P::STORE == STORE_F32ADD || P::STORE == STORE_F32) {
    F5B_CHECK((g.ldc & 3) == 0, "f5b_gemm: f32 output pitch %d must be a multiple of 4", g.ldc);
    if constexpr (P::F32ADD_DIRECT) F5B_CHECK((g.N & 31) == 0, "f5b_gemm: the direct-reduction epilogue needs N %% 32 == 0 (N %d)", g.N);
    if (make_tmap_2d(&tmC, g.out, 4, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc * 4, 32, 32, true)) return -1;
  }
  CUtensorMap tmD = tmC;
  if constexpr (P::DUAL) {
    F5B_CHECK(g.out2 != nullptr && (g.ldc2 & 7) == 0, "f5b_gemm: BF16_DUAL needs out2 with a pitch that is a multiple of 8");
    if (make_tmap_2d(&tmD, g.out2, 2, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc2 * 2, 64, 32, true)) return -1;
  }
  if constexpr (BN == 256) {
    if (g_gemm_pair_mode) return launch_engine<P, true>(tmA, tmB, tmC, p, p.n_tiles * ((p.m_tiles + 1) / 2), stream, &tmD);
  }
  return launch_engine(tmA, tmB, tmC, p, p.n_tiles * ((p.m_tiles + 1) / 2), stream, &tmD);
}

template <int BN, bool TF32>
static int dispatch(const CUtensorMap& a, const CUtensorMap& b, const F5bGemmArgs& g, cudaStream_t s) {
#define F5B_CASE(E, A) \
  if (g.epi == E && g.act == A) return launch_linear<BN, E, A, TF32>(a, b, g, s);
  F5B_CASE(F5B_EPI_BF16, F5B_ACT_NONE)
  F5B_CASE(F5B_EPI_BF16, F5B_ACT_GELU_TANH)
  F5B_CASE(F5B_EPI_BF16, F5B_ACT_GELU_ERF)
  F5B_CASE(F5B_EPI_BF16, F5B_ACT_SILU)
  F5B_CASE(F5B_EPI_F32, F5B_ACT_NONE)
  F5B_CASE(F5B_EPI_QKV_ROPE, F5B_ACT_NONE)
  F5B_CASE(F5B_EPI_GATE_RESID, F5B_ACT_NONE)
  if constexpr (!TF32) {
    F5B_CASE(F5B_EPI_BF16_DUAL, F5B_ACT_GELU_TANH)
  }
#undef F5B_CASE
  set_error("f5b_gemm: unsupported epilogue/activation combination (%d, %d)", g.epi, g.act);
  return -1;
}

int gemm(const void* A, int lda, const void* W, int ldw, const F5bGemmArgs& g, cudaStream_t stream) {
  F5B_CHECK(A != nullptr && W != nullptr, "f5b_gemm: null operand");
  F5B_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "f5b_gemm: empty problem %d x %d x %d", g.M, g.N, g.K);
  F5B_CHECK(g.out != nullptr, "f5b_gemm: null output");
  const bool tf32 = g.tf32 != 0;
  const int pitch_mask = tf32 ? 3 : 7;  // 16-byte rows for TMA
  F5B_CHECK(lda >= g.K && ldw >= g.K && (lda & pitch_mask) == 0 && (ldw & pitch_mask) == 0,
            "f5b_gemm: row pitches must be >= K and multiples of %d elements (lda %d ldw %d K %d)", pitch_mask + 1, lda, ldw, g.K);
  if (g.epi == F5B_EPI_QKV_ROPE) {
    F5B_CHECK(g.rope && g.rows_per_batch > 0 && g.heads > 0 && g.N == 3 * g.heads * 64 && g.M % g.rows_per_batch == 0,
              "f5b_gemm: bad QKV_ROPE arguments (N %d heads %d n %d)", g.N, g.heads, g.rows_per_batch);
  }
  if (g.epi == F5B_EPI_GATE_RESID)
    F5B_CHECK(g.rows_per_batch > 0 && (g.gate_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(g.gate) & 15) == 0,
              "f5b_gemm: GATE_RESID needs rows_per_batch > 0 and a 16-byte aligned gate with a stride that is a multiple of 4");
  if (g.bias != nullptr) F5B_CHECK((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "f5b_gemm: bias must be 16-byte aligned");
  const double esz = tf32 ? 4.0 : 2.0;
  const double out_bytes = (g.epi == F5B_EPI_BF16 || g.epi == F5B_EPI_QKV_ROPE) ? esz
                           : (g.epi == F5B_EPI_BF16_DUAL ? 2.0 * esz : (g.epi == F5B_EPI_GATE_RESID ? 8.0 : 4.0));
  LaunchScope scope(K_GEMM, stream, 2.0 * g.M * g.N * g.K, esz * ((double)g.M * g.K + (double)g.N * g.K) + out_bytes * g.M * g.N);
  const int m_tiles = (g.M + BM - 1) / BM;
  // 256-wide tiles (CTA pairs) once they fill about two thirds of the machine; below that 128-wide tiles give twice the units to
  // spread.  Measured on the B = 1 shapes (M = 1880): threshold 148 tiles 75.4 ms per utterance, 98 -> 72.8 ms (FF1, 120 tiles, moves to
  // pairs), 60 -> 76.5 ms (out-proj / FF2, 60 tiles, are better off with 128-wide tiles).  F5B_GEMM_BN256_MIN_TILES overrides it.
  // (A wave-quantisation rule — units / clusters rounded up, a 128-wide tile costed at 0.6 of a pair tile — differs from this one for
  // QKV at M = 1880 only, 96 pair units = 2 waves against 192 narrow units = 3 waves, and measured WORSE: 67.3 vs 66.6 ms per utterance.)
  static const int min_tiles256 = [] {
    const char* e = getenv("F5B_GEMM_BN256_MIN_TILES");
    return e ? atoi(e) : sm_count() * 2 / 3;
  }();
  const int bn = (g.N % 256 == 0 && m_tiles * (g.N / 256) >= min_tiles256) ? 256 : 128;
  CUtensorMap tmA, tmB;
  if (tf32) {
    if (make_tmap_2d(&tmA, A, 4, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)lda * 4, 32, BM, true)) return -1;
    if (make_tmap_2d(&tmB, W, 4, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)ldw * 4, 32, bn / 2, true)) return -1;
    return bn == 256 ? dispatch<256, true>(tmA, tmB, g, stream) : dispatch<128, true>(tmA, tmB, g, stream);
  }
  if (make_tmap_2d(&tmA, A, 2, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)lda * 2, BK, BM, true)) return -1;
  if (make_tmap_2d(&tmB, W, 2, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)ldw * 2, BK, bn / 2, true)) return -1;  // half tile per CTA
  return bn == 256 ? dispatch<256, false>(tmA, tmB, g, stream) : dispatch<128, false>(tmA, tmB, g, stream);
}

static F5bGemmArgs base_args(int M, int N, int K, int epi, int act, const float* bias, void* out, int ldc) {
  F5bGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.epi = epi; g.act = act; g.bias = bias; g.out = out; g.ldc = ldc;
  return g;
}

int linear_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M, int N, int K,
                int act, cudaStream_t s, bool tf32) {
  F5bGemmArgs g = base_args(M, N, K, F5B_EPI_BF16, act, bias, out, ldc);
  g.tf32 = tf32;
  return gemm(A, lda, W, ldw, g, s);
}

int linear_bf16_dual(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, void* out_act, int ldc_act,
                     int M, int N, int K, int act, cudaStream_t s) {
  F5bGemmArgs g = base_args(M, N, K, F5B_EPI_BF16_DUAL, act, bias, out, ldc);
  g.out2 = out_act; g.ldc2 = ldc_act;
  return gemm(A, lda, W, ldw, g, s);
}

int linear_f32(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldc, int M, int N, int K,
               int act, const float* addsrc, int ld_add, void* out_bf16, int ld_bf16, cudaStream_t s, bool tf32) {
  F5bGemmArgs g = base_args(M, N, K, F5B_EPI_F32, act, bias, out, ldc);
  g.addsrc = addsrc; g.ld_add = ld_add; g.out2 = out_bf16; g.ldc2 = ld_bf16;
  g.tf32 = tf32;
  return gemm(A, lda, W, ldw, g, s);
}

int linear_gate_resid(const void* A, int lda, const void* W, int ldw, const float* bias, float* x, int ldc, int M, int N,
                      int K, int rows_per_batch, const float* gate, int64_t gate_bstride, const int32_t* lens,
                      int batch_mod, cudaStream_t s, bool tf32) {
  F5bGemmArgs g = base_args(M, N, K, F5B_EPI_GATE_RESID, F5B_ACT_NONE, bias, x, ldc);
  g.rows_per_batch = rows_per_batch; g.gate = gate; g.gate_bstride = gate_bstride; g.lens = lens; g.batch_mod = batch_mod;
  g.tf32 = tf32;
  return gemm(A, lda, W, ldw, g, s);
}

}  // namespace f5b

extern "C" int f5b_gemm(const void* A, int lda, const void* W, int ldw, const F5bGemmArgs* args, f5b_stream_t stream) {
  if (args == nullptr) {
    f5b::set_error("f5b_gemm: null args");
    return -1;
  }
  return f5b::gemm(A, lda, W, ldw, *args, static_cast<cudaStream_t>(stream));
}

extern "C" void f5b_debug_gemm_pair_mode(int v) { f5b::g_gemm_pair_mode = v; }
