// Transposed-operand GEMMs for the backward pass of every nn.Linear on the path — groundwork of the training step
// (/root/reference/src/f5_tts/model/trainer.py:1280 accelerator.backward -> autograd of the Linears in model/modules.py).
//   C[m, n] (+)= sum_k A(m, k) * B(n, k),   A(m,k) = A[m*lda + k]  (K-major)  or  A[k*lda + m]  (MN-major), same for B.
//   dgrad  dX[M,K] = dY[M,N] W[N,K]   : A = dY K-major,            B = W  MN-major (reduction dim N is W's row index)
//   wgrad  dW[N,K] = dY[M,N]^T X[M,K] : A = dY MN-major (rows = m), B = X  MN-major, reduction over the M tokens, split into
//                                       `splits` K-ranges whose partial tiles are TMA-reduce-added into the fp32 gradient.
// No transposed copy of any activation or weight is made: MN-major operands are TMA-loaded as [64 reduction rows x 64] boxes
// from the row-major source and described to tcgen05.mma with the MN-major descriptor form (tile_engine.cuh).
#include "f5b_internal.h"
#include "tile_engine.cuh"

namespace f5b {

struct GradArgs {
  int M, N, K;          // output M x N, reduction K
  int m_tiles, n_tiles, kb_per_split, splits;
  float* bias_unused;
  // implicit grouped-conv weight gradient (conv_cpb > 0; A and B MN-major, 3-D tensor maps [channel, position, batch row]):
  // the reduction runs over (batch row, 64-position chunk), conv_cpb chunks per batch row, and 64-column chunk c of the B tile is the
  // layer input shifted by tap (n_blk * BN / 64 + c) - conv_pad positions — im2col by TMA coordinates, zero fill = the conv's padding
  int conv_cpb, conv_pad;
};

// PAIR: the unit is a vertical PAIR of m-blocks run by a 2-CTA cluster as one tcgen05 CTA pair (cta_group::2, 256 x BN tile):
// g.m_tiles is then padded to even (the block past the end loads zeros and its stores are clipped).
template <int BN_, bool A_MN, bool B_MN, int STORE_, bool PAIR = false>
struct GradProblem {
  static constexpr int BN = BN_;
  static constexpr int STORE = STORE_;   // STORE_BF16 or STORE_F32ADD
  static constexpr int CLUSTER = PAIR ? 2 : 1;
  static constexpr bool TF32 = false;
  GradArgs g;

  struct RowCtx {
    int row, n_base;
    bool valid;
  };
  __device__ __forceinline__ int num_units() const { return (PAIR ? g.m_tiles / 2 : g.m_tiles) * g.n_tiles * g.splits; }
  __device__ __forceinline__ int unit_tile(int unit, uint32_t rank) const {
    if constexpr (!PAIR) return unit;
    const int n_blk = unit % g.n_tiles, t = unit / g.n_tiles;
    const int mp = t % (g.m_tiles / 2), split = t / (g.m_tiles / 2);
    return (split * g.m_tiles + 2 * mp + (int)rank) * g.n_tiles + n_blk;
  }
  __device__ __forceinline__ uint32_t idesc2() const { return idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0); }
  // pair mode: own A rows + this CTA's half of the B tile, both completing on the leader CTA's barrier
  __device__ __forceinline__ void load2(int unit, int kb, uint8_t* sA, uint8_t* sB, uint32_t leader_bar, const CUtensorMap* tmA,
                                        const CUtensorMap* tmB, uint32_t rank) const {
    int split, m_blk, n_blk;
    decode(unit, split, m_blk, n_blk);
    const int k0 = (split * g.kb_per_split + kb) * BK;
    if constexpr (A_MN) {
#pragma unroll
      for (int c = 0; c < BM / 64; ++c) tma_load_2d_2sm(sA + c * 8192, tmA, leader_bar, m_blk * BM + c * 64, k0);
    } else {
      tma_load_2d_2sm(sA, tmA, leader_bar, k0, m_blk * BM);
    }
    const int n0 = n_blk * BN + (int)rank * (BN / 2);
    if constexpr (B_MN) {
#pragma unroll
      for (int c = 0; c < BN / 128; ++c) tma_load_2d_2sm(sB + c * 8192, tmB, leader_bar, n0 + c * 64, k0);
    } else {
      tma_load_2d_2sm(sB, tmB, leader_bar, k0, n0);  // tensor map box: BN/2 rows
    }
  }
  __device__ __forceinline__ int num_kblocks() const { return g.kb_per_split; }
  __device__ __forceinline__ uint32_t umma_n() const { return BN; }
  __device__ __forceinline__ uint32_t b_tx_bytes() const { return EngCfg<BN>::B_BYTES; }
  __device__ __forceinline__ uint32_t idesc() const { return idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0); }
  __device__ __forceinline__ uint64_t a_desc(uint32_t addr, int k) const { return A_MN ? desc_mnmajor(addr, k) : desc_kmajor(addr, k); }
  __device__ __forceinline__ uint64_t b_desc(uint32_t addr, int k) const { return B_MN ? desc_mnmajor(addr, k) : desc_kmajor(addr, k); }
  __device__ __forceinline__ void decode(int unit, int& split, int& m_blk, int& n_blk) const {
    n_blk = unit % g.n_tiles;
    const int t = unit / g.n_tiles;
    m_blk = t % g.m_tiles;
    split = t / g.m_tiles;
  }
  __device__ __forceinline__ int tile_cols(int unit) const {
    const int left = g.N - (unit % g.n_tiles) * BN;
    return left < BN ? left : BN;
  }
  __device__ __forceinline__ int out_col0(int unit) const { return (unit % g.n_tiles) * BN; }
  __device__ __forceinline__ int out_row0(int unit) const { return ((unit / g.n_tiles) % g.m_tiles) * BM; }
  __device__ __forceinline__ void load(int unit, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar, const CUtensorMap* tmA,
                                       const CUtensorMap* tmB, uint32_t) const {
    int split, m_blk, n_blk;
    decode(unit, split, m_blk, n_blk);
    if constexpr (A_MN && B_MN) {
      if (g.conv_cpb > 0) {
        const int kbg = split * g.kb_per_split + kb;
        const int b = kbg / g.conv_cpb, p0 = (kbg - b * g.conv_cpb) * BK;  // b past the last batch row: zero fill
#pragma unroll
        for (int c = 0; c < BM / 64; ++c) tma_load_3d(sA + c * 8192, tmA, bar, m_blk * BM + c * 64, p0, b);
#pragma unroll
        for (int c = 0; c < BN / 64; ++c) tma_load_3d(sB + c * 8192, tmB, bar, 0, p0 + n_blk * (BN / 64) + c - g.conv_pad, b);
        return;
      }
    }
    const int k0 = (split * g.kb_per_split + kb) * BK;  // past-the-end K ranges are zero-filled by TMA
    if constexpr (A_MN) {
#pragma unroll
      for (int c = 0; c < BM / 64; ++c) tma_load_2d(sA + c * 8192, tmA, bar, m_blk * BM + c * 64, k0);
    } else {
      tma_load_2d(sA, tmA, bar, k0, m_blk * BM);
    }
    if constexpr (B_MN) {
#pragma unroll
      for (int c = 0; c < BN / 64; ++c) tma_load_2d(sB + c * 8192, tmB, bar, n_blk * BN + c * 64, k0);
    } else {
      tma_load_2d(sB, tmB, bar, k0, n_blk * BN);
    }
  }
  __device__ __forceinline__ RowCtx row_ctx(int unit, int r) const {
    RowCtx c;
    c.row = out_row0(unit) + r;
    c.n_base = out_col0(unit);
    c.valid = c.row < g.M;
    return c;
  }
  __device__ __forceinline__ void compute(const RowCtx& c, int c0, const uint32_t (&r)[32], float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = c.valid ? __uint_as_float(r[i]) : 0.f;
  }
  __device__ __forceinline__ void epilogue(const RowCtx&, int, const uint32_t (&)[32]) const {}
};

template <int BN, bool A_MN, bool B_MN, int STORE, bool PAIR = false>
static int launch_grad(const void* A, int lda, const void* B, int ldb, void* out, int ldc, int M, int N, int K, int splits,
                       cudaStream_t stream) {
  using P = GradProblem<BN, A_MN, B_MN, STORE, PAIR>;
  P p;
  p.g.M = M; p.g.N = N; p.g.K = K;
  p.g.m_tiles = (M + BM - 1) / BM;
  if (PAIR) p.g.m_tiles = (p.g.m_tiles + 1) / 2 * 2;
  p.g.n_tiles = (N + BN - 1) / BN;
  const int kblocks = (K + BK - 1) / BK;
  p.g.splits = splits;
  p.g.kb_per_split = (kblocks + splits - 1) / splits;
  p.g.conv_cpb = 0;
  p.g.conv_pad = 0;
  CUtensorMap tmA, tmB, tmC;
  if (A_MN) {
    if (make_tmap_2d(&tmA, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, 64, true)) return -1;
  } else {
    if (make_tmap_2d(&tmA, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM, true)) return -1;
  }
  if (B_MN) {
    if (make_tmap_2d(&tmB, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, 64, true)) return -1;
  } else {
    if (make_tmap_2d(&tmB, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, PAIR ? BN / 2 : BN, true)) return -1;
  }
  if (STORE == STORE_BF16) {
    if (make_tmap_2d(&tmC, out, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 2, 64, 32, true)) return -1;
  } else {
    if (make_tmap_2d(&tmC, out, 4, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 4, 32, 32, true)) return -1;
  }
  const int units = (PAIR ? p.g.m_tiles / 2 : p.g.m_tiles) * p.g.n_tiles * splits;
  return launch_engine<P, PAIR>(tmA, tmB, tmC, p, units, stream);
}

// Weight gradient of one group of a grouped Conv1d (ConvPositionEmbedding, model/modules.py:167-190) WITHOUT an im2col buffer:
//   dW_g[co, k * cpg + ci] += sum_{b, t} dY[b, t, co] * X[b, t + k - pad, ci]        (cpg = 64 channels per group)
// = the MN-major x MN-major split-K GEMM above with M = cpg, N = ks * 64, the reduction over (b, t), where the 64-column chunk of
// tap k of the B operand is simply the TMA box of X at position offset k - pad (3-D maps clip / zero-fill per batch row).  The
// materialised k-major im2col this replaces wrote and re-read 152 MB per group (4.9 GB per step at cfg-5).
int conv_wgrad_implicit(const void* dy_g, const void* x_g, int ld, float* out, int ldc, int B, int n, int cpg, int ks, int splits,
                        cudaStream_t stream) {
  F5B_CHECK(cpg == 64 && (ld & 7) == 0 && (ldc & 3) == 0, "conv_wgrad_implicit: 64 channels per group expected (cpg %d)", cpg);
  using P = GradProblem<256, true, true, STORE_F32ADD, false>;
  P p;
  const int N = ks * 64;
  p.g.M = cpg; p.g.N = N;
  p.g.conv_cpb = (n + BK - 1) / BK;
  p.g.conv_pad = ks / 2;
  const int kblocks = B * p.g.conv_cpb;
  p.g.K = kblocks * BK;
  p.g.m_tiles = 1;
  p.g.n_tiles = (N + 255) / 256;
  if (splits > kblocks) splits = kblocks;
  p.g.splits = splits;
  p.g.kb_per_split = (kblocks + splits - 1) / splits;
  LaunchScope scope(K_GEMM, stream, 2.0 * cpg * N * (double)B * n, 2.0 * 2.0 * (double)B * n * cpg + 8.0 * cpg * N);
  CUtensorMap tmA, tmB, tmC;
  const uint64_t pitch = (uint64_t)ld * 2;
  if (make_tmap_3d(&tmA, dy_g, 2, (uint64_t)cpg, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, 64, 1, true)) return -1;
  if (make_tmap_3d(&tmB, x_g, 2, (uint64_t)cpg, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, 64, 1, true)) return -1;
  if (make_tmap_2d(&tmC, out, 4, (uint64_t)N, (uint64_t)cpg, (uint64_t)ldc * 4, 32, 32, true)) return -1;
  return launch_engine<P, false>(tmA, tmB, tmC, p, p.g.n_tiles * splits, stream);
}

int g_grad_pair_mode = 1;  // 1 (default): 256-wide tiles with at least two m-blocks run as tcgen05 CTA pairs

}  // namespace f5b

using namespace f5b;

extern "C" int f5b_gemm_tn(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* out, int ldc,
                           int out_f32_accumulate, int M, int N, int K, int splits, f5b_stream_t stream) {
  F5B_CHECK(A && B && out && M > 0 && N > 0 && K > 0, "f5b_gemm_tn: bad argument");
  F5B_CHECK((lda & 7) == 0 && (ldb & 7) == 0, "f5b_gemm_tn: operand pitches must be multiples of 8 elements");
  F5B_CHECK(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K), "f5b_gemm_tn: pitch smaller than the row length");
  F5B_CHECK(splits >= 1 && (splits == 1 || out_f32_accumulate), "f5b_gemm_tn: split-K needs the f32 accumulate output");
  F5B_CHECK(out_f32_accumulate ? (ldc & 3) == 0 : (ldc & 7) == 0, "f5b_gemm_tn: bad output pitch %d", ldc);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LaunchScope scope(K_GEMM, s, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K) + (out_f32_accumulate ? 8.0 : 2.0) * M * N);
#define F5B_GRAD_CASE(AM, BMj)                                                                                            \
  if ((a_mn_major != 0) == AM && (b_mn_major != 0) == BMj) {                                                               \
    if (N >= 256 && M > 128 && g_grad_pair_mode)  /* wide tiles as CTA pairs */                                             \
      return out_f32_accumulate ? launch_grad<256, AM, BMj, STORE_F32ADD, true>(A, lda, B, ldb, out, ldc, M, N, K, splits, s)  \
                                : launch_grad<256, AM, BMj, STORE_BF16, true>(A, lda, B, ldb, out, ldc, M, N, K, splits, s);   \
    if (N >= 256)  /* wide tiles: half the operand traffic per flop */                                                     \
      return out_f32_accumulate ? launch_grad<256, AM, BMj, STORE_F32ADD>(A, lda, B, ldb, out, ldc, M, N, K, splits, s)  \
                                : launch_grad<256, AM, BMj, STORE_BF16>(A, lda, B, ldb, out, ldc, M, N, K, splits, s);   \
    return out_f32_accumulate ? launch_grad<128, AM, BMj, STORE_F32ADD>(A, lda, B, ldb, out, ldc, M, N, K, splits, s)    \
                              : launch_grad<128, AM, BMj, STORE_BF16>(A, lda, B, ldb, out, ldc, M, N, K, splits, s);     \
  }
  F5B_GRAD_CASE(false, false)
  F5B_GRAD_CASE(false, true)
  F5B_GRAD_CASE(true, false)
  F5B_GRAD_CASE(true, true)
#undef F5B_GRAD_CASE
  return -1;
}

extern "C" void f5b_debug_grad_pair_mode(int v) { f5b::g_grad_pair_mode = v; }
