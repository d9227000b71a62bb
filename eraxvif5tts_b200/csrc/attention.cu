// Flash-style attention forward for the DiT blocks: O = softmax(Q K^T / sqrt(64), keys < len[b]) V, non-causal, d_head 64.
// Replaces AttnProcessor's mask expansion + F.scaled_dot_product_attention + head merge
// (/root/reference/src/f5_tts/model/modules.py:483-493; dropout_p = 0, see DESIGN.md "oracle adjustments").
//
// sm_100a design (one CTA = one 128-query tile of one (batch, head); the kernel is latency-bound per CTA — measured: one
// resident CTA/SM 320 TFLOP/s, two 577 — so it is built for THREE co-resident CTAs: 40 KB smem (72 KB on the shared-memory P path),
// 128 + 32 TMEM columns, <= 128 regs):
//   warp 4 (one elected lane): TMA producer + tcgen05.mma issuer.  KV is consumed in tiles of 64 keys:
//       S_j[128 x 64] = Q K_j^T    (K-major bf16 tiles, 128B swizzle; fp32 accumulator in TMEM columns 0..63).  The softmax
//                                   threads pull S_j into registers and release the buffer at once (bar_sfree), so S_{j+1} is
//                                   computed while the exponentials of tile j run.
//       O[128 x 64]  += P_j V_j    (default, PT: P_j is written by the softmax warps into the CTA's own 32 tensor-memory columns — a
//                                   second tcgen05.alloc — and read as a TMEM A operand; the dropout / split-KV kernels keep the
//                                   round-1 path, P_j double-buffered in swizzled smem.  V_j [64 keys x 64] is fed as an MN-MAJOR B
//                                   operand straight from the head-major layout the QKV epilogue writes — no transposed copy of V
//                                   exists; O stays resident in TMEM columns 64..127)
//   warps 0..3: online softmax, thread = query row (tcgen05.ld 32x32b -> no cross-lane reductions), exp2 with the
//       1/sqrt(d)*log2(e) scale folded in, row sum in a register.  The running maximum is updated lazily: O and the row sum
//       are rescaled only when a row's maximum grew by more than 2^8, and the exponentials are issued speculatively
//       before the tile's maximum is known (the rare miss rescales O in TMEM and recomputes the tile's P).
//   K lives in a 2-stage TMA ring, V in one stage (its reload hides behind the next tile's softmax).  Key-padding is a per-batch
//   length bound: KV tiles past len[b] are never loaded, the last tile is masked by index.
#include "common.cuh"
#include "dropout.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
constexpr int ATT_THREADS = 160;
#ifndef ATT_CTAS
#define ATT_CTAS 3
#endif
constexpr int ATT_CTAS_PER_SM = ATT_CTAS;
constexpr uint32_t ATT_Q_BYTES = ATT_BQ * 64 * 2;          // 16 KB
constexpr uint32_t ATT_K_BYTES = ATT_BKV * 64 * 2;         // 8 KB per stage, 2 stages
constexpr uint32_t ATT_V_BYTES = ATT_BKV * 64 * 2;         // 8 KB, 1 stage: [64 kv rows x 64 d]
constexpr uint32_t ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;     // 16 KB per buffer, 2 buffers
#ifndef ATT_P_TMEM
#define ATT_P_TMEM 1  // default kernel: 1 = P_j goes to the CTA's OWN 32 tensor-memory columns (tcgen05.st) and P.V reads it as a TMEM A
                      // operand; 0 = P_j through double-buffered swizzled shared memory (the round-1 path; dropout and split-KV kernels
                      // always use it).  A/B at cfg-2's shape: 726 vs 711 TFLOP/s stand-alone, attention class -1.9 % in situ.
#endif
#ifndef ATT_EXTRA_SMEM
#define ATT_EXTRA_SMEM 0  // experiments: pad shared memory to lower the number of resident CTAs
#endif
constexpr uint32_t att_smem(bool pt) { return ATT_Q_BYTES + 2 * ATT_K_BYTES + ATT_V_BYTES + (pt ? 0 : 2 * ATT_P_BYTES) + 1024 + 128 + ATT_EXTRA_SMEM; }
constexpr bool ATT_PT = ATT_P_TMEM != 0;
constexpr uint32_t ATT_TMEM_COLS = 128;  // S: 0..63, O: 64..127
constexpr uint32_t ATT_TMEM_P_COLS = 32; // PT kernels: a SECOND allocation for P (64 keys x bf16 = 32 columns): 160 columns per CTA, 480 for three
                                         // co-resident CTAs — a 256-column power-of-two allocation would cap the SM at two CTAs
constexpr float ATT_RESCALE_LOG2 = 8.0f;

struct AttnParams {
  __nv_bfloat16* out;
  float* lse;  // optional [B, H, n]: log2-domain log-sum-exp of the scaled scores (training); +inf for padded query rows
  const int32_t* lens;
  int lens_mod, B, H, n;
  float scale_log2;
  long long* trace;  // debug only (ATT_TRACE builds)
  AttnDrop dr;       // DROP kernels: mask stream of this layer's SDPA dropout (model/modules.py:490)
  const uint32_t* dr_seed;  // optional device word mixed into the stream's key (a captured graph replays with a fresh mask per step)
  int n8;            // ceil(n / 8): mask groups per (batch, head, query) row
};

// P chunk: 32 columns -> exp2 -> bf16 -> swizzled smem row; returns the chunk's row-sum contribution.
// DROP: the dropped elements of the probabilities that go into P.V are zeroed (an AND on the packed pairs; the 1/(1-p) factor is
// folded into the final normalisation), while the row sum keeps the un-dropped values — dropout acts on the NORMALISED probabilities
// (torch SDPA semantics).  g0 = mask group of the chunk's first key (8 keys per group, dropout.cuh).
template <bool MASKED, bool DROP = false>
__device__ __forceinline__ float softmax_chunk(const uint32_t (&s)[32], float sl2, float mb, int lim, uint8_t* prow, int cbase, int rx,
                                               const AttnDrop* dr = nullptr, uint64_t g0 = 0) {
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = fmaf(__uint_as_float(s[q * 8 + i]), sl2, -mb);
      e[i] = ex2_approx(x);
      if constexpr (MASKED) {
        if (q * 8 + i >= lim) e[i] = 0.f;
      }
    }
    sum0 += (e[0] + e[1]) + (e[2] + e[3]);
    sum1 += (e[4] + e[5]) + (e[6] + e[7]);
    uint4 pk;
    pk.x = pack_bf16(e[0], e[1]);
    pk.y = pack_bf16(e[2], e[3]);
    pk.z = pack_bf16(e[4], e[5]);
    pk.w = pack_bf16(e[6], e[7]);
    if constexpr (DROP) {
      uint32_t w0, w1;
      attn_drop_words(*dr, g0 + q, w0, w1);
      pk.x &= attn_drop_pair_mask(w0, 0);
      pk.y &= attn_drop_pair_mask(w0, 1);
      pk.z &= attn_drop_pair_mask(w1, 0);
      pk.w &= attn_drop_pair_mask(w1, 1);
    }
    *reinterpret_cast<uint4*>(prow + (((cbase + q) ^ rx) << 4)) = pk;
  }
  return sum0 + sum1;
}

// same, P kept in registers (16 packed bf16 pairs) for the tensor-memory path
template <bool MASKED>
__device__ __forceinline__ float softmax_chunk_reg(const uint32_t (&s)[32], float sl2, float mb, int lim, uint32_t* pk) {
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = fmaf(__uint_as_float(s[q * 8 + i]), sl2, -mb);
      e[i] = ex2_approx(x);
      if constexpr (MASKED) {
        if (q * 8 + i >= lim) e[i] = 0.f;
      }
    }
    sum0 += (e[0] + e[1]) + (e[2] + e[3]);
    sum1 += (e[4] + e[5]) + (e[6] + e[7]);
    pk[q * 4 + 0] = pack_bf16(e[0], e[1]);
    pk[q * 4 + 1] = pack_bf16(e[2], e[3]);
    pk[q * 4 + 2] = pack_bf16(e[4], e[5]);
    pk[q * 4 + 3] = pack_bf16(e[6], e[7]);
  }
  return sum0 + sum1;
}

// row maximum of the first `valid` of 32 raw scores (4 independent chains when the chunk is full)
__device__ __forceinline__ float row_max32(const uint32_t (&a)[32], int valid) {
  if (valid >= 32) {
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      m0 = fmaxf(fmaxf(m0, __uint_as_float(a[i])), __uint_as_float(a[i + 1]));
      m1 = fmaxf(fmaxf(m1, __uint_as_float(a[i + 2])), __uint_as_float(a[i + 3]));
      m2 = fmaxf(fmaxf(m2, __uint_as_float(a[i + 4])), __uint_as_float(a[i + 5]));
      m3 = fmaxf(fmaxf(m3, __uint_as_float(a[i + 6])), __uint_as_float(a[i + 7]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < valid) m = fmaxf(m, __uint_as_float(a[i]));
  return m;
}

// SPLIT = 2 (small grids, e.g. the B = 1 serving shape: 8 query tiles x 32 batch-heads = 256 CTAs for 444 resident slots, each walking
// all 15 key tiles serially): the key range of a query tile is cut in two, the two halves run as a CLUSTER of 2 CTAs, and rank 1
// hands its un-normalised (O, m, l) to rank 0 through distributed shared memory, which merges and writes — flash-decoding's split-KV
// without a workspace or a combine kernel.  SPLIT = 1 compiles to the single-CTA kernel.
// PT: P in tensor memory.  The score tile S_j is pulled into registers and its columns released exactly as on the shared-memory path
// (S_{j+1} overlaps the exponentials of tile j); P_j is written by tcgen05.st into the P columns and P.V runs as a TS-mode MMA, which
// takes the P store (16 KB), the P operand read (16 KB) and the proxy fence out of the 80 KB of shared-memory traffic per key tile and
// shrinks the CTA to 40 KB.  P is single-buffered: before storing P_j a softmax thread makes sure P_{j-1} V_{j-1} has retired (it was
// issued a whole tile of exponentials earlier, so the probe succeeds at once).
template <int SPLIT, bool DROP = false, bool PT = false>
__global__ void __launch_bounds__(ATT_THREADS, ATT_CTAS_PER_SM)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  static_assert(SPLIT == 1 || (SPLIT == 2 && !PT), "split-KV is built for two halves on the shared-memory P path");
  static_assert(!DROP || (SPLIT == 1 && !PT), "attention dropout is built for the single-CTA shared-memory P path");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_Q_BYTES;
  uint8_t* sV = sK + 2 * ATT_K_BYTES;
  uint8_t* sP = sV + ATT_V_BYTES;  // 16K + 16K + 8K = 40K: 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (PT ? 0 : 2 * ATT_P_BYTES));
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;      // [2] K_j landed
  uint64_t* bar_v = bars + 3;      // V_j landed
  uint64_t* bar_s = bars + 4;      // S_j in TMEM
  uint64_t* bar_sfree = bars + 5;  // S_j pulled into registers by all 128 softmax threads
  uint64_t* bar_p = bars + 6;      // [2] P_j in smem (128 arrivals)
  // [2] P_j V_j retired, tile j on barrier j & 1.  TWO barriers because the softmax threads do not wait for every tile's P V: with a
  // single barrier whose phase advances once per tile, a softmax thread that reaches the epilogue (or the rescale path) while the
  // barrier is still TWO phases behind sees "the phase of this parity has completed" and reads O before the last products landed
  // (observed: rows off by the weight of the last tiles, only when one warp runs a full tile ahead of the slowest).  With two
  // alternating barriers a waiter would have to be four tiles ahead to alias, which the S pipeline (depth 1) rules out.
  uint64_t* bar_pv = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = (SPLIT > 1 ? blockIdx.x / SPLIT : blockIdx.x) * ATT_BQ;
  const uint32_t srank = SPLIT > 1 ? cluster_ctarank() : 0u;  // which half of the key range (cluster = SPLIT CTAs along x)
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  int kvlen = p.n;
  if (p.lens != nullptr) kvlen = min(p.n, __ldg(p.lens + (p.lens_mod > 0 ? b % p.lens_mod : b)));
  const int D = p.H * 64;

  if (kvlen <= 0 || q0 >= kvlen) {
    // whole tile is padding: the reference zeroes these rows after to_out (model/modules.py:499-501)
    griddep_wait();
    if (warp < 4) {
      const int pos = q0 + warp * 32 + lane;
      if (pos < p.n) {
        uint4* o = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.n + pos) * D + h * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
        if (p.lse != nullptr) p.lse[(size_t)bh * p.n + pos] = INFINITY;
      }
    }
    return;
  }
  const int T_all = (kvlen + ATT_BKV - 1) / ATT_BKV;
  int t_begin = 0, T = T_all;  // this CTA's key tiles: [t_begin, t_begin + T)
  if constexpr (SPLIT > 1) {
    const int per = (T_all + SPLIT - 1) / SPLIT;
    t_begin = (int)srank * per;
    T = max(0, min(T_all, t_begin + per) - t_begin);  // may be empty for rank 1 when there is a single key tile
  }

  if (warp == 4) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmK);
      prefetch_tmap(&tmV);
      mbar_init(bar_q, 1);
      mbar_init(&bar_k[0], 1);
      mbar_init(&bar_k[1], 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_sfree, 128);
      mbar_init(&bar_p[0], 128);
      mbar_init(&bar_p[1], 128);
      mbar_init(&bar_pv[0], 1);
      mbar_init(&bar_pv[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (PT) {
      tmem_alloc_keep(tmem_slot, ATT_TMEM_COLS);
      tmem_alloc(tmem_slot + 1, ATT_TMEM_P_COLS);
    } else {
      tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 64;
  const uint32_t tmem_P = PT ? tmem_slot[1] : 0u;
  griddep_wait();  // PDL: q / k / v (the QKV GEMM's output) are first read below
  griddep_launch_dependents();

  if (warp == 4) {
    if (T > 0 && elect_one()) {  // (not `lane == 0`: see tile_engine.cuh — ptxas emits plain tcgen05 / TMA sequences in an elected region)
      const uint32_t idesc = idesc_bf16(128, 64, 0, 0);     // S = Q K^T: both operands K-major
      const uint32_t idesc_pv = idesc_bf16(128, 64, 0, 1);  // O += P V: V is MN-major (d contiguous, keys strided)
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      auto load_k = [&](int j) {
        mbar_arrive_expect_tx(&bar_k[j & 1], ATT_K_BYTES);
        tma_load_3d(sK + (j & 1) * ATT_K_BYTES, &tmK, &bar_k[j & 1], h * 64, (t_begin + j) * ATT_BKV, b);
      };
      auto load_v = [&](int j) {
        mbar_arrive_expect_tx(bar_v, ATT_V_BYTES);
        tma_load_3d(sV, &tmV, bar_v, h * 64, (t_begin + j) * ATT_BKV, b);
      };
      auto issue_s = [&](int j) {
        mbar_wait(&bar_k[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t kb = k_addr + (j & 1) * ATT_K_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, smem_desc_sw128(q_addr + k * 32, 1024, 16), smem_desc_sw128(kb + k * 32, 1024, 16), idesc, k != 0);
        umma_commit(bar_s);
      };
      // prologue
      mbar_arrive_expect_tx(bar_q, ATT_Q_BYTES);
      tma_load_3d(sQ, &tmQ, bar_q, h * 64, q0, b);
      load_k(0);
      if (T > 1) load_k(1);
      load_v(0);
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        // S_j sits in registers: its TMEM buffer and K stage are free -> S_{j+1} overlaps the exponentials of tile j
        mbar_wait(bar_sfree, j & 1);
        tc_fence_after();
        if (j + 1 < T) issue_s(j + 1);
        if (j + 2 < T) load_k(j + 2);  // stage j&1 held K_j
        mbar_wait(&bar_p[j & 1], (j >> 1) & 1);  // P_j in place (and O rescaled if needed)
        tc_fence_after();
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
        if constexpr (PT) {
          // A operand from tensor memory (lane = query row, a 32-bit column = two consecutive keys, a K-step of 16 keys = 8 columns)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tmem_O, tmem_P + kk * 8, smem_desc_sw128(v_addr + kk * 2048, 1024, 8192), idesc_pv, (j | kk) != 0);
        } else {
          const uint32_t pa = p_addr + (j & 1) * ATT_P_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            // MN-major SW128 operand: rows = keys (128 B of d each), 8-row groups 1024 B apart (SBO); one UMMA K-step = 16 keys
            umma_bf16(tmem_O, smem_desc_sw128(pa + kk * 32, 1024, 16), smem_desc_sw128(v_addr + kk * 2048, 1024, 8192), idesc_pv,
                      (j | kk) != 0);
        }
        umma_commit(&bar_pv[j & 1]);
        if (j + 1 < T) {
          mbar_wait(&bar_pv[j & 1], (j >> 1) & 1);  // the single V stage is free once P_j V_j retired
          load_v(j + 1);
        }
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    AttnDrop dr = p.dr;
    if constexpr (DROP) {
      if (p.dr_seed != nullptr) {  // per-launch word from device memory (graph replays: a fresh mask per ODE step)
        const uint32_t sw = __ldg(p.dr_seed);
        dr.key ^= sw * 0x9E3779B9u;
        dr.chi += sw;
      }
    }
    float m_used = -INFINITY;  // log2-domain maximum the exponentials are taken against
    float l_run = 0.f;         // row sum of exp2(s - m_used)
    const float sl2 = p.scale_log2;
    const int rx = r & 7;
#ifdef ATT_TRACE
    long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#define ATT_MARK(i) { const long long tn = clock64(); tr[i] += tn - tprev; tprev = tn; }
#else
#define ATT_MARK(i)
#endif

    for (int j = 0; j < T; ++j) {
      uint8_t* p_row = sP + (j & 1) * ATT_P_BYTES + r * 128;
      const int valid = min(ATT_BKV, kvlen - (t_begin + j) * ATT_BKV);  // CTA-uniform, >= 1
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      ATT_MARK(0)
      uint32_t s0[32], s1[32];
      tmem_ld32(tmem_base + lane_addr, s0);
      tmem_ld32(tmem_base + lane_addr + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_sfree);
      ATT_MARK(1)
      if (j == 0) m_used = fmaxf(row_max32(s0, min(32, valid)), row_max32(s1, min(32, valid - 32))) * sl2;
      // Exponentials are speculative w.r.t. this tile's maximum (see header).  Shared-memory path: bar_s(j) was committed after
      // P_{j-2} V_{j-2} had been issued (tcgen05.commit covers every earlier MMA of the issuing thread), so P buffer (j & 1) is free.
      // PT: P_j stays in registers (16 packed pairs per 32 keys) until it is final.
      float ts;
      // mask group of this tile's first key in this thread's (batch, head, query) row
      const uint64_t g0 = DROP ? ((uint64_t)bh * p.n + (uint64_t)(q0 + r)) * p.n8 + (uint64_t)((t_begin + j) * (ATT_BKV / 8)) : 0;
      // PT: the tile's P goes to tensor memory in two halves of 32 keys (16 packed columns each), so that only 16 packed registers are
      // live next to the 64 scores (which stay live for the rare recomputation below): with all 32 the kernel sat at the 128-register
      // cap with ~6 spill reloads per warp and tile (profiles/r02_ncu_attention_pt.txt).  The P columns are single-buffered: before
      // the first store P_{j-1} V_{j-1} (issued a whole tile of exponentials ago) must have retired.
      auto make_p = [&](bool first) {
        if constexpr (PT) {
          uint32_t ppk[16];
          if (valid == ATT_BKV) ts = softmax_chunk_reg<false>(s0, sl2, m_used, 32, ppk);
          else ts = softmax_chunk_reg<true>(s0, sl2, m_used, valid, ppk);
          if (first && j > 0) {
            mbar_wait(&bar_pv[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tc_fence_after();
          }
          tmem_st16(tmem_P + lane_addr, ppk);
          if (valid == ATT_BKV) ts += softmax_chunk_reg<false>(s1, sl2, m_used, 32, ppk);
          else ts += softmax_chunk_reg<true>(s1, sl2, m_used, valid - 32, ppk);
          tmem_st16(tmem_P + lane_addr + 16, ppk);
        } else {
          if (valid == ATT_BKV) {
            ts = softmax_chunk<false, DROP>(s0, sl2, m_used, 32, p_row, 0, rx, &dr, g0);
            ts += softmax_chunk<false, DROP>(s1, sl2, m_used, 32, p_row, 4, rx, &dr, g0 + 4);
          } else {
            ts = softmax_chunk<true, DROP>(s0, sl2, m_used, valid, p_row, 0, rx, &dr, g0);
            ts += softmax_chunk<true, DROP>(s1, sl2, m_used, valid - 32, p_row, 4, rx, &dr, g0 + 4);
          }
        }
      };
      make_p(true);
      ATT_MARK(2)
      // "the row maximum grew by more than 2^8" implies that some exponential of this tile exceeds 2^8, hence so does the tile's row
      // sum: the 64-way maximum is only evaluated when that cheap necessary condition holds for some row of the warp
      // (!(ts <= 256) also catches an overflowed sum)
      if (j > 0 && __any_sync(0xffffffffu, !(ts <= 256.0f))) {
        const float mt = fmaxf(row_max32(s0, min(32, valid)), row_max32(s1, min(32, valid - 32))) * sl2;
        // warp-uniform decision; tcgen05.ld/st are warp-collective
        if (__any_sync(0xffffffffu, mt > m_used + ATT_RESCALE_LOG2)) {
          const float m_new = fmaxf(m_used, mt);
          const float f = ex2_approx(m_used - m_new);
          m_used = m_new;
          l_run *= f;
          mbar_wait(&bar_pv[(j - 1) & 1], ((j - 1) >> 1) & 1);  // P_{j-1} V_{j-1} has landed in O
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(tmem_O + lane_addr + c * 32, o);
          }
          tmem_st_wait();
          make_p(false);  // against the new reference (PT: overwrites the P columns; P_j V_j has not been issued yet)
        }
      }
      if constexpr (PT) tmem_st_wait();
      l_run += ts;
      ATT_MARK(3)
      if constexpr (!PT) fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_p[j & 1]);
      ATT_MARK(4)
    }
#ifdef ATT_TRACE
    if (p.trace != nullptr && blockIdx.x == 3 && blockIdx.y == 5 && lane == 0) {
      for (int i = 0; i < 6; ++i) p.trace[warp * 8 + i] = tr[i];
      p.trace[warp * 8 + 6] = T;
    }
#endif
    // epilogue: O / l
    const int pos = q0 + r;
    if constexpr (SPLIT == 1) {
      mbar_wait(&bar_pv[(T - 1) & 1], ((T - 1) >> 1) & 1);
      tc_fence_after();
      float inv = (pos < kvlen) ? 1.f / l_run : 0.f;
      if constexpr (DROP) inv *= dr.scale;  // the kept probabilities' 1 / (1 - p)
      if (p.lse != nullptr && pos < p.n) p.lse[(size_t)bh * p.n + pos] = (pos < kvlen) ? m_used + log2f(l_run) : INFINITY;
      __nv_bfloat16* orow = p.out + ((size_t)b * p.n + (pos < p.n ? pos : 0)) * D + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tmem_O + lane_addr + c * 32, o);
        tmem_ld_wait();
        if (pos < p.n) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk;
            pk.x = pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            pk.y = pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            pk.z = pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            pk.w = pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            reinterpret_cast<uint4*>(orow + c * 32)[q] = pk;
          }
        }
      }
    } else {
      // the tile buffers (Q, K, V, P: 72 KB from `smem`) are dead once the last product has retired: rank 1 parks its partial there,
      // column-major [64][128] so that the 128 row-threads store and load without bank conflicts, then (m, l)
      float* xo = reinterpret_cast<float*>(smem);
      float* xm = xo + 64 * ATT_BQ;
      float* xl = xm + ATT_BQ;
      if (T > 0) {
        mbar_wait(&bar_pv[(T - 1) & 1], ((T - 1) >> 1) & 1);
        tc_fence_after();
      }
      uint32_t o0[32], o1[32];
      if (T > 0) {
        tmem_ld32(tmem_O + lane_addr, o0);
        tmem_ld32(tmem_O + lane_addr + 32, o1);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o0[i] = o1[i] = 0u;
      }
      if (srank == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          xo[i * ATT_BQ + r] = __uint_as_float(o0[i]);
          xo[(32 + i) * ATT_BQ + r] = __uint_as_float(o1[i]);
        }
        xm[r] = m_used;
        xl[r] = l_run;
      }
      cluster_sync_all();  // (warp 4 takes part below) rank 1's partial is visible cluster-wide
      if (srank == 0) {
        const uint32_t ro = mapa_u32(xo, 1), rm = mapa_u32(xm, 1), rl = mapa_u32(xl, 1);
        const float m1 = ld_shared_cluster_f32(rm + r * 4), l1 = ld_shared_cluster_f32(rl + r * 4);
        const float m = fmaxf(m_used, m1);  // rank 0 always owns >= 1 key tile, so m is finite
        const float f0 = ex2_approx(m_used - m), f1 = (l1 > 0.f) ? ex2_approx(m1 - m) : 0.f;
        const float l = l_run * f0 + l1 * f1;
        const float inv = (pos < kvlen) ? 1.f / l : 0.f;
        const float a0 = f0 * inv, a1 = f1 * inv;
        __nv_bfloat16* orow = p.out + ((size_t)b * p.n + (pos < p.n ? pos : 0)) * D + h * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float mine = __uint_as_float(c == 0 ? o0[i] : o1[i]);
            v[i] = mine * a0 + ld_shared_cluster_f32(ro + (uint32_t)(((c * 32 + i) * ATT_BQ + r) * 4)) * a1;
          }
          if (pos < p.n) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 pk;
              pk.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
              pk.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
              pk.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
              pk.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
              reinterpret_cast<uint4*>(orow + c * 32)[q] = pk;
            }
          }
        }
      }
      cluster_sync_all();  // rank 1's shared memory stays alive until rank 0 has read it
    }
    tc_fence_before();
  }
  if constexpr (SPLIT > 1) {
    if (warp == 4) {  // the cluster barrier counts every thread of both CTAs
      cluster_sync_all();
      cluster_sync_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    if constexpr (PT) tmem_dealloc(tmem_P, ATT_TMEM_P_COLS);
  }
}

long long* g_attn_trace = nullptr;
int g_attn_split = [] { const char* e = getenv("F5B_ATTN_SPLIT"); return e ? atoi(e) : 0; }();  // 1: split-KV wherever it is allowed
int g_attn_variant = 0;  // 0: this file; 1: the experimental persistent kernel (attention_fa.cu, only in F5B_WITH_ATTN_FA builds)
#ifdef F5B_WITH_ATTN_FA
int attn_fwd_fa(const void* q, const void* k, const void* v, int ld, void* out, float* lse, const int32_t* lens, int lens_mod, int B,
                int H, int n, float scale, cudaStream_t stream);
#endif

int attn_fwd(const void* q, const void* k, const void* v, int ld, void* out, float* lse, const int32_t* lens, int lens_mod, int B, int H,
             int n, float scale, cudaStream_t stream, const AttnDrop* drop, const uint32_t* drop_seed_dev) {
  F5B_CHECK(q && k && v && out, "f5b_attn_fwd: null pointer");
  F5B_CHECK(B > 0 && H > 0 && n > 0 && ld >= H * 64 && (ld & 7) == 0, "f5b_attn_fwd: bad shape B %d H %d n %d ld %d", B, H, n, ld);
  LaunchScope scope(K_ATTN, stream, 4.0 * B * H * (double)n * n * 64, 2.0 * 4 * B * H * (double)n * 64);
#ifdef F5B_WITH_ATTN_FA
  if (g_attn_variant == 1) return attn_fwd_fa(q, k, v, ld, out, lse, lens, lens_mod, B, H, n, scale, stream);
#endif
  // q, k, v are token-major matrices [B*n, ld] (e.g. the three column sections of the fused QKV GEMM output); head h of
  // batch row b is the strided box (cols h*64.., rows b*n + pos) — TMA gathers it, no head-major copy exists
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t hw = (uint64_t)H * 64, pitch = (uint64_t)ld * 2;
  if (make_tmap_3d(&tmQ, q, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, ATT_BQ, 1, true)) return -1;
  if (make_tmap_3d(&tmK, k, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, ATT_BKV, 1, true)) return -1;
  if (make_tmap_3d(&tmV, v, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, ATT_BKV, 1, true)) return -1;
  static bool configured = false;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<1, false, ATT_PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem(ATT_PT)));
    F5B_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem(false)));
    F5B_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem(false)));
    configured = true;
  }
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.lens = lens;
  p.lens_mod = lens_mod;
  p.B = B;
  p.H = H;
  p.n = n;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  p.dr = drop ? *drop : AttnDrop{0u, 1.f, 0.f, 0u, 0u};
  p.n8 = (n + 7) / 8;
  p.dr_seed = drop_seed_dev;
  dim3 grid((n + ATT_BQ - 1) / ATT_BQ, B * H);
  if (p.dr.addc != 0) {  // SDPA dropout (training): the forward's mask is regenerated by attn_bwd from the same Drop
    F5B_CUDA(launch_dep(attn_fwd_kernel<1, true>, grid, dim3(ATT_THREADS), att_smem(false), stream, 1, tmQ, tmK, tmV, p));
    F5B_CUDA(cudaGetLastError());
    return 0;
  }
  // split-KV (see the kernel's header) is OPT-IN: F5B_ATTN_SPLIT=1 / f5b_debug_attn_split(1).  Measured on the shape it was built
  // for (cfg-1: 8 x 32 CTAs, 15 key tiles): attention 16.4 -> 21.9 ms per utterance, the step 68.0 -> 69.7 ms — the cluster
  // co-scheduling and the distributed-shared-memory merge cost more than the halved key loop saves, so it is not used by default.
  if (lse == nullptr && n >= 4 * ATT_BKV && g_attn_split == 1) {
    grid.x *= 2;
    F5B_CUDA(launch_dep(attn_fwd_kernel<2>, grid, dim3(ATT_THREADS), att_smem(false), stream, 2, tmQ, tmK, tmV, p));
    F5B_CUDA(cudaGetLastError());
    return 0;
  }
  F5B_CUDA(launch_dep(attn_fwd_kernel<1, false, ATT_PT>, grid, dim3(ATT_THREADS), att_smem(ATT_PT), stream, 1, tmQ, tmK, tmV, p));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f5b

extern "C" int f5b_attn_fwd(const void* q, const void* k, const void* v, int ld, void* out, const int32_t* lens, int lens_mod,
                            int B, int H, int n, float scale, f5b_stream_t stream) {
  return f5b::attn_fwd(q, k, v, ld, out, nullptr, lens, lens_mod, B, H, n, scale, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

extern "C" int f5b_attn_fwd_lse(const void* q, const void* k, const void* v, int ld, void* out, float* lse, const int32_t* lens,
                                int lens_mod, int B, int H, int n, float scale, f5b_stream_t stream) {
  return f5b::attn_fwd(q, k, v, ld, out, lse, lens, lens_mod, B, H, n, scale, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

extern "C" void f5b_debug_set_attn_trace(long long* buf) { f5b::g_attn_trace = buf; }
extern "C" void f5b_debug_attn_variant(int v) { f5b::g_attn_variant = v; }
extern "C" void f5b_debug_attn_split(int v) { f5b::g_attn_split = v; }
#ifndef F5B_WITH_ATTN_FA
extern "C" void f5b_debug_attn_poly(int) {}
#endif
