// Attention forward, second generation: O = softmax(Q K^T / sqrt(64), keys < len[b]) V, non-causal, d_head 64.
// Replaces AttnProcessor's mask expansion + F.scaled_dot_product_attention + head merge + padded-row zeroing
// (/root/reference/src/f5_tts/model/modules.py:483-501; dropout_p = 0, see DESIGN.md "oracle adjustments").
//
// The first-generation kernel (attention.cu: one 128-query tile per CTA, three CTAs per SM, 64-key tiles, P through shared memory)
// sits at ~0.6 of its MUFU.EX2 bound (ncu: XU 68 %, tensor 34 %).  This one is built like a Blackwell flash-attention pipeline:
//
//   * PERSISTENT, one CTA per SM, 384 threads (setmaxnreg moves the registers of the third warpgroup to the softmax warps).  A work item is a PAIR of 128-query tiles of one (batch, head) — 256 query rows that
//     share every K / V tile — walked over the keys in tiles of 128.
//   * warp 9: TMA producer.  Q pair double-buffered (the next item's Q is in flight while this one computes), K and V in
//     3-stage rings of [128 keys x 64] tiles gathered straight from the token-major QKV matrix (3-D boxes, 128B swizzle).
//   * warps 8 / 10: the tcgen05.mma issuers, one per query tile.  Per key tile and query tile i:  O_i += P_i V_j  with the A operand (P) read from TENSOR
//     MEMORY and V as an MN-major shared-memory operand.  TMEM (all 512 columns): S_0 | S_1 (2 x 128 fp32) | O_0 | O_1 (2 x 64) |
//     P_0 | P_1 (2 x 64 columns of bf16 pairs).  P has its OWN columns, so S_i of key tile j+1 is issued as soon as the softmax group
//     has pulled S_i of tile j into registers — under its exponentials — instead of behind P_i V_j (first version: P over S, the
//     chain softmax -> P V -> next S -> softmax measured 1090 clocks of tensor round trip per tile on top of the softmax).
//   * warps 0-3 / 4-7: one softmax warpgroup per query tile, thread = query row (tcgen05.ld 32x32b: no cross-lane reduction).
//     The two warpgroups PING-PONG: while group 0 takes the exponentials of S_0, the tensor pipe produces S_1 / consumes P_1, and
//     each SM sub-partition always has one warp in its MUFU phase while the other waits for / loads / reduces its next tile.
//     Row maximum with 3-input FMNMX, scale-and-subtract and the row sum on packed FFMA2 / FADD2, lazy rescale of O (only when a
//     row's maximum grows by more than 2^8; O is rescaled in tensor memory by the softmax threads themselves, race-free because the
//     commit that publishes S_{j+1} also covers P_j V_j), and a fraction of the exponentials (ATT_POLY of every 16 pairs) evaluated
//     on the FMA pipe — Cody-Waite split + degree-3 polynomial, exponent spliced in with an integer multiply-add — to relieve the
//     16-lane MUFU pipe, which is the roofline of d_head 64 attention (one exp per 256 tensor FLOPs).
//   * epilogue per query tile by its own warpgroup (O / l -> bf16 token-major rows, optional log-sum-exp), overlapped with the
//     other tile's last key tiles and with the next item's first S products.
// Key padding is a per-batch length bound: key tiles past len[b] are never loaded, the last one is masked by index; query tiles past
// len[b] are zero-filled (the reference zeroes those rows after to_out).
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int FA_BQ = 128;
constexpr int FA_BK = 128;
constexpr int FA_NS = 3;          // K / V ring depth
constexpr int FA_THREADS = 384;   // 8 softmax warps + a third warpgroup: two MMA warps (one per query tile), the TMA warp, one idle
// register budget per SM sub-partition (16384): two softmax warps at 216 + one warp of the third group at 72
constexpr int FA_REGS_SOFTMAX = 216, FA_REGS_OTHER = 72;
constexpr uint32_t FA_TILE = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr uint32_t FA_SMEM = (4 + 2 * FA_NS) * FA_TILE + 1024 + 512;
constexpr uint32_t FA_TMEM_COLS = 512;
constexpr float FA_RESCALE_LOG2 = 8.0f;
#ifndef ATT_POLY
#define ATT_POLY 4  // of every 16 element pairs, this many take their exponentials on the FMA pipe (0 = all on MUFU)
#endif
int g_attn_poly = ATT_POLY;  // f5b_debug_attn_poly(): A/B switch between the instantiated fractions {0, 2, 4, 6, 8} / 16

struct FaParams {
  __nv_bfloat16* out;
  float* lse;  // optional [B, H, n]: log2-domain log-sum-exp of the scaled scores (training); +inf for padded query rows
  const int32_t* lens;
  int lens_mod, B, H, n, n_pairs, num_items;
  float scale_log2;
  long long* trace;  // debug only (FA_TRACE builds): clock64 stamps of CTA 0's first items
};
#ifdef FA_TRACE
// slot layout: [role 0..2 = softmax group 0, softmax group 1, MMA][event index 0..255][stamp]
#define FA_STAMP(role, idx, k) { if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 64) p.trace[((role) * 64 + (idx)) * 8 + (k)] = clock64(); }
#else
#define FA_STAMP(role, idx, k) {}
#endif

// bounded mbarrier wait without the printf of common.cuh's mbar_wait (a call in the 40-register warps would spill everything live):
// a protocol bug still traps instead of hanging the GPU
__device__ __forceinline__ void fa_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 6000000000LL) {  // ~3 s
#ifdef FA_DEBUG
        printf("fa_wait timeout: block %d warp %d lane %d bar@%u parity %u\n", blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, smem_u32(bar) & 1023u, parity);
#endif
        __trap();
      }
    }
  }
}
// named barriers 1 / 2: the MUFU "token" the two softmax groups pass back and forth (ping-pong): a group takes its exponentials only
// while it holds the token, so the other group's wait / tcgen05.ld / maximum / P store always run in the shadow of a MUFU phase
__device__ __forceinline__ void named_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_arrive(int id) { asm volatile("bar.arrive %0, 256;" ::"r"(id) : "memory"); }
#ifndef FA_PINGPONG
#define FA_PINGPONG 1
#endif
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 2^x for x <= ~100 on the FMA pipe: n = round(x) through the 1.5 * 2^23 magic add, r = x - n in [-0.5, 0.5], degree-3 minimax
// polynomial of 2^r (max relative error 7.5e-5, far below the bf16 rounding of P), exponent spliced in as an integer multiply-add
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 xf = __fadd2_rn(x, magic);
  const float2 nn = __fadd2_rn(xf, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __ffma2_rn(nn, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(make_float2(0.0551716648f, 0.0551716648f), r, make_float2(0.2426111251f, 0.2426111251f));
  p = __ffma2_rn(p, r, make_float2(0.6932609677f, 0.6932609677f));
  p = __ffma2_rn(p, r, make_float2(0.9999280572f, 0.9999280572f));
  float2 e;
  e.x = __int_as_float(__float_as_int(xf.x) * (1 << 23) + __float_as_int(p.x));
  e.y = __int_as_float(__float_as_int(xf.y) * (1 << 23) + __float_as_int(p.y));
  return e;
}

// 32 scores of one row -> exp2(s * sl2 - m) -> 16 packed bf16 pairs; returns the updated (packed) running row sum
template <int NPOLY>
__device__ __forceinline__ float2 fa_chunk(const uint32_t (&s)[32], float2 sc, float2 negm, float2 sum, uint32_t (&pk)[16]) {
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * p]), __uint_as_float(s[2 * p + 1])), sc, negm);
    float2 e;
    // spread the NPOLY polynomial pairs evenly over the 16 (Bresenham), so their FMA chains fill the MUFU latency of the others
    if (((p + 1) * NPOLY) / 16 != (p * NPOLY) / 16) {
      e = exp2_poly2(x);
    } else {
      e.x = ex2_approx(x.x);
      e.y = ex2_approx(x.y);
    }
    sum = __fadd2_rn(sum, e);
    pk[p] = pack_bf16(e.x, e.y);
  }
  return sum;
}

__device__ __forceinline__ float fa_rowmax(const uint32_t (&s)[4][32]) {
  float m[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float a = fmax3(__uint_as_float(s[c][0]), __uint_as_float(s[c][1]), __uint_as_float(s[c][2]));
    float b = fmax3(__uint_as_float(s[c][3]), __uint_as_float(s[c][4]), __uint_as_float(s[c][5]));
#pragma unroll
    for (int i = 6; i + 3 < 32; i += 4) {
      a = fmax3(a, __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
      b = fmax3(b, __uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3]));
    }
    m[c] = fmax3(a, b, fmaxf(__uint_as_float(s[c][30]), __uint_as_float(s[c][31])));
  }
  return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}

struct FaItem {
  int b, h, pair, kv, T;
  bool two;
};
__device__ __forceinline__ FaItem fa_item(const FaParams& p, int w) {
  FaItem it;
  const int bh = w / p.n_pairs;
  it.pair = w - bh * p.n_pairs;
  it.b = bh / p.H;
  it.h = bh - it.b * p.H;
  it.kv = p.n;
  if (p.lens != nullptr) it.kv = max(0, min(p.n, __ldg(p.lens + (p.lens_mod > 0 ? it.b % p.lens_mod : it.b))));
  const bool a0 = it.pair * 2 * FA_BQ < it.kv;
  it.two = (it.pair * 2 + 1) * FA_BQ < it.kv;
  it.T = a0 ? (it.kv + FA_BK - 1) / FA_BK : 0;
  return it;
}
// the same for a role whose whole warp runs converged (MMA / TMA warps): every field provably warp-uniform for the compiler
// (the key length comes from a global load, which it would otherwise treat as per-lane and wrap each tcgen05 / TMA instruction in
// a per-value "waterfall" loop of R2UR + ELECT + BRA — measured: ~65 clocks per tcgen05.mma, twice its execution time)
__device__ __forceinline__ FaItem fa_item_uniform(const FaParams& p, int w) {
  FaItem it = fa_item(p, w);
  it.kv = __shfl_sync(0xffffffffu, it.kv, 0);
  const bool a0 = it.pair * 2 * FA_BQ < it.kv;
  it.two = (it.pair * 2 + 1) * FA_BQ < it.kv;
  it.T = a0 ? (it.kv + FA_BK - 1) / FA_BK : 0;
  return it;
}

template <int NPOLY>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_fa_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // [2 buffers][2 tiles]
  uint8_t* sK = sQ + 4 * FA_TILE;           // [FA_NS]
  uint8_t* sV = sK + FA_NS * FA_TILE;       // [FA_NS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FA_NS * FA_TILE);
  uint64_t* q_full = bars + 0;              // [2]
  uint64_t* q_free = bars + 2;              // [2]
  uint64_t* k_full = bars + 4;              // [FA_NS]
  uint64_t* k_free = k_full + FA_NS;
  uint64_t* v_full = k_free + FA_NS;
  uint64_t* v_free = v_full + FA_NS;
  uint64_t* s_full = v_free + FA_NS;        // [2] S_i of the next key tile is in tensor memory
  uint64_t* s_free = s_full + 2;            // [2] S_i pulled into registers by its 128 softmax threads: the columns may be overwritten
  uint64_t* p_full = s_free + 2;            // [2] P_i written (and O_i rescaled if needed): 128 arrivals
  uint64_t* p_free = p_full + 2;            // [2] P_i V retired: P_i may be overwritten, O_i is at rest
  uint64_t* o_full = p_free + 2;            // [2] last P_i V of the item retired
  uint64_t* o_free = o_full + 2;            // [2] O_i drained by the epilogue: 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmK);
      prefetch_tmap(&tmV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1);
        mbar_init(&q_free[i], 2);  // both MMA warps
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 128);
        mbar_init(&p_full[i], 128);
        mbar_init(&p_free[i], 1);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 128);
      }
      for (int i = 0; i < FA_NS; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_free[i], 2);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_free[i], 2);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // PDL: q / k / v (the QKV GEMM's output) are first read below
  griddep_launch_dependents();

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FA_REGS_OTHER));
    if (warp == 9) {
    // ------------------------------------------------------------------------------------------------ TMA producer
    {
      // whole warp converged, one elected lane issues (see the MMA warp below for why not `lane == 0`)
      uint32_t kv_it = 0, q_it = 0;
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
        const FaItem it = fa_item_uniform(p, w);
        if (it.T == 0) continue;
        const uint32_t qs = q_it & 1;
        fa_wait(&q_free[qs], ((q_it >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[qs], it.two ? 2 * FA_TILE : FA_TILE);
          tma_load_3d(sQ + (qs * 2 + 0) * FA_TILE, &tmQ, &q_full[qs], it.h * 64, it.pair * 2 * FA_BQ, it.b);
          if (it.two) tma_load_3d(sQ + (qs * 2 + 1) * FA_TILE, &tmQ, &q_full[qs], it.h * 64, (it.pair * 2 + 1) * FA_BQ, it.b);
        }
        __syncwarp();
        ++q_it;
        for (int j = 0; j < it.T; ++j, ++kv_it) {
          const uint32_t st = kv_it % FA_NS, ph = (kv_it / FA_NS) & 1;
          fa_wait(&k_free[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&k_full[st], FA_TILE);
            tma_load_3d(sK + st * FA_TILE, &tmK, &k_full[st], it.h * 64, j * FA_BK, it.b);
          }
          __syncwarp();
          fa_wait(&v_free[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&v_full[st], FA_TILE);
            tma_load_3d(sV + st * FA_TILE, &tmV, &v_full[st], it.h * 64, j * FA_BK, it.b);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
    } else if (warp == 8 || warp == 10) {
    // ------------------------------------------------------------------------------------------------ MMA issuers
    {
      // One issuing warp PER QUERY TILE (warp 8: tile 0, warp 10: tile 1): each walks its own tile's schedule in order and therefore
      // answers that tile's events at once (a single in-order issuer was measured busy ~70 % of the time — blocked in the shallow
      // tcgen05 queue and in ~90-clock barrier probes — so P_i waited up to 800 clocks for its P V).  The two warps touch disjoint
      // tensor-memory columns (S_i, P_i, O_i); the K / V / Q stages they share are released by count-2 barriers.
      // The WHOLE warp runs converged (all lanes poll the barriers) and one lane, chosen by elect.sync, issues: ptxas then knows the
      // region is single-threaded and emits plain UTCHMMA sequences from uniform registers (under a `lane == 0` predicate it wraps
      // EVERY tcgen05 instruction in an R2UR / ELECT / BRA.U.ANY loop: ~65 clocks per MMA, twice its execution time).
      const int ti = warp == 8 ? 0 : 1;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t tS = tmem_u + ti * 128, tO = tmem_u + 256 + ti * 64, tP = tmem_u + 384 + ti * 64;
      const uint32_t idesc_s = idesc_bf16(128, 128, 0, 0);   // S = Q K^T: both operands K-major, N = 128 keys
      const uint32_t idesc_pv = idesc_bf16(128, 64, 0, 1);   // O += P V: P from tensor memory, V MN-major (d contiguous)
      const uint32_t q_addr = smem_u32(sQ) + ti * FA_TILE, k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      // One step = one key tile of one item of this query tile.  Issue order per step: S of the NEXT step (it only needs the previous
      // scores pulled into registers: s_free), then P V of this step.
      struct Step {
        bool valid, two, first, last, early;
        uint32_t qs, qph, kv;  // Q buffer + phase, running key-tile counter (ring stage = kv % FA_NS)
      };
      uint32_t s_iss = 0, p_cnt = 0, o_cnt = 0, kv_it = 0, q_it = 0;
      int w = blockIdx.x, j = 0, T = 0;
      bool two = false;
      auto next_step = [&](bool start) -> Step {
        Step st = {};
        // `early`: this step's S may be issued BEFORE the previous step's P V.  Not across an item this warp skips: the producer
        // reaches the item after it only through the skipped item's K / V loads, whose ring stages wait for that very P V.
        st.early = true;
        if (!start) {
          ++j;
          ++kv_it;
          if (j < T) {
            st.valid = true;
          } else {
            ++q_it;
            w += gridDim.x;
          }
        }
        if (!st.valid) {
          for (; w < p.num_items; w += gridDim.x) {
            const FaItem it = fa_item_uniform(p, w);
            if (it.T > 0 && (ti == 0 || it.two)) {
              T = it.T;
              two = it.two;
              break;
            }
            if (it.T > 0) {  // an item without a second query tile: its key tiles and Q buffer still advance the rings
              kv_it += it.T;
              ++q_it;
              st.early = false;
            }
          }
          if (w >= p.num_items) return st;
          j = 0;
          st.valid = true;
        }
        st.two = two;
        st.first = j == 0;
        st.last = j == T - 1;
        st.qs = q_it & 1;
        st.qph = (q_it >> 1) & 1;
        st.kv = kv_it;
        return st;
      };
      auto issue_s = [&](const Step& st) {
        if (st.first) fa_wait(&q_full[st.qs], st.qph);
        fa_wait(&k_full[st.kv % FA_NS], (st.kv / FA_NS) & 1);
        fa_wait(&s_free[ti], (s_iss & 1) ^ 1);  // the previous S_i has been pulled into registers
        ++s_iss;
        tc_fence_after();
        const uint32_t a = q_addr + st.qs * 2 * FA_TILE, bb = k_addr + (st.kv % FA_NS) * FA_TILE;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tS, smem_desc_sw128(a + k * 32, 1024, 16), smem_desc_sw128(bb + k * 32, 1024, 16), idesc_s, k != 0);
          umma_commit(&s_full[ti]);
          umma_commit(&k_free[st.kv % FA_NS]);
          if (!st.two) umma_commit(&k_free[st.kv % FA_NS]);  // count-2 barrier, no second query tile in this item
          if (st.last) {  // the item's last read of its Q buffer
            umma_commit(&q_free[st.qs]);
            if (!st.two) umma_commit(&q_free[st.qs]);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](const Step& st) {
        if (lane == 0) FA_STAMP(2, 2 * p_cnt + ti, 0)
        fa_wait(&p_full[ti], p_cnt & 1);
        if (lane == 0) FA_STAMP(2, 2 * p_cnt + ti, 1)
        ++p_cnt;
        if (st.first) {
          fa_wait(&o_free[ti], (o_cnt & 1) ^ 1);  // the previous item's epilogue has drained O_i
          ++o_cnt;
        }
        fa_wait(&v_full[st.kv % FA_NS], (st.kv / FA_NS) & 1);
        tc_fence_after();
        if (lane == 0) FA_STAMP(2, 2 * (p_cnt - 1) + ti, 3)
        const uint32_t vb = v_addr + (st.kv % FA_NS) * FA_TILE;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            // MN-major SW128 operand: rows = keys (128 B of d each), 8-row groups 1024 B apart; one UMMA K-step = 16 keys
            umma_bf16_ts(tO, tP + kk * 8, smem_desc_sw128(vb + kk * 2048, 1024, 8192), idesc_pv, (!st.first) || kk != 0);
          umma_commit(&p_free[ti]);
          umma_commit(&v_free[st.kv % FA_NS]);
          if (!st.two) umma_commit(&v_free[st.kv % FA_NS]);
          if (st.last) umma_commit(&o_full[ti]);
        }
        __syncwarp();
        if (lane == 0) FA_STAMP(2, 2 * (p_cnt - 1) + ti, 4)
      };
      Step cur = next_step(true);
      if (cur.valid) issue_s(cur);
      while (cur.valid) {
        const Step nxt = next_step(false);
        if (nxt.valid && nxt.early) issue_s(nxt);
        issue_pv(cur);
        if (nxt.valid && !nxt.early) issue_s(nxt);
        cur = nxt;
      }
      // this warp's last commits arrive on the CTA's shared memory: see the last one land before the CTA may exit
      if (o_cnt > 0) fa_wait(&o_full[ti], (o_cnt - 1) & 1);
    }
    __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FA_REGS_SOFTMAX));
    const int wg = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + wg * 128 + lane_addr;
    const uint32_t tO = tmem_base + 256 + wg * 64 + lane_addr;
    const float sl2 = p.scale_log2;
    const float2 sc = make_float2(sl2, sl2);
    const int D = p.H * 64;
    uint32_t s_cnt = 0, o_cnt = 0, p_cnt = 0;
    const uint32_t tP = tmem_base + 384 + wg * 64 + lane_addr;
    if (FA_PINGPONG && wg == 1) named_arrive(1);  // group 0 holds the token first
    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
      const FaItem it = fa_item(p, w);
      const int q0 = (it.pair * 2 + wg) * FA_BQ;
      const int pos = q0 + r;
      if (q0 >= it.kv) {
        // the whole tile is padding: the reference zeroes these rows after to_out (model/modules.py:499-501)
        if (pos < p.n) {
          uint4* o = reinterpret_cast<uint4*>(p.out + ((size_t)it.b * p.n + pos) * D + it.h * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
          if (p.lse != nullptr) p.lse[((size_t)it.b * p.H + it.h) * p.n + pos] = INFINITY;
        }
        continue;
      }
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < it.T; ++j) {
        if (lane == 0 && (warp & 3) == 0) FA_STAMP(wg, s_cnt, 0)
        fa_wait(&s_full[wg], s_cnt & 1);
        if (lane == 0 && (warp & 3) == 0) FA_STAMP(wg, s_cnt, 1)
        const uint32_t tr_idx = s_cnt;
        ++s_cnt;
        tc_fence_after();
        uint32_t s[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[wg]);  // S_i is in registers: the tensor pipe may produce the next key tile's scores over it
        if (lane == 0 && (warp & 3) == 0) FA_STAMP(wg, tr_idx, 2)
        const int valid = it.kv - j * FA_BK;  // CTA-uniform, >= 1
        if (valid < FA_BK) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf: excluded from the maximum, exp2 -> 0
        }
        const float mt = fa_rowmax(s) * sl2;
        // P_i of the previous key tile has been consumed (and O_i is at rest) once its P V retired
        fa_wait(&p_free[wg], (p_cnt & 1) ^ 1);
        ++p_cnt;
        tc_fence_after();
        if (j == 0) {
          m_used = mt;
        } else if (__any_sync(0xffffffffu, mt > m_used + FA_RESCALE_LOG2)) {
          // lazy rescale (warp-uniform: tcgen05.ld / st are warp-collective); O_i is at rest until P_j is published below
          const float m_new = fmaxf(m_used, mt);
          const float f = ex2_approx(m_used - m_new);
          m_used = m_new;
          l_run *= f;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(tO + c * 32, o);
          }
        }
        if (lane == 0 && (warp & 3) == 0) FA_STAMP(wg, tr_idx, 3)
        const float2 negm = make_float2(-m_used, -m_used);
        float2 sum = make_float2(0.f, 0.f);
        if (FA_PINGPONG && it.two) named_sync(1 + wg);  // take the MUFU token
        if (valid < FA_BK) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
            sum = fa_chunk<0>(s[c], sc, negm, sum, pk);  // masked tile: all exponentials on MUFU (exp2(-inf) = 0 exactly)
            tmem_st16(tP + c * 16, pk);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
            sum = fa_chunk<NPOLY>(s[c], sc, negm, sum, pk);
            tmem_st16(tP + c * 16, pk);
          }
        }
        if (FA_PINGPONG && it.two) named_arrive(2 - wg);  // hand it to the other group
        l_run += sum.x + sum.y;
        if (lane == 0 && (warp & 3) == 0) FA_STAMP(wg, tr_idx, 4)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[wg]);
        if (lane == 0 && (warp & 3) == 0) FA_STAMP(wg, tr_idx, 5)
      }
      // epilogue: O / l -> bf16 rows
      fa_wait(&o_full[wg], o_cnt & 1);
      ++o_cnt;
      tc_fence_after();
      const float inv = (pos < it.kv) ? 1.f / l_run : 0.f;
      if (p.lse != nullptr && pos < p.n)
        p.lse[((size_t)it.b * p.H + it.h) * p.n + pos] = (pos < it.kv) ? m_used + log2f(l_run) : INFINITY;
      __nv_bfloat16* orow = p.out + ((size_t)it.b * p.n + (pos < p.n ? pos : 0)) * D + it.h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();
          mbar_arrive(&o_free[wg]);  // O_i is in registers: the next item's first P V may overwrite it
        }
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FA_TMEM_COLS);
  }
}

int attn_fwd_fa(const void* q, const void* k, const void* v, int ld, void* out, float* lse, const int32_t* lens, int lens_mod, int B, int H,
                int n, float scale, cudaStream_t stream) {
  // q, k, v are token-major matrices [B*n, ld] (e.g. the three column sections of the fused QKV GEMM output); head h of
  // batch row b is the strided box (cols h*64.., rows b*n + pos) — TMA gathers it, no head-major copy exists
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t hw = (uint64_t)H * 64, pitch = (uint64_t)ld * 2;
  if (make_tmap_3d(&tmQ, q, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, FA_BQ, 1, true)) return -1;
  if (make_tmap_3d(&tmK, k, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, FA_BK, 1, true)) return -1;
  if (make_tmap_3d(&tmV, v, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, FA_BK, 1, true)) return -1;
  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, FaParams) = attn_fa_kernel<ATT_POLY>;
  switch (g_attn_poly) {
    case 0: kern = attn_fa_kernel<0>; break;
    case 2: kern = attn_fa_kernel<2>; break;
    case 4: kern = attn_fa_kernel<4>; break;
    case 6: kern = attn_fa_kernel<6>; break;
    case 8: kern = attn_fa_kernel<8>; break;
    default: break;
  }
  static bool configured = false;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(attn_fa_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM));
    F5B_CUDA(cudaFuncSetAttribute(attn_fa_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM));
    F5B_CUDA(cudaFuncSetAttribute(attn_fa_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM));
    F5B_CUDA(cudaFuncSetAttribute(attn_fa_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM));
    F5B_CUDA(cudaFuncSetAttribute(attn_fa_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM));
    configured = true;
  }
  FaParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.lens = lens;
  p.lens_mod = lens_mod;
  p.B = B;
  p.H = H;
  p.n = n;
  p.n_pairs = (n + 2 * FA_BQ - 1) / (2 * FA_BQ);
  p.num_items = B * H * p.n_pairs;
  p.scale_log2 = scale * 1.4426950408889634f;
  extern long long* g_attn_trace;
  p.trace = g_attn_trace;
  const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
  F5B_CUDA(launch_dep(kern, dim3(grid), dim3(FA_THREADS), FA_SMEM, stream, 1, tmQ, tmK, tmV, p));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f5b

extern "C" void f5b_debug_attn_poly(int v) { f5b::g_attn_poly = v; }
