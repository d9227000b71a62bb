// Flash-style attention BACKWARD for the DiT blocks (training path, CFM.forward -> loss.backward in the reference,
// /root/reference/src/f5_tts/model/cfm.py:210-283 driving AttnProcessor /root/reference/src/f5_tts/model/modules.py:442-503
// through autograd; dropout_p = 0, DESIGN.md "oracle adjustments").
//
//   given  Q, K, V (token-major bf16, the fused QKV GEMM output), dO, the forward's log2-sum-exp L and delta = rowsum(dO . O):
//     P  = exp2(Q K^T * c - L)            c = scale * log2(e), keys >= len[b] masked
//     dV = P^T dO          dP = dO V^T          dS = scale * P . (dP - delta)
//     dK = dS^T Q          dQ = dS K
//
// sm_100a design: one CTA owns one 128-key tile of one (batch, head) and walks the 128-query tiles.  Everything is computed in the
// TRANSPOSED orientation (TMEM lane = key), so that P^T and dS^T come out of the softmax threads row-major in exactly the
// layout the tcgen05 A operand wants, and all five products run without a single transposed copy:
//     S^T  = K   Q^T    A = K   (K-major)   B = Q   (K-major)        TMEM cols   0..127
//     dP^T = V   dO^T   A = V   (K-major)   B = dO  (K-major)        TMEM cols 128..255
//     dV  += P^T  dO    A = P^T (TMEM, 448..511) B = dO (MN-major)     TMEM cols 256..319   (resident for the whole CTA)
//     dK  += dS^T Q     A = dS^T(K-major)   B = Q   (MN-major)       TMEM cols 320..383   (resident)
//     dQ_j = dS   K     A = dS^T tile read MN-MAJOR, B = K (MN-major)  TMEM cols 384..447 -> fp32 red.add into the dQ workspace
//   warp 0: TMA producer (K, V once; Q_j, dO_j double-buffered) + per-tile L / delta staging; warp 1: tcgen05.mma issuer;
//   warps 2..9: 256 softmax threads (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4);
//   warps 10..13: dQ drain (one per TMEM lane quarter): TMEM -> swizzled staging -> bulk TMA reduce-add, off the softmax path.
// dK and dV leave as bf16 straight into the dQKV matrix the QKV dgrad/wgrad GEMMs consume (RoPE's transpose applied to dK on the
// way out); dQ is accumulated across the key-tile CTAs in fp32 and converted (+ RoPE transpose) by attn_dq_finish_kernel.
#include "common.cuh"
#include "dropout.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int AB_T = 128;                       // query tile == key tile
constexpr int AB_THREADS = 448;  // 14 warps: four of them share an SM sub-partition's 16 K registers, hence the 128-register cap (a __maxnreg__(144)
                                  // build — 64512 registers per CTA on paper — fails to launch); the softmax threads spill ~70 loads / stores
constexpr uint32_t AB_TILE = AB_T * 64 * 2;     // 16 KB: [128 rows x 64 d] bf16, SW128
constexpr uint32_t AB_PT = AB_T * AB_T * 2;     // 32 KB: [2 q-atoms][128 keys x 64 q] bf16, SW128
constexpr uint32_t AB_STG = 8 * 4096;            // dQ staging: one [32 rows x 32 f32] SW128 box per softmax warp
constexpr uint32_t AB_MASK = 2 * 4 * AB_T * 4;   // SDPA-dropout keep bits of two tiles: [2][4 query chunks of 32][128 keys] words
// Q / dO (+ L, delta) ring depth.  Three, not two: with two, the load of tile j+1 can only be issued once the products of tile j-1
// have retired (same stage), i.e. about one TMA round trip before S^T_{j+1} is wanted — the in-order MMA thread then sat ~640 clocks
// per tile inside issue_sdp(j+1) waiting for the data, with the products of tile j (whose operands were ready) queued behind it.
constexpr int AB_NS = 3;
constexpr uint32_t AB_SMEM = 2 * AB_TILE /*K,V*/ + 2 * AB_NS * AB_TILE /*Q,dO ring*/ + AB_PT /*dS^T*/ + AB_STG + 2 * AB_NS * AB_T * 4 /*L, delta*/ + AB_MASK + 256 + 1024;
constexpr uint32_t AB_TMEM_COLS = 512;

struct AttnBwdParams {
  const float* lse;     // [B, H, n] log2 domain; +inf for padded query rows
  const float* delta;   // [B, H, n]
  float* dq;            // [B*n, H*64] fp32, zero-initialised
  __nv_bfloat16* dqkv;  // [B*n, ld_d]: dK at column H*64, dV at 2*H*64
  int ld_d;
  const int32_t* lens;
  int lens_mod, B, H, n;
  float scale, scale_log2;
  const float* rope;    // [n, 32] (cos, sin)
  int rope_heads;
  long long* trace;     // debug only (AB_TRACE builds)
  AttnDrop dr;          // DROP kernels: the forward's SDPA dropout mask stream (regenerated here, never stored)
  int n8;               // ceil(n / 8)
};

#ifndef AB_POLY
#define AB_POLY 0  // of the 8 groups of 4 scores in a 32-column chunk, this many evaluate their second pair of exponentials on the FMA pipe
#endif
// 2^x on the FMA pipe (x <= ~100): n = round(x) through the 1.5 * 2^23 magic add, r = x - n in [-0.5, 0.5], degree-3 minimax
// polynomial of 2^r (max relative error 7.5e-5, far below the bf16 rounding of P), exponent spliced in as an integer multiply-add
__device__ __forceinline__ float2 ab_exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 xf = __fadd2_rn(x, magic);
  const float2 nn = __fadd2_rn(xf, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __ffma2_rn(nn, make_float2(-1.0f, -1.0f), x);
  float2 q = __ffma2_rn(make_float2(0.0551716648f, 0.0551716648f), r, make_float2(0.2426111251f, 0.2426111251f));
  q = __ffma2_rn(q, r, make_float2(0.6932609677f, 0.6932609677f));
  q = __ffma2_rn(q, r, make_float2(0.9999280572f, 0.9999280572f));
  float2 e;
  e.x = __int_as_float(__float_as_int(xf.x) * (1 << 23) + __float_as_int(q.x));
  e.y = __int_as_float(__float_as_int(xf.y) * (1 << 23) + __float_as_int(q.y));
  return e;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// DROP: the forward multiplied the normalised probabilities by a mask m = keep / (1-p) before P.V, so
//   dV = (P . m)^T dO,   dS = scale * P . (m . dP - delta)   with delta = rowsum(dO . O) of the dropped forward's O.
// With P' = P / (1-p) (the producer warp stages L - log2(1/(1-p))) and delta' = delta (1-p) this is dV = (P' . keep)^T dO,
// dS = scale * P' . (keep . dP - delta'): the softmax threads only need the keep BIT of their elements.  The forward's Philox blocks
// cover 8 keys of one query, while a thread here owns one key and 64 queries, so the four dQ-drain warps regenerate the tile's bits
// (lane = query, one block per key octet) and transpose them across the warp into shared memory words [query chunk][key] whose
// bit e is query e of the chunk: every block is computed once per CTA, and a softmax thread reads two words per tile.
template <bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdKV, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
#ifdef AB_TRACE
  const long long t_entry = clock64();  // CTA timeline (trace[8..13], written by softmax warp 2 of the sampled CTA)
  long long t_setup = 0, t_sdp0 = 0, t_pds0 = 0, t_loop = 0;
#endif
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;            // AB_NS stages
  uint8_t* sdO = sQ + AB_NS * AB_TILE;   // AB_NS stages
  uint8_t* sdST = sdO + AB_NS * AB_TILE;
  uint8_t* sStg = sdST + AB_PT;                         // [8][4096]
  float* sL = reinterpret_cast<float*>(sStg + AB_STG);  // [AB_NS][128]
  float* sDl = sL + AB_NS * AB_T;                       // [AB_NS][128]
  uint32_t* sMask = reinterpret_cast<uint32_t*>(sDl + AB_NS * AB_T);  // [2][4][128] (DROP only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + AB_MASK / 4);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_sdp = bars + 1;     // S^T_j and dP^T_j in TMEM
  uint64_t* bar_pds = bars + 2;     // P^T_j, dS^T_j in smem; S / dP / dQ TMEM drained (256 arrivals)
  uint64_t* bar_dq = bars + 3;      // dV, dK, dQ_j products retired
  uint64_t* bar_sfree = bars + 4;   // S^T_j / dP^T_j pulled into registers by all 256 threads
  uint64_t* bar_dqfree = bars + 5;  // dQ_j pulled out of TMEM by the 4 drain warps (128 arrivals)
  uint64_t* bar_mask = bars + 6;    // [2] DROP: keep bits of the tile that uses buffer s are in shared memory (128 arrivals)
  uint64_t* bar_qdo = bars + 8;             // [AB_NS] Q_j, dO_j landed
  uint64_t* bar_ld = bar_qdo + AB_NS;       // [AB_NS] L_j, delta_j staged (32 arrivals)
  uint64_t* bar_free = bar_ld + AB_NS;      // [AB_NS] products of the tile that used stage s retired -> Q / dO / L / delta of that stage reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_free + AB_NS);
  static_assert((8 + 3 * AB_NS) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * AB_T;
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  int kvlen = p.n;
  if (p.lens != nullptr) kvlen = min(p.n, __ldg(p.lens + (p.lens_mod > 0 ? b % p.lens_mod : b)));
  const int D = p.H * 64;

  if (kvlen <= 0 || k0 >= kvlen) {
    // masked keys receive no gradient
    griddep_wait();
    if (warp >= 2 && warp < 10) {
      const int t = threadIdx.x - 64;  // 0..255
      const int row = t >> 1, half = t & 1;
      const int pos = k0 + row;
      if (pos < p.n) {
        __nv_bfloat16* base = p.dqkv + ((size_t)b * p.n + pos) * p.ld_d + h * 64 + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          reinterpret_cast<uint4*>(base + D)[i] = make_uint4(0u, 0u, 0u, 0u);
          reinterpret_cast<uint4*>(base + 2 * D)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    return;
  }
  const int Tq = (kvlen + AB_T - 1) / AB_T;  // query rows >= len carry dO = 0 and L = +inf: skipped

  if (warp == 1) {
    if (lane == 0) {
      mbar_init(bar_kv, 1);
      for (int i = 0; i < AB_NS; ++i) {
        mbar_init(&bar_qdo[i], 1);
        mbar_init(&bar_ld[i], 32);
        mbar_init(&bar_free[i], 1);
      }
      mbar_init(bar_sdp, 1);
      mbar_init(bar_pds, 256);
      mbar_init(bar_dq, 1);
      mbar_init(bar_sfree, 256);
      mbar_init(bar_dqfree, 128);
      mbar_init(&bar_mask[0], 128);
      mbar_init(&bar_mask[1], 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, AB_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // PDL (common.cuh): prologue under the previous kernel's tail
  griddep_launch_dependents();
#ifdef AB_TRACE
  t_setup = clock64();
#endif
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dV = tmem_base + 256, tm_dK = tmem_base + 320, tm_dQ = tmem_base + 384,
                 tm_PT = tmem_base + 448;  // P^T as bf16 pairs: 64 columns = 128 queries (A operand of dV, read from tensor memory)

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------------ producer
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmK);
      prefetch_tmap(&tmV);
      prefetch_tmap(&tmdO);
      prefetch_tmap(&tmdKV);
      mbar_arrive_expect_tx(bar_kv, 2 * AB_TILE);
      tma_load_3d(sK, &tmK, bar_kv, h * 64, k0, b);
      tma_load_3d(sV, &tmV, bar_kv, h * 64, k0, b);
    }
    const size_t row0 = (size_t)bh * p.n;
    for (int j = 0; j < Tq; ++j) {
      const int st = j % AB_NS;
      if (j >= AB_NS) mbar_wait(&bar_free[st], ((j / AB_NS) - 1) & 1);  // tile j - AB_NS retired (a per-stage barrier cannot run a phase ahead)
      if (elect_one()) {  // (not `lane == 0`: plain UTMALDG instead of a per-instruction ELECT / BRA.U.ANY loop, see tile_engine.cuh)
        mbar_arrive_expect_tx(&bar_qdo[st], 2 * AB_TILE);
        tma_load_3d(sQ + st * AB_TILE, &tmQ, &bar_qdo[st], h * 64, j * AB_T, b);
        tma_load_3d(sdO + st * AB_TILE, &tmdO, &bar_qdo[st], h * 64, j * AB_T, b);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 32 + lane;
        const int pos = j * AB_T + r;
        float lv = pos < p.n ? __ldg(p.lse + row0 + pos) : INFINITY;
        float dlv = pos < p.n ? __ldg(p.delta + row0 + pos) * p.scale : 0.f;  // pre-scaled: dS = P (dP scale - delta scale)
        if constexpr (DROP) {  // P' = P / (1-p), delta' = delta (1-p): see the kernel's header
          lv -= p.dr.log2_scale;
          dlv *= 1.f / p.dr.scale;
        }
        sL[st * AB_T + r] = lv;
        sDl[st * AB_T + r] = dlv;
      }
      mbar_arrive(&bar_ld[st]);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------------ MMA issuer
    // the issuing lane is chosen by elect.sync, not `lane == 0`: ptxas then emits plain UTCHMMA sequences (32 small MMAs per tile pair
    // here; under a lane predicate each sat in its own R2UR + ELECT + BRA.U.ANY loop, ~65 clocks apiece against 32-64 of execution)
    if (elect_one()) {
      const uint32_t id_sq = idesc_bf16(128, 128, 0, 0);   // S^T, dP^T
      const uint32_t id_acc = idesc_bf16(128, 64, 0, 1);   // dV, dK: B is MN-major
      const uint32_t id_dq = idesc_bf16(128, 64, 1, 1);    // dQ: A (dS^T tile) and B (K) both MN-major
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), do_addr = smem_u32(sdO);
      const uint32_t ds_addr = smem_u32(sdST);
      auto issue_sdp = [&](int j) {
        const int st = j % AB_NS;
        mbar_wait(&bar_qdo[st], (j / AB_NS) & 1);
        tc_fence_after();
        const uint32_t qa = q_addr + st * AB_TILE, da = do_addr + st * AB_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm_S, smem_desc_sw128(k_addr + k * 32, 1024, 16), smem_desc_sw128(qa + k * 32, 1024, 16), id_sq, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm_dP, smem_desc_sw128(v_addr + k * 32, 1024, 16), smem_desc_sw128(da + k * 32, 1024, 16), id_sq, k != 0);
        umma_commit(bar_sdp);
      };
      mbar_wait(bar_kv, 0);
      issue_sdp(0);
#ifdef AB_TRACE
      long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      long long tprev = clock64();
#define AB_MARK(i) { const long long tn = clock64(); tr[i] += tn - tprev; tprev = tn; }
#else
#define AB_MARK(i)
#endif
      for (int j = 0; j < Tq; ++j) {
        const int st = j % AB_NS;
        // S^T / dP^T of the next tile as soon as this tile's scores sit in registers (overlaps the exponentials)
        mbar_wait(bar_sfree, j & 1);
        tc_fence_after();
        AB_MARK(0)
        if (j + 1 < Tq) issue_sdp(j + 1);
        AB_MARK(1)
        mbar_wait(bar_pds, j & 1);
        tc_fence_after();
        AB_MARK(2)
        const uint32_t qa = q_addr + st * AB_TILE, da = do_addr + st * AB_TILE;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // reduction over the 128 queries of the tile, 16 per step
          // A = P^T from TENSOR MEMORY (lane = key, a K-step of 16 queries = 8 packed columns): no P^T stores / operand reads on the
          // shared-memory port, which is what bounds this kernel
          umma_bf16_ts(tm_dV, tm_PT + kk * 8, smem_desc_sw128(da + kk * 2048, 1024, 8192), id_acc, (j | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t aoff = (kk >> 2) * (AB_PT / 2) + (kk & 3) * 32;
          umma_bf16(tm_dK, smem_desc_sw128(ds_addr + aoff, 1024, 16), smem_desc_sw128(qa + kk * 2048, 1024, 8192), id_acc, (j | kk) != 0);
        }
        if (j > 0) {
          mbar_wait(bar_dqfree, (j - 1) & 1);  // dQ_{j-1} has left TMEM
          tc_fence_after();
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // reduction over the 128 keys: dS^T rows are the K dimension here (MN-major A, 2 q-atoms)
          umma_bf16(tm_dQ, smem_desc_sw128(ds_addr + kk * 2048, 1024, AB_PT / 2), smem_desc_sw128(k_addr + kk * 2048, 1024, 8192), id_dq,
                    kk != 0);
        umma_commit(bar_dq);
        umma_commit(&bar_free[st]);
        AB_MARK(3)
#ifdef AB_TRACE
        mbar_wait(bar_dq, j & 1);  // measurement only: serialises the issue loop
        AB_MARK(4)
#endif
      }
#ifdef AB_TRACE
      if (p.trace != nullptr && blockIdx.x == 3 && blockIdx.y == 37) {
        for (int i = 0; i < 5; ++i) p.trace[i] = tr[i];
        p.trace[6] = Tq;
      }
#endif
    }
    __syncwarp();
  } else if (warp >= 10) {
    // ------------------------------------------------------------------------------------------------ dQ drain warps
    // dQ_j leaves through a swizzled staging box and bulk TMA reduce-adds (fp32, accumulated in L2 at full-line granularity; rows
    // past the utterance are clipped by the 3-D tensor map).  Per-thread red.global of the same data scattered 32 half-sectors
    // per warp instruction; draining from the softmax warps put ~1000 cycles per tile on the critical path.
    const int lq = warp & 3;
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    uint8_t* my_stg = sStg + (warp - 10) * 8192;  // two [32 rows x 32 f32] SW128 boxes
    // DROP: keep bits of query chunk (warp - 10) of tile jt for the CTA's 128 keys -> sMask[jt & 1][warp - 10][key].  Lane = query:
    // four Philox blocks give the lane its 32 keep bits of a 32-key block (row q of a 32 x 32 bit matrix), five shuffle stages
    // transpose the matrix across the warp, and lane k stores the word of key k.  (A first version used one warp ballot per key:
    // 128 ballots per tile whose lane-select chains ran at 0.2 IPC and made this the kernel's critical path, 18 -> 32 ms.)
    auto gen_mask = [&](int jt) {
      const int qc = warp - 10;
      const uint64_t g = ((uint64_t)bh * p.n + (uint64_t)(jt * AB_T + qc * 32 + lane)) * p.n8 + (uint64_t)(k0 >> 3);
      uint32_t* dst = sMask + ((jt & 1) * 4 + qc) * AB_T;
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb) {
        uint32_t row = 0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          uint32_t w0, w1;
          attn_drop_words(p.dr, g + kb * 4 + o, w0, w1);
          row |= attn_drop_nibble(w0) << (8 * o);
          row |= attn_drop_nibble(w1) << (8 * o + 4);
        }
        dst[kb * 32 + lane] = warp_bit_transpose(row, lane);
      }
      mbar_arrive(&bar_mask[jt & 1]);  // release: this lane's words; 128 arrivals complete the buffer
    };
    if constexpr (DROP) {
      gen_mask(0);
      if (Tq > 1) gen_mask(1);
    }
    for (int j = 0; j < Tq; ++j) {
      mbar_wait(bar_dq, j & 1);
      tc_fence_after();
      uint32_t a0[32], a1[32];
      tmem_ld32(tm_dQ + lane_addr, a0);
      tmem_ld32(tm_dQ + lane_addr + 32, a1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_dqfree);
      if (elect_one()) bulk_wait_read0();  // the previous boxes have been read out of the staging buffer
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int sw = (q ^ (lane & 7)) << 4;
        *reinterpret_cast<uint4*>(my_stg + lane * 128 + sw) = make_uint4(a0[4 * q], a0[4 * q + 1], a0[4 * q + 2], a0[4 * q + 3]);
        *reinterpret_cast<uint4*>(my_stg + 4096 + lane * 128 + sw) = make_uint4(a1[4 * q], a1[4 * q + 1], a1[4 * q + 2], a1[4 * q + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
#ifndef AB_EXPERIMENT_NO_DQ_REDUCE
      if (elect_one()) {
        tma_reduce_add_3d(&tmdQ, my_stg, h * 64, j * AB_T + lq * 32, b);
        tma_reduce_add_3d(&tmdQ, my_stg + 4096, h * 64 + 32, j * AB_T + lq * 32, b);
        bulk_commit();
      }
#endif
      // bar_dq(j) above implies that every softmax thread has finished with the bits of tile j: their buffer takes tile j + 2
      if constexpr (DROP) {
        if (j + 2 < Tq) gen_mask(j + 2);
      }
    }
    if (elect_one()) bulk_wait0();  // the reductions have landed before the CTA retires its shared memory
  } else {
    // ------------------------------------------------------------------------------------------------ softmax / gradient threads
    const int lq = warp & 3;          // TMEM lane quarter this warp may touch
    const int ch = (warp - 2) >> 2;   // column half
    const int r = lq * 32 + lane;     // TMEM lane: key row (S^T, dP^T, dV, dK)
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    const int rx = r & 7;
    const bool key_ok = (k0 + r) < kvlen;
    const float c2 = p.scale_log2, sc = p.scale;
    uint8_t* ds_row = sdST + ch * (AB_PT / 2) + r * 128;

    for (int j = 0; j < Tq; ++j) {
      const int st = j % AB_NS;
      mbar_wait(&bar_ld[st], (j / AB_NS) & 1);
      mbar_wait(bar_sdp, j & 1);
      tc_fence_after();
#ifdef AB_TRACE
      if (j == 0) t_sdp0 = clock64();
#endif
      const float4* L4 = reinterpret_cast<const float4*>(sL + st * AB_T + ch * 64);
      const float4* D4 = reinterpret_cast<const float4*>(sDl + st * AB_T + ch * 64);
      uint32_t ppk[32], dpk[32];  // P^T and dS^T rows of this thread, packed bf16
      uint32_t kb0 = 0xffffffffu, kb1 = 0xffffffffu;  // DROP: keep bits of this key for the 2 x 32 queries of this column half
      if constexpr (DROP) {
        mbar_wait(&bar_mask[j & 1], (j >> 1) & 1);
        kb0 = sMask[((j & 1) * 4 + ch * 2) * AB_T + r];
        kb1 = sMask[((j & 1) * 4 + ch * 2 + 1) * AB_T + r];
      }
      auto half = [&](const uint32_t (&sv)[32], const uint32_t (&gv)[32], int c) {
        const uint32_t kbits = c ? kb1 : kb0;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 l = L4[c * 8 + q4];
          const float4 dl = D4[c * 8 + q4];  // delta * scale
          const float ls[4] = {l.x, l.y, l.z, l.w};
          const float dls[4] = {dl.x, dl.y, dl.z, dl.w};
          float pv[4], dv[4];
          if (AB_POLY > 0 && ((q4 + 1) * AB_POLY) / 8 != (q4 * AB_POLY) / 8) {  // (compile-time: q4 is an unrolled index)
            const float2 e2 = ab_exp2_poly2(make_float2(fmaf(__uint_as_float(sv[q4 * 4 + 2]), c2, -ls[2]),
                                                        fmaf(__uint_as_float(sv[q4 * 4 + 3]), c2, -ls[3])));
            pv[2] = e2.x;
            pv[3] = e2.y;
          } else {
            pv[2] = ex2_approx(fmaf(__uint_as_float(sv[q4 * 4 + 2]), c2, -ls[2]));
            pv[3] = ex2_approx(fmaf(__uint_as_float(sv[q4 * 4 + 3]), c2, -ls[3]));
          }
          pv[0] = ex2_approx(fmaf(__uint_as_float(sv[q4 * 4 + 0]), c2, -ls[0]));
          pv[1] = ex2_approx(fmaf(__uint_as_float(sv[q4 * 4 + 1]), c2, -ls[1]));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = q4 * 4 + i;
            if constexpr (DROP) {
              const bool keep = (kbits >> e) & 1u;  // pv is P' = P / (1-p), dls is delta' (staged by the producer warp)
              dv[i] = pv[i] * fmaf(keep ? __uint_as_float(gv[e]) : 0.f, sc, -dls[i]);
              if (!keep) pv[i] = 0.f;  // P^T that feeds dV is the dropped one
            } else {
              dv[i] = pv[i] * fmaf(__uint_as_float(gv[e]), sc, -dls[i]);
            }
          }
          ppk[c * 16 + q4 * 2] = pack_bf16(pv[0], pv[1]);
          ppk[c * 16 + q4 * 2 + 1] = pack_bf16(pv[2], pv[3]);
          dpk[c * 16 + q4 * 2] = pack_bf16(dv[0], dv[1]);
          dpk[c * 16 + q4 * 2 + 1] = pack_bf16(dv[2], dv[3]);
        }
      };
      // the scores are pulled into registers in two halves; once the second half is in, the TMEM buffers go back to the tensor
      // pipe, which computes S^T / dP^T of tile j+1 while the exponentials run
      {
        uint32_t sv[32], gv[32];
        tmem_ld32(tm_S + lane_addr + ch * 64, sv);
        tmem_ld32(tm_dP + lane_addr + ch * 64, gv);
        tmem_ld_wait();
        half(sv, gv, 0);
      }
      {
        uint32_t sv[32], gv[32];
        tmem_ld32(tm_S + lane_addr + ch * 64 + 32, sv);
        tmem_ld32(tm_dP + lane_addr + ch * 64 + 32, gv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_sfree);
        half(sv, gv, 1);
      }
      if (!key_ok) {  // masked key row (tail of the utterance): no probability mass, no gradient
#pragma unroll
        for (int i = 0; i < 32; ++i) ppk[i] = dpk[i] = 0u;
      }
      if (j > 0) mbar_wait(bar_dq, (j - 1) & 1);  // products of tile j-1 retired: the P^T / dS^T buffers are free
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int chunk = (q ^ rx) << 4;
        *reinterpret_cast<uint4*>(ds_row + chunk) = make_uint4(dpk[q * 4], dpk[q * 4 + 1], dpk[q * 4 + 2], dpk[q * 4 + 3]);
      }
      tmem_st32(tm_PT + lane_addr + ch * 32, ppk);  // this thread's 64 probabilities (its key row, its column half)
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pds);
#ifdef AB_TRACE
      if (j == 0) t_pds0 = clock64();
#endif
    }
    mbar_wait(bar_dq, (Tq - 1) & 1);
    tc_fence_after();
#ifdef AB_TRACE
    t_loop = clock64();
#endif
    // dV, dK of this key tile (all products retired: bar_dq of the last tile) leave as bf16 through swizzled staging tiles (the
    // dS^T buffer is free now) and two TMA stores straight into the dQKV matrix; rows past the utterance are clipped by the 3-D map.
    // (Per-thread 16-byte stores scattered 32 half-sectors per warp instruction: the drain took ~3700 clocks of a ~36 000-clock CTA.)
    const int pos = k0 + r;
    uint8_t* stV = sdST + r * 128;
    uint8_t* stK = sdST + AB_TILE + r * 128;
    {
      uint32_t a[32];
      tmem_ld32(tm_dV + lane_addr + ch * 32, a);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pk;
        pk.x = pack_bf16(__uint_as_float(a[q * 8 + 0]), __uint_as_float(a[q * 8 + 1]));
        pk.y = pack_bf16(__uint_as_float(a[q * 8 + 2]), __uint_as_float(a[q * 8 + 3]));
        pk.z = pack_bf16(__uint_as_float(a[q * 8 + 4]), __uint_as_float(a[q * 8 + 5]));
        pk.w = pack_bf16(__uint_as_float(a[q * 8 + 6]), __uint_as_float(a[q * 8 + 7]));
        *reinterpret_cast<uint4*>(stV + (((ch * 4 + q) ^ rx) << 4)) = pk;
      }
      tmem_ld32(tm_dK + lane_addr + ch * 32, a);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]);
      if (h < p.rope_heads && pos < p.n) {
        // transpose of the forward rotation (y0 = x0 c - x1 s, y1 = x1 c + x0 s)
        const float4* cs = reinterpret_cast<const float4*>(p.rope) + (size_t)pos * 16 + ch * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = __ldg(cs + i);
          const float y0 = v[4 * i], y1 = v[4 * i + 1], y2 = v[4 * i + 2], y3 = v[4 * i + 3];
          v[4 * i] = y0 * t.x + y1 * t.y;
          v[4 * i + 1] = y1 * t.x - y0 * t.y;
          v[4 * i + 2] = y2 * t.z + y3 * t.w;
          v[4 * i + 3] = y3 * t.z - y2 * t.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pk;
        pk.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
        pk.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
        pk.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
        pk.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(stK + (((ch * 4 + q) ^ rx) << 4)) = pk;
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 softmax warps: both staging tiles are complete
    if (warp == 2 && elect_one()) {
      tma_store_3d(&tmdKV, sdST, 2 * D + h * 64, k0, b);
      tma_store_3d(&tmdKV, sdST + AB_TILE, D + h * 64, k0, b);
      bulk_commit();
      bulk_wait_read0();  // the staging tiles have been read before the CTA retires its shared memory
    }
    tc_fence_before();
#ifdef AB_TRACE
    if (p.trace != nullptr && blockIdx.x == 3 && blockIdx.y == 37 && warp == 2 && lane == 0) {
      const long long t_end = clock64();
      p.trace[8] = t_setup - t_entry;   // barrier init, TMEM alloc, __syncthreads, griddepcontrol.wait
      p.trace[9] = t_sdp0 - t_setup;    // K / V / Q / dO loads + S^T_0, dP^T_0
      p.trace[10] = t_pds0 - t_sdp0;    // first softmax pass
      p.trace[11] = t_loop - t_pds0;    // the rest of the loop up to the last products
      p.trace[12] = t_end - t_loop;     // dV / dK drain
      p.trace[13] = t_end - t_entry;
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AB_TMEM_COLS);
  }
}

// delta[b, h, pos] = sum_c dO[row, h*64 + c] * O[row, h*64 + c]   (one warp per token row, all heads)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, int ld, float* __restrict__ delta,
                                  int B, int H, int n) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * n) return;
  const int b = row / n, pos = row - b * n;
  const __nv_bfloat162* po = reinterpret_cast<const __nv_bfloat162*>(o + (size_t)row * ld);
  const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(dout + (size_t)row * ld);
  for (int h = 0; h < H; ++h) {
    const float2 a = __bfloat1622float2(po[h * 32 + lane]);
    const float2 g = __bfloat1622float2(pd[h * 32 + lane]);
    float s = a.x * g.x + a.y * g.y;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) delta[((size_t)b * H + h) * n + pos] = s;
  }
}

// dQ fp32 accumulator -> bf16 columns [0, D) of dQKV, with the transpose of RoPE on the first rope_heads heads
__global__ void attn_dq_finish_kernel(const float* __restrict__ dq, __nv_bfloat16* __restrict__ dqkv, int ld_d, const float* __restrict__ rope,
                                      int rope_heads, long long rows, int n, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread = 8 consecutive columns
  const int per_row = D >> 3;
  if (i >= rows * per_row) return;
  const long long row = i / per_row;
  const int c0 = (int)(i - row * per_row) * 8;
  const float4 a = *reinterpret_cast<const float4*>(dq + row * D + c0);
  const float4 c = *reinterpret_cast<const float4*>(dq + row * D + c0 + 4);
  float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
  if ((c0 >> 6) < rope_heads) {
    const int pos = (int)(row % n);
    const float4* cs = reinterpret_cast<const float4*>(rope) + (size_t)pos * 16 + ((c0 & 63) >> 2);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float4 t = __ldg(cs + k);
      const float y0 = v[4 * k], y1 = v[4 * k + 1], y2 = v[4 * k + 2], y3 = v[4 * k + 3];
      v[4 * k] = y0 * t.x + y1 * t.y;
      v[4 * k + 1] = y1 * t.x - y0 * t.y;
      v[4 * k + 2] = y2 * t.z + y3 * t.w;
      v[4 * k + 3] = y3 * t.z - y2 * t.w;
    }
  }
  uint4 pk;
  pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]); pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(dqkv + row * ld_d + c0) = pk;
}

long long* g_attn_bwd_trace = nullptr;

int attn_bwd(const void* q, const void* k, const void* v, int ld, const void* out, const void* dout, int ld_o, const float* lse,
             float* delta, float* dq_ws, void* dqkv, int ld_d, const int32_t* lens, int lens_mod, int B, int H, int n, float scale,
             const float* rope, int rope_heads, cudaStream_t stream, const AttnDrop* drop) {
  F5B_CHECK(q && k && v && out && dout && lse && delta && dq_ws && dqkv, "f5b_attn_bwd: null pointer");
  F5B_CHECK(B > 0 && H > 0 && n > 0 && ld >= H * 64 && (ld & 7) == 0 && ld_o >= H * 64 && (ld_o & 7) == 0 && ld_d >= 3 * H * 64 && (ld_d & 7) == 0,
            "f5b_attn_bwd: bad shape B %d H %d n %d ld %d ld_o %d ld_d %d", B, H, n, ld, ld_o, ld_d);
  F5B_CHECK(rope_heads == 0 || rope != nullptr, "f5b_attn_bwd: rope table missing");
  const int D = H * 64;
  const long long rows = (long long)B * n;
  LaunchScope scope(K_ATTN, stream, 10.0 * B * H * (double)n * n * 64, 2.0 * 8 * B * H * (double)n * 64 + 4.0 * rows * D * 2, 3);
  F5B_CUDA(cudaMemsetAsync(dq_ws, 0, sizeof(float) * rows * D, stream));
  attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(out),
                                                                    reinterpret_cast<const __nv_bfloat16*>(dout), ld_o, delta, B, H, n);
  F5B_CUDA(cudaGetLastError());
  CUtensorMap tmQ, tmK, tmV, tmdO, tmdQ, tmdKV;
  const uint64_t hw = (uint64_t)H * 64, pitch = (uint64_t)ld * 2, pitch_o = (uint64_t)ld_o * 2;
  if (make_tmap_3d(&tmQ, q, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, AB_T, 1, true)) return -1;
  if (make_tmap_3d(&tmK, k, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, AB_T, 1, true)) return -1;
  if (make_tmap_3d(&tmV, v, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, AB_T, 1, true)) return -1;
  if (make_tmap_3d(&tmdO, dout, 2, hw, (uint64_t)n, (uint64_t)B, pitch_o, (uint64_t)n * pitch_o, 64, AB_T, 1, true)) return -1;
  if (make_tmap_3d(&tmdQ, dq_ws, 4, hw, (uint64_t)n, (uint64_t)B, hw * 4, (uint64_t)n * hw * 4, 32, 32, 1, true)) return -1;
  // dK / dV leave through this map over the whole [B, n, 3D] dQKV matrix ([128 rows x 64 columns] boxes at column D + 64h / 2D + 64h)
  if (make_tmap_3d(&tmdKV, dqkv, 2, 3 * hw, (uint64_t)n, (uint64_t)B, (uint64_t)ld_d * 2, (uint64_t)n * ld_d * 2, 64, AB_T, 1, true)) return -1;
  static bool configured = false;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM));
    F5B_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM));
    configured = true;
  }
  AttnBwdParams p;
  p.lse = lse;
  p.delta = delta;
  p.dq = dq_ws;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.ld_d = ld_d;
  p.lens = lens;
  p.lens_mod = lens_mod;
  p.B = B;
  p.H = H;
  p.n = n;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.rope = rope;
  p.rope_heads = rope_heads;
  p.trace = g_attn_bwd_trace;
  p.dr = drop ? *drop : AttnDrop{0u, 1.f, 0.f, 0u, 0u};
  p.n8 = (n + 7) / 8;
  dim3 grid((n + AB_T - 1) / AB_T, B * H);
  if (p.dr.addc != 0) F5B_CUDA(launch_dep(attn_bwd_kernel<true>, grid, dim3(AB_THREADS), AB_SMEM, stream, 1, tmQ, tmK, tmV, tmdO, tmdQ, tmdKV, p));
  else F5B_CUDA(launch_dep(attn_bwd_kernel<false>, grid, dim3(AB_THREADS), AB_SMEM, stream, 1, tmQ, tmK, tmV, tmdO, tmdQ, tmdKV, p));
  F5B_CUDA(cudaGetLastError());
  const long long items = rows * (D >> 3);
  attn_dq_finish_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(dq_ws, p.dqkv, ld_d, rope, rope_heads, rows, n, D);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f5b

extern "C" int f5b_attn_bwd(const void* q, const void* k, const void* v, int ld, const void* out, const void* dout, int ld_o,
                            const float* lse, float* delta_ws, float* dq_ws, void* dqkv, int ld_d, const int32_t* lens, int lens_mod,
                            int B, int H, int n, float scale, const float* rope, int rope_heads, f5b_stream_t stream) {
  return f5b::attn_bwd(q, k, v, ld, out, dout, ld_o, lse, delta_ws, dq_ws, dqkv, ld_d, lens, lens_mod, B, H, n, scale, rope, rope_heads,
                       static_cast<cudaStream_t>(stream));
}

extern "C" void f5b_debug_set_attn_bwd_trace(long long* buf) { f5b::g_attn_bwd_trace = buf; }
