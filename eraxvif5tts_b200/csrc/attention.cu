// Flash-style attention forward for the DiT blocks: O = softmax(Q K^T / sqrt(64), keys < len[b]) V, non-causal, d_head 64.
// Replaces AttnProcessor's mask expansion + F.scaled_dot_product_attention + head merge
// (/root/reference/src/f5_tts/model/modules.py:483-493; dropout_p = 0, see DESIGN.md "oracle adjustments").
//
// sm_100a design (one CTA = one 128-query tile of one (batch, head); two CTAs co-resident per SM):
//   warp 4 (one elected lane): TMA producer + tcgen05.mma issuer.  KV is consumed in tiles of 64 keys:
//       S_j[128 x 64] = Q K_j^T    (K-major bf16 tiles, 128B swizzle; fp32 accumulator in TMEM, DOUBLE-buffered: S_{j+1}
//                                   and S_{j+2} are computed while the softmax warps still work on S_j)
//       O[128 x 80]  += P_j V'_j   (P_j written to swizzled smem (double-buffered) by the softmax warps; V' = V^T from the
//                                   QKV epilogue's transposed store plus a constant row of ones, so column 64 of O
//                                   accumulates the softmax row sum on the tensor pipe; O stays resident in TMEM)
//   warps 0..3: online softmax, thread = query row (tcgen05.ld 32x32b -> no cross-lane reductions), exp2 with the
//       1/sqrt(d)*log2(e) scale folded in.  The running maximum is updated lazily: O (and with it the row sum) is rescaled in
//       TMEM only when a row's maximum grew by more than 2^8, so most KV tiles cost no accumulator round trip and the
//       softmax warps never wait for the tensor pipe in steady state (the kernel is MUFU.EX2-bound).
//   K and V' live in 3-stage TMA rings.  Key-padding is a per-batch length bound: KV tiles past len[b] are never loaded,
//   the last tile is masked by index.
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
constexpr int ATT_SM_WARPS = 4;   // softmax warps: one per TMEM lane quadrant, thread = query row
constexpr int ATT_THREADS = (ATT_SM_WARPS + 1) * 32;
constexpr int ATT_NV = 80;                                  // 64 value columns + the ones row + zero padding to N % 16 == 0
constexpr int ATT_KV_STAGES = 3;
constexpr uint32_t ATT_Q_BYTES = ATT_BQ * 64 * 2;          // 16 KB
constexpr uint32_t ATT_K_BYTES = ATT_BKV * 64 * 2;         // 8 KB per stage
constexpr uint32_t ATT_V_BYTES = ATT_NV * 128;             // 10 KB per stage: 80 rows x (64 kv x 2 B)
constexpr uint32_t ATT_V_TX = 64 * 128;                    // bytes TMA writes per V' tile (the 64 real rows)
constexpr uint32_t ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;     // 16 KB per buffer, 2 buffers
#ifndef ATT_EXTRA_SMEM
#define ATT_EXTRA_SMEM 0  // experiments: pad shared memory to force one CTA per SM
#endif
constexpr uint32_t ATT_SMEM = ATT_Q_BYTES + ATT_KV_STAGES * (ATT_K_BYTES + ATT_V_BYTES) + 2 * ATT_P_BYTES + 1024 + 256 + ATT_EXTRA_SMEM;
constexpr uint32_t ATT_TMEM_COLS = 256;  // S0: 0..63, S1: 64..127, O: 128..207
constexpr float ATT_RESCALE_LOG2 = 8.0f;

struct AttnParams {
  __nv_bfloat16* out;
  const int32_t* lens;
  int lens_mod, B, H, n;
  float scale_log2;
  long long* trace;  // debug only (ATT_TRACE builds)
};

// 2^x on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max rel. error 7.5e-5, far below
// the bf16 rounding of P): a fixed fraction of the exponentials is computed this way so the MUFU pipe, the bottleneck of the
// softmax, gets fewer of them.  Valid for x <= ~100; inputs below -126 are clamped (result 2^-126 instead of 0).
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05517143756151199f, 0.24261081218719482f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999281167984009f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#ifndef ATT_POLY_MASK
#define ATT_POLY_MASK 0x00  // which of every 8 consecutive columns use exp2_poly (bit i = column i).  Measured on B200
                            // (profiles/r01_attention_notes.md): 0x00 577, 0x88 543, 0x92 518, 0xAA 486 TFLOP/s — the softmax warps
                            // are issue/latency-bound, not MUFU-bound, at d_head 64, so the offload is off.
#endif

template <bool MASKED>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&s)[32], float sl2, float mb, int lim, uint8_t* prow, int cbase, int rx) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = fmaf(__uint_as_float(s[q * 8 + i]), sl2, -mb);
      e[i] = ((ATT_POLY_MASK >> i) & 1) ? exp2_poly(x) : ex2_approx(x);
      if constexpr (MASKED) {
        if (q * 8 + i >= lim) e[i] = 0.f;
      }
    }
    uint4 pk;
    pk.x = pack_bf16(e[0], e[1]);
    pk.y = pack_bf16(e[2], e[3]);
    pk.z = pack_bf16(e[4], e[5]);
    pk.w = pack_bf16(e[6], e[7]);
    *reinterpret_cast<uint4*>(prow + (((cbase + q) ^ rx) << 4)) = pk;
  }
}

// named barrier shared by the two softmax warps of one TMEM lane quadrant (ids 1..4, 64 threads)
__device__ __forceinline__ void pair_sync(int quad) {
  if (quad == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
  else if (quad == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
  else if (quad == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
  else asm volatile("bar.sync 4, 64;" ::: "memory");
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

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_Q_BYTES;
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_K_BYTES;
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_V_BYTES;  // offset 16K + 24K + 30K = 70K: 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_P_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;   // [3] K_j landed
  uint64_t* bar_v = bars + 4;   // [3] V'_j landed
  uint64_t* bar_s = bars + 7;   // [2] S_j in TMEM
  uint64_t* bar_p = bars + 9;   // [2] P_j in smem, S_j consumed (128 arrivals)
  uint64_t* bar_pv = bars + 11; // [2] P_j V'_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  int kvlen = p.n;
  if (p.lens != nullptr) kvlen = min(p.n, __ldg(p.lens + (p.lens_mod > 0 ? b % p.lens_mod : b)));
  const int D = p.H * 64;

  if (kvlen <= 0 || q0 >= kvlen) {
    // whole tile is padding: the reference zeroes these rows after to_out (model/modules.py:499-501)
    if (warp < 4) {
      const int pos = q0 + warp * 32 + lane;
      if (pos < p.n) {
        uint4* o = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.n + pos) * D + h * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    return;
  }
  const int T = (kvlen + ATT_BKV - 1) / ATT_BKV;

  if (warp == ATT_SM_WARPS) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmK);
      prefetch_tmap(&tmV);
      mbar_init(bar_q, 1);
      for (int i = 0; i < ATT_KV_STAGES; ++i) {
        mbar_init(&bar_k[i], 1);
        mbar_init(&bar_v[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_s[i], 1);
        mbar_init(&bar_p[i], ATT_SM_WARPS * 32);
        mbar_init(&bar_pv[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  } else if (warp < 4) {
    // constant rows 64..79 of every V' stage: row 64 = ones (bf16 1.0), rows 65..79 = 0   (3 x 2 KB, 128 threads x 16 B)
    const int t = threadIdx.x;  // 0..127
    const uint32_t v = (t < 8) ? 0x3F803F80u : 0u;  // first 8 x 16 B = row 64 (identical chunks, swizzle-invariant)
#pragma unroll
    for (int st = 0; st < ATT_KV_STAGES; ++st)
      *reinterpret_cast<uint4*>(sV + st * ATT_V_BYTES + 64 * 128 + t * 16) = make_uint4(v, v, v, v);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == ATT_SM_WARPS) {
    if (lane == 0) {
      const uint32_t idesc_s = idesc_bf16(128, ATT_BKV, 0, 0);
      const uint32_t idesc_o = idesc_bf16(128, ATT_NV, 0, 0);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      auto load_k = [&](int j) {
        const int st = j % ATT_KV_STAGES;
        mbar_arrive_expect_tx(&bar_k[st], ATT_K_BYTES);
        tma_load_3d(sK + st * ATT_K_BYTES, &tmK, &bar_k[st], 0, j * ATT_BKV, bh);
      };
      auto load_v = [&](int j) {
        const int st = j % ATT_KV_STAGES;
        mbar_arrive_expect_tx(&bar_v[st], ATT_V_TX);
        tma_load_3d(sV + st * ATT_V_BYTES, &tmV, &bar_v[st], j * ATT_BKV, 0, bh);
      };
      auto issue_s = [&](int j) {
        const int st = j % ATT_KV_STAGES;
        mbar_wait(&bar_k[st], (j / ATT_KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t kb = k_addr + st * ATT_K_BYTES;
        const uint32_t d = tmem_base + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(d, smem_desc_sw128(q_addr + k * 32, 1024, 16), smem_desc_sw128(kb + k * 32, 1024, 16), idesc_s, k != 0);
        umma_commit(&bar_s[j & 1]);
      };
      // prologue
      mbar_arrive_expect_tx(bar_q, ATT_Q_BYTES);
      tma_load_3d(sQ, &tmQ, bar_q, 0, q0, bh);
      for (int j = 0; j < ATT_KV_STAGES && j < T; ++j) load_k(j);
      for (int j = 0; j < 2 && j < T; ++j) load_v(j);
      mbar_wait(bar_q, 0);
      issue_s(0);
      if (T > 1) issue_s(1);
      for (int j = 0; j < T; ++j) {
        const int pb = j & 1;
        mbar_wait(&bar_p[pb], (j >> 1) & 1);  // P_j in smem, S_j consumed, O rescaled if needed
        tc_fence_after();
        const int vst = j % ATT_KV_STAGES;
        mbar_wait(&bar_v[vst], (j / ATT_KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t pa = p_addr + pb * ATT_P_BYTES;
        const uint32_t vb = v_addr + vst * ATT_V_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_O, smem_desc_sw128(pa + kk * 32, 1024, 16), smem_desc_sw128(vb + kk * 32, 1024, 16), idesc_o, (j | kk) != 0);
        umma_commit(&bar_pv[pb]);
        if (j + 2 < T) issue_s(j + 2);  // S buffer pb is free (softmax consumed S_j before arriving on bar_p)
        if (j + ATT_KV_STAGES < T) load_k(j + ATT_KV_STAGES);  // K stage of tile j is free (S_j retired long ago)
        if (j + 2 < T) {
          // V' stage (j+2)%3 was last read by P_{j-1} V'_{j-1}
          if (j >= 1) mbar_wait(&bar_pv[(j - 1) & 1], ((j - 1) >> 1) & 1);
          load_v(j + 2);
        }
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float m_used = -INFINITY;  // log2-domain maximum the exponentials are taken against
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
      const int pb = j & 1;
      const uint32_t tS = tmem_base + lane_addr + pb * 64;
      uint8_t* p_row = sP + pb * ATT_P_BYTES + r * 128;
      const int valid = min(ATT_BKV, kvlen - j * ATT_BKV);  // CTA-uniform, >= 1
      mbar_wait(&bar_s[pb], (j >> 1) & 1);
      tc_fence_after();
      ATT_MARK(0)
      uint32_t s0[32], s1[32];
      tmem_ld32(tS, s0);
      tmem_ld32(tS + 32, s1);
      tmem_ld_wait();
      ATT_MARK(1)
      // The exponentials are taken against the running reference maximum m_used, which only moves when a row maximum grows
      // by more than 2^8 (lazy rescale).  They are therefore issued SPECULATIVELY, before this tile's maximum is known: the
      // max reduction (FMNMX3) has no consumer inside the block and fills the issue slots in the shadow of the MUFU
      // instructions.  In the rare case that the check fails, O is rescaled in TMEM and the tile's P is recomputed.
      if (j == 0) m_used = fmaxf(row_max32(s0, min(32, valid)), row_max32(s1, min(32, valid - 32))) * sl2;
      ATT_MARK(2)
      uint8_t* p_row_ = p_row;
      if (valid == ATT_BKV) {
        softmax_chunk<false>(s0, sl2, m_used, 32, p_row_, 0, rx);
        softmax_chunk<false>(s1, sl2, m_used, 32, p_row_, 4, rx);
      } else {
        softmax_chunk<true>(s0, sl2, m_used, valid, p_row_, 0, rx);
        softmax_chunk<true>(s1, sl2, m_used, valid - 32, p_row_, 4, rx);
      }
      ATT_MARK(3)
      if (j > 0) {
        const float mt = fmaxf(row_max32(s0, min(32, valid)), row_max32(s1, min(32, valid - 32))) * sl2;
        // warp-uniform decision; tcgen05.ld/st are warp-collective
        if (__any_sync(0xffffffffu, mt > m_used + ATT_RESCALE_LOG2)) {
          const float m_new = fmaxf(m_used, mt);
          const float f = ex2_approx(m_used - m_new);
          m_used = m_new;
          mbar_wait(&bar_pv[(j - 1) & 1], ((j - 1) >> 1) & 1);  // P_{j-1} V'_{j-1} has landed in O
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(tmem_O + lane_addr + c * 32, o);
          }
          tmem_st_wait();
          if (valid == ATT_BKV) {
            softmax_chunk<false>(s0, sl2, m_used, 32, p_row_, 0, rx);
            softmax_chunk<false>(s1, sl2, m_used, 32, p_row_, 4, rx);
          } else {
            softmax_chunk<true>(s0, sl2, m_used, valid, p_row_, 0, rx);
            softmax_chunk<true>(s1, sl2, m_used, valid - 32, p_row_, 4, rx);
          }
        }
      }
      ATT_MARK(4)
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_p[pb]);
      ATT_MARK(5)
    }
#ifdef ATT_TRACE
    if (p.trace != nullptr && blockIdx.x == 3 && blockIdx.y == 5 && lane == 0) {
      for (int i = 0; i < 6; ++i) p.trace[warp * 8 + i] = tr[i];
      p.trace[warp * 8 + 6] = T;
    }
#endif
    // epilogue: O[:, :64] / O[:, 64]
    mbar_wait(&bar_pv[(T - 1) & 1], ((T - 1) >> 1) & 1);
    tc_fence_after();
    const int pos = q0 + r;
    float inv;
    {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_addr + 64, o);
      tmem_ld_wait();
      inv = (pos < kvlen) ? 1.f / __uint_as_float(o[0]) : 0.f;
    }
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
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ATT_SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

long long* g_attn_trace = nullptr;

int attn_fwd(const void* q, const void* k, const void* vt, void* out, const int32_t* lens, int lens_mod, int B, int H, int n,
             int n_pad, float scale, cudaStream_t stream) {
  F5B_CHECK(q && k && vt && out, "f5b_attn_fwd: null pointer");
  F5B_CHECK(B > 0 && H > 0 && n > 0 && n_pad >= n && (n_pad & 7) == 0, "f5b_attn_fwd: bad shape B %d H %d n %d n_pad %d", B, H, n,
            n_pad);
  LaunchScope scope(K_ATTN, stream, 4.0 * B * H * (double)n * n * 64, 2.0 * 4 * B * H * (double)n * 64);
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t bh = (uint64_t)B * H;
  if (make_tmap_3d(&tmQ, q, 2, 64, (uint64_t)n, bh, 128, (uint64_t)n * 128, 64, ATT_BQ, 1, true)) return -1;
  if (make_tmap_3d(&tmK, k, 2, 64, (uint64_t)n, bh, 128, (uint64_t)n * 128, 64, ATT_BKV, 1, true)) return -1;
  if (make_tmap_3d(&tmV, vt, 2, (uint64_t)n, 64, bh, (uint64_t)n_pad * 2, (uint64_t)n_pad * 128, ATT_BKV, 64, 1, true)) return -1;
  static bool configured = false;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    configured = true;
  }
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lens = lens;
  p.lens_mod = lens_mod;
  p.B = B;
  p.H = H;
  p.n = n;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  dim3 grid((n + ATT_BQ - 1) / ATT_BQ, B * H);
  attn_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f5b

extern "C" int f5b_attn_fwd(const void* q, const void* k, const void* vt, void* out, const int32_t* lens, int lens_mod, int B,
                            int H, int n, int n_pad, float scale, f5b_stream_t stream) {
  return f5b::attn_fwd(q, k, vt, out, lens, lens_mod, B, H, n, n_pad, scale, static_cast<cudaStream_t>(stream));
}

extern "C" void f5b_debug_set_attn_trace(long long* buf) { f5b::g_attn_trace = buf; }
