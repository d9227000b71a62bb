// Attention forward, multi-query-tile variant: same math and interface as attention.cu, different schedule and data path.
// Replaces AttnProcessor's mask expansion + F.scaled_dot_product_attention + head merge
// (/root/reference/src/f5_tts/model/modules.py:483-493; dropout_p = 0, see DESIGN.md "oracle adjustments").
//
// attention.cu (one 128-query tile per CTA, every MMA operand from shared memory, 3 CTAs / SM) is shared-memory-bandwidth
// bound: Q 16 KB + K 8 KB + P 16 KB + V 8 KB of operand reads, 16 KB of P stores and 16 KB of TMA fills per 128 x 64 tile = 80 KB
// = 640 clocks of the 128 B/clk port, more than the 512 MUFU clocks of the tile's exponentials (profiles/r01_attention_notes.md).
// Here ONE CTA per SM runs THREE query tiles ("contexts") of the same (batch, head) side by side:
//   * K_j / V_j are fetched once and serve all three contexts (TMA fills per tile and context: 16 KB -> 5.3 KB);
//   * P goes from the softmax registers straight to TENSOR MEMORY (tcgen05.st) and O += P V reads it from there
//     (tcgen05.mma with the A operand in TMEM): no P stores, no P operand reads on the shared-memory port;
//   -> 16 + 8 + 8 + 5.3 = 37 KB = 300 clocks per tile and context: the MUFU, not shared memory, is the limit again;
//   * each context keeps the register-pull pipeline of attention.cu (S_j is pulled into registers and its TMEM buffer released
//     at once, so S_{j+1} is computed while the exponentials of tile j run), with 12 softmax warps per SM as before.
// TMEM (512 columns allocated, 480 used): context c at 160 c: S 0..63 | P 64..95 | O 96..159.
// Warps 0..11: softmax (context = warp / 4, TMEM lane quarter = warp % 4); warps 12..14: one tcgen05.mma issuer per context (a
// single issuing thread for all three contexts was issue-bound: 391 TFLOP/s); warp 15: TMA producer (lane 0: K ring, lane 1: V
// ring; 4 stages each, a stage is recycled when every context's MMAs on it have retired).
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int MQ_NQ = 3;
constexpr int MQ_BQ = 128;
constexpr int MQ_BKV = 64;
constexpr int MQ_THREADS = (MQ_NQ * 4 + MQ_NQ + 1) * 32;  // 512: 12 softmax warps, 3 MMA warps, 1 TMA warp
constexpr int MQ_STAGES = 4;
constexpr uint32_t MQ_Q_BYTES = MQ_BQ * 64 * 2;    // 16 KB per context
constexpr uint32_t MQ_KV_BYTES = MQ_BKV * 64 * 2;  // 8 KB per stage
constexpr uint32_t MQ_SMEM = MQ_NQ * MQ_Q_BYTES + 2 * MQ_STAGES * MQ_KV_BYTES + 1024 + 512;
constexpr uint32_t MQ_TMEM_COLS = 512;
constexpr uint32_t MQ_CTX_COLS = 160;
constexpr float MQ_RESCALE_LOG2 = 8.0f;

struct AttnMqParams {
  __nv_bfloat16* out;
  float* lse;
  const int32_t* lens;
  int lens_mod, B, H, n;
  float scale_log2;
};

template <bool MASKED>
__device__ __forceinline__ float mq_chunk(const uint32_t (&s)[32], float sl2, float mb, int lim, uint32_t* pk) {
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      e[i] = ex2_approx(fmaf(__uint_as_float(s[q * 8 + i]), sl2, -mb));
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

__device__ __forceinline__ float mq_row_max32(const uint32_t (&a)[32], int valid) {
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

__global__ void __launch_bounds__(MQ_THREADS, 1)
attn_fwd_mq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnMqParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                         // [NQ] tiles
  uint8_t* sK = sQ + MQ_NQ * MQ_Q_BYTES;      // 4 stages
  uint8_t* sV = sK + MQ_STAGES * MQ_KV_BYTES; // 4 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + MQ_STAGES * MQ_KV_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;        // [4] K_j landed
  uint64_t* bar_v = bars + 5;        // [4]
  uint64_t* bar_kfree = bars + 9;    // [4] S_j of every context retired
  uint64_t* bar_vfree = bars + 13;   // [4] P_j V_j of every context retired
  uint64_t* bar_s = bars + 17;       // [NQ] S_j in TMEM
  uint64_t* bar_sfree = bars + 20;   // [NQ] S_j pulled into registers (128 arrivals)
  uint64_t* bar_p = bars + 23;       // [NQ] P_j in TMEM (128 arrivals)
  uint64_t* bar_pv = bars + 26;      // [NQ] P_j V_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_base = blockIdx.x * (MQ_NQ * MQ_BQ);
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  int kvlen = p.n;
  if (p.lens != nullptr) kvlen = min(p.n, __ldg(p.lens + (p.lens_mod > 0 ? b % p.lens_mod : b)));
  const int D = p.H * 64;
  // contexts whose query tile starts inside the valid keys do attention; the others are padding (zeros; the reference zeroes
  // those rows after to_out, model/modules.py:499-501)
  int n_act = 0;
#pragma unroll
  for (int c = 0; c < MQ_NQ; ++c)
    if (kvlen > 0 && q_base + c * MQ_BQ < kvlen) n_act = c + 1;

  if (warp < MQ_NQ * 4 && (warp >> 2) >= n_act) {
    const int pos = q_base + (warp >> 2) * MQ_BQ + (warp & 3) * 32 + lane;
    if (pos < p.n) {
      uint4* o = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.n + pos) * D + h * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
      if (p.lse != nullptr) p.lse[(size_t)bh * p.n + pos] = INFINITY;
    }
  }
  if (n_act == 0) return;
  const int T = (kvlen + MQ_BKV - 1) / MQ_BKV;

  if (warp == 12) {
    if (lane == 0) {
      mbar_init(bar_q, 1);
      for (int i = 0; i < MQ_STAGES; ++i) {
        mbar_init(&bar_k[i], 1);
        mbar_init(&bar_v[i], 1);
        mbar_init(&bar_kfree[i], n_act);  // one tcgen05.commit per active context
        mbar_init(&bar_vfree[i], n_act);
      }
      for (int c = 0; c < MQ_NQ; ++c) {
        mbar_init(&bar_s[c], 1);
        mbar_init(&bar_sfree[c], 128);
        mbar_init(&bar_p[c], 128);
        mbar_init(&bar_pv[c], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, MQ_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 15) {
    // ------------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmK);
      mbar_arrive_expect_tx(bar_q, n_act * MQ_Q_BYTES);
      for (int c = 0; c < n_act; ++c) tma_load_3d(sQ + c * MQ_Q_BYTES, &tmQ, bar_q, h * 64, q_base + c * MQ_BQ, b);
      for (int j = 0; j < T; ++j) {
        const int st = j & (MQ_STAGES - 1);
        if (j >= MQ_STAGES) mbar_wait(&bar_kfree[st], ((j / MQ_STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&bar_k[st], MQ_KV_BYTES);
        tma_load_3d(sK + st * MQ_KV_BYTES, &tmK, &bar_k[st], h * 64, j * MQ_BKV, b);
      }
    } else if (lane == 1) {
      prefetch_tmap(&tmV);
      for (int j = 0; j < T; ++j) {
        const int st = j & (MQ_STAGES - 1);
        if (j >= MQ_STAGES) mbar_wait(&bar_vfree[st], ((j / MQ_STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&bar_v[st], MQ_KV_BYTES);
        tma_load_3d(sV + st * MQ_KV_BYTES, &tmV, &bar_v[st], h * 64, j * MQ_BKV, b);
      }
    }
    __syncwarp();
  } else if (warp >= 12) {
    // ------------------------------------------------------------------------------------------------ MMA issue (context warp - 12)
    const int c = warp - 12;
    if (lane == 0 && c < n_act) {
      const uint32_t idesc_s = idesc_bf16(128, 64, 0, 0);   // S = Q K^T: both operands K-major in smem
      const uint32_t idesc_pv = idesc_bf16(128, 64, 0, 1);  // O += P V: A (TMEM) K-major, V MN-major
      const uint32_t qa = smem_u32(sQ) + c * MQ_Q_BYTES, k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t tc = tmem_base + c * MQ_CTX_COLS;
      auto issue_s = [&](int j) {
        const int st = j & (MQ_STAGES - 1);
        mbar_wait(&bar_k[st], (j / MQ_STAGES) & 1);
        tc_fence_after();
        const uint32_t kb = k_addr + st * MQ_KV_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tc, smem_desc_sw128(qa + k * 32, 1024, 16), smem_desc_sw128(kb + k * 32, 1024, 16), idesc_s, k != 0);
        umma_commit(&bar_s[c]);
        umma_commit(&bar_kfree[st]);  // this context is done with K_j once the MMAs retire
      };
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        // S_j sits in registers: its TMEM buffer is free -> S_{j+1} overlaps the exponentials of tile j
        mbar_wait(&bar_sfree[c], j & 1);
        tc_fence_after();
        if (j + 1 < T) issue_s(j + 1);
        mbar_wait(&bar_p[c], j & 1);  // P_j in TMEM (and O rescaled if needed)
        tc_fence_after();
        const int st = j & (MQ_STAGES - 1);
        mbar_wait(&bar_v[st], (j / MQ_STAGES) & 1);
        tc_fence_after();
        const uint32_t vb = v_addr + st * MQ_KV_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ts(tc + 96, tc + 64 + kk * 8, smem_desc_sw128(vb + kk * 2048, 1024, 8192), idesc_pv, (j | kk) != 0);
        umma_commit(&bar_pv[c]);
        umma_commit(&bar_vfree[st]);
      }
    }
    __syncwarp();
  } else if ((warp >> 2) < n_act) {
    // ------------------------------------------------------------------------------------------------ softmax
    const int c = warp >> 2;
    const int lq = warp & 3;
    const int r = lq * 32 + lane;  // query row in the context's tile == TMEM lane
    const int q0 = q_base + c * MQ_BQ;
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    const uint32_t tm_S = tmem_base + c * MQ_CTX_COLS + lane_addr;
    const uint32_t tm_P = tm_S + 64, tm_O = tm_S + 96;
    float m_used = -INFINITY;  // log2-domain maximum the exponentials are taken against
    float l_run = 0.f;         // row sum of exp2(s - m_used)
    const float sl2 = p.scale_log2;
#ifndef MQ_STAGGER_NS
#define MQ_STAGGER_NS 170
#endif
    // the three contexts start a third of a tile time apart: in lockstep all 12 warps hit the MUFU together and then all wait
    // together (measured: MUFU 53 % busy, stall_mio on every EX2); the K / V ring lets the offset persist
    if (c > 0) __nanosleep(c * MQ_STAGGER_NS);

    for (int j = 0; j < T; ++j) {
      const int valid = min(MQ_BKV, kvlen - j * MQ_BKV);  // CTA-uniform, >= 1
      mbar_wait(&bar_s[c], j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tm_S, s0);
      tmem_ld32(tm_S + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bar_sfree[c]);  // S_{j+1} may overwrite the buffer: it is computed while the exponentials below run
      if (j == 0) m_used = fmaxf(mq_row_max32(s0, min(32, valid)), mq_row_max32(s1, min(32, valid - 32))) * sl2;
      uint32_t ppk[32];
      float ts;
      auto make_p = [&]() {
        if (valid == MQ_BKV) {
          ts = mq_chunk<false>(s0, sl2, m_used, 32, ppk);
          ts += mq_chunk<false>(s1, sl2, m_used, 32, ppk + 16);
        } else {
          ts = mq_chunk<true>(s0, sl2, m_used, valid, ppk);
          ts += mq_chunk<true>(s1, sl2, m_used, valid - 32, ppk + 16);
        }
      };
      make_p();  // speculative w.r.t. this tile's maximum
      if (j > 0) {
        const float mt = fmaxf(mq_row_max32(s0, min(32, valid)), mq_row_max32(s1, min(32, valid - 32))) * sl2;
        const bool grow = __any_sync(0xffffffffu, mt > m_used + MQ_RESCALE_LOG2);  // warp-uniform (tcgen05.ld/st are collective)
        mbar_wait(&bar_pv[c], (j - 1) & 1);  // P_{j-1} V_{j-1} retired: the P columns are free and O is current
        tc_fence_after();
        if (grow) {
          const float m_new = fmaxf(m_used, mt);
          const float f = ex2_approx(m_used - m_new);
          m_used = m_new;
          l_run *= f;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t o[32];
            tmem_ld32(tm_O + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(tm_O + cc * 32, o);
          }
          make_p();
        }
      }
      l_run += ts;
      tmem_st32(tm_P, ppk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bar_p[c]);
    }
    // epilogue: O / l
    mbar_wait(&bar_pv[c], (T - 1) & 1);
    tc_fence_after();
    const int pos = q0 + r;
    const float inv = (pos < kvlen) ? 1.f / l_run : 0.f;
    if (p.lse != nullptr && pos < p.n) p.lse[(size_t)bh * p.n + pos] = (pos < kvlen) ? m_used + log2f(l_run) : INFINITY;
    __nv_bfloat16* orow = p.out + ((size_t)b * p.n + (pos < p.n ? pos : 0)) * D + h * 64;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t o[32];
      tmem_ld32(tm_O + cc * 32, o);
      tmem_ld_wait();
      if (pos < p.n) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          pk.x = pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
          pk.y = pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
          pk.z = pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
          pk.w = pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
          reinterpret_cast<uint4*>(orow + cc * 32)[q] = pk;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, MQ_TMEM_COLS);
  }
}

int attn_fwd_ts(const void* q, const void* k, const void* v, int ld, void* out, float* lse, const int32_t* lens, int lens_mod, int B,
                int H, int n, float scale, cudaStream_t stream) {
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t hw = (uint64_t)H * 64, pitch = (uint64_t)ld * 2;
  if (make_tmap_3d(&tmQ, q, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, MQ_BQ, 1, true)) return -1;
  if (make_tmap_3d(&tmK, k, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, MQ_BKV, 1, true)) return -1;
  if (make_tmap_3d(&tmV, v, 2, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 64, MQ_BKV, 1, true)) return -1;
  static bool configured = false;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(attn_fwd_mq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MQ_SMEM));
    configured = true;
  }
  AttnMqParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.lens = lens;
  p.lens_mod = lens_mod;
  p.B = B;
  p.H = H;
  p.n = n;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((n + MQ_NQ * MQ_BQ - 1) / (MQ_NQ * MQ_BQ), B * H);
  attn_fwd_mq_kernel<<<grid, MQ_THREADS, MQ_SMEM, stream>>>(tmQ, tmK, tmV, p);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f5b
