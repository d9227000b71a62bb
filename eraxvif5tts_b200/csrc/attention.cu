// Flash-style attention forward for the DiT blocks: O = softmax(Q K^T / sqrt(64), keys < len[b]) V, non-causal, d_head 64.
// Replaces AttnProcessor's mask expansion + F.scaled_dot_product_attention + head merge
// (/root/reference/src/f5_tts/model/modules.py:483-493; dropout_p = 0, see DESIGN.md "oracle adjustments").
//
// sm_100a design (one CTA = one 128-query tile of one (batch, head); two CTAs co-resident per SM so that one CTA's
// softmax overlaps the other's tensor work):
//   warp 4 (one elected lane): TMA producer + tcgen05.mma issuer.
//       S[128 x 128]  = Q K_j^T    (Q, K_j K-major bf16 tiles, 128B swizzle; fp32 accumulator in TMEM columns 0..127)
//       O[128 x 80]  += P_j V'_j   (P_j written to swizzled smem by the softmax warps; V' = V^T from the QKV epilogue's
//                                   transposed store plus a constant row of ones, so column 64 of O accumulates the softmax
//                                   row sum on the tensor pipe; accumulator stays resident in TMEM columns 128..207)
//   warps 0..3: online softmax, thread = query row (tcgen05.ld 32x32b -> no cross-lane reductions), exp2 with the
//       1/sqrt(d)*log2(e) scale folded in.  The running maximum is updated lazily: O (and with it the row sum) is rescaled in
//       TMEM only when a row's maximum grew by more than 2^8, so most KV tiles cost no accumulator round trip.
//   K is double-buffered, V single-buffered (its reload hides behind the next tile's softmax).
//   Key-padding is a per-batch length bound: KV tiles past len[b] are never loaded, the last tile is masked by index.
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_THREADS = 160;
constexpr int ATT_NV = 80;                                  // 64 value columns + the ones row + zero padding to N % 16 == 0
constexpr uint32_t ATT_Q_BYTES = ATT_BQ * 64 * 2;          // 16 KB
constexpr uint32_t ATT_K_BYTES = ATT_BKV * 64 * 2;         // 16 KB per stage, 2 stages
constexpr uint32_t ATT_VC_BYTES = ATT_NV * 128;            // one 64-kv chunk of V': 80 rows x 128 B = 10 KB
constexpr uint32_t ATT_V_BYTES = 2 * ATT_VC_BYTES;         // 20 KB
constexpr uint32_t ATT_V_TX = 2 * 64 * 128;                // bytes TMA writes per tile (the 64 real rows of both chunks)
constexpr uint32_t ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;     // 32 KB (two [128 q x 64 kv] atoms)
constexpr uint32_t ATT_SMEM = ATT_Q_BYTES + 2 * ATT_K_BYTES + ATT_V_BYTES + ATT_P_BYTES + 1024 + 128;
constexpr uint32_t ATT_TMEM_COLS = 256;  // S: 128, O: 80
constexpr float ATT_RESCALE_LOG2 = 8.0f;

struct AttnParams {
  __nv_bfloat16* out;
  const int32_t* lens;
  int lens_mod, B, H, n;
  float scale_log2;
};

template <bool MASKED>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&s)[32], float sl2, float mb, int lim, uint8_t* atom, int cbase, int rx) {
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
    uint4 pk;
    pk.x = pack_bf16(e[0], e[1]);
    pk.y = pack_bf16(e[2], e[3]);
    pk.z = pack_bf16(e[4], e[5]);
    pk.w = pack_bf16(e[6], e[7]);
    *reinterpret_cast<uint4*>(atom + (((cbase + q) ^ rx) << 4)) = pk;
  }
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_Q_BYTES;
  uint8_t* sV = sK + 2 * ATT_K_BYTES;
  uint8_t* sP = sV + ATT_V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + ATT_P_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;  // [2]
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_p = bars + 5;
  uint64_t* bar_o = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

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
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  } else {
    // constant rows 64..79 of both V' chunks: row 64 = ones (bf16 1.0), rows 65..79 = 0.  128 threads x 2 x 16 B x 8.
    const int t = threadIdx.x;  // 0..127
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint8_t* base = sV + c * ATT_VC_BYTES + 64 * 128;
      const uint32_t one2 = 0x3F803F80u;
      const uint32_t v = (t < 8) ? one2 : 0u;  // the first 8 x 16 B = row 64 (identical chunks, swizzle-invariant)
      *reinterpret_cast<uint4*>(base + t * 16) = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_s = idesc_bf16(128, ATT_BKV, 0, 0);
      const uint32_t idesc_o = idesc_bf16(128, ATT_NV, 0, 0);
      // prologue loads
      mbar_arrive_expect_tx(bar_q, ATT_Q_BYTES);
      tma_load_3d(sQ, &tmQ, bar_q, 0, q0, bh);
      mbar_arrive_expect_tx(&bar_k[0], ATT_K_BYTES);
      tma_load_3d(sK, &tmK, &bar_k[0], 0, 0, bh);
      mbar_arrive_expect_tx(bar_v, ATT_V_TX);
      tma_load_3d(sV, &tmV, bar_v, 0, 0, bh);
      tma_load_3d(sV + ATT_VC_BYTES, &tmV, bar_v, 64, 0, bh);
      if (T > 1) {
        mbar_arrive_expect_tx(&bar_k[1], ATT_K_BYTES);
        tma_load_3d(sK + ATT_K_BYTES, &tmK, &bar_k[1], 0, ATT_BKV, bh);
      }
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      // S_0
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[0], 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_S, smem_desc_sw128(q_addr + k * 32, 1024, 16), smem_desc_sw128(k_addr + k * 32, 1024, 16), idesc_s,
                  k != 0);
      umma_commit(bar_s);
      for (int j = 0; j < T; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(bar_p, ph);  // P_j in smem, S_j consumed, O rescaled if needed
        tc_fence_after();
        // K buffer (j&1) is free (S_j retired before the softmax warps saw bar_s): prefetch K_{j+2}
        if (j + 2 < T) {
          mbar_arrive_expect_tx(&bar_k[j & 1], ATT_K_BYTES);
          tma_load_3d(sK + (j & 1) * ATT_K_BYTES, &tmK, &bar_k[j & 1], 0, (j + 2) * ATT_BKV, bh);
        }
        mbar_wait(bar_v, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t a = p_addr + (kk >> 2) * 16384 + (kk & 3) * 32;
          const uint32_t bb = v_addr + (kk >> 2) * ATT_VC_BYTES + (kk & 3) * 32;
          umma_bf16(tmem_O, smem_desc_sw128(a, 1024, 16), smem_desc_sw128(bb, 1024, 16), idesc_o, (j | kk) != 0);
        }
        umma_commit(bar_o);
        if (j + 1 < T) {
          // S_{j+1} queues behind P_j V_j on the tensor pipe
          const int nb = (j + 1) & 1;
          mbar_wait(&bar_k[nb], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint32_t kb_addr = k_addr + nb * ATT_K_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_S, smem_desc_sw128(q_addr + k * 32, 1024, 16), smem_desc_sw128(kb_addr + k * 32, 1024, 16),
                      idesc_s, k != 0);
          umma_commit(bar_s);
          // V buffer is free once P_j V_j retired
          mbar_wait(bar_o, ph);
          mbar_arrive_expect_tx(bar_v, ATT_V_TX);
          tma_load_3d(sV, &tmV, bar_v, (j + 1) * ATT_BKV, 0, bh);
          tma_load_3d(sV + ATT_VC_BYTES, &tmV, bar_v, (j + 1) * ATT_BKV + 64, 0, bh);
        }
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float m_used = -INFINITY;  // log2-domain maximum the exponentials are taken against
    const float sl2 = p.scale_log2;
    uint8_t* p_row = sP + r * 128;
    const int rx = r & 7;

    for (int j = 0; j < T; ++j) {
      const uint32_t ph = j & 1;
      const int valid = min(ATT_BKV, kvlen - j * ATT_BKV);  // CTA-uniform, >= 1
      mbar_wait(bar_s, ph);
      tc_fence_after();
      // pass 1: row max (raw scores)
      float m_tile = -INFINITY;
      if (valid == ATT_BKV) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t s[32];
          tmem_ld32(tmem_S + lane_addr + c * 32, s);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) m_tile = fmaxf(fmaxf(m_tile, __uint_as_float(s[i])), __uint_as_float(s[i + 1]));
        }
      } else {
#pragma unroll 1
        for (int c = 0; c * 32 < valid; ++c) {
          uint32_t s[32];
          tmem_ld32(tmem_S + lane_addr + c * 32, s);
          tmem_ld_wait();
          const int lim = valid - c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < lim) m_tile = fmaxf(m_tile, __uint_as_float(s[i]));
        }
      }
      const float mt = m_tile * sl2;
      // lazy rescale (warp-uniform decision; tcgen05.ld/st are warp-collective)
      if (j == 0) {
        m_used = mt;
      } else if (__any_sync(0xffffffffu, mt > m_used + ATT_RESCALE_LOG2)) {
        const float m_new = fmaxf(m_used, mt);
        const float f = ex2_approx(m_used - m_new);
        m_used = m_new;
        mbar_wait(bar_o, (j - 1) & 1);  // P_{j-1} V_{j-1} has landed in O
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
      }
      // pass 2: P = exp2(S*sl2 - m_used) -> bf16 -> swizzled smem
      const float mb = m_used;
      if (valid == ATT_BKV) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t s[32];
          tmem_ld32(tmem_S + lane_addr + c * 32, s);
          tmem_ld_wait();
          softmax_chunk<false>(s, sl2, mb, 32, p_row + (c >> 1) * 16384, (c & 1) * 4, rx);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t s[32];
          const int lim = valid - c * 32;  // may be <= 0 -> all zeros
          if (lim > 0) {
            tmem_ld32(tmem_S + lane_addr + c * 32, s);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] = 0u;
          }
          softmax_chunk<true>(s, sl2, mb, lim, p_row + (c >> 1) * 16384, (c & 1) * 4, rx);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // epilogue: O[:, :64] / O[:, 64]
    mbar_wait(bar_o, (T - 1) & 1);
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
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

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
  if (make_tmap_3d(&tmV, vt, 2, (uint64_t)n, 64, bh, (uint64_t)n_pad * 2, (uint64_t)n_pad * 128, 64, 64, 1, true)) return -1;
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
