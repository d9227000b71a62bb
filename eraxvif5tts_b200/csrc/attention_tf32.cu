// Attention forward of the tf32 operand mode: O = softmax(Q K^T / sqrt(64), keys < len[b]) V with Q, K, V, P as tf32 tensor-core
// operands (fp32 words, 10-bit mantissa, rounded to nearest by their producers) and fp32 accumulation / softmax — the attention of
// the precision mode that holds 1e-3 of the fp32 reference (AttnProcessor, /root/reference/src/f5_tts/model/modules.py:483-493).
// Same structure as attention.cu (one CTA = 128 queries of one (batch, head), 64-key tiles, S and O in tensor memory, thread = query
// row online softmax with the lazy 2^8 rescale), with every operand tile twice as wide in bytes:
//   Q  [128 x 64] fp32 = two 128B-swizzled K-major halves of 32 channels (2 x 16 KB);  K_j [64 x 64] the same (2 x 8 KB, 2 stages);
//   V_j: kind::tf32 takes K-major operands only (measured on B200: with the B-transpose bit of the instruction descriptor set the
//        MMA leaves the accumulator at zero; as with wgmma, MN-major is a 16-bit-type feature), so V is first transposed to
//        V^T [B, H*64, n] by a small tiled-transpose kernel and V^T_j [64 d x 64 keys] is loaded as two K-major halves of 32 keys;
//   P_j [128 x 64 keys] fp32 in shared memory = two K-major halves of 32 keys, ONE buffer (so two CTAs fit an SM: 113 KB each).
// kind::tf32 runs at half the kind::f16 rate per byte; this kernel is the accuracy path, not the throughput path.
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int AT_BQ = 128;
constexpr int AT_BKV = 64;
constexpr int AT_THREADS = 160;
constexpr uint32_t AT_Q_BYTES = AT_BQ * 64 * 4;    // 32 KB
constexpr uint32_t AT_K_BYTES = AT_BKV * 64 * 4;   // 16 KB per stage, 2 stages
constexpr uint32_t AT_V_BYTES = AT_BKV * 64 * 4;   // 16 KB
constexpr uint32_t AT_P_BYTES = AT_BQ * AT_BKV * 4;  // 32 KB
// tiles (112 KB) + 1 KB that covers the 1024-byte alignment of the tile base AND the barriers (the dynamic shared window is at
// least 128-byte aligned, so at least 128 bytes of the kilobyte remain behind the tiles): 2 x (113 KB + 1 KB reserved) = 228 KB
constexpr uint32_t AT_TILE_BYTES = AT_Q_BYTES + 2 * AT_K_BYTES + AT_V_BYTES + AT_P_BYTES;
constexpr uint32_t AT_SMEM = AT_TILE_BYTES + 1024;
constexpr uint32_t AT_TMEM_COLS = 128;  // S: 0..63, O: 64..127
constexpr float AT_RESCALE_LOG2 = 8.0f;

struct AttnTf32Params {
  float* out;
  const int32_t* lens;
  int lens_mod, B, H, n;
  float scale_log2;
};

__device__ __forceinline__ float row_max_n(const uint32_t (&a)[32], int valid) {
  float m = -INFINITY;
  if (valid >= 32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(a[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) m = fmaxf(m, __uint_as_float(a[i]));
  }
  return m;
}

// 32 scores -> exp2 -> tf32 -> one 128-byte swizzled row of a K-major P half; returns the row-sum contribution
__device__ __forceinline__ float p_chunk_tf32(const uint32_t (&s)[32], float sl2, float mb, int lim, uint8_t* prow, int rx) {
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      e[i] = ex2_approx(fmaf(__uint_as_float(s[q * 4 + i]), sl2, -mb));
      if (q * 4 + i >= lim) e[i] = 0.f;
    }
    sum += (e[0] + e[1]) + (e[2] + e[3]);
    *reinterpret_cast<float4*>(prow + ((q ^ rx) << 4)) = make_float4(tf32_rn(e[0]), tf32_rn(e[1]), tf32_rn(e[2]), tf32_rn(e[3]));
  }
  return sum;
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_tf32_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnTf32Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_Q_BYTES;
  uint8_t* sV = sK + 2 * AT_K_BYTES;
  uint8_t* sP = sV + AT_V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AT_P_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;      // [2]
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;      // S_j in TMEM
  uint64_t* bar_sfree = bars + 5;  // S_j pulled into registers (128 arrivals)
  uint64_t* bar_p = bars + 6;      // P_j in smem (128 arrivals)
  uint64_t* bar_pv = bars + 7;     // [2] P_j V_j retired, tile j on barrier j & 1 (two: see attention.cu on parity aliasing)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ;
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  int kvlen = p.n;
  if (p.lens != nullptr) kvlen = min(p.n, __ldg(p.lens + (p.lens_mod > 0 ? b % p.lens_mod : b)));
  const int D = p.H * 64;

  if (kvlen <= 0 || q0 >= kvlen) {  // whole tile is padding: zeros (model/modules.py:499-501 zeroes these rows after to_out)
    if (warp < 4) {
      const int pos = q0 + warp * 32 + lane;
      if (pos < p.n) {
        float4* o = reinterpret_cast<float4*>(p.out + ((size_t)b * p.n + pos) * D + h * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }
  const int T = (kvlen + AT_BKV - 1) / AT_BKV;

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
      mbar_init(bar_p, 128);
      mbar_init(&bar_pv[0], 1);
      mbar_init(&bar_pv[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, AT_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 64;

  if (warp == 4) {
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(128, 64, 0, 0);     // S = Q K^T: both operands K-major
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      auto load_k = [&](int j) {
        uint8_t* dst = sK + (j & 1) * AT_K_BYTES;
        mbar_arrive_expect_tx(&bar_k[j & 1], AT_K_BYTES);
        tma_load_3d(dst, &tmK, &bar_k[j & 1], h * 64, j * AT_BKV, b);
        tma_load_3d(dst + AT_K_BYTES / 2, &tmK, &bar_k[j & 1], h * 64 + 32, j * AT_BKV, b);
      };
      auto load_v = [&](int j) {
        mbar_arrive_expect_tx(bar_v, AT_V_BYTES);
        tma_load_3d(sV, &tmV, bar_v, j * AT_BKV, h * 64, b);  // V^T: keys contiguous; keys >= n are zero-filled
        tma_load_3d(sV + AT_V_BYTES / 2, &tmV, bar_v, j * AT_BKV + 32, h * 64, b);
      };
      auto issue_s = [&](int j) {
        mbar_wait(&bar_k[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t kb = k_addr + (j & 1) * AT_K_BYTES;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int k = 0; k < 4; ++k)  // K-step = 8 channels = 32 bytes inside a 32-channel half
            umma_tf32(tmem_base, smem_desc_sw128(q_addr + a * (AT_Q_BYTES / 2) + k * 32, 1024, 16),
                      smem_desc_sw128(kb + a * (AT_K_BYTES / 2) + k * 32, 1024, 16), idesc, (a | k) != 0);
        umma_commit(bar_s);
      };
      mbar_arrive_expect_tx(bar_q, AT_Q_BYTES);
      tma_load_3d(sQ, &tmQ, bar_q, h * 64, q0, b);
      tma_load_3d(sQ + AT_Q_BYTES / 2, &tmQ, bar_q, h * 64 + 32, q0, b);
      load_k(0);
      if (T > 1) load_k(1);
      load_v(0);
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        mbar_wait(bar_sfree, j & 1);  // S_j sits in registers: its TMEM buffer and K stage are free
        tc_fence_after();
        if (j + 1 < T) issue_s(j + 1);
        if (j + 2 < T) load_k(j + 2);
        mbar_wait(bar_p, j & 1);  // P_j in smem (and O rescaled if needed)
        tc_fence_after();
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          // A: P half kk / 4, B: V^T half kk / 4 (both K-major, keys contiguous); K-step (kk & 3) * 32 bytes = 8 keys
          umma_tf32(tmem_O, smem_desc_sw128(p_addr + (kk >> 2) * (AT_P_BYTES / 2) + (kk & 3) * 32, 1024, 16),
                    smem_desc_sw128(v_addr + (kk >> 2) * (AT_V_BYTES / 2) + (kk & 3) * 32, 1024, 16), idesc, (j | kk) != 0);
        umma_commit(&bar_pv[j & 1]);
        if (j + 1 < T) {
          mbar_wait(&bar_pv[j & 1], (j >> 1) & 1);  // the single V stage (and the single P buffer) is free once P_j V_j retired
          load_v(j + 1);
        }
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float m_used = -INFINITY;
    float l_run = 0.f;
    const float sl2 = p.scale_log2;
    const int rx = r & 7;
    uint8_t* p_row = sP + r * 128;

    for (int j = 0; j < T; ++j) {
      const int valid = min(AT_BKV, kvlen - j * AT_BKV);  // CTA-uniform, >= 1
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tmem_base + lane_addr, s0);
      tmem_ld32(tmem_base + lane_addr + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_sfree);
      const float mt = fmaxf(row_max_n(s0, min(32, valid)), row_max_n(s1, min(32, valid - 32))) * sl2;
      if (j == 0) m_used = mt;
      if (j > 0) {
        mbar_wait(&bar_pv[(j - 1) & 1], ((j - 1) >> 1) & 1);  // P_{j-1} V_{j-1} has landed in O; the P buffer is free
        tc_fence_after();
        if (__any_sync(0xffffffffu, mt > m_used + AT_RESCALE_LOG2)) {  // warp-uniform: tcgen05.ld / st are warp-collective
          const float m_new = fmaxf(m_used, mt);
          const float f = ex2_approx(m_used - m_new);
          m_used = m_new;
          l_run *= f;
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
        }
      }
      float ts = p_chunk_tf32(s0, sl2, m_used, valid, p_row, rx);
      ts += p_chunk_tf32(s1, sl2, m_used, valid - 32, p_row + AT_P_BYTES / 2, rx);
      l_run += ts;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(&bar_pv[(T - 1) & 1], ((T - 1) >> 1) & 1);
    tc_fence_after();
    const int pos = q0 + r;
    const float inv = (pos < kvlen) ? 1.f / l_run : 0.f;
    float* orow = p.out + ((size_t)b * p.n + (pos < p.n ? pos : 0)) * D + h * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_addr + c * 32, o);
      tmem_ld_wait();
      if (pos < p.n) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<float4*>(orow + c * 32)[q] =
              make_float4(tf32_rn(__uint_as_float(o[q * 4]) * inv), tf32_rn(__uint_as_float(o[q * 4 + 1]) * inv),
                          tf32_rn(__uint_as_float(o[q * 4 + 2]) * inv), tf32_rn(__uint_as_float(o[q * 4 + 3]) * inv));
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// V [B*n, ld] (head h = columns h*64..) -> V^T [B, HD, n_pad] (keys contiguous), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_v_kernel(const float* __restrict__ v, int ld, float* __restrict__ vt, int n, int n_pad,
                                                          int HD) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = p0 + ty + i * 8;
    tile[ty + i * 8][tx] = pos < n ? v[((size_t)b * n + pos) * ld + c0 + tx] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, pos = p0 + tx;
    if (pos < n_pad) vt[((size_t)b * HD + c) * n_pad + pos] = tile[tx][ty + i * 8];
  }
}

size_t attn_tf32_ws_floats(int B, int H, int n) { return (size_t)B * H * 64 * ((n + 3) / 4 * 4); }

int attn_fwd_tf32(const float* q, const float* k, const float* v, int ld, float* out, float* vt_ws, const int32_t* lens, int lens_mod,
                  int B, int H, int n, float scale, cudaStream_t stream) {
  F5B_CHECK(q && k && v && out && vt_ws, "f5b_attn_fwd_tf32: null pointer");
  F5B_CHECK(B > 0 && H > 0 && n > 0 && ld >= H * 64 && (ld & 3) == 0, "f5b_attn_fwd_tf32: bad shape B %d H %d n %d ld %d", B, H, n, ld);
  LaunchScope scope(K_ATTN, stream, 4.0 * B * H * (double)n * n * 64, 4.0 * 6 * B * H * (double)n * 64, 2);
  const int n_pad = (n + 3) / 4 * 4;  // 16-byte row pitch for the TMA map of V^T
  transpose_v_kernel<<<dim3((n_pad + 31) / 32, H * 2, B), 256, 0, stream>>>(v, ld, vt_ws, n, n_pad, H * 64);
  F5B_CUDA(cudaGetLastError());
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t hw = (uint64_t)H * 64, pitch = (uint64_t)ld * 4;
  if (make_tmap_3d(&tmQ, q, 4, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 32, AT_BQ, 1, true)) return -1;
  if (make_tmap_3d(&tmK, k, 4, hw, (uint64_t)n, (uint64_t)B, pitch, (uint64_t)n * pitch, 32, AT_BKV, 1, true)) return -1;
  if (make_tmap_3d(&tmV, vt_ws, 4, (uint64_t)n, hw, (uint64_t)B, (uint64_t)n_pad * 4, hw * n_pad * 4, 32, 64, 1, true)) return -1;
  static bool configured = false;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(attn_fwd_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
    configured = true;
  }
  AttnTf32Params p;
  p.out = out;
  p.lens = lens;
  p.lens_mod = lens_mod;
  p.B = B;
  p.H = H;
  p.n = n;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((n + AT_BQ - 1) / AT_BQ, B * H);
  attn_fwd_tf32_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f5b

extern "C" int f5b_attn_fwd_tf32(const float* q, const float* k, const float* v, int ld, float* out, float* vt_ws, const int32_t* lens,
                                 int lens_mod, int B, int H, int n, float scale, f5b_stream_t stream) {
  return f5b::attn_fwd_tf32(q, k, v, ld, out, vt_ws, lens, lens_mod, B, H, n, scale, static_cast<cudaStream_t>(stream));
}
extern "C" size_t f5b_attn_tf32_ws_floats(int B, int H, int n) { return f5b::attn_tf32_ws_floats(B, H, n); }
