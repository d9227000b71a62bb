// ConvPositionEmbedding's grouped Conv1d(D, D, k=31, groups=16, pad 15) + Mish as an implicit GEMM on tcgen05
// (/root/reference/src/f5_tts/model/modules.py:167-190; the residual add of model/backbones/dit.py:96 is fused into the
// second layer's epilogue).  It is a GROUPED conv (64 in/out channels per group for dim 1024), 8.1 MFLOP per token, so
// it runs on the tensor cores through the same tile engine as the linear layers:
//   output tile = 128 consecutive positions of one batch row x the NP (<= 64) output channels of one group;
//   k-block kb = tap kb: A = x[b, n0 + kb - pad : +128, g*cpg : +64] fetched straight from the token-major bf16
//   activation by a 3-D TMA box (negative / past-the-end positions are zero-filled by TMA = the conv's zero padding, per
//   batch row), B = the tap's [NP x 64] weight slab from the pre-packed weight tensor.  31 taps accumulate in TMEM.
#include "f5b_internal.h"
#include "tile_engine.cuh"

namespace f5b {

__device__ __forceinline__ float mish(float x) {
  // x * tanh(softplus(x)); tanh(ln(1+e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = t / (t + 2), t = e (e + 2).  Branch-free, nine instructions
  // (two on the SFU): the argument of the exponential is clamped at 20, where t / (t + 2) already rounds to 1 — the earlier
  // `if (x > 20) return x` cost a BSSY / BRA / BSYNC triple per ELEMENT in an epilogue that bounded the kernel
  const float e = ex2_approx(fminf(x, 20.f) * 1.4426950408889634f);
  const float t = e * (e + 2.f);
  return x * (t * rcp_approx(t + 2.f));
}
__device__ __forceinline__ float mish_precise(float x) {  // tf32 operand mode: libm exp and an IEEE division
  if (x > 20.f) return x;
  const float e = expf(x);
  const float t = e * (e + 2.f);
  return x * (t / (t + 2.f));
}

// TF32_ (the tf32 operand mode): x and the packed weights are fp32 words rounded to tf32, 32 channels per 128-byte stage row, so a
// tap takes `kper` = ceil(cpg / 32) k-blocks (input-channel halves); the mode-0 / mode-2 output is fp32 rounded to tf32.
template <int MODE, bool TF32_ = false>
struct ConvPosProblem {
  static constexpr int BN = 64;
  static constexpr int STORE = STORE_DIRECT;
  static constexpr int CLUSTER = 1;
  static constexpr bool TF32 = TF32_;
  int B, n, D, groups, cpg, NP, ksize, pad, n_tiles_seq;
  const float* bias;
  void* out;  // bf16, or fp32 in the tf32 mode
  float* resid;
  int kper;   // k-blocks per tap (1 unless TF32 and cpg > 32)

  struct RowCtx {
    size_t row_off;  // (b*n + pos) * D + g*cpg
    int ch0;         // g*cpg
    bool valid;
  };

  __device__ __forceinline__ int num_units() const { return B * n_tiles_seq * groups; }
  __device__ __forceinline__ int unit_tile(int unit, uint32_t) const { return unit; }
  __device__ __forceinline__ int num_kblocks() const { return ksize * kper; }
  __device__ __forceinline__ uint32_t umma_n() const { return NP; }
  __device__ __forceinline__ uint32_t idesc() const { return TF32 ? idesc_tf32(BM, umma_n(), 0, 0) : idesc_bf16(BM, umma_n(), 0, 0); }
  __device__ __forceinline__ uint64_t a_desc(uint32_t addr, int k) const { return desc_kmajor(addr, k); }
  __device__ __forceinline__ uint64_t b_desc(uint32_t addr, int k) const { return desc_kmajor(addr, k); }
  __device__ __forceinline__ uint32_t b_tx_bytes() const { return NP * 128; }
  __device__ __forceinline__ int tile_cols(int) const { return cpg; }
  __device__ __forceinline__ int out_col0(int) const { return 0; }
  __device__ __forceinline__ int out_row0(int) const { return 0; }
  __device__ __forceinline__ void compute(const RowCtx&, int, const uint32_t (&)[32], float (&)[32]) const {}
  __device__ __forceinline__ void decode(int tile, int& b, int& nt, int& g) const {
    g = tile % groups;
    const int t2 = tile / groups;
    nt = t2 % n_tiles_seq;
    b = t2 / n_tiles_seq;
  }
  __device__ __forceinline__ void load(int tile, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar, const CUtensorMap* tmA,
                                       const CUtensorMap* tmB, uint32_t) const {
    int b, nt, g;
    decode(tile, b, nt, g);
    if constexpr (TF32) {
      const int tap = kb / kper, ch = kb - tap * kper;
      tma_load_3d(sA, tmA, bar, g * cpg + ch * 32, nt * BM + tap - pad, b);
      tma_load_2d(sB, tmB, bar, 0, ((g * ksize + tap) * kper + ch) * NP);
    } else {
      tma_load_3d(sA, tmA, bar, g * cpg, nt * BM + kb - pad, b);
      tma_load_2d(sB, tmB, bar, 0, (g * ksize + kb) * NP);
    }
  }
  __device__ __forceinline__ RowCtx row_ctx(int tile, int r) const {
    int b, nt, g;
    decode(tile, b, nt, g);
    RowCtx c;
    const int pos = nt * BM + r;
    c.valid = pos < n;
    c.ch0 = g * cpg;
    c.row_off = ((size_t)b * n + pos) * D + c.ch0;
    return c;
  }
  __device__ __forceinline__ void epilogue(const RowCtx& c, int c0, const uint32_t (&r)[32]) const {
    if (!c.valid) return;
    const int left = cpg - c0;  // > 0; 32 for whole chunks, 16 for the second chunk of a 48-channel group
    const bool v4 = ((D | cpg) & 3) == 0;  // 16-byte aligned fp32 groups of 4 (bias, residual)
    float v[32];
    {
      float bb[32];
      if (v4) {
        load_bias32_groups(bias, c.ch0 + c0, left, bb);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bb[i] = (bias != nullptr && i < left) ? __ldg(bias + c.ch0 + c0 + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float t = __uint_as_float(r[i]) + bb[i];
        v[i] = MODE == 2 ? t : (TF32 ? mish_precise(t) : mish(t));  // mode 2: pre-activation output (training forward / transposed conv of the backward)
      }
    }
    if constexpr ((MODE == 0 || MODE == 2) && TF32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
      store_row32_f32(reinterpret_cast<float*>(out) + c.row_off + c0, v, left, v4);
    } else if constexpr (MODE == 0 || MODE == 2) {
      store_row32_bf16(reinterpret_cast<__nv_bfloat16*>(out) + c.row_off + c0, v, left, ((D | cpg) & 7) == 0);
    } else {
      float* o = resid + c.row_off + c0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (v4 && q * 4 + 4 <= left) {
          // residual += mish(conv): a 16-byte reduction at the L2 instead of load-add-store (each element is touched once per launch;
          // the read-modify-write epilogue ran at half the speed of the store-only one: 452 vs 222 us at cfg-2's shape)
          red_add_v4_f32(o + q * 4, v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
#pragma unroll
          for (int i = q * 4; i < q * 4 + 4; ++i)
            if (i < left) o[i] += v[i];
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Halo variant (the one that runs): the per-tap A tiles of an output tile are the SAME activation rows shifted by one row per
// tap, and the [NP x 64] weight slab of a tap is the same for every position tile of a group.  So one CTA work unit is a
// SUPER-TILE of up to 4 x 128 consecutive positions of one (batch row, group):
//   * its halo [4*128 + k - 1 rows x 64 channels] is fetched ONCE (4 TMA boxes of 136 rows, contiguous in smem);
//   * tap t of sub-tile s is fed to tcgen05.mma through a shared-memory descriptor whose start address is simply advanced by
//     (128 s + t) rows — measured on B200: the 128B swizzle is a function of the absolute smem address, so a start that is not
//     1024-byte aligned needs NO descriptor base offset (setting it corrupts the result);
//   * each tap's weight slab is streamed once per super-tile and feeds 4 x 4 MMAs (one barrier round trip per 512 tensor
//     clocks instead of per 128), accumulating into 4 x 64 TMEM columns, double-buffered across super-tiles (512 columns).
// The per-tap implicit GEMM on the generic engine (above) ran at the L2 -> SM cap with the tensor pipe 19 % busy.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HALO_SUB = 4;
constexpr int HALO_BOX_ROWS = 136;                       // 17 x 8 rows: every box starts on a 1024-byte swizzle boundary
constexpr int HALO_BOXES = 4;                            // 544 rows >= 4*128 + 30
constexpr uint32_t HALO_A_BYTES = HALO_BOXES * HALO_BOX_ROWS * 128;  // 68 KB
constexpr int HALO_B_STAGES = 8;
constexpr uint32_t HALO_B_BYTES = 64 * 128;              // 8 KB per stage (NP <= 64 rows)
constexpr uint32_t HALO_SMEM = 2 * HALO_A_BYTES + HALO_B_STAGES * HALO_B_BYTES + 1024 + 256;
// warps: 0 TMA producer, 1 and 10 tcgen05.mma issuers (even / odd sub-tiles: ONE issuing thread cannot keep the tensor pipe busy with
// 33-clock 128 x 64 x 16 MMAs, two threads share the pipe without loss), 2..9 epilogue
constexpr int HALO_THREADS = ENGINE_THREADS + 32;
constexpr int HALO_MMA_WARP2 = ENGINE_THREADS / 32;

// KSTEPS: K-steps of 16 input channels per tap that hold real channels (compile time, so the issue loop carries no predicates): the
// stage rows are 64 channels wide, but a group of 48 (F5TTS_Small, D = 768) fills only three of the four slices — the fourth would
// multiply the neighbour group's channels by packed zeros.
template <int MODE, int KSTEPS>
__global__ void __launch_bounds__(HALO_THREADS, 1)
convpos_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvPosProblem<MODE> p,
                    const int n_super) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * HALO_A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + HALO_B_STAGES * HALO_B_BYTES);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + HALO_B_STAGES;
  uint64_t* tfull = b_empty + HALO_B_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 2);  // one tcgen05.commit per issuing warp
        mbar_init(&tfull[i], 2);
        mbar_init(&tempty[i], EPI_WARPS);
      }
      for (int i = 0; i < HALO_B_STAGES; ++i) {
        mbar_init(&b_full[i], 1);
        mbar_init(&b_empty[i], 2);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // PDL (common.cuh): the prologue above overlaps the previous kernel's tail
  griddep_launch_dependents();
  const int total = p.B * n_super * p.groups;  // super-tiles
  const int ks = p.ksize;
  // super-tile u -> (b, st, g), g fastest; number of live 128-row sub-tiles
  auto decode_super = [&](int u, int& b, int& st, int& g, int& nsub) {
    g = u % p.groups;
    const int t2 = u / p.groups;
    st = t2 % n_super;
    b = t2 / n_super;
    const int left = p.n - st * HALO_SUB * BM;
    nsub = min(HALO_SUB, (left + BM - 1) / BM);
  };

  // issuing lanes chosen by elect.sync, not `lane == 0` (see tile_engine.cuh: plain UTMALDG / UTCHMMA sequences instead of one
  // R2UR + ELECT + BRA.U.ANY loop per instruction)
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x, ++it) {
        int b, st, g, nsub;
        decode_super(u, b, st, g, nsub);
        const int ab = it & 1;
        const int rows = nsub * BM + ks - 1;
        const int nbox = (rows + HALO_BOX_ROWS - 1) / HALO_BOX_ROWS;
        mbar_wait(&a_empty[ab], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[ab], nbox * HALO_BOX_ROWS * 128);
        for (int q = 0; q < nbox; ++q)
          tma_load_3d(sA + ab * HALO_A_BYTES + q * HALO_BOX_ROWS * 128, &tmA, &a_full[ab], g * p.cpg,
                      st * HALO_SUB * BM - p.pad + q * HALO_BOX_ROWS, b);
        for (int k = 0; k < ks; ++k) {
          mbar_wait(&b_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&b_full[stage], p.NP * 128);
          tma_load_2d(sB + stage * HALO_B_BYTES, &tmB, &b_full[stage], 0, (g * ks + k) * p.NP);
          if (++stage == HALO_B_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == HALO_MMA_WARP2) {
    const int mw = warp == 1 ? 0 : 1;  // this issuer's sub-tiles: sblk & 1 == mw
    if (elect_one()) {
      const uint32_t idesc = idesc_bf16(BM, p.NP, 0, 0);
      // ISSUE COST.  A 128 x 64 x 16 MMA executes in ~33 clocks, so the issuing side has about 30 instructions per MMA; building both
      // 64-bit shared-memory descriptors from addresses each time (shift / mask / or on the uniform datapath plus R2UR moves) was ~12
      // instructions and three R2UR per MMA.  All descriptors of a super-tile differ only in the 14-bit start-address field, which never
      // carries (shared memory < 256 KB): the low word is advanced by constants — +8 per tap row (128 B), +1024 per 128-row sub-tile,
      // +2 per K-step (32 B) — and the high word is fixed; the sub-tiles are split over two issuing warps.  (Measured: worth 5 %; what
      // bounded the kernel was its epilogue, profiles/r02_convpos_epilogue.md.)
      const uint64_t d0 = smem_desc_sw128(0, 1024, 16);
      const uint32_t d_hi = (uint32_t)(d0 >> 32), d_lo0 = (uint32_t)d0;
      auto mk = [&](uint32_t lo) { return ((uint64_t)d_hi << 32) | (uint64_t)lo; };
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x, ++it) {
        int b, st, g, nsub;
        decode_super(u, b, st, g, nsub);
        const int acc = it & 1, ab = it & 1;
        const uint32_t ph2 = (it >> 1) & 1;
        mbar_wait(&tempty[acc], ph2 ^ 1);
        mbar_wait(&a_full[ab], ph2);
        tc_fence_after();
        const uint32_t a_lo = d_lo0 + (smem_u32(sA + ab * HALO_A_BYTES) >> 4);
        const uint32_t d_acc = tmem_base + acc * (HALO_SUB * 64);
        for (int k = 0; k < ks; ++k) {
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          const uint32_t b_lo = d_lo0 + (smem_u32(sB + stage * HALO_B_BYTES) >> 4);
          const uint32_t a_lo_k = a_lo + k * 8;  // tap k: the halo rows shifted down by k
#pragma unroll
          for (int half = 0; half < HALO_SUB / 2; ++half) {
            const int sblk = 2 * half + mw;
            if (sblk < nsub) {
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk)
                umma_bf16(d_acc + sblk * 64, mk(a_lo_k + sblk * (BM * 8) + kk * 2), mk(b_lo + kk * 2), idesc, (k | kk) != 0);
            }
          }
          umma_commit(&b_empty[stage]);
          if (++stage == HALO_B_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&a_empty[ab]);
        umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int chunks = (p.cpg + 31) / 32;
    int it = 0;
    for (int u = blockIdx.x; u < total; u += gridDim.x, ++it) {
      int b, st, g, nsub;
      decode_super(u, b, st, g, nsub);
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int w = half; w < nsub * chunks; w += 2) {  // (sub-tile, 32-column chunk) pairs, alternating between the warp pair
        const int sblk = w / chunks, c = w - sblk * chunks;
        typename ConvPosProblem<MODE>::RowCtx ctx;  // (built from the super-tile's coordinates: no tile-index divisions per visit)
        const int pos = (st * HALO_SUB + sblk) * BM + quad * 32 + lane;
        ctx.valid = pos < p.n;
        ctx.ch0 = g * p.cpg;
        ctx.row_off = ((size_t)b * p.n + pos) * p.D + ctx.ch0;
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + lane_base + acc * (HALO_SUB * 64) + sblk * 64 + c * 32, r);
        tmem_ld_wait();
        p.epilogue(ctx, c * 32, r);
      }
      __syncwarp();
      tc_fence_before();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_convpos_halo = 1;  // 0: per-tap loads on the generic engine (kept for A/B measurements), 1: super-tile halo kernel

template <int MODE, int KSTEPS>
static int launch_halo_k(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvPosProblem<MODE>& p, cudaStream_t stream) {
  static bool configured = false;
  auto kern = convpos_halo_kernel<MODE, KSTEPS>;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HALO_SMEM));
    configured = true;
  }
  const int n_super = (p.n + HALO_SUB * BM - 1) / (HALO_SUB * BM);
  const int total = p.B * n_super * p.groups;
  const int grid = total < sm_count() ? total : sm_count();
  F5B_CUDA(launch_dep(kern, dim3(grid), dim3(HALO_THREADS), HALO_SMEM, stream, 1, tmA, tmB, p, n_super));
  F5B_CUDA(cudaGetLastError());
  return 0;
}
template <int MODE>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvPosProblem<MODE>& p, cudaStream_t stream) {
  // (any KSTEPS >= ceil(cpg / 16) is correct: the packed weights are zero beyond cpg)
  return p.cpg <= 48 ? launch_halo_k<MODE, 3>(tmA, tmB, p, stream) : launch_halo_k<MODE, 4>(tmA, tmB, p, stream);
}

__global__ void pack_convpos_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpk, int D, int groups, int ksize,
                                    int cpg, int NP, int transpose_flip) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = (int64_t)groups * ksize * NP * 64;
  if (i >= tot) return;
  const int ci = (int)(i & 63);
  int64_t t = i >> 6;
  const int co = (int)(t % NP);
  t /= NP;
  const int k = (int)(t % ksize);
  const int g = (int)(t / ksize);
  float v = 0.f;
  if (co < cpg && ci < cpg)
    // transpose_flip: the conv that maps an output gradient back to the input (swap in/out channel, reverse the taps)
    v = transpose_flip ? w[((size_t)(g * cpg + ci) * cpg + co) * ksize + (ksize - 1 - k)] : w[((size_t)(g * cpg + co) * cpg + ci) * ksize + k];
  wpk[i] = __float2bfloat16(v);
}

static inline int round16(int x) { return (x + 15) / 16 * 16; }

int convpos(const void* x, const void* wpk, const float* bias, void* out, float* resid, int B, int n, int D, int groups,
            int ksize, int mode, cudaStream_t stream) {
  F5B_CHECK(x && wpk && (bias || mode == 2), "f5b_convpos: null pointer");
  F5B_CHECK(B > 0 && n > 0 && D > 0 && groups > 0 && D % groups == 0 && (ksize & 1) == 1, "f5b_convpos: bad shape");
  const int cpg = D / groups;
  F5B_CHECK(cpg <= 64 && (D & 7) == 0, "f5b_convpos: channels per group %d must be <= 64 and D a multiple of 8", cpg);
  F5B_CHECK(mode >= 0 && mode <= 2 && (mode != 1 ? out != nullptr : resid != nullptr), "f5b_convpos: null output for mode %d", mode);
  const int NP = round16(cpg);
  LaunchScope scope(K_CONVPOS, stream, 2.0 * B * n * (double)D * cpg * ksize, (double)B * n * D * (mode != 1 ? 4.0 : 10.0));
  CUtensorMap tmA, tmB;
  const bool halo = g_convpos_halo != 0 && HALO_SUB * BM + ksize - 1 <= HALO_BOXES * HALO_BOX_ROWS;
  if (make_tmap_3d(&tmA, x, 2, (uint64_t)D, (uint64_t)n, (uint64_t)B, (uint64_t)D * 2, (uint64_t)n * D * 2, 64, halo ? HALO_BOX_ROWS : BM, 1,
                   true))
    return -1;
  if (make_tmap_2d(&tmB, wpk, 2, 64, (uint64_t)groups * ksize * NP, 128, 64, NP, true)) return -1;
  const int nts = (n + BM - 1) / BM;
  const int total = B * nts * groups;
  if (halo) {
    if (mode == 0) {
      ConvPosProblem<0> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, 1};
      return launch_halo(tmA, tmB, p, stream);
    }
    if (mode == 2) {
      ConvPosProblem<2> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, 1};
      return launch_halo(tmA, tmB, p, stream);
    }
    ConvPosProblem<1> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, 1};
    return launch_halo(tmA, tmB, p, stream);
  }
  if (mode == 0) {
    ConvPosProblem<0> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, 1};
    return launch_engine(tmA, tmB, tmA, p, total, stream);
  }
  if (mode == 2) {
    ConvPosProblem<2> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, 1};
    return launch_engine(tmA, tmB, tmA, p, total, stream);
  }
  ConvPosProblem<1> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, 1};
  return launch_engine(tmA, tmB, tmA, p, total, stream);
}

// ---- tf32 operand mode: per-tap loads on the generic engine, kind::tf32 (a precision mode, not a throughput one) -------------
__global__ void pack_convpos_tf32_kernel(const float* __restrict__ w, float* __restrict__ wpk, int groups, int ksize, int cpg, int NP,
                                         int kper) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = (int64_t)groups * ksize * kper * NP * 32;
  if (i >= tot) return;
  const int cl = (int)(i & 31);
  int64_t t = i >> 5;
  const int co = (int)(t % NP);
  t /= NP;
  const int ch = (int)(t % kper);
  t /= kper;
  const int k = (int)(t % ksize);
  const int g = (int)(t / ksize);
  const int ci = ch * 32 + cl;
  float v = 0.f;
  if (co < cpg && ci < cpg) v = w[((size_t)(g * cpg + co) * cpg + ci) * ksize + k];
  wpk[i] = tf32_rn(v);
}

static inline int tf32_kper(int cpg) { return (cpg + 31) / 32; }

int convpos_tf32(const float* x, const float* wpk, const float* bias, float* out, float* resid, int B, int n, int D, int groups, int ksize,
                 int mode, cudaStream_t stream) {
  F5B_CHECK(x && wpk && bias, "f5b_convpos_tf32: null pointer");
  F5B_CHECK(B > 0 && n > 0 && D > 0 && groups > 0 && D % groups == 0 && (ksize & 1) == 1, "f5b_convpos_tf32: bad shape");
  const int cpg = D / groups;
  F5B_CHECK(cpg <= 64 && (D & 3) == 0, "f5b_convpos_tf32: channels per group %d must be <= 64 and D a multiple of 4", cpg);
  F5B_CHECK((mode == 0 && out != nullptr) || (mode == 1 && resid != nullptr), "f5b_convpos_tf32: mode 0 needs out, mode 1 needs resid");
  const int NP = round16(cpg), kper = tf32_kper(cpg);
  LaunchScope scope(K_CONVPOS, stream, 2.0 * B * n * (double)D * cpg * ksize, (double)B * n * D * (mode != 1 ? 8.0 : 12.0));
  CUtensorMap tmA, tmB;
  if (make_tmap_3d(&tmA, x, 4, (uint64_t)D, (uint64_t)n, (uint64_t)B, (uint64_t)D * 4, (uint64_t)n * D * 4, 32, BM, 1, true)) return -1;
  if (make_tmap_2d(&tmB, wpk, 4, 32, (uint64_t)groups * ksize * kper * NP, 128, 32, NP, true)) return -1;
  const int nts = (n + BM - 1) / BM;
  const int total = B * nts * groups;
  if (mode == 0) {
    ConvPosProblem<0, true> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, kper};
    return launch_engine(tmA, tmB, tmA, p, total, stream);
  }
  ConvPosProblem<1, true> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, out, resid, kper};
  return launch_engine(tmA, tmB, tmA, p, total, stream);
}

}  // namespace f5b

extern "C" {

int f5b_convpos_tf32(const float* x, const float* wpk, const float* bias, float* out, float* resid, int B, int n, int D, int groups,
                     int ksize, int mode, f5b_stream_t stream) {
  return f5b::convpos_tf32(x, wpk, bias, out, resid, B, n, D, groups, ksize, mode, static_cast<cudaStream_t>(stream));
}
size_t f5b_convpos_packed_elems_tf32(int D, int groups, int ksize) {
  if (groups <= 0 || D % groups != 0) return 0;
  const int cpg = D / groups;
  return (size_t)groups * ksize * f5b::tf32_kper(cpg) * f5b::round16(cpg) * 32;
}
int f5b_pack_convpos_weight_tf32(const float* w, float* wpk, int D, int groups, int ksize, f5b_stream_t stream) {
  using namespace f5b;
  F5B_CHECK(w && wpk && groups > 0 && D % groups == 0 && D / groups <= 64, "f5b_pack_convpos_weight_tf32: bad shape");
  const int cpg = D / groups, NP = round16(cpg), kper = tf32_kper(cpg);
  const int64_t tot = (int64_t)groups * ksize * kper * NP * 32;
  LaunchScope scope(K_ELEMENTWISE, static_cast<cudaStream_t>(stream), 0, 8.0 * tot);
  pack_convpos_tf32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, wpk, groups, ksize, cpg, NP, kper);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_convpos(const void* x, const void* wpk, const float* bias, void* out, float* resid, int B, int n, int D, int groups,
                int ksize, int mode, f5b_stream_t stream) {
  return f5b::convpos(x, wpk, bias, out, resid, B, n, D, groups, ksize, mode, static_cast<cudaStream_t>(stream));
}

size_t f5b_convpos_packed_elems(int D, int groups, int ksize) {
  if (groups <= 0 || D % groups != 0) return 0;
  return (size_t)groups * ksize * f5b::round16(D / groups) * 64;
}

static int pack_convpos(const float* w, void* wpk, int D, int groups, int ksize, int transpose_flip, f5b_stream_t stream) {
  using namespace f5b;
  F5B_CHECK(w && wpk && groups > 0 && D % groups == 0 && D / groups <= 64, "f5b_pack_convpos_weight: bad shape");
  const int cpg = D / groups, NP = round16(cpg);
  const int64_t tot = (int64_t)groups * ksize * NP * 64;
  LaunchScope scope(K_ELEMENTWISE, static_cast<cudaStream_t>(stream), 0, 6.0 * tot);
  pack_convpos_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<__nv_bfloat16*>(wpk), D, groups, ksize, cpg, NP, transpose_flip);
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_pack_convpos_weight(const float* w, void* wpk, int D, int groups, int ksize, f5b_stream_t stream) {
  return pack_convpos(w, wpk, D, groups, ksize, 0, stream);
}
int f5b_pack_convpos_weight_t(const float* w, void* wpk, int D, int groups, int ksize, f5b_stream_t stream) {
  return pack_convpos(w, wpk, D, groups, ksize, 1, stream);
}

}  // extern "C"

extern "C" void f5b_debug_convpos_mode(int halo) { f5b::g_convpos_halo = halo; }
