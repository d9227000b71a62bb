// The tcgen05 tile engine shared by the linear GEMM (gemm.cu) and the grouped position-embedding convolution
// (convpos.cu).
//
// sm_100a design: persistent warp-specialised kernel, one CTA per SM, 192 threads.
//   warp 0      : TMA producer — A tile [128 x 64] and B tile [umma_n x 64] (both K-major bf16) land in a ring of
//                 128B-swizzled shared-memory stages, mbarrier complete_tx.
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected thread), UMMA 128 x N x 16 (kind::f16, bf16 in,
//                 fp32 accumulate in TMEM), accumulator double-buffered (2 x BN columns) so the epilogue of tile i
//                 overlaps the main loop of tile i+1.
//   warps 2..5  : epilogue — tcgen05.ld 32x32b (thread = accumulator row), problem-specific fused math, global stores.
// A "Problem" policy supplies the tile -> coordinate mapping (which TMA boxes feed k-block kb of tile t) and the
// epilogue, so the same pipeline serves  C = A W^T  and the 31-tap implicit-GEMM convolution.
#pragma once
#include "common.cuh"

namespace f5b {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int ENGINE_THREADS = 192;

template <int BN>
struct EngCfg {
  static constexpr int STAGES = (BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8);
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <class P>
__global__ void __launch_bounds__(ENGINE_THREADS, 1)
engine_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const P p) {
  constexpr int BN = P::BN;
  using Cfg = EngCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total = p.num_tiles();
  const int kblocks = p.num_kblocks();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = Cfg::A_BYTES + p.b_tx_bytes();
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], tx);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          p.load(tile, kb, sa, sa + Cfg::A_BYTES, &full[stage], &tmA, &tmB);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BM, p.umma_n(), 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = smem_desc_sw128(a_addr + k * 32, 1024, 16);
            const uint64_t bd = smem_desc_sw128(b_addr + k * 32, 1024, 16);
            umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int ncols = p.tile_cols(tile);  // warp-uniform
      typename P::RowCtx ctx = p.row_ctx(tile, quad * 32 + lane);
#pragma unroll 1
      for (int c = 0; c * 32 < ncols; ++c) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c * 32, r);
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
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <class P>
static int launch_engine(const CUtensorMap& tmA, const CUtensorMap& tmB, const P& p, int total_tiles, cudaStream_t stream) {
  using Cfg = EngCfg<P::BN>;
  static bool configured = false;
  auto kern = engine_kernel<P>;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    configured = true;
  }
  const int grid = total_tiles < sm_count() ? total_tiles : sm_count();
  kern<<<grid, ENGINE_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

// ---- shared epilogue helpers: 32 consecutive columns of one accumulator row ------------------------------------
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* o, const float (&v)[32], int ncols_left, bool vec_ok) {
  if (ncols_left >= 32 && vec_ok) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 pk;
      pk.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
      pk.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
      pk.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
      pk.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
      reinterpret_cast<uint4*>(o)[q] = pk;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols_left) o[i] = __float2bfloat16(v[i]);
  }
}
__device__ __forceinline__ void store_row32_f32(float* o, const float (&v)[32], int ncols_left, bool vec_ok) {
  if (ncols_left >= 32 && vec_ok) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      reinterpret_cast<float4*>(o)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols_left) o[i] = v[i];
  }
}

}  // namespace f5b
