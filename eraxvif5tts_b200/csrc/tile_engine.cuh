// The tcgen05 tile engine shared by the linear GEMM (gemm.cu) and the grouped position-embedding convolution
// (convpos.cu).
//
// sm_100a design: persistent warp-specialised kernel, one CTA per SM, 320 threads.
//   warp 0      : TMA producer — A tile [128 x 64] and B tile [umma_n x 64] (both K-major bf16) land in a ring of
//                 128B-swizzled shared-memory stages, mbarrier complete_tx.  With P::CLUSTER == 2 two CTAs that work on
//                 vertically adjacent output tiles (same weight columns) form a cluster: each loads HALF of the B tile and
//                 TMA-multicasts it into both CTAs' stages, which cuts the L2 -> SM traffic per FLOP by a third (the
//                 linear GEMMs ran at the ~12 TB/s L2 cap without it); stage release is a multicast tcgen05.commit.
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected thread), UMMA 128 x N x 16 (kind::f16, bf16 in,
//                 fp32 accumulate in TMEM), accumulator double-buffered (2 x BN columns) so the epilogue of tile i
//                 overlaps the main loop of tile i+1.
//   warps 2..9  : epilogue — two warps per TMEM lane quadrant, each taking alternate column granules:
//                 tcgen05.ld 32x32b (thread = accumulator row) -> problem-specific fused math -> one of
//                   STORE_DIRECT : per-thread vectorised global stores (scatter epilogues: QKV head split, conv);
//                   STORE_BF16   : bf16 tile -> 128B-swizzled per-warp staging smem -> TMA store (coalesced, async);
//                   STORE_F32ADD : f32 tile -> staging smem -> TMA reduce-add: the residual stream x += tile is
//                                  accumulated in L2, x is never read into the SM;
//                   STORE_F32    : f32 tile -> staging smem -> TMA store (the tf32 operand mode's activation outputs).
// P::TF32 selects kind::tf32 (fp32 words in the stages, 32 per 128-byte row) instead of kind::f16 (64 bf16 per row).
// A "Problem" policy supplies the tile -> coordinate mapping (which TMA boxes feed k-block kb of tile t) and the
// epilogue math, so the same pipeline serves  C = A W^T  and the 31-tap implicit-GEMM convolution.
#pragma once
#include "common.cuh"

namespace f5b {

#ifndef F5B_F32ADD_DIRECT
#define F5B_F32ADD_DIRECT 0
#endif
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int ENGINE_THREADS = 64 + EPI_WARPS * 32;
enum { STORE_DIRECT = 0, STORE_BF16 = 1, STORE_F32ADD = 2, STORE_F32 = 3 };

template <int BN>
struct EngCfg {
  static constexpr int STAGES = (BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8);
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t STAGING_BYTES = EPI_WARPS * 4096;  // one [32 rows x 128 B] tile per epilogue warp
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// CTA-pair mode (cta_group::2): each CTA stages A [128 x 64] and HALF of the B tile [BN/2 x 64] per k-block
template <int BN>
struct EngCfg2 {
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 6 : 8;
  static constexpr uint32_t STAGING_BYTES = EPI_WARPS * 4096;
  static constexpr uint32_t TMEM_COLS = EngCfg<BN>::TMEM_COLS;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
};
template <bool TWOSM, int BN> struct EngCfgSel { using type = EngCfg<BN>; };
template <int BN> struct EngCfgSel<true, BN> { using type = EngCfg2<BN>; };

// store 8 consecutive 16-byte chunks (this lane's 128-byte row) into a [32 x 128 B] SWIZZLE_128B staging tile
__device__ __forceinline__ void stage_store16(uint8_t* stg, int lane, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(stg + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = v;
}

// P::DUAL (optional member, STORE_BF16 problems): the epilogue writes TWO bf16 tiles per accumulator tile — compute()'s values through
// tmC and P::second() of the same (bf16-rounded) values through tmD — e.g. the pre-activation AND the activated output of a training
// forward GEMM, which otherwise costs a separate sweep that re-reads the pre-activation from HBM.
// P::F32ADD_DIRECT (optional member, STORE_F32ADD problems): see the epilogue
template <class P, class = void> struct engine_f32add_direct { static constexpr bool value = false; };
template <class P> struct engine_f32add_direct<P, decltype((void)P::F32ADD_DIRECT)> { static constexpr bool value = P::F32ADD_DIRECT; };
template <class P, class = void> struct engine_is_dual { static constexpr bool value = false; };
template <class P> struct engine_is_dual<P, decltype((void)P::DUAL)> { static constexpr bool value = P::DUAL; };

// TWOSM: the two CTAs of the cluster form one tcgen05 CTA pair: a 256-row x BN tile per pair, the MMA issued by the leader CTA
// reads each operand half from each CTA's shared memory, so the shared-memory port of an SM sees half the operand traffic per flop
// (with cta_group::1 and a 128 x 256 tile the TMA fills + MMA operand reads add up to ~192 B/clk against a 128 B/clk port).
template <class P, bool TWOSM = false>
__global__ void __launch_bounds__(ENGINE_THREADS, 1)
engine_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const P p) {
  constexpr int BN = P::BN;
  constexpr bool DUAL = engine_is_dual<P>::value;
  static_assert(!DUAL || P::STORE == STORE_BF16, "dual-output epilogues are built for the bf16 TMA-store path");
  using Cfg = typename EngCfgSel<TWOSM, BN>::type;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int CL = P::CLUSTER;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  const int unit0 = (CL > 1) ? (int)(blockIdx.x / CL) : (int)blockIdx.x;      // first work unit of this CTA (cluster)
  const int ustride = (CL > 1) ? (int)(gridDim.x / CL) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if constexpr (P::STORE != STORE_DIRECT) prefetch_tmap(&tmC);
    if constexpr (DUAL) prefetch_tmap(&tmD);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], TWOSM ? 1 : CL);  // pair mode: ONE multicast commit of the leader arrives in each CTA
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], TWOSM ? 2 * EPI_WARPS : EPI_WARPS);  // pair mode: both CTAs' epilogue warps report to the leader
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (TWOSM) tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // PDL: everything above overlapped the previous kernel's tail; nothing global has been touched yet
  griddep_launch_dependents();

  const int total = p.num_units();  // work units: tiles, or vertical tile pairs when CL == 2
  const int kblocks = p.num_kblocks();

  // The single issuing lane of the producer / MMA warps is chosen by elect.sync, NOT by `lane == 0`: ptxas then knows the region is
  // single-threaded and emits plain UTMALDG / UTCHMMA / UTCBAR sequences from uniform registers; under a lane predicate it wraps EVERY
  // such instruction in an R2UR + ELECT + BRA.U.ANY "waterfall" loop (~65 clocks per tcgen05.mma — twice the execution time of a
  // 128 x 64 x 16 MMA; measured in profiles/r02_attention_fa_notes.md).
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = Cfg::A_BYTES + p.b_tx_bytes();
      for (int unit = unit0; unit < total; unit += ustride) {
        const int tile = p.unit_tile(unit, crank);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          if constexpr (TWOSM) {
            // both CTAs' loads complete on the LEADER's barrier (it alone issues the MMAs); bytes that land before the leader's
            // expect_tx only drive the pending count negative for a moment
            if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
            p.load2(tile, kb, sa, sa + Cfg::A_BYTES, mapa_u32(&full[stage], 0), &tmA, &tmB, crank);
          } else {
            mbar_arrive_expect_tx(&full[stage], tx);
            p.load(tile, kb, sa, sa + Cfg::A_BYTES, &full[stage], &tmA, &tmB, crank);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if ((!TWOSM || crank == 0) && elect_one()) {
      uint32_t idesc;
      if constexpr (TWOSM) idesc = p.idesc2();
      else idesc = p.idesc();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = unit0; unit < total; unit += ustride, ++it) {
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
            // (P::TF32: the stage holds 32 fp32 elements per 128-byte row, one MMA covers K = 8 of them — the same 32-byte K-step)
            if constexpr (P::TF32) {
              if constexpr (TWOSM) umma_tf32_2sm(d_tmem, p.a_desc(a_addr, k), p.b_desc(b_addr, k), idesc, (kb | k) != 0);
              else umma_tf32(d_tmem, p.a_desc(a_addr, k), p.b_desc(b_addr, k), idesc, (kb | k) != 0);
            } else {
              if constexpr (TWOSM) umma_bf16_2sm(d_tmem, p.a_desc(a_addr, k), p.b_desc(b_addr, k), idesc, (kb | k) != 0);
              else umma_bf16(d_tmem, p.a_desc(a_addr, k), p.b_desc(b_addr, k), idesc, (kb | k) != 0);
            }
          }
          if constexpr (TWOSM) umma_commit2_mcast(&empty[stage], (uint16_t)0x3);
          else if constexpr (CL > 1) umma_commit_mcast(&empty[stage], (uint16_t)((1u << CL) - 1));
          else umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (TWOSM) umma_commit2_mcast(&tfull[acc], (uint16_t)0x3);
        else umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;   // which of the two warps of the quadrant
    uint8_t* stg = staging + ew * 4096;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    int it = 0;
    for (int unit = unit0; unit < total; unit += ustride, ++it) {
      const int tile = p.unit_tile(unit, crank);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ncols = p.tile_cols(tile);  // warp-uniform
      typename P::RowCtx ctx = p.row_ctx(tile, quad * 32 + lane);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + lane_base + acc * BN;
      if constexpr (P::STORE == STORE_DIRECT) {
#pragma unroll 1
        for (int c = half; c * 32 < ncols; c += 2) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(t_acc + c * 32, r);
          tmem_ld_wait();
          p.epilogue(ctx, c * 32, r);
        }
      } else if constexpr (P::STORE == STORE_BF16) {
        // 64-column granules: two TMEM chunks -> one [32 x 64] bf16 staging tile -> one TMA store
#pragma unroll 1
        for (int g = half; g * 64 < ncols; g += 2) {
          if (elect_one()) bulk_wait_read0();  // the previous store has drained the staging tile
          __syncwarp();
          uint4 first[DUAL ? 8 : 1];  // DUAL: the packed first output, input of the second
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c0 = g * 64 + cc * 32;
            if (c0 < ncols) {  // warp-uniform
              uint32_t r[32];
              tmem_ld32(t_acc + c0, r);
              tmem_ld_wait();
              float v[32];
              p.compute(ctx, c0, r, v);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 pk;
                pk.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
                pk.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                pk.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
                pk.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
                stage_store16(stg, lane, cc * 4 + q, pk);
                if constexpr (DUAL) first[cc * 4 + q] = pk;
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_2d(&tmC, stg, p.out_col0(tile) + g * 64, p.out_row0(tile) + quad * 32);
            bulk_commit();
          }
          if constexpr (DUAL) {
            // second output = P::second(what the first output holds, i.e. the bf16-rounded values) through the same staging tile
            uint4 sec[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (g * 64 + (q >> 2) * 32 < ncols) {
                const uint32_t w[4] = {first[q].x, first[q].y, first[q].z, first[q].w};
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                  o[k] = pack_bf16(p.second(f.x), p.second(f.y));
                }
                sec[q] = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
            if (elect_one()) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (g * 64 + (q >> 2) * 32 < ncols) stage_store16(stg, lane, q, sec[q]);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              tma_store_2d(&tmD, stg, p.out_col0(tile) + g * 64, p.out_row0(tile) + quad * 32);
              bulk_commit();
            }
          }
        }
      } else {
        // 32-column f32 granules -> TMA reduce-add into the residual stream (STORE_F32ADD) or a plain TMA store (STORE_F32)
#pragma unroll 1
        for (int g = half; g * 32 < ncols; g += 2) {
          uint32_t r[32];
          tmem_ld32(t_acc + g * 32, r);
          tmem_ld_wait();
          float v[32];
          p.compute(ctx, g * 32, r, v);
          if constexpr (engine_f32add_direct<P>::value) {
            // experiment (-DF5B_F32ADD_DIRECT=1, off): the tile leaves as 16-byte reductions straight from the registers (no staging
            // tile, no TMA) — 128 KB less shared-memory traffic per 128 x 256 tile, but eight times as many, smaller, reduction requests
            // at the L2, and that rate is what binds: out-proj at cfg-2's shape 129 -> 198 us, FF2 203 -> 227 us, the cfg-2 step
            // 23.8 k -> 22.8 k frames/s (tools/kernel_ab.py gate).  The full-line TMA reduce-add stays.
            float* o = p.f32add_row(ctx, tile, g * 32);
            if (o != nullptr) {
#pragma unroll
              for (int q = 0; q < 8; ++q) red_add_v4_f32(o + q * 4, v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
            continue;
          }
          if (elect_one()) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            stage_store16(stg, lane, q,
                          make_uint4(__float_as_uint(v[q * 4]), __float_as_uint(v[q * 4 + 1]), __float_as_uint(v[q * 4 + 2]),
                                     __float_as_uint(v[q * 4 + 3])));
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            if constexpr (P::STORE == STORE_F32ADD) tma_reduce_add_2d(&tmC, stg, p.out_col0(tile) + g * 32, p.out_row0(tile) + quad * 32);
            else tma_store_2d(&tmC, stg, p.out_col0(tile) + g * 32, p.out_row0(tile) + quad * 32);
            bulk_commit();
          }
        }
      }
      __syncwarp();
      tc_fence_before();
      if (lane == 0) {
        if constexpr (TWOSM) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));  // the leader CTA issues the next MMAs into this buffer
        else mbar_arrive(&tempty[acc]);
      }
    }
    if constexpr (P::STORE != STORE_DIRECT) {
      if (elect_one()) bulk_wait0();  // (elect.sync with the full mask picks the same lane every time: the one that committed the groups)
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (TWOSM) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <class P, bool TWOSM = false>
static int launch_engine(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const P& p, int total_units,
                         cudaStream_t stream, const CUtensorMap* tmD_opt = nullptr) {
  const CUtensorMap& tmD = tmD_opt ? *tmD_opt : tmC;  // second output map of a DUAL problem
  using Cfg = typename EngCfgSel<TWOSM, P::BN>::type;
  static_assert(!TWOSM || P::CLUSTER == 2, "pair mode needs a 2-CTA cluster");
  static bool configured = false;
  auto kern = engine_kernel<P, TWOSM>;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    configured = true;
  }
  constexpr int CL = P::CLUSTER;
  const int max_clusters = sm_count() / CL;
  const int nclusters = total_units < max_clusters ? total_units : max_clusters;
  F5B_CUDA(launch_dep(kern, dim3(nclusters * CL), dim3(ENGINE_THREADS), Cfg::SMEM_BYTES, stream, CL, tmA, tmB, tmC, tmD, p));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

// ---- operand descriptors for UMMA K-step k (16 elements of the reduction dimension) of one [rows x 64] bf16 stage tile ------
// K-major tile (reduction dim contiguous): rows of 128 B, 8-row groups 1024 B apart; a K-step advances the start by 32 B.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int k) { return smem_desc_sw128(tile_addr + k * 32, 1024, 16); }
// MN-major tile (the M / N dim contiguous): 64-wide chunks of [64 reduction rows x 128 B] 8192 B apart (LBO); inside a chunk the
// 8-row groups are 1024 B apart (SBO); a K-step = 16 reduction rows = 2048 B.  (Form verified by the attention kernel's P.V.)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int k) { return smem_desc_sw128(tile_addr + k * 2048, 1024, 8192); }

// ---- shared epilogue helpers: 32 consecutive columns of one accumulator row ------------------------------------
// (partial chunks keep the 16-byte stores for every whole group of 8 / 4 columns: a 48-channel conv group ends with a 16-column chunk)
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* o, const float (&v)[32], int ncols_left, bool vec_ok) {
  if (vec_ok) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q * 8 + 8 <= ncols_left) {
        uint4 pk;
        pk.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
        pk.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
        pk.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
        pk.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
        reinterpret_cast<uint4*>(o)[q] = pk;
      } else {
#pragma unroll
        for (int i = q * 8; i < q * 8 + 8; ++i)
          if (i < ncols_left) o[i] = __float2bfloat16(v[i]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols_left) o[i] = __float2bfloat16(v[i]);
  }
}
__device__ __forceinline__ void store_row32_f32(float* o, const float (&v)[32], int ncols_left, bool vec_ok) {
  if (vec_ok) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q * 4 + 4 <= ncols_left) {
        reinterpret_cast<float4*>(o)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else {
#pragma unroll
        for (int i = q * 4; i < q * 4 + 4; ++i)
          if (i < ncols_left) o[i] = v[i];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols_left) o[i] = v[i];
  }
}

// bias for 32 consecutive columns (vectorised when the whole chunk is in range)
__device__ __forceinline__ void load_bias32(const float* bias, int n0, int left, float (&b)[32]) {
  if (bias == nullptr) {
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = 0.f;
  } else if (left >= 32) {
    const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = __ldg(b4 + q);
      b[q * 4] = t.x; b[q * 4 + 1] = t.y; b[q * 4 + 2] = t.z; b[q * 4 + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = i < left ? __ldg(bias + n0 + i) : 0.f;
  }
}
// same for a chunk that may be partial, 16-byte loads for every whole group of 4 (bias + n0 must be 16-byte aligned)
__device__ __forceinline__ void load_bias32_groups(const float* bias, int n0, int left, float (&b)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (bias != nullptr && q * 4 + 4 <= left) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(bias + n0) + q);
      b[q * 4] = t.x; b[q * 4 + 1] = t.y; b[q * 4 + 2] = t.z; b[q * 4 + 3] = t.w;
    } else {
#pragma unroll
      for (int i = q * 4; i < q * 4 + 4; ++i) b[i] = (bias != nullptr && i < left) ? __ldg(bias + n0 + i) : 0.f;
    }
  }
}

}  // namespace f5b
