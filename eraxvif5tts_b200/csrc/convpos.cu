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
  // x * tanh(softplus(x)); tanh(ln(1+e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = t / (t + 2), t = e (e + 2)
  if (x > 20.f) return x;
  const float e = __expf(x);
  const float t = e * (e + 2.f);
  return x * __fdividef(t, t + 2.f);
}

template <int MODE>
struct ConvPosProblem {
  static constexpr int BN = 64;
  static constexpr int STORE = STORE_DIRECT;
  static constexpr int CLUSTER = 1;
  int B, n, D, groups, cpg, NP, ksize, pad, n_tiles_seq;
  const float* bias;
  __nv_bfloat16* out;
  float* resid;

  struct RowCtx {
    size_t row_off;  // (b*n + pos) * D + g*cpg
    int ch0;         // g*cpg
    bool valid;
  };

  __device__ __forceinline__ int num_units() const { return B * n_tiles_seq * groups; }
  __device__ __forceinline__ int unit_tile(int unit, uint32_t) const { return unit; }
  __device__ __forceinline__ int num_kblocks() const { return ksize; }
  __device__ __forceinline__ uint32_t umma_n() const { return NP; }
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
    tma_load_3d(sA, tmA, bar, g * cpg, nt * BM + kb - pad, b);
    tma_load_2d(sB, tmB, bar, 0, (g * ksize + kb) * NP);
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
    const int left = cpg - c0;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float bb = 0.f;
      if (i < left) bb = __ldg(bias + c.ch0 + c0 + i);
      v[i] = mish(__uint_as_float(r[i]) + bb);
    }
    if constexpr (MODE == 0) {
      store_row32_bf16(out + c.row_off + c0, v, left, ((D | cpg) & 7) == 0);
    } else {
      float* o = resid + c.row_off + c0;
      if (left >= 32 && ((D | cpg) & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 x = reinterpret_cast<float4*>(o)[q];
          x.x += v[q * 4];
          x.y += v[q * 4 + 1];
          x.z += v[q * 4 + 2];
          x.w += v[q * 4 + 3];
          reinterpret_cast<float4*>(o)[q] = x;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < left) o[i] += v[i];
      }
    }
  }
};

__global__ void pack_convpos_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpk, int D, int groups, int ksize,
                                    int cpg, int NP) {
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
  if (co < cpg && ci < cpg) v = w[((size_t)(g * cpg + co) * cpg + ci) * ksize + k];
  wpk[i] = __float2bfloat16(v);
}

static inline int round16(int x) { return (x + 15) / 16 * 16; }

int convpos(const void* x, const void* wpk, const float* bias, void* out, float* resid, int B, int n, int D, int groups,
            int ksize, int mode, cudaStream_t stream) {
  F5B_CHECK(x && wpk && bias, "f5b_convpos: null pointer");
  F5B_CHECK(B > 0 && n > 0 && D > 0 && groups > 0 && D % groups == 0 && (ksize & 1) == 1, "f5b_convpos: bad shape");
  const int cpg = D / groups;
  F5B_CHECK(cpg <= 64 && (D & 7) == 0, "f5b_convpos: channels per group %d must be <= 64 and D a multiple of 8", cpg);
  F5B_CHECK(mode == 0 ? out != nullptr : resid != nullptr, "f5b_convpos: null output for mode %d", mode);
  const int NP = round16(cpg);
  LaunchScope scope(K_CONVPOS, stream, 2.0 * B * n * (double)D * cpg * ksize, (double)B * n * D * (mode == 0 ? 4.0 : 10.0));
  CUtensorMap tmA, tmB;
  if (make_tmap_3d(&tmA, x, 2, (uint64_t)D, (uint64_t)n, (uint64_t)B, (uint64_t)D * 2, (uint64_t)n * D * 2, 64, BM, 1, true))
    return -1;
  if (make_tmap_2d(&tmB, wpk, 2, 64, (uint64_t)groups * ksize * NP, 128, 64, NP, true)) return -1;
  const int nts = (n + BM - 1) / BM;
  const int total = B * nts * groups;
  if (mode == 0) {
    ConvPosProblem<0> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, reinterpret_cast<__nv_bfloat16*>(out), resid};
    return launch_engine(tmA, tmB, tmA, p, total, stream);
  }
  ConvPosProblem<1> p{B, n, D, groups, cpg, NP, ksize, ksize / 2, nts, bias, reinterpret_cast<__nv_bfloat16*>(out), resid};
  return launch_engine(tmA, tmB, tmA, p, total, stream);
}

}  // namespace f5b

extern "C" {

int f5b_convpos(const void* x, const void* wpk, const float* bias, void* out, float* resid, int B, int n, int D, int groups,
                int ksize, int mode, f5b_stream_t stream) {
  return f5b::convpos(x, wpk, bias, out, resid, B, n, D, groups, ksize, mode, static_cast<cudaStream_t>(stream));
}

size_t f5b_convpos_packed_elems(int D, int groups, int ksize) {
  if (groups <= 0 || D % groups != 0) return 0;
  return (size_t)groups * ksize * f5b::round16(D / groups) * 64;
}

int f5b_pack_convpos_weight(const float* w, void* wpk, int D, int groups, int ksize, f5b_stream_t stream) {
  using namespace f5b;
  F5B_CHECK(w && wpk && groups > 0 && D % groups == 0 && D / groups <= 64, "f5b_pack_convpos_weight: bad shape");
  const int cpg = D / groups, NP = round16(cpg);
  const int64_t tot = (int64_t)groups * ksize * NP * 64;
  LaunchScope scope(K_ELEMENTWISE, static_cast<cudaStream_t>(stream), 0, 6.0 * tot);
  pack_convpos_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<__nv_bfloat16*>(wpk), D, groups, ksize, cpg, NP);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
