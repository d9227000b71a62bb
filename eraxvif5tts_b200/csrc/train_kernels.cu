// Backward-pass building blocks of the training step (autograd of CFM.forward, /root/reference/src/f5_tts/model/cfm.py:210-283,
// through DiTBlock /root/reference/src/f5_tts/model/modules.py:627-641): everything that is not a GEMM or attention.
// All of these are HBM-bound row / column sweeps; the per-batch-row modulation gradients and bias gradients are column
// reductions done in the same sweep that produces the tensor the next GEMM consumes (one pass over the activation, fp32 atomics
// for the few thousand column sums).
#include "common.cuh"
#include "f5b_internal.h"
#include "dropout.cuh"

namespace f5b {

constexpr int CT_ROWS = 64;     // rows per CTA of the column-thread kernels
constexpr int CT_THREADS = 256; // each thread owns column pairs {2t, 2t+1} + k*512

__device__ __forceinline__ float act_eval(int act, float x) {
  if (act == F5B_ACT_GELU_TANH) {
    const float u = 0.7978845608028654f * x * fmaf(0.044715f * x, x, 1.0f);
    return 0.5f * x * (1.0f + tanh_approx(u));
  } else if (act == F5B_ACT_GELU_ERF) {
    return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
  } else if (act == F5B_ACT_SILU) {
    return x / (1.0f + __expf(-x));
  } else if (act == F5B_ACT_MISH) {
    const float sp = x > 20.f ? x : log1pf(__expf(x));
    return x * tanhf(sp);
  }
  return x;
}
__device__ __forceinline__ float act_grad(int act, float x) {
  if (act == F5B_ACT_GELU_TANH) {
    const float k = 0.7978845608028654f;
    const float u = k * x * fmaf(0.044715f * x, x, 1.0f);
    const float t = tanh_approx(u);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k * fmaf(3.0f * 0.044715f * x, x, 1.0f);
  } else if (act == F5B_ACT_GELU_ERF) {
    return 0.5f * (1.0f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
  } else if (act == F5B_ACT_SILU) {
    const float s = 1.0f / (1.0f + __expf(-x));
    return s * (1.0f + x * (1.0f - s));
  } else if (act == F5B_ACT_MISH) {
    const float sp = x > 20.f ? x : log1pf(__expf(x));
    const float t = tanhf(sp);
    const float s = 1.0f / (1.0f + __expf(-x));
    return t + x * (1.0f - t * t) * s;
  }
  return 1.0f;
}

// out[r, :] = x[r, :] + gate[b, :] * z[r, :]   (rows with pos >= len[b] keep x: the reference zeroes the attention branch there,
// model/modules.py:499-501).  The un-fused training form of the gate-residual GEMM epilogue: z is kept for the gate gradient.
__global__ void gate_add_kernel(const float* x, const __nv_bfloat16* __restrict__ z, const float* __restrict__ gate,
                                int64_t gate_bstride, const int32_t* __restrict__ lens, float* out, int n, int C) {  // out may alias x
  const int b = blockIdx.y;
  const int pos = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pos >= n) return;
  const size_t row = (size_t)b * n + pos;
  const bool live = lens == nullptr || pos < __ldg(lens + b);
  const float* g = gate ? gate + (size_t)b * gate_bstride : nullptr;
  for (int c = (threadIdx.x & 31) * 2; c < C; c += 64) {
    float2 v = *reinterpret_cast<const float2*>(x + row * C + c);
    if (live) {
      const float2 zz = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(z + row * C + c));
      const float g0 = g ? __ldg(g + c) : 1.f, g1 = g ? __ldg(g + c + 1) : 1.f;
      v.x = fmaf(g0, zz.x, v.x);
      v.y = fmaf(g1, zz.y, v.y);
    }
    *reinterpret_cast<float2*>(out + row * C + c) = v;
  }
}

// dz[r, :] = bf16(gate[b, :] * dx[r, :]) (0 on masked rows);  dgate[b, :] += sum_r dx[r, :] * z[r, :];  dbias[:] += sum_r dz[r, :]
// thread = 4 adjacent columns (16-byte fp32 / 8-byte bf16 accesses), 64 rows per CTA
__global__ void __launch_bounds__(CT_THREADS) gate_bwd_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ z,
                                                              const float* __restrict__ gate, int64_t gate_bstride,
                                                              const int32_t* __restrict__ lens, __nv_bfloat16* __restrict__ dz,
                                                              float* __restrict__ dgate, float* __restrict__ dbias, int n, int C,
                                                              const Drop dr) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * CT_ROWS;
  const int p1 = min(n, p0 + CT_ROWS);
  const int live_end = lens ? min(p1, __ldg(lens + b)) : p1;
  const float* g = gate ? gate + (size_t)b * gate_bstride : nullptr;
  for (int c = threadIdx.x * 4; c < C; c += 4 * CT_THREADS) {
    float gv[4] = {1.f, 1.f, 1.f, 1.f};
    if (g) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(g + c));
      gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w;
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f}, s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int pos = p0; pos < p1; ++pos) {
      const size_t off = ((size_t)b * n + pos) * C + c;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (pos < live_end) {
        const float4 d4 = *reinterpret_cast<const float4*>(dx + off);
        float d[4] = {d4.x, d4.y, d4.z, d4.w};
        if (dr.thr16) {  // z went through dropout before the gate: d(x_out)/d(z) = gate * mask / (1 - p)
          float m[4];
          drop_mult4(dr, off, m);
#pragma unroll
          for (int i = 0; i < 4; ++i) d[i] *= m[i];
        }
        if (z != nullptr) {
          const uint2 zz = *reinterpret_cast<const uint2*>(z + off);
          const float2 z01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zz.x));
          const float2 z23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zz.y));
          a[0] = fmaf(d[0], z01.x, a[0]); a[1] = fmaf(d[1], z01.y, a[1]); a[2] = fmaf(d[2], z23.x, a[2]); a[3] = fmaf(d[3], z23.y, a[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[i] = gv[i] * d[i];
          s[i] += o[i];
        }
      }
      *reinterpret_cast<uint2*>(dz + off) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    }
    if (dgate && z != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(dgate + (size_t)b * gate_bstride + c + i, a[i]);
    }
    if (dbias) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(dbias + c + i, s[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 16-byte column-strip forms of gate_bwd / act_bwd (C a multiple of 8).  CTA = 8 warps on a [256 rows x 256 columns] block: a warp
// streams 32 consecutive rows of the block, lane = 8 adjacent columns (16-byte bf16 / 2 x 16-byte fp32 accesses, 512 B - 1 KB
// contiguous per warp per row).  Every batch of 4 rows is LOADED IN FULL before anything is computed or stored (act_bwd runs in
// place — dh aliases du — so without the explicit batching each row's store fences the next row's loads).  The column sums of the 8
// warps are folded in shared memory and leave the CTA as ONE fp32 atomic per column: rows / 256 atomics per column instead of
// rows / 64 — measured (ncu, r02) the r01 kernels were bound by the same-address atomics at the L2 (a 32-row variant with twice the
// atomics ran 2.8 x SLOWER at the same traffic, `drain` as the top stall).
// ---------------------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 256;   // rows per CTA
constexpr int CS_WROWS = 32;   // consecutive rows per warp
constexpr int CS_COLS = 256;   // columns per CTA = 32 lanes x 8
constexpr int CS_BATCH = 4;

__device__ __forceinline__ void bf8_to_f32(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
// fold the 8 warps' per-lane column sums and add them to out[c0 + 0..255] (one atomic per column)
__device__ __forceinline__ void cs_fold_atomic(float (*red)[CS_COLS], const float (&s)[8], float* out, int c0, int C, bool enabled) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  *reinterpret_cast<float4*>(&red[warp][lane * 8]) = make_float4(s[0], s[1], s[2], s[3]);
  *reinterpret_cast<float4*>(&red[warp][lane * 8 + 4]) = make_float4(s[4], s[5], s[6], s[7]);
  __syncthreads();
  const int t = threadIdx.x;
  if (enabled && c0 + t < C) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][t];
    atomicAdd(out + c0 + t, a);
  }
}

__global__ void __launch_bounds__(256) gate_bwd8_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ z,
                                                        const float* __restrict__ gate, int64_t gate_bstride,
                                                        const int32_t* __restrict__ lens, __nv_bfloat16* __restrict__ dz,
                                                        float* __restrict__ dgate, float* __restrict__ dbias, int n, int C,
                                                        const Drop dr) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  __shared__ __align__(16) float red[8][CS_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.z * CS_COLS;
  const int c = c0 + lane * 8;
  const bool col_ok = c < C;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * CS_ROWS + warp * CS_WROWS;
  const int p1 = min(n, p0 + CS_WROWS);
  const int live_end = lens ? min(p1, __ldg(lens + b)) : p1;
  float gv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gv[i] = 1.f;
  if (gate && col_ok) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * gate_bstride + c));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * gate_bstride + c) + 1);
    gv[0] = t0.x; gv[1] = t0.y; gv[2] = t0.z; gv[3] = t0.w; gv[4] = t1.x; gv[5] = t1.y; gv[6] = t1.z; gv[7] = t1.w;
  }
  float a[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = s[i] = 0.f;
  if (col_ok) {
    for (int pb = p0; pb < p1; pb += CS_BATCH) {
      float4 d0[CS_BATCH], d1[CS_BATCH];
      uint4 zz[CS_BATCH];
#pragma unroll
      for (int k = 0; k < CS_BATCH; ++k) {
        const size_t off = ((size_t)b * n + pb + k) * C + c;
        if (pb + k < live_end) {
          d0[k] = *reinterpret_cast<const float4*>(dx + off);
          d1[k] = *reinterpret_cast<const float4*>(dx + off + 4);
          if (z != nullptr) zz[k] = *reinterpret_cast<const uint4*>(z + off);
        }
      }
#pragma unroll
      for (int k = 0; k < CS_BATCH; ++k) {
        if (pb + k < p1) {
          const size_t off = ((size_t)b * n + pb + k) * C + c;
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = 0.f;
          if (pb + k < live_end) {
            float d[8] = {d0[k].x, d0[k].y, d0[k].z, d0[k].w, d1[k].x, d1[k].y, d1[k].z, d1[k].w};
            if (dr.thr16) {  // z went through dropout before the gate: d(x_out)/d(z) = gate * mask / (1 - p)
              float m0[4], m1[4];
              drop_mult4(dr, off, m0);
              drop_mult4(dr, off + 4, m1);
#pragma unroll
              for (int i = 0; i < 4; ++i) { d[i] *= m0[i]; d[4 + i] *= m1[i]; }
            }
            if (z != nullptr) {
              float zf[8];
              bf8_to_f32(zz[k], zf);
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = fmaf(d[i], zf[i], a[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i] = gv[i] * d[i];
              s[i] += o[i];
            }
          }
          *reinterpret_cast<uint4*>(dz + off) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
      }
    }
  }
  if (dgate != nullptr && z != nullptr) cs_fold_atomic(red, a, dgate + (size_t)b * gate_bstride, c0, C, true);
  if (dbias != nullptr) cs_fold_atomic(red, s, dbias, c0, C, true);
}

// ACT: compile-time activation (F5B_ACT_*), or -1 = the run-time `act`
template <int ACT>
__global__ void __launch_bounds__(256) act_bwd8_kernel(const __nv_bfloat16* du, const __nv_bfloat16* __restrict__ h,
                                                       __nv_bfloat16* dh /* may alias du */, float* __restrict__ dbias, int64_t rows, int C,
                                                       int ld, int act, const Drop dr) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  __shared__ __align__(16) float red[8][CS_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * CS_COLS;
  const int c = c0 + lane * 8;
  const int64_t r0 = (int64_t)blockIdx.x * CS_ROWS + warp * CS_WROWS;
  const int64_t r1 = min(rows, r0 + CS_WROWS);
  const int a_ = ACT >= 0 ? ACT : act;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (c < C) {
    for (int64_t rb = r0; rb < r1; rb += CS_BATCH) {
      uint4 dv[CS_BATCH], hv[CS_BATCH];
#pragma unroll
      for (int k = 0; k < CS_BATCH; ++k) {
        if (rb + k < r1) {
          const size_t off = (size_t)(rb + k) * ld + c;
          dv[k] = *reinterpret_cast<const uint4*>(du + off);
          if (h != nullptr) hv[k] = *reinterpret_cast<const uint4*>(h + off);
        }
      }
#pragma unroll
      for (int k = 0; k < CS_BATCH; ++k) {
        if (rb + k < r1) {
          const size_t off = (size_t)(rb + k) * ld + c;
          float d[8];
          bf8_to_f32(dv[k], d);
          if (h != nullptr) {
            float hf[8];
            bf8_to_f32(hv[k], hf);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] *= act_grad(a_, hf[i]);
            if (dr.thr16) {  // the forward's mask (requires ld == C: the element index is the forward's)
              float m0[4], m1[4];
              drop_mult4(dr, off, m0);
              drop_mult4(dr, off + 4, m1);
#pragma unroll
              for (int i = 0; i < 4; ++i) { d[i] *= m0[i]; d[4 + i] *= m1[i]; }
            }
          }
          if (dh != nullptr) {
            const uint4 pk = make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
            *reinterpret_cast<uint4*>(dh + off) = pk;
            bf8_to_f32(pk, d);  // the bias sees what the GEMMs see
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] += d[i];
        }
      }
    }
  }
  if (dbias != nullptr) cs_fold_atomic(red, s, dbias, c0, C, true);
}

__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out, int64_t n8, int act, const Drop dr) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread = 8 elements (16-byte accesses)
  if (i >= n8) return;
  const uint4 v = reinterpret_cast<const uint4*>(h)[i];
  const uint32_t in[4] = {v.x, v.y, v.z, v.w};
  float m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (dr.thr16) {
    float a[4], c[4];
    drop_mult4(dr, (uint64_t)i * 8, a);
    drop_mult4(dr, (uint64_t)i * 8 + 4, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = a[k]; m[4 + k] = c[k]; }
  }
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&in[k]));
    o[k] = pack_bf16(act_eval(act, f.x) * m[2 * k], act_eval(act, f.y) * m[2 * k + 1]);
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// dh[r, :] = bf16(du[r, :] * act'(h[r, :]));  dbias[:] += sum_r dh[r, :]      (act NONE + dh NULL = plain column sum of du)
// thread = 4 adjacent columns (8-byte accesses), 64 rows per CTA
__global__ void __launch_bounds__(CT_THREADS) act_bwd_kernel(const __nv_bfloat16* du, const __nv_bfloat16* __restrict__ h,
                                                             __nv_bfloat16* dh /* may alias du */, float* __restrict__ dbias, int64_t rows,
                                                             int C, int ld, int act, const Drop dr) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  const int64_t r0 = (int64_t)blockIdx.x * CT_ROWS;
  const int64_t r1 = min(rows, r0 + CT_ROWS);
  for (int c = threadIdx.x * 4; c < C; c += 4 * CT_THREADS) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int64_t r = r0; r < r1; ++r) {
      const size_t off = (size_t)r * ld + c;
      const uint2 dv = *reinterpret_cast<const uint2*>(du + off);
      const float2 d01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv.x));
      const float2 d23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv.y));
      float d[4] = {d01.x, d01.y, d23.x, d23.y};
      if (h != nullptr) {
        const uint2 hv = *reinterpret_cast<const uint2*>(h + off);
        const float2 h01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hv.x));
        const float2 h23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hv.y));
        d[0] *= act_grad(act, h01.x); d[1] *= act_grad(act, h01.y); d[2] *= act_grad(act, h23.x); d[3] *= act_grad(act, h23.y);
        if (dr.thr16) {  // the forward's mask (requires ld == C: the element index is the forward's)
          float m[4];
          drop_mult4(dr, off, m);
#pragma unroll
          for (int i = 0; i < 4; ++i) d[i] *= m[i];
        }
      }
      if (dh != nullptr) {
        uint2 pk;
        pk.x = pack_bf16(d[0], d[1]);
        pk.y = pack_bf16(d[2], d[3]);
        *reinterpret_cast<uint2*>(dh + off) = pk;
        const float2 r01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));  // the bias sees what the GEMMs see
        const float2 r23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
        d[0] = r01.x; d[1] = r01.y; d[2] = r23.x; d[3] = r23.y;
      }
      s[0] += d[0]; s[1] += d[1]; s[2] += d[2]; s[3] += d[3];
    }
    if (dbias) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(dbias + c + i, s[i]);
    }
  }
}

// Backward of y = LN(x) * (1 + scale[b]) + shift[b] (no affine; AdaLayerNorm model/modules.py:310-315):
//   dshift[b] += sum_r dy;  dscale[b] += sum_r dy * xhat;  dx (+)= rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * (1 + scale[b])
// One warp per row, 4 rows per warp.  The row lives in registers and ALL of its loads (x, dy and, when accumulating, the old dx)
// are issued before the first reduction: ~10 KB in flight per warp (the r01 kernel streamed the row four times through L1 / L2 to
// stay at 48 registers — 30 % / 58 % hit rates, one dependent stream after the other, 3.5 TB/s).  The column sums live in a per-warp
// shared-memory slice (plain 128-bit load-add-store: shared-memory fp32 atomics compile to CAS loops on sm_100a), folded at the end
// into one fp32 atomic per column per CTA.
template <int VEC>
__global__ void __launch_bounds__(256, 2) ln_mod_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ scale, int64_t mod_bstride, float* __restrict__ dx,
                                                            int accumulate, float* __restrict__ dscale, float* __restrict__ dshift, int n,
                                                            int D, float eps, int affine) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  extern __shared__ float4 ln_acc4[];  // [8 warps][2][D/4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int nvec = D >> 2;
  float4* my_sc = ln_acc4 + (size_t)warp * 2 * nvec;
  float4* my_sh = my_sc + nvec;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    if (idx < nvec) my_sc[idx] = my_sh[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4* scv = scale ? reinterpret_cast<const float4*>(scale + (size_t)b * mod_bstride) : nullptr;
  const float o1 = affine ? 0.f : 1.f;  // affine: `scale` is LayerNorm's weight itself
  const int p0 = blockIdx.x * 32 + warp * 4;
  for (int pos = p0; pos < min(n, p0 + 4); ++pos) {
    const size_t row = (size_t)b * n + pos;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + row * D);
    float4* o = reinterpret_cast<float4*>(dx + row * D);
    float4 v[VEC], pz[VEC];
    uint2 dd[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int idx = lane + j * 32;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      pz[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      dd[j] = make_uint2(0u, 0u);
      if (idx < nvec) {
        v[j] = __ldg(xr + idx);
        dd[j] = __ldg(dr + idx);
        if (accumulate) pz[j] = o[idx];
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (lane + j * 32 < nvec) {
        const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int idx = lane + j * 32;
      if (idx < nvec) {
        const float2 d0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dd[j].x));
        const float2 d1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dd[j].y));
        // xhat replaces x in the register copy of the row
        v[j].x = (v[j].x - mean) * rstd; v[j].y = (v[j].y - mean) * rstd; v[j].z = (v[j].z - mean) * rstd; v[j].w = (v[j].w - mean) * rstd;
        float4 a = my_sh[idx];
        a.x += d0.x; a.y += d0.y; a.z += d1.x; a.w += d1.y;
        my_sh[idx] = a;
        a = my_sc[idx];
        a.x = fmaf(d0.x, v[j].x, a.x); a.y = fmaf(d0.y, v[j].y, a.y); a.z = fmaf(d1.x, v[j].z, a.z); a.w = fmaf(d1.y, v[j].w, a.w);
        my_sc[idx] = a;
        float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
        if (scv != nullptr) {
          const float4 sc = __ldg(scv + idx);
          m = make_float4(o1 + sc.x, o1 + sc.y, o1 + sc.z, o1 + sc.w);
        }
        m.x *= d0.x; m.y *= d0.y; m.z *= d1.x; m.w *= d1.y;  // g = dy * (1 + scale)
        s1 += (m.x + m.y) + (m.z + m.w);
        s2 += (m.x * v[j].x + m.y * v[j].y) + (m.z * v[j].z + m.w * v[j].w);
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int idx = lane + j * 32;
      if (idx < nvec) {
        const float2 d0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dd[j].x));
        const float2 d1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dd[j].y));
        float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
        if (scv != nullptr) {
          const float4 sc = __ldg(scv + idx);
          m = make_float4(o1 + sc.x, o1 + sc.y, o1 + sc.z, o1 + sc.w);
        }
        float4 r;
        r.x = rstd * (d0.x * m.x - s1 - v[j].x * s2) + pz[j].x;
        r.y = rstd * (d0.y * m.y - s1 - v[j].y * s2) + pz[j].y;
        r.z = rstd * (d1.x * m.z - s1 - v[j].z * s2) + pz[j].z;
        r.w = rstd * (d1.y * m.w - s1 - v[j].w * s2) + pz[j].w;
        o[idx] = r;
      }
    }
  }
  __syncthreads();
  const float* acc = reinterpret_cast<const float*>(ln_acc4);
  for (int i = threadIdx.x; i < D; i += 256) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) {
      a += acc[(size_t)w8 * 2 * D + i];
      c += acc[(size_t)w8 * 2 * D + D + i];
    }
    if (dscale) atomicAdd(dscale + (size_t)b * mod_bstride + i, a);
    if (dshift) atomicAdd(dshift + (size_t)b * mod_bstride + i, c);
  }
}

// d loss / d pred of the flow-matching loss (cfm.py:280-283): 2 (pred - flow) / (count * C) on masked rows, bf16 [rows, ld] zero padded
__global__ void mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ flow, const uint8_t* __restrict__ mask,
                                const float* __restrict__ loss2, __nv_bfloat16* __restrict__ out, int64_t rows, int C, int ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const int64_t r = i / ld;
  const int c = (int)(i - r * ld);
  float v = 0.f;
  if (c < C && mask[r]) v = 2.f * (pred[r * C + c] - flow[r * C + c]) / fmaxf(loss2[1], 1.f);
  out[i] = __float2bfloat16(v);
}

// ---- distillation losses (train/distil_reload.py:1066-1093): with m = rand_span_mask [rows], cnt = max(sum m, 1) FRAMES (the
// channel axis is summed, not averaged -- unlike CFM.forward's loss):
//   student = sum m (p - flow)^2 / cnt;  distill = sum m (p - T)^2 / cnt  (or |p - T| for "l1");  spec_l1 = sum m |p - T| / cnt
//   total = (1 - alpha) student + alpha distill + w spec_l1        (T = the teacher's prediction, detached)
__global__ void __launch_bounds__(256) distill_loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ flow,
                                                                   const float* __restrict__ teacher, const uint8_t* __restrict__ mask,
                                                                   float* __restrict__ partial, int rows, int C) {
  __shared__ float red[3][8];
  float s = 0.f, d2 = 0.f, d1 = 0.f, cnt = 0.f;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    if (!mask[r]) continue;  // block-uniform
    if (threadIdx.x == 0) cnt += 1.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float p = pred[(size_t)r * C + c];
      const float a = p - flow[(size_t)r * C + c], b = p - teacher[(size_t)r * C + c];
      s += a * a;
      d2 += b * b;
      d1 += fabsf(b);
    }
  }
  s = warp_sum(s); d2 = warp_sum(d2); d1 = warp_sum(d1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = d2; red[2][threadIdx.x >> 5] = d1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
    partial[blockIdx.x * 4] = a; partial[blockIdx.x * 4 + 1] = b; partial[blockIdx.x * 4 + 2] = c; partial[blockIdx.x * 4 + 3] = cnt;
  }
}
// out5 = (total, student, distill, spec_l1, cnt)
__global__ void distill_loss_final_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ out, int l1, float alpha, float w) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0, n = 0.0;  // fixed summation order: deterministic
    for (int i = 0; i < nblk; ++i) { a += partial[4 * i]; b += partial[4 * i + 1]; c += partial[4 * i + 2]; n += partial[4 * i + 3]; }
    const double cnt = n > 1.0 ? n : 1.0;
    const double student = a / cnt, distill = (l1 ? c : b) / cnt, spec = w > 0.f ? c / cnt : 0.0;
    out[0] = (float)((1.0 - alpha) * student + alpha * distill + w * spec);
    out[1] = (float)student; out[2] = (float)distill; out[3] = (float)spec; out[4] = (float)cnt;
  }
}
// d total / d pred, bf16 [rows, ld] zero padded
__global__ void distill_grad_kernel(const float* __restrict__ pred, const float* __restrict__ flow, const float* __restrict__ teacher,
                                    const uint8_t* __restrict__ mask, const float* __restrict__ out5, __nv_bfloat16* __restrict__ out,
                                    int64_t rows, int C, int ld, int l1, float alpha, float w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const int64_t r = i / ld;
  const int c = (int)(i - r * ld);
  float v = 0.f;
  if (c < C && mask[r]) {
    const float p = pred[r * C + c];
    const float a = p - flow[r * C + c], b = p - teacher[r * C + c];
    const float sg = b > 0.f ? 1.f : (b < 0.f ? -1.f : 0.f);
    v = ((1.f - alpha) * 2.f * a + alpha * (l1 ? sg : 2.f * b) + w * sg) / out5[4];
  }
  out[i] = __float2bfloat16(v);
}

// ---- GRN backward (model/modules.py:225-234) fused with the GELU(erf) backward in front of it (ConvNeXtV2Block :263-265) ----
// t3 = gamma * t2 * nx + beta + t2,  nx[b,c] = gx[b,c] / (mean_c gx[b,:] + 1e-6),  gx[b,c] = ||t2[b,:,c]||_2,  t2 = gelu(p1)
// stats[b][0..2][C]: pass 1 writes (sum t2^2, sum dt3*t2, sum dt3); the per-batch-row kernel turns them into (mult, coef):
//   dt2 = dt3 * mult + t2 * coef,  mult = gamma*nx + 1,  coef = (gamma*S/(m+eps) - A) / gx,  A = mean_c(gamma*S*gx) / (m+eps)^2
// CTA = 128 channels x all rows of one batch item: 16 column lanes (8 channels = 16 bytes each) x 16 row lanes, 4 rows in flight per
// thread (round 2: the 4-byte / 4-row-lane form had 256 CTAs with one dependent load stream per thread and ran at 1.1 TB/s).
__global__ void __launch_bounds__(256) grn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ d3, const __nv_bfloat16* __restrict__ t2,
                                                            float* __restrict__ stats, int n, int C) {
  __shared__ float red[3][16][128 + 4];
  const int cl = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int c = blockIdx.x * 128 + cl * 8;
  const int b = blockIdx.y;
  float q[8], sd[8], r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = sd[i] = r[i] = 0.f;
  if (c < C) {  // (C is a multiple of 8: checked by the launcher)
    const size_t base = (size_t)b * n * C + c;
    for (int r0 = rl; r0 < n; r0 += 64) {
      uint4 tv[4], dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = r0 + 16 * k;
        if (row < n) {
          tv[k] = *reinterpret_cast<const uint4*>(t2 + base + (size_t)row * C);
          dv[k] = *reinterpret_cast<const uint4*>(d3 + base + (size_t)row * C);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (r0 + 16 * k < n) {
          float t[8], d[8];
          bf8_to_f32(tv[k], t);
          bf8_to_f32(dv[k], d);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            q[i] = fmaf(t[i], t[i], q[i]);
            sd[i] = fmaf(d[i], t[i], sd[i]);
            r[i] += d[i];
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][rl][cl * 8 + i] = q[i];
    red[1][rl][cl * 8 + i] = sd[i];
    red[2][rl][cl * 8 + i] = r[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * 128; t += 256) {
    const int k = t / 128, cc = t - k * 128;
    if (blockIdx.x * 128 + cc < C) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) a += red[k][j][cc];
      stats[((size_t)b * 3 + k) * C + blockIdx.x * 128 + cc] = a;
    }
  }
}

__global__ void __launch_bounds__(256) grn_bwd_mid_kernel(float* __restrict__ stats, const float* __restrict__ gamma,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  __shared__ float red[2][8];
  __shared__ float sh[2];
  const int b = blockIdx.x;
  float* Q = stats + (size_t)b * 3 * C;
  float* S = Q + C;
  float* R = S + C;
  float sg = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) sg += sqrtf(Q[c]);
  sg = warp_sum(sg);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = sg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[0][i];
    sh[0] = 1.f / (t / (float)C + 1e-6f);
  }
  __syncthreads();
  const float inv = sh[0];
  float sa = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) sa += gamma[c] * S[c] * sqrtf(Q[c]);
  sa = warp_sum(sa);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = sa;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[1][i];
    sh[1] = t / (float)C * inv * inv;
  }
  __syncthreads();
  const float A = sh[1];
  for (int c = threadIdx.x; c < C; c += 256) {
    const float gx = sqrtf(Q[c]);
    const float nx = gx * inv;
    if (dgamma) atomicAdd(dgamma + c, S[c] * nx);
    if (dbeta) atomicAdd(dbeta + c, R[c]);
    const float coef = gx > 0.f ? (gamma[c] * S[c] * inv - A) / gx : 0.f;
    Q[c] = gamma[c] * nx + 1.f;  // mult
    S[c] = coef;
  }
}

__global__ void __launch_bounds__(CT_THREADS) grn_bwd_apply_kernel(const __nv_bfloat16* d3, const __nv_bfloat16* __restrict__ t2,
                                                                   const __nv_bfloat16* __restrict__ p1, const float* __restrict__ stats,
                                                                   __nv_bfloat16* dp1 /* may alias d3 */, float* __restrict__ dbias, int n,
                                                                   int C) {
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * CT_ROWS, r1 = min(n, r0 + CT_ROWS);
  const float* mult = stats + (size_t)b * 3 * C;
  const float* coef = mult + C;
  for (int c = threadIdx.x * 2; c < C; c += 2 * CT_THREADS) {
    const float m0 = mult[c], m1 = mult[c + 1], k0 = coef[c], k1 = coef[c + 1];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const size_t off = ((size_t)b * n + r) * C + c;
      const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d3 + off));
      const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(t2 + off));
      const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p1 + off));
      const float g0 = (d.x * m0 + t.x * k0) * act_grad(F5B_ACT_GELU_ERF, p.x);
      const float g1 = (d.y * m1 + t.y * k1) * act_grad(F5B_ACT_GELU_ERF, p.y);
      const uint32_t pk = pack_bf16(g0, g1);
      *reinterpret_cast<uint32_t*>(dp1 + off) = pk;
      const float2 rr = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk));
      s0 += rr.x;
      s1 += rr.y;
    }
    if (dbias) {
      atomicAdd(dbias + c, s0);
      atomicAdd(dbias + c + 1, s1);
    }
  }
}

// depth-wise Conv1d(k=7, pad 3) backward: dx[b,p,c] += sum_k w[c,k] dy[b,p-k+3,c];  dw[c,k] += sum dy[b,p,c] x[b,p+k-3,c];  db[c] += sum dy
// thread = 4 adjacent channels (16-byte accesses), 64 rows per CTA walked with register SLIDING WINDOWS of dy and x (7 rows each):
// one new row of each per step instead of 14 reloads through L1 (round 2; the scalar form ran 315 us for 315 MB).
__global__ void __launch_bounds__(128) dwconv7_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                                                          float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int n, int C) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * CT_ROWS, p1 = min(n, p0 + CT_ROWS);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = (blockIdx.z * 128 + threadIdx.x) * 4; c < C; c += 4 * 128 * gridDim.z) {
    float wk[4][7], aw[4][7];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 7; ++k) { wk[j][k] = __ldg(w + (size_t)(c + j) * 7 + k); aw[j][k] = 0.f; }
    float4 ab = zero;
    const size_t base = (size_t)b * n * C + c;
    auto ld = [&](const float* ptr, int row) { return (row >= 0 && row < n) ? *reinterpret_cast<const float4*>(ptr + base + (size_t)row * C) : zero; };
    // windows: gw[i] = dy[p - 3 + i], xw[i] = x[p - 3 + i]
    float4 gw[7], xw[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) { gw[i] = ld(dy, p0 - 3 + i); xw[i] = ld(x, p0 - 3 + i); }
    for (int p = p0; p < p1; ++p) {
      const float4 gn = ld(dy, p + 4), xn = ld(x, p + 4);  // next step's new rows: issued before this step's math
      const float4 g = gw[3];
      ab.x += g.x; ab.y += g.y; ab.z += g.z; ab.w += g.w;
      float4 acc = zero;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        // dw[c,k] += dy[p] * x[p + k - 3]   (rows outside the utterance are zero in the window)
        aw[0][k] = fmaf(g.x, xw[k].x, aw[0][k]); aw[1][k] = fmaf(g.y, xw[k].y, aw[1][k]);
        aw[2][k] = fmaf(g.z, xw[k].z, aw[2][k]); aw[3][k] = fmaf(g.w, xw[k].w, aw[3][k]);
        // dx[p] += w[c,k] * dy[p - k + 3]
        const float4 gr = gw[6 - k];
        acc.x = fmaf(wk[0][k], gr.x, acc.x); acc.y = fmaf(wk[1][k], gr.y, acc.y);
        acc.z = fmaf(wk[2][k], gr.z, acc.z); acc.w = fmaf(wk[3][k], gr.w, acc.w);
      }
      float4* o = reinterpret_cast<float4*>(dx + base + (size_t)p * C);
      float4 cur = *o;
      cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
      *o = cur;
#pragma unroll
      for (int i = 0; i < 6; ++i) { gw[i] = gw[i + 1]; xw[i] = xw[i + 1]; }
      gw[6] = gn;
      xw[6] = xn;
    }
    if (dw) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 7; ++k) atomicAdd(dw + (size_t)(c + j) * 7 + k, aw[j][k]);
    }
    if (db) {
      atomicAdd(db + c, ab.x); atomicAdd(db + c + 1, ab.y); atomicAdd(db + c + 2, ab.z); atomicAdd(db + c + 3, ab.w);
    }
  }
}

// nn.Embedding backward of TextEmbedding (model/backbones/dit.py:49-60): dtable[token(row)] += dh[row]; consecutive rows that hit
// the same token (the filler tail of every utterance) are summed in registers before the atomic
__global__ void __launch_bounds__(256) text_lookup_bwd_kernel(const float* __restrict__ dh, const int64_t* __restrict__ ids, int nt,
                                                              float* __restrict__ dtable, int n, int C, int vocab_rows, int drop_text) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * CT_ROWS, p1 = min(n, p0 + CT_ROWS);
  for (int c = threadIdx.x; c < C; c += 256) {
    long long cur = -1;
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) {
      long long tok = 0;
      if (p < nt && !drop_text) tok = ids[(size_t)b * nt + p] + 1;
      if (tok < 0 || tok >= vocab_rows) __trap();  // the forward lookup reports the offending id; never scatter out of bounds
      if (tok != cur) {
        if (cur >= 0) atomicAdd(dtable + (size_t)cur * C + c, acc);
        cur = tok;
        acc = 0.f;
      }
      acc += dh[((size_t)b * n + p) * C + c];
    }
    if (cur >= 0) atomicAdd(dtable + (size_t)cur * C + c, acc);
  }
}

// x_out = x + gate[b] * z (rows >= lens[b] keep x)  AND  out = LN(x_out) * (1 + scale[b]) + shift[b] in one sweep: the gated
// residual of one branch and the AdaLN of the next share the read of the residual stream (training forward; inference has both
// fused into GEMM epilogues).  One warp per row, the row lives in registers (D <= 1024).
template <int VEC>
__global__ void __launch_bounds__(256) gate_add_ln_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ z,
                                                          const float* __restrict__ gate, int64_t gate_bstride,
                                                          const int32_t* __restrict__ lens, float* __restrict__ x_out,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          int64_t mod_bstride, __nv_bfloat16* __restrict__ out, int n, int D, float eps,
                                                          const Drop dr) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int pos = blockIdx.x * 8 + warp;
  if (pos >= n) return;
  const size_t row = (size_t)b * n + pos;
  const bool live = lens == nullptr || pos < __ldg(lens + b);
  const int nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  const uint2* zr = reinterpret_cast<const uint2*>(z + row * D);
  const float4* gv = gate ? reinterpret_cast<const float4*>(gate + (size_t)b * gate_bstride) : nullptr;
  float4* xo = reinterpret_cast<float4*>(x_out + row * D);
  // ALL loads of the row (x, z, gate) are issued before the first use: the first version loaded, combined and stored one 16-byte chunk
  // after the other (the `live` branch kept ptxas from hoisting the later loads), i.e. eight dependent round trips per row — 103 us at
  // cfg-5's shape for 472 MB (4.6 TB/s)
  float4 v[VEC], gt[VEC];
  uint2 zz[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    gt[j] = make_float4(1.f, 1.f, 1.f, 1.f);
    zz[j] = make_uint2(0u, 0u);
    if (idx < nvec) {
      v[j] = xr[idx];
      if (live) {
        zz[j] = zr[idx];
        if (gv) gt[j] = __ldg(gv + idx);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    if (idx < nvec) {
      if (live) {
        const float2 z01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zz[j].x));
        const float2 z23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zz[j].y));
        float4 g = gt[j];
        if (dr.thr16) {
          float m[4];
          drop_mult4(dr, (uint64_t)row * D + (uint64_t)idx * 4, m);
          g.x *= m[0]; g.y *= m[1]; g.z *= m[2]; g.w *= m[3];
        }
        v[j].x = fmaf(g.x, z01.x, v[j].x); v[j].y = fmaf(g.y, z01.y, v[j].y);
        v[j].z = fmaf(g.z, z23.x, v[j].z); v[j].w = fmaf(g.w, z23.y, v[j].w);
      }
      xo[idx] = v[j];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  // the modulation vectors of the output stage are fetched under the two reductions (z / gate registers are dead by now)
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_bstride);
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_bstride);
  float4 g4[VEC], h4[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    g4[j] = h4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < nvec) {
      g4[j] = __ldg(sc + idx);
      h4[j] = __ldg(sh + idx);
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j)
    if (lane + j * 32 < nvec) {
      const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  uint2* o = reinterpret_cast<uint2*>(out + row * D);
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    if (idx < nvec) {
      const float4 g = g4[j], h = h4[j];
      o[idx] = make_uint2(pack_bf16((v[j].x - mean) * rstd * (1.f + g.x) + h.x, (v[j].y - mean) * rstd * (1.f + g.y) + h.y),
                          pack_bf16((v[j].z - mean) * rstd * (1.f + g.z) + h.z, (v[j].w - mean) * rstd * (1.f + g.w) + h.w));
    }
  }
}

}  // namespace f5b

using namespace f5b;
#define ST(s) static_cast<cudaStream_t>(s)

namespace f5b {
static float g_drop_p = 0.f, g_attn_drop_p = 0.f;
static uint64_t g_drop_seed = 0;
static Drop drop_off() { return Drop{0u, 1.f, 0ull}; }
// site 0 = FeedForward's Dropout, 1 = the Dropout behind to_out (site 2, SDPA's own dropout: attn_drop_for_layer)
Drop drop_for_site(int layer, int site) {
  const float pr = g_drop_p;
  if (!(pr > 0.f)) return drop_off();
  uint32_t thr = (uint32_t)(pr * 65536.0f + 0.5f);
  if (thr > 65535u) thr = 65535u;
  return Drop{thr, 65536.0f / (65536.0f - (float)thr), g_drop_seed * 0xD1342543DE82EF95ull + (uint64_t)(layer * 8 + site + 1) * 0x9E3779B97F4A7C15ull};
}
static Drop drop_for(int layer, int site) { return drop_for_site(layer, site); }
bool train_dropout_on() { return g_drop_p > 0.f; }
// SDPA's dropout (site 2) draws from its own Philox stream (dropout.cuh); same seed, same per-(layer, site) key derivation
AttnDrop make_attn_drop(float p, uint64_t seed, int layer) {
  if (!(p > 0.f)) return AttnDrop{0u, 1.f, 0.f, 0u, 0u};
  int t7 = (int)(p * 128.0f + 0.5f);
  t7 = t7 < 1 ? 1 : (t7 > 127 ? 127 : t7);
  const float scale = 128.0f / (float)(128 - t7);
  const uint64_t key = seed * 0xD1342543DE82EF95ull + (uint64_t)(layer * 8 + 3) * 0x9E3779B97F4A7C15ull;
  return AttnDrop{(uint32_t)(128 - t7) * 0x01010101u, scale, log2f(scale), (uint32_t)key, (uint32_t)(key >> 32)};
}
AttnDrop attn_drop_for_layer(int layer) { return make_attn_drop(g_attn_drop_p, g_drop_seed, layer); }
}  // namespace f5b

extern "C" {

int f5b_gate_add(const float* x, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens, float* out, int B,
                 int n, int C, f5b_stream_t stream) {
  F5B_CHECK(x && z_bf16 && out && B > 0 && n > 0 && C > 0 && (C & 1) == 0, "f5b_gate_add: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 10.0 * B * n * C);
  gate_add_kernel<<<dim3((n + 7) / 8, B), 256, 0, ST(stream)>>>(x, reinterpret_cast<const __nv_bfloat16*>(z_bf16), gate, gate_bstride, lens,
                                                               out, n, C);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

static int gate_add_ln_launch(const float* x, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens,
                              float* x_out, const float* scale, const float* shift, int64_t mod_bstride, void* out_bf16, int B, int n,
                              int D, float eps, const Drop dr, f5b_stream_t stream) {
  F5B_CHECK(x && z_bf16 && x_out && scale && shift && out_bf16 && B > 0 && n > 0, "f5b_gate_add_ln_modulate: bad argument");
  F5B_CHECK(D > 0 && (D & 3) == 0 && D <= 1024 && (gate_bstride & 3) == 0 && (mod_bstride & 3) == 0,
            "f5b_gate_add_ln_modulate: D=%d must be a multiple of 4 and <= 1024, strides multiples of 4", D);
  LaunchScope scope(K_NORM, ST(stream), 0, 12.0 * B * n * D);
  const dim3 grid((n + 7) / 8, B);
  auto* zz = reinterpret_cast<const __nv_bfloat16*>(z_bf16);
  auto* oo = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  const int nvec = D / 4;
  if (nvec <= 64) F5B_CUDA(launch_dep(gate_add_ln_kernel<2>, grid, dim3(256), 0, ST(stream), 1, x, zz, gate, gate_bstride, lens, x_out, scale, shift, mod_bstride, oo, n, D, eps, dr));
  else if (nvec <= 128) F5B_CUDA(launch_dep(gate_add_ln_kernel<4>, grid, dim3(256), 0, ST(stream), 1, x, zz, gate, gate_bstride, lens, x_out, scale, shift, mod_bstride, oo, n, D, eps, dr));
  else F5B_CUDA(launch_dep(gate_add_ln_kernel<8>, grid, dim3(256), 0, ST(stream), 1, x, zz, gate, gate_bstride, lens, x_out, scale, shift, mod_bstride, oo, n, D, eps, dr));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_gate_add_ln_modulate(const float* x, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens,
                             float* x_out, const float* scale, const float* shift, int64_t mod_bstride, void* out_bf16, int B, int n,
                             int D, float eps, f5b_stream_t stream) {
  return gate_add_ln_launch(x, z_bf16, gate, gate_bstride, lens, x_out, scale, shift, mod_bstride, out_bf16, B, n, D, eps, drop_off(), stream);
}

static int gate_bwd_launch(const float* dx, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens, void* dz_bf16,
                           float* dgate, float* dbias, int B, int n, int C, const Drop dr, f5b_stream_t stream) {
  F5B_CHECK(dx && dz_bf16 && B > 0 && n > 0 && C > 0 && (C & 3) == 0 && (gate_bstride & 3) == 0, "f5b_gate_bwd: C and the gate stride must be multiples of 4");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 8.0 * B * n * C);
  if ((C & 7) == 0 && (gate_bstride & 7) == 0 && B <= 65535) {
    F5B_CUDA(launch_dep(gate_bwd8_kernel, dim3((n + CS_ROWS - 1) / CS_ROWS, B, (C + CS_COLS - 1) / CS_COLS), dim3(256), 0, ST(stream), 1, dx,
                        reinterpret_cast<const __nv_bfloat16*>(z_bf16), gate, gate_bstride, lens, reinterpret_cast<__nv_bfloat16*>(dz_bf16),
                        dgate, dbias, n, C, dr));
    F5B_CUDA(cudaGetLastError());
    return 0;
  }
  F5B_CUDA(launch_dep(gate_bwd_kernel, dim3((n + CT_ROWS - 1) / CT_ROWS, B), dim3(CT_THREADS), 0, ST(stream), 1, dx,
                      reinterpret_cast<const __nv_bfloat16*>(z_bf16), gate, gate_bstride, lens, reinterpret_cast<__nv_bfloat16*>(dz_bf16),
                      dgate, dbias, n, C, dr));
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_gate_bwd(const float* dx, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens, void* dz_bf16,
                 float* dgate, float* dbias, int B, int n, int C, f5b_stream_t stream) {
  return gate_bwd_launch(dx, z_bf16, gate, gate_bstride, lens, dz_bf16, dgate, dbias, B, n, C, drop_off(), stream);
}

static int act_fwd_launch(const void* h_bf16, void* out_bf16, int64_t count, int act, const Drop dr, f5b_stream_t stream) {
  F5B_CHECK(h_bf16 && out_bf16 && count > 0 && (count & 7) == 0, "f5b_act_fwd: count must be a multiple of 8");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 4.0 * count);
  F5B_CUDA(launch_dep(act_fwd_kernel, dim3((unsigned)((count / 8 + 255) / 256)), dim3(256), 0, ST(stream), 1,
                      reinterpret_cast<const __nv_bfloat16*>(h_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), count / 8, act, dr));
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_act_fwd(const void* h_bf16, void* out_bf16, int64_t count, int act, f5b_stream_t stream) {
  return act_fwd_launch(h_bf16, out_bf16, count, act, drop_off(), stream);
}

static int act_bwd_launch(const void* du_bf16, const void* h_bf16, void* dh_bf16, float* dbias, int64_t rows, int C, int ld, int act,
                          const Drop dr, f5b_stream_t stream) {
  F5B_CHECK(du_bf16 && rows > 0 && C > 0 && (C & 3) == 0 && (ld & 3) == 0 && ld >= C, "f5b_act_bwd: C and ld must be multiples of 4");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, (h_bf16 ? 2.0 : 0.0) * rows * C + (dh_bf16 ? 2.0 : 0.0) * rows * C + 2.0 * rows * C);
  if ((C & 7) == 0 && (ld & 7) == 0) {
    const dim3 grid((unsigned)((rows + CS_ROWS - 1) / CS_ROWS), (C + CS_COLS - 1) / CS_COLS);
    auto* du = reinterpret_cast<const __nv_bfloat16*>(du_bf16);
    auto* hh = reinterpret_cast<const __nv_bfloat16*>(h_bf16);
    auto* dh = reinterpret_cast<__nv_bfloat16*>(dh_bf16);
    if (act == F5B_ACT_GELU_TANH) F5B_CUDA(launch_dep(act_bwd8_kernel<F5B_ACT_GELU_TANH>, grid, dim3(256), 0, ST(stream), 1, du, hh, dh, dbias, rows, C, ld, act, dr));
    else if (act == F5B_ACT_NONE) F5B_CUDA(launch_dep(act_bwd8_kernel<F5B_ACT_NONE>, grid, dim3(256), 0, ST(stream), 1, du, hh, dh, dbias, rows, C, ld, act, dr));
    else F5B_CUDA(launch_dep(act_bwd8_kernel<-1>, grid, dim3(256), 0, ST(stream), 1, du, hh, dh, dbias, rows, C, ld, act, dr));
    F5B_CUDA(cudaGetLastError());
    return 0;
  }
  F5B_CUDA(launch_dep(act_bwd_kernel, dim3((unsigned)((rows + CT_ROWS - 1) / CT_ROWS)), dim3(CT_THREADS), 0, ST(stream), 1,
                      reinterpret_cast<const __nv_bfloat16*>(du_bf16), reinterpret_cast<const __nv_bfloat16*>(h_bf16),
                      reinterpret_cast<__nv_bfloat16*>(dh_bf16), dbias, rows, C, ld, act, dr));
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_act_bwd(const void* du_bf16, const void* h_bf16, void* dh_bf16, float* dbias, int64_t rows, int C, int ld, int act,
                f5b_stream_t stream) {
  return act_bwd_launch(du_bf16, h_bf16, dh_bf16, dbias, rows, C, ld, act, drop_off(), stream);
}

/* train-mode dropout of the DiT blocks for the f5b_dit_train_* drivers (FeedForward's Dropout after GELU and the Dropout behind
 * to_out; p = 0 switches it off).  The keep decision is a pure function of (seed, block, site, element): call it once per
 * optimizer micro-step with a fresh seed, BEFORE f5b_dit_train_forward, and leave it unchanged until the matching backward. */
int f5b_train_set_dropout(float p, uint64_t seed) {
  F5B_CHECK(p >= 0.f && p < 1.f, "f5b_train_set_dropout: p must be in [0, 1)");
  g_drop_p = p;
  g_drop_seed = seed;
  return 0;
}
/* the third dropout site of a DiT block: F.scaled_dot_product_attention(dropout_p=...) (model/modules.py:490; the fork hard-codes
 * 0.1).  Applied to the attention probabilities after the softmax normalisation, kept values scaled by 1 / (1 - p); same
 * seed as the other two sites, own Philox-2x32 stream with the probability quantised to 1/128 (dropout.cuh: AttnDrop). */
int f5b_train_set_attn_dropout(float p) {
  F5B_CHECK(p >= 0.f && p < 1.f, "f5b_train_set_attn_dropout: p must be in [0, 1)");
  g_attn_drop_p = p;
  return 0;
}
/* the training drivers' versions of the four sweeps: site 0 = FeedForward dropout, site 1 = attention-output dropout */
int f5b_act_fwd_site(const void* h_bf16, void* out_bf16, int64_t count, int act, int layer, int site, f5b_stream_t stream) {
  return act_fwd_launch(h_bf16, out_bf16, count, act, drop_for(layer, site), stream);
}
int f5b_act_bwd_site(const void* du_bf16, const void* h_bf16, void* dh_bf16, float* dbias, int64_t rows, int C, int ld, int act, int layer,
                     int site, f5b_stream_t stream) {
  F5B_CHECK(ld == C || !(g_drop_p > 0.f), "f5b_act_bwd_site: dropout needs a dense [rows, C] tensor");
  return act_bwd_launch(du_bf16, h_bf16, dh_bf16, dbias, rows, C, ld, act, drop_for(layer, site), stream);
}
int f5b_gate_add_ln_modulate_site(const float* x, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens,
                                  float* x_out, const float* scale, const float* shift, int64_t mod_bstride, void* out_bf16, int B,
                                  int n, int D, float eps, int layer, int site, f5b_stream_t stream) {
  return gate_add_ln_launch(x, z_bf16, gate, gate_bstride, lens, x_out, scale, shift, mod_bstride, out_bf16, B, n, D, eps,
                            drop_for(layer, site), stream);
}
int f5b_gate_bwd_site(const float* dx, const void* z_bf16, const float* gate, int64_t gate_bstride, const int32_t* lens, void* dz_bf16,
                      float* dgate, float* dbias, int B, int n, int C, int layer, int site, f5b_stream_t stream) {
  return gate_bwd_launch(dx, z_bf16, gate, gate_bstride, lens, dz_bf16, dgate, dbias, B, n, C, drop_for(layer, site), stream);
}

static int ln_bwd_launch(const void* dy_bf16, const float* x, const float* scale, int64_t mod_bstride, float* dx, int accumulate,
                         float* dscale, float* dshift, int B, int n, int D, float eps, int affine, f5b_stream_t stream) {
  F5B_CHECK(dy_bf16 && x && dx && B > 0 && n > 0, "f5b_ln_modulate_bwd: bad argument");
  F5B_CHECK(D > 0 && (D & 3) == 0 && D <= 1024, "f5b_ln_modulate_bwd: D=%d must be a multiple of 4 and <= 1024", D);
  LaunchScope scope(K_NORM, ST(stream), 0, (accumulate ? 14.0 : 10.0) * B * n * D);
  const dim3 grid((n + 31) / 32, B);
  const size_t sm = 16 * (size_t)D * sizeof(float);  // 8 per-warp slices of (dscale, dshift)
  static bool configured = false;
  if (!configured) {
    F5B_CUDA(cudaFuncSetAttribute(ln_mod_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
    F5B_CUDA(cudaFuncSetAttribute(ln_mod_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
    F5B_CUDA(cudaFuncSetAttribute(ln_mod_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
    configured = true;
  }
  auto* d = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  const int nvec = D / 4;
  if (nvec <= 64) F5B_CUDA(launch_dep(ln_mod_bwd_kernel<2>, grid, dim3(256), sm, ST(stream), 1, d, x, scale, mod_bstride, dx, accumulate, dscale, dshift, n, D, eps, affine));
  else if (nvec <= 128) F5B_CUDA(launch_dep(ln_mod_bwd_kernel<4>, grid, dim3(256), sm, ST(stream), 1, d, x, scale, mod_bstride, dx, accumulate, dscale, dshift, n, D, eps, affine));
  else F5B_CUDA(launch_dep(ln_mod_bwd_kernel<8>, grid, dim3(256), sm, ST(stream), 1, d, x, scale, mod_bstride, dx, accumulate, dscale, dshift, n, D, eps, affine));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_ln_modulate_bwd(const void* dy_bf16, const float* x, const float* scale, int64_t mod_bstride, float* dx, int accumulate,
                        float* dscale, float* dshift, int B, int n, int D, float eps, f5b_stream_t stream) {
  return ln_bwd_launch(dy_bf16, x, scale, mod_bstride, dx, accumulate, dscale, dshift, B, n, D, eps, 0, stream);
}
/* backward of LayerNorm(D, affine): dx = LN-backward(dy * w), dw += sum dy * xhat, db += sum dy */
int f5b_ln_affine_bwd(const void* dy_bf16, const float* x, const float* w, float* dx, int accumulate, float* dw, float* db, int B,
                      int n, int D, float eps, f5b_stream_t stream) {
  F5B_CHECK(w != nullptr, "f5b_ln_affine_bwd: null weight");
  return ln_bwd_launch(dy_bf16, x, w, 0, dx, accumulate, dw, db, B, n, D, eps, 1, stream);
}

int f5b_grn_gelu_bwd(const void* dt3_bf16, const void* t2_bf16, const void* p1_bf16, const float* gamma, void* dp1_bf16, float* dgamma,
                     float* dbeta, float* dbias1, float* stats_ws, int B, int n, int C, f5b_stream_t stream) {
  F5B_CHECK(dt3_bf16 && t2_bf16 && p1_bf16 && gamma && dp1_bf16 && stats_ws && B > 0 && n > 0 && C > 0 && (C & 7) == 0,
            "f5b_grn_gelu_bwd: bad argument (C must be a multiple of 8)");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 12.0 * B * n * C, 3);
  auto* d3 = reinterpret_cast<const __nv_bfloat16*>(dt3_bf16);
  auto* t2 = reinterpret_cast<const __nv_bfloat16*>(t2_bf16);
  grn_bwd_stats_kernel<<<dim3((C + 127) / 128, B), 256, 0, ST(stream)>>>(d3, t2, stats_ws, n, C);
  F5B_CUDA(cudaGetLastError());
  grn_bwd_mid_kernel<<<B, 256, 0, ST(stream)>>>(stats_ws, gamma, dgamma, dbeta, C);
  F5B_CUDA(cudaGetLastError());
  grn_bwd_apply_kernel<<<dim3((n + CT_ROWS - 1) / CT_ROWS, B), CT_THREADS, 0, ST(stream)>>>(
      d3, t2, reinterpret_cast<const __nv_bfloat16*>(p1_bf16), stats_ws, reinterpret_cast<__nv_bfloat16*>(dp1_bf16), dbias1, n, C);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_dwconv7_bwd(const float* dy, const float* x, const float* w, float* dx_accum, float* dw, float* db, int B, int n, int C,
                    f5b_stream_t stream) {
  F5B_CHECK(dy && x && w && dx_accum && B > 0 && n > 0 && C > 0 && (C & 3) == 0, "f5b_dwconv7_bwd: bad argument (C must be a multiple of 4)");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 16.0 * B * n * C);
  dwconv7_bwd_kernel<<<dim3((n + CT_ROWS - 1) / CT_ROWS, B, (C / 4 + 127) / 128), 128, 0, ST(stream)>>>(dy, x, w, dx_accum, dw, db, n, C);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_text_lookup_bwd(const float* dh, const int64_t* ids, int nt, float* dtable, int B, int n, int C, int vocab_rows, int drop_text,
                        f5b_stream_t stream) {
  F5B_CHECK(dh && ids && dtable && B > 0 && n > 0 && C > 0 && nt > 0 && vocab_rows > 0, "f5b_text_lookup_bwd: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 4.0 * B * n * C);
  text_lookup_bwd_kernel<<<dim3((n + CT_ROWS - 1) / CT_ROWS, B), 256, 0, ST(stream)>>>(dh, ids, nt, dtable, n, C, vocab_rows, drop_text);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_mse_grad(const float* pred, const float* flow, const uint8_t* mask, const float* loss2, void* out_bf16, int64_t rows, int C, int ld,
                 f5b_stream_t stream) {
  F5B_CHECK(pred && flow && mask && loss2 && out_bf16 && rows > 0 && C > 0 && ld >= C, "f5b_mse_grad: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 8.0 * rows * C + 2.0 * rows * ld);
  mse_grad_kernel<<<(unsigned)((rows * ld + 255) / 256), 256, 0, ST(stream)>>>(pred, flow, mask, loss2, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                                               rows, C, ld);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_distill_loss(const float* pred, const float* flow, const float* teacher, const uint8_t* mask, float* ws /*[4*1024]*/, float* out5,
                     int rows, int C, int l1, float alpha, float spec_l1_weight, f5b_stream_t stream) {
  F5B_CHECK(pred && flow && teacher && mask && ws && out5 && rows > 0 && C > 0, "f5b_distill_loss: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 12.0 * rows * C);
  const int nblk = rows < 1024 ? rows : 1024;
  distill_loss_partial_kernel<<<nblk, 256, 0, ST(stream)>>>(pred, flow, teacher, mask, ws, rows, C);
  F5B_CUDA(cudaGetLastError());
  distill_loss_final_kernel<<<1, 32, 0, ST(stream)>>>(ws, nblk, out5, l1, alpha, spec_l1_weight);
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_distill_grad(const float* pred, const float* flow, const float* teacher, const uint8_t* mask, const float* out5, void* out_bf16,
                     int64_t rows, int C, int ld, int l1, float alpha, float spec_l1_weight, f5b_stream_t stream) {
  F5B_CHECK(pred && flow && teacher && mask && out5 && out_bf16 && rows > 0 && C > 0 && ld >= C, "f5b_distill_grad: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 12.0 * rows * C + 2.0 * rows * ld);
  distill_grad_kernel<<<(unsigned)((rows * ld + 255) / 256), 256, 0, ST(stream)>>>(
      pred, flow, teacher, mask, out5, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, C, ld, l1, alpha, spec_l1_weight);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
