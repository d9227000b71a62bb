// Memory-bound kernels of the F5-TTS path: LayerNorm + AdaLN modulation, depth-wise conv + LayerNorm, GRN, text-token
// lookup, timestep sinusoid, CFG + Euler update, small packing helpers.  All are coalesced / vectorised, one warp per
// token row where a row reduction is needed; HBM traffic is the roofline (algorithmic bytes per row in DESIGN.md).
// Reference citations are relative to /root/reference/src/f5_tts/.
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

// Activation outputs are bf16 (kind::f16 operands of the next GEMM) or, in the tf32 operand mode, fp32 words rounded to tf32.
__device__ __forceinline__ void store4(__nv_bfloat16* o, int idx, float a, float b, float c, float d) {
  reinterpret_cast<uint2*>(o)[idx] = make_uint2(pack_bf16(a, b), pack_bf16(c, d));
}
__device__ __forceinline__ void store4(float* o, int idx, float a, float b, float c, float d) {
  reinterpret_cast<float4*>(o)[idx] = make_float4(tf32_rn(a), tf32_rn(b), tf32_rn(c), tf32_rn(d));
}
__device__ __forceinline__ float2 load2(const __nv_bfloat16* p) {
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
  return make_float2(__low2float(v), __high2float(v));
}
__device__ __forceinline__ float2 load2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(tf32_rn(a), tf32_rn(b)); }
__device__ __forceinline__ void store1(__nv_bfloat16* p, float a) { *p = __float2bfloat16(a); }
__device__ __forceinline__ void store1(float* p, float a) { *p = tf32_rn(a); }

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm (no affine, eps) * (1 + scale[b]) + shift[b] -> bf16      AdaLayerNorm.forward model/modules.py:310-315,
// DiTBlock ff norm :637, AdaLayerNorm_Final :331-336.  One warp per row, row cached in registers (D <= 2048).
// ---------------------------------------------------------------------------------------------------------------
template <int VEC, class OutT>  // float4 per lane
__global__ void __launch_bounds__(256) ln_modulate_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int64_t mod_bstride,
                                                          int batch_mod, OutT* __restrict__ out, int rows,
                                                          int rows_per_batch, int D, float eps) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  const int nvec = D >> 2;
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    v[j] = idx < nvec ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    if (idx < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  int b = row / rows_per_batch;
  if (batch_mod > 0) b %= batch_mod;
  const float4* sc = scale ? reinterpret_cast<const float4*>(scale + (size_t)b * mod_bstride) : nullptr;
  const float4* sh = shift ? reinterpret_cast<const float4*>(shift + (size_t)b * mod_bstride) : nullptr;
  OutT* o = out + (size_t)row * D;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    if (idx < nvec) {
      float4 g = sc ? __ldg(sc + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 h = sh ? __ldg(sh + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float a = (v[j].x - mean) * rstd * (1.f + g.x) + h.x;
      const float bb = (v[j].y - mean) * rstd * (1.f + g.y) + h.y;
      const float c = (v[j].z - mean) * rstd * (1.f + g.z) + h.z;
      const float d = (v[j].w - mean) * rstd * (1.f + g.w) + h.w;
      store4(o, idx, a, bb, c, d);
    }
  }
}

template <class OutT>
static int ln_modulate_t(const float* x, const float* scale, const float* shift, int64_t mod_bstride, int batch_mod, OutT* o, int rows,
                         int rows_per_batch, int D, float eps, cudaStream_t s) {
  const int grid = (rows + 7) / 8;
  const int nvec = D / 4;
  if (nvec <= 32 * 2) F5B_CUDA(launch_dep(ln_modulate_kernel<2, OutT>, dim3(grid), dim3(256), 0, s, 1, x, scale, shift, mod_bstride, batch_mod, o, rows, rows_per_batch, D, eps));
  else if (nvec <= 32 * 4) F5B_CUDA(launch_dep(ln_modulate_kernel<4, OutT>, dim3(grid), dim3(256), 0, s, 1, x, scale, shift, mod_bstride, batch_mod, o, rows, rows_per_batch, D, eps));
  else if (nvec <= 32 * 8) F5B_CUDA(launch_dep(ln_modulate_kernel<8, OutT>, dim3(grid), dim3(256), 0, s, 1, x, scale, shift, mod_bstride, batch_mod, o, rows, rows_per_batch, D, eps));
  else F5B_CUDA(launch_dep(ln_modulate_kernel<16, OutT>, dim3(grid), dim3(256), 0, s, 1, x, scale, shift, mod_bstride, batch_mod, o, rows, rows_per_batch, D, eps));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int ln_modulate(const float* x, const float* scale, const float* shift, int64_t mod_bstride, int batch_mod, void* out,
                int rows, int rows_per_batch, int D, float eps, cudaStream_t s, bool tf32) {
  F5B_CHECK(rows > 0 && D > 0 && (D & 3) == 0 && D <= 2048, "f5b_ln_modulate: D=%d must be a multiple of 4 and <= 2048", D);
  F5B_CHECK((mod_bstride & 3) == 0, "f5b_ln_modulate: modulation stride must be a multiple of 4");
  F5B_CHECK(rows_per_batch > 0, "f5b_ln_modulate: rows_per_batch");
  LaunchScope scope(K_NORM, s, 0, (tf32 ? 8.0 : 6.0) * rows * D);
  if (tf32) return ln_modulate_t(x, scale, shift, mod_bstride, batch_mod, reinterpret_cast<float*>(out), rows, rows_per_batch, D, eps, s);
  return ln_modulate_t(x, scale, shift, mod_bstride, batch_mod, reinterpret_cast<__nv_bfloat16*>(out), rows, rows_per_batch, D, eps, s);
}

// ---------------------------------------------------------------------------------------------------------------
// depth-wise Conv1d(k=7, pad 3) + bias + LayerNorm(C, affine, eps) -> bf16    ConvNeXtV2Block model/modules.py:259-262
// (and the identical Vocos ConvNeXtBlock front).  One warp per output row; the 7 neighbour rows come from L1/L2.
// ---------------------------------------------------------------------------------------------------------------
template <int VEC, class OutT>
__global__ void __launch_bounds__(256) dwconv7_ln_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                         const float* __restrict__ ln_b, OutT* __restrict__ out,
                                                         int B, int n, int C, float eps, float* __restrict__ y_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B * n) return;
  const int b = row / n, pos = row - b * n;
  const int nvec = C >> 2;
  float4 acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    acc[j] = idx < nvec ? __ldg(reinterpret_cast<const float4*>(bias) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int p = pos + k - 3;
    if (p < 0 || p >= n) continue;  // warp-uniform
    const float4* xr = reinterpret_cast<const float4*>(x + ((size_t)b * n + p) * C);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int idx = lane + j * 32;
      if (idx < nvec) {
        const float4 xv = xr[idx];
        const float* wc = w + (size_t)idx * 4 * 7 + k;
        acc[j].x += xv.x * __ldg(wc);
        acc[j].y += xv.y * __ldg(wc + 7);
        acc[j].z += xv.z * __ldg(wc + 14);
        acc[j].w += xv.w * __ldg(wc + 21);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j)
    if (lane + j * 32 < nvec) {
      s += acc[j].x + acc[j].y + acc[j].z + acc[j].w;
      if (y_out != nullptr) reinterpret_cast<float4*>(y_out + (size_t)row * C)[lane + j * 32] = acc[j];  // training: keep the conv output
    }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j)
    if (lane + j * 32 < nvec) {
      const float a = acc[j].x - mean, bb = acc[j].y - mean, c = acc[j].z - mean, d = acc[j].w - mean;
      q += a * a + bb * bb + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  OutT* o = out + (size_t)row * C;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    if (idx < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(ln_w) + idx);
      const float4 h = __ldg(reinterpret_cast<const float4*>(ln_b) + idx);
      store4(o, idx, (acc[j].x - mean) * rstd * g.x + h.x, (acc[j].y - mean) * rstd * g.y + h.y,
             (acc[j].z - mean) * rstd * g.z + h.z, (acc[j].w - mean) * rstd * g.w + h.w);
    }
  }
}

template <class OutT>
static int dwconv7_ln_t(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b, OutT* o, int B, int n, int C,
                        float eps, cudaStream_t s, float* y_out) {
  const int grid = (B * n + 7) / 8;
  const int nvec = C / 4;
  if (nvec <= 32) dwconv7_ln_kernel<1, OutT><<<grid, 256, 0, s>>>(x, w, b, ln_w, ln_b, o, B, n, C, eps, y_out);
  else if (nvec <= 64) dwconv7_ln_kernel<2, OutT><<<grid, 256, 0, s>>>(x, w, b, ln_w, ln_b, o, B, n, C, eps, y_out);
  else if (nvec <= 128) dwconv7_ln_kernel<4, OutT><<<grid, 256, 0, s>>>(x, w, b, ln_w, ln_b, o, B, n, C, eps, y_out);
  else dwconv7_ln_kernel<8, OutT><<<grid, 256, 0, s>>>(x, w, b, ln_w, ln_b, o, B, n, C, eps, y_out);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int dwconv7_ln(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b, void* out, int B, int n,
               int C, float eps, cudaStream_t s, float* y_out, bool tf32) {
  F5B_CHECK(B > 0 && n > 0 && C > 0 && (C & 3) == 0 && C <= 1024, "f5b_dwconv7_ln: C=%d must be a multiple of 4 and <= 1024", C);
  LaunchScope scope(K_NORM, s, 0, (tf32 ? 8.0 : 6.0) * B * n * C);
  if (tf32) return dwconv7_ln_t(x, w, b, ln_w, ln_b, reinterpret_cast<float*>(out), B, n, C, eps, s, y_out);
  return dwconv7_ln_t(x, w, b, ln_w, ln_b, reinterpret_cast<__nv_bfloat16*>(out), B, n, C, eps, s, y_out);
}

// ---------------------------------------------------------------------------------------------------------------
// GRN  model/modules.py:225-234: Gx = ||h||_2 over the SEQUENCE axis, Nx = Gx / (mean_c Gx + 1e-6),
// out = gamma * (h * Nx) + beta + h.   Pass 1: deterministic column norms; pass 2: apply.
// ---------------------------------------------------------------------------------------------------------------
// CTA = 128 channels x all rows of one batch item: 16 column lanes x 16 row lanes, 8 channels per thread, 4 rows in flight
// (round 2: the 2-channel / 4-row-lane form ran one dependent load stream per thread, 0.7 TB/s).
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <class T>
__global__ void __launch_bounds__(256) grn_colnorm_kernel(const T* __restrict__ h, float* __restrict__ gx, int n, int C) {
  __shared__ float red[16][128 + 4];
  const int cl = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int c = blockIdx.x * 128 + cl * 8;
  const int b = blockIdx.y;
  float q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = 0.f;
  if (c < C) {
    const T* base = h + (size_t)b * n * C + c;
    for (int r0 = rl; r0 < n; r0 += 64) {
      float v[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = r0 + 16 * k;
        if (row < n) load8(base + (size_t)row * C, v[k]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[k][i] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = fmaf(v[k][i], v[k][i], q[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][cl * 8 + i] = q[i];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < C) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) t += red[j][threadIdx.x];
      gx[(size_t)b * C + cc] = sqrtf(t);
    }
  }
}

template <class T>
__global__ void __launch_bounds__(256) grn_apply_kernel(const T* __restrict__ h, const float* __restrict__ gx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        T* __restrict__ out, int n, int C, int rows_per_block) {
  __shared__ float red[8];
  __shared__ float s_mean;
  const int b = blockIdx.y;
  const float* g = gx + (size_t)b * C;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += g[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    s_mean = t / (float)C;
  }
  __syncthreads();
  const float inv = 1.f / (s_mean + 1e-6f);
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  const int pairs = C >> 1;
  for (int r = r0; r < r1; ++r) {
    const T* hr = h + ((size_t)b * n + r) * C;
    T* o = out + ((size_t)b * n + r) * C;
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
      const float2 v = load2(hr + 2 * p);
      const float a = v.x, d = v.y;
      const int c = p * 2;
      const float ya = __ldg(gamma + c) * (a * (g[c] * inv)) + __ldg(beta + c) + a;
      const float yd = __ldg(gamma + c + 1) * (d * (g[c + 1] * inv)) + __ldg(beta + c + 1) + d;
      store2(o + 2 * p, ya, yd);
    }
  }
}

template <class T>
static int grn_t(const T* h, const float* gamma, const float* beta, T* out, float* ws, int B, int n, int C, cudaStream_t s) {
  grn_colnorm_kernel<T><<<dim3((C + 127) / 128, B), 256, 0, s>>>(h, ws, n, C);
  F5B_CUDA(cudaGetLastError());
  const int rpb = 8;
  grn_apply_kernel<T><<<dim3((n + rpb - 1) / rpb, B), 256, 0, s>>>(h, ws, gamma, beta, out, n, C, rpb);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int grn(const void* h, const float* gamma, const float* beta, void* out, float* ws, int B, int n, int C, cudaStream_t s, bool tf32) {
  F5B_CHECK(B > 0 && n > 0 && C > 0 && (C & 7) == 0, "f5b_grn: C=%d must be a multiple of 8", C);
  LaunchScope scope(K_ELEMENTWISE, s, 0, (tf32 ? 12.0 : 6.0) * B * n * C, 2);
  if (tf32) return grn_t(reinterpret_cast<const float*>(h), gamma, beta, reinterpret_cast<float*>(out), ws, B, n, C, s);
  return grn_t(reinterpret_cast<const __nv_bfloat16*>(h), gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), ws, B, n, C, s);
}

// ---------------------------------------------------------------------------------------------------------------
// TextEmbedding front, model/backbones/dit.py:49-72
// ---------------------------------------------------------------------------------------------------------------
__global__ void text_lookup_kernel(const int64_t* __restrict__ ids, int nt, const float* __restrict__ table,
                                   const float* __restrict__ pos, float* __restrict__ out, uint8_t* __restrict__ mask_out,
                                   int n, int C, int vocab_rows, int drop_text, int add_pos) {
  const int row = blockIdx.x;  // b*n + p
  const int b = row / n, p = row - b * n;
  long long tok = 0;
  if (p < nt) tok = ids[(size_t)b * nt + p] + 1;  // -1 padding -> filler 0
  if (tok < 0 || tok >= vocab_rows) {  // nn.Embedding asserts on the device; never read the table out of bounds
    if (threadIdx.x == 0) printf("f5b_text_lookup: token id %lld (row %d, position %d) outside the embedding table [0, %d)\n", tok - 1, b, p, vocab_rows - 1);
    __trap();
  }
  if (mask_out != nullptr && threadIdx.x == 0) mask_out[row] = (tok == 0) ? 1 : 0;
  if (drop_text) tok = 0;
  const float* tr = table + (size_t)tok * C;
  const float* pr = pos + (size_t)min(p, 4095) * C;  // get_pos_embed_indices clamps to max_pos-1 (modules.py:210-219)
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)row * C + c] = tr[c] + (add_pos ? pr[c] : 0.f);
}

__global__ void mask_rows_kernel(float* __restrict__ x, const uint8_t* __restrict__ mask, int C) {
  const int row = blockIdx.x;
  if (!mask[row]) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) x[(size_t)row * C + c] = 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// SinusPositionEmbedding(256), model/modules.py:149-161: emb = 1000 * t * exp(-ln(1e4)/(128-1) * k); cat(sin, cos)
// ---------------------------------------------------------------------------------------------------------------
template <class OutT>
__global__ void time_sinus_kernel(const float* __restrict__ t, OutT* __restrict__ out, int M) {
  const int m = blockIdx.x;
  const int k = threadIdx.x;  // 0..127
  const float f = expf((float)k * -(9.210340371976184f / 127.0f));
  const float a = 1000.0f * t[m] * f;
  store1(out + (size_t)m * 256 + k, sinf(a));
  store1(out + (size_t)m * 256 + 128 + k, cosf(a));
}

__global__ void silu_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = x[i];
    out[i] = __float2bfloat16(v / (1.f + __expf(-v)));
  }
}

template <class OutT>
__global__ void pack_bf16_kernel(const float* __restrict__ x, int ld_in, OutT* __restrict__ out, int ld_out, int rows, int cols,
                                 int width) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * width) return;
  const int r = (int)(i / width), c = (int)(i - (int64_t)r * width);
  store1(out + (size_t)r * ld_out + c, c < cols ? x[(size_t)r * ld_in + c] : 0.f);
}

// LayerNorm(D, affine, eps) with f32 and/or bf16 output (Vocos backbone.norm / final_layer_norm); one warp per row
template <int VEC>
__global__ void __launch_bounds__(256) ln_affine_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out_f32,
                                                        __nv_bfloat16* __restrict__ out_bf16, int rows, int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  const int nvec = D >> 2;
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    v[j] = idx < nvec ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j)
    if (lane + j * 32 < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int idx = lane + j * 32;
    if (idx < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(w) + idx);
      const float4 h = __ldg(reinterpret_cast<const float4*>(bias) + idx);
      float4 y;
      y.x = (v[j].x - mean) * rstd * g.x + h.x;
      y.y = (v[j].y - mean) * rstd * g.y + h.y;
      y.z = (v[j].z - mean) * rstd * g.z + h.z;
      y.w = (v[j].w - mean) * rstd * g.w + h.w;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)row * D)[idx] = y;
      if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + (size_t)row * D)[idx] = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
    }
  }
}

int ln_affine(const float* x, const float* w, const float* b, float* out_f32, void* out_bf16, int rows, int D, float eps,
              cudaStream_t s) {
  F5B_CHECK(rows > 0 && D > 0 && (D & 3) == 0 && D <= 1024, "f5b_ln_affine: D=%d must be a multiple of 4 and <= 1024", D);
  const int grid = (rows + 7) / 8;
  LaunchScope scope(K_NORM, s, 0, (4.0 + (out_f32 ? 4.0 : 0.0) + (out_bf16 ? 2.0 : 0.0)) * rows * D);
  auto* o = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  const int nvec = D / 4;
  if (nvec <= 32) ln_affine_kernel<1><<<grid, 256, 0, s>>>(x, w, b, out_f32, o, rows, D, eps);
  else if (nvec <= 64) ln_affine_kernel<2><<<grid, 256, 0, s>>>(x, w, b, out_f32, o, rows, D, eps);
  else if (nvec <= 128) ln_affine_kernel<4><<<grid, 256, 0, s>>>(x, w, b, out_f32, o, rows, D, eps);
  else ln_affine_kernel<8><<<grid, 256, 0, s>>>(x, w, b, out_f32, o, rows, D, eps);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// CFG combine + Euler step: fn closure model/cfm.py:159-173 (pred + (pred - null) * cfg) and torchdiffeq's fixed-grid
// euler update y += dt * v; also refreshes the zero-padded bf16 copy of the state that feeds the next input GEMM.
// ---------------------------------------------------------------------------------------------------------------
__global__ void cfg_euler_kernel(float* __restrict__ y, const float* __restrict__ pc, const float* __restrict__ pu, float cfg,
                                 float dt, const float* __restrict__ dev_params, __nv_bfloat16* __restrict__ ybf, int ld_bf,
                                 float* __restrict__ vel, int rows, int C) {
  griddep_wait();  // PDL (common.cuh)
  griddep_launch_dependents();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * ld_bf) return;
  if (dev_params != nullptr) {  // (cfg, dt) live in device memory so a captured CUDA graph can be replayed for every ODE step
    cfg = __ldg(dev_params);
    dt = __ldg(dev_params + 1);
  }
  const int r = (int)(i / ld_bf), c = (int)(i - (int64_t)r * ld_bf);
  float yn = 0.f;
  if (c < C) {
    const size_t j = (size_t)r * C + c;
    const float p = pc[j];
    const float v = pu ? p + (p - pu[j]) * cfg : p;
    if (vel) vel[j] = v;
    yn = y[j] + dt * v;
    y[j] = yn;
  }
  if (ybf) ybf[i] = __float2bfloat16(yn);
}

// x_transformers RotaryEmbedding.forward_from_seq_len (call site model/backbones/dit.py:215), dim_head 64
__global__ void rope_table_kernel(float2* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 32) return;
  const int pos = i >> 5, j = i & 31;
  const float inv = 1.0f / powf(10000.0f, (float)(2 * j) / 64.0f);
  const float a = (float)pos * inv;
  out[i] = make_float2(cosf(a), sinf(a));
}

// Vocos embed Conv1d(n_mels -> C, k=7, pad 3) as a GEMM: im2col row = [k=0..6][c], zero outside the utterance
__global__ void im2col7_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ out, int T, int n_mels, int ld) {
  const int row = blockIdx.x;  // b*T + t
  const int b = row / T, t = row - b * T;
  for (int j = threadIdx.x; j < ld; j += blockDim.x) {
    float v = 0.f;
    if (j < 7 * n_mels) {
      const int k = j / n_mels, c = j - k * n_mels;
      const int p = t + k - 3;
      if (p >= 0 && p < T) v = mel[((size_t)b * T + p) * n_mels + c];
    }
    out[(size_t)row * ld + j] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Flow-matching training inputs and loss (CFM.forward, model/cfm.py:255-283):
//   phi = (1 - t) x0 + t x1,  flow = x1 - x0,  cond = where(span_mask, 0, x1);   loss = mean over masked rows of (pred - flow)^2
// ---------------------------------------------------------------------------------------------------------------
__global__ void fm_prepare_kernel(const float* __restrict__ x1, const float* __restrict__ x0, const float* __restrict__ time,
                                  const uint8_t* __restrict__ span, float* __restrict__ phi, float* __restrict__ flow,
                                  float* __restrict__ cond, int n, int C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t row = i / C;
  const int b = (int)(row / n);
  const float t = time[b];
  const float a = x1[i], z = x0[i];
  phi[i] = (1.f - t) * z + t * a;
  flow[i] = a - z;
  cond[i] = span[row] ? 0.f : a;
}

__global__ void __launch_bounds__(256) masked_mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ flow,
                                                                 const uint8_t* __restrict__ mask, float* __restrict__ partial,
                                                                 int rows, int C) {
  __shared__ float red[2][8];
  float s = 0.f, cnt = 0.f;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    if (!mask[r]) continue;  // block-uniform
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float d = pred[(size_t)r * C + c] - flow[(size_t)r * C + c];
      s += d * d;
      cnt += 1.f;
    }
  }
  s = warp_sum(s);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    partial[blockIdx.x * 2] = a;
    partial[blockIdx.x * 2 + 1] = b;
  }
}
__global__ void masked_mse_final_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;  // fixed summation order: deterministic
    for (int i = 0; i < nblk; ++i) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    out[0] = (float)(a / (b > 0.0 ? b : 1.0));
    out[1] = (float)b;
  }
}

// One fold step of the chunk cross-fade (infer/f5tts_wrapper.py:556-572, infer/utils_infer.py:519-543): the last cfs samples of the
// accumulated wave are blended with the first cfs of the next chunk, fade_out = linspace(1, 0, cfs), fade_in = linspace(0, 1, cfs)
// evaluated in fp64 as numpy does (start + t * step, last point exact), and the rest of the chunk is appended.
__global__ void crossfade_append_kernel(float* __restrict__ acc, int64_t acc_len, const float* __restrict__ next, int64_t next_len, int cfs) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= next_len) return;
  const int64_t o = acc_len - cfs + t;
  if (t < cfs) {
    double fo, fi;
    if (cfs == 1) { fo = 1.0; fi = 0.0; }
    else if (t == cfs - 1) { fo = 0.0; fi = 1.0; }
    else {
      // numpy: arange(num) * step + start, each operation rounded separately (no FMA contraction)
      fo = __dadd_rn(__dmul_rn((double)t, -1.0 / (double)(cfs - 1)), 1.0);
      fi = __dmul_rn((double)t, 1.0 / (double)(cfs - 1));
    }
    acc[o] = (float)__dadd_rn(__dmul_rn((double)acc[o], fo), __dmul_rn((double)next[t], fi));
  } else {
    acc[o] = next[t];
  }
}
// np.int16(x * 32767) (socket_server.py:54, finetune_gradio.py:704): fp32 product, truncation toward zero; saturated instead of
// wrapping for |x| > 1
__global__ void pcm16_kernel(const float* __restrict__ x, int16_t* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = __float2int_rz(x[i] * 32767.f);
  out[i] = (int16_t)max(-32768, min(32767, v));
}

}  // namespace f5b

using namespace f5b;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int f5b_ln_modulate(const float* x, const float* scale, const float* shift, int64_t mod_bstride, int batch_mod, void* out,
                    int rows, int rows_per_batch, int D, float eps, f5b_stream_t stream) {
  return ln_modulate(x, scale, shift, mod_bstride, batch_mod, out, rows, rows_per_batch, D, eps, ST(stream));
}

int f5b_ln_modulate_tf32(const float* x, const float* scale, const float* shift, int64_t mod_bstride, int batch_mod, float* out,
                         int rows, int rows_per_batch, int D, float eps, f5b_stream_t stream) {
  return ln_modulate(x, scale, shift, mod_bstride, batch_mod, out, rows, rows_per_batch, D, eps, ST(stream), true);
}

int f5b_dwconv7_ln(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b, void* out, int B,
                   int n, int C, float eps, f5b_stream_t stream) {
  return dwconv7_ln(x, w, b, ln_w, ln_b, out, B, n, C, eps, ST(stream));
}

int f5b_grn(const void* h, const float* gamma, const float* beta, void* out, float* ws, int B, int n, int C,
            f5b_stream_t stream) {
  return grn(h, gamma, beta, out, ws, B, n, C, ST(stream));
}

int f5b_text_lookup(const int64_t* ids, int nt, const float* table, const float* pos, float* out, uint8_t* mask_out, int B,
                    int n, int C, int vocab_rows, int drop_text, int add_pos, f5b_stream_t stream) {
  F5B_CHECK(B > 0 && n > 0 && C > 0 && nt >= 0 && vocab_rows > 0, "f5b_text_lookup: bad shape");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 12.0 * B * n * C);
  text_lookup_kernel<<<B * n, 128, 0, ST(stream)>>>(ids, nt, table, pos, out, mask_out, n, C, vocab_rows, drop_text, add_pos);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_mask_rows_f32(float* x, const uint8_t* mask, int rows, int C, f5b_stream_t stream) {
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 1.0 * rows);
  mask_rows_kernel<<<rows, 128, 0, ST(stream)>>>(x, mask, C);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_time_sinus(const float* t, void* out, int M, f5b_stream_t stream) {
  F5B_CHECK(M > 0, "f5b_time_sinus: M");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 516.0 * M);
  time_sinus_kernel<<<M, 128, 0, ST(stream)>>>(t, reinterpret_cast<__nv_bfloat16*>(out), M);
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_time_sinus_tf32(const float* t, float* out, int M, f5b_stream_t stream) {
  F5B_CHECK(M > 0, "f5b_time_sinus_tf32: M");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 1028.0 * M);
  time_sinus_kernel<<<M, 128, 0, ST(stream)>>>(t, out, M);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_silu_bf16(const float* x, void* out, int64_t n, f5b_stream_t stream) {
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 6.0 * n);
  silu_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(out), n);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_ln_affine(const float* x, const float* w, const float* b, float* out_f32, void* out_bf16, int rows, int D, float eps,
                  f5b_stream_t stream) {
  return ln_affine(x, w, b, out_f32, out_bf16, rows, D, eps, ST(stream));
}

int f5b_pack_bf16(const float* x, int ld_in, void* out, int ld_out, int rows, int cols, int width, f5b_stream_t stream) {
  F5B_CHECK(rows > 0 && cols >= 0 && cols <= width && width <= ld_out && (x != nullptr || cols == 0), "f5b_pack_bf16: bad shape");
  const int64_t tot = (int64_t)rows * width;
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 4.0 * rows * cols + 2.0 * tot);
  pack_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ST(stream)>>>(x, ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out,
                                                                          rows, cols, width);
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_pack_tf32(const float* x, int ld_in, float* out, int ld_out, int rows, int cols, int width, f5b_stream_t stream) {
  F5B_CHECK(rows > 0 && cols >= 0 && cols <= width && width <= ld_out && (x != nullptr || cols == 0), "f5b_pack_tf32: bad shape");
  const int64_t tot = (int64_t)rows * width;
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 4.0 * rows * cols + 4.0 * tot);
  pack_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ST(stream)>>>(x, ld_in, out, ld_out, rows, cols, width);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_crossfade_append(float* acc, int64_t acc_len, const float* next, int64_t next_len, int cfs, f5b_stream_t stream) {
  F5B_CHECK(acc && next && acc_len >= 0 && next_len > 0 && cfs >= 0 && cfs <= acc_len && cfs <= next_len, "f5b_crossfade_append: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 8.0 * next_len);
  crossfade_append_kernel<<<(unsigned)((next_len + 255) / 256), 256, 0, ST(stream)>>>(acc, acc_len, next, next_len, cfs);
  F5B_CUDA(cudaGetLastError());
  return 0;
}
int f5b_pcm16(const float* x, int16_t* out, int64_t n, f5b_stream_t stream) {
  F5B_CHECK(x && out && n > 0, "f5b_pcm16: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 6.0 * n);
  pcm16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, out, n);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_cfg_euler(float* y, const float* pc, const float* pu, float cfg, float dt, void* y_bf16, int ld_bf, float* vel_out,
                  int rows, int C, f5b_stream_t stream) {
  F5B_CHECK(rows > 0 && C > 0 && ld_bf >= C, "f5b_cfg_euler: bad shape");
  const int64_t tot = (int64_t)rows * ld_bf;
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, (pu ? 16.0 : 12.0) * rows * C + (y_bf16 ? 2.0 * tot : 0.0));
  F5B_CUDA(launch_dep(cfg_euler_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, ST(stream), 1, y, pc, pu, cfg, dt,
                      (const float*)nullptr, reinterpret_cast<__nv_bfloat16*>(y_bf16), ld_bf, vel_out, rows, C));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_cfg_euler_dev(float* y, const float* pc, const float* pu, const float* params_dev, void* y_bf16, int ld_bf, float* vel_out,
                      int rows, int C, f5b_stream_t stream) {
  F5B_CHECK(rows > 0 && C > 0 && ld_bf >= C && params_dev != nullptr, "f5b_cfg_euler_dev: bad argument");
  const int64_t tot = (int64_t)rows * ld_bf;
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, (pu ? 16.0 : 12.0) * rows * C + (y_bf16 ? 2.0 * tot : 0.0));
  F5B_CUDA(launch_dep(cfg_euler_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, ST(stream), 1, y, pc, pu, 0.f, 0.f, params_dev,
                      reinterpret_cast<__nv_bfloat16*>(y_bf16), ld_bf, vel_out, rows, C));
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_fm_prepare(const float* x1, const float* x0, const float* time, const uint8_t* span_mask, float* phi, float* flow, float* cond,
                   int B, int n, int C, f5b_stream_t stream) {
  F5B_CHECK(x1 && x0 && time && span_mask && phi && flow && cond && B > 0 && n > 0 && C > 0, "f5b_fm_prepare: bad argument");
  const int64_t tot = (int64_t)B * n * C;
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 20.0 * tot);
  fm_prepare_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ST(stream)>>>(x1, x0, time, span_mask, phi, flow, cond, n, C, tot);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_masked_mse(const float* pred, const float* flow, const uint8_t* mask, float* ws /*[2*1024]*/, float* out2, int rows, int C,
                   f5b_stream_t stream) {
  F5B_CHECK(pred && flow && mask && ws && out2 && rows > 0 && C > 0, "f5b_masked_mse: bad argument");
  const int nblk = rows < 1024 ? rows : 1024;
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 8.0 * rows * C, 2);
  masked_mse_partial_kernel<<<nblk, 256, 0, ST(stream)>>>(pred, flow, mask, ws, rows, C);
  F5B_CUDA(cudaGetLastError());
  masked_mse_final_kernel<<<1, 32, 0, ST(stream)>>>(ws, nblk, out2);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_rope_table(float* out, int n, f5b_stream_t stream) {
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 256.0 * n);
  rope_table_kernel<<<(n * 32 + 255) / 256, 256, 0, ST(stream)>>>(reinterpret_cast<float2*>(out), n);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_im2col7(const float* mel, void* out, int B, int T, int n_mels, int ld, f5b_stream_t stream) {
  F5B_CHECK(ld >= 7 * n_mels && (ld & 7) == 0, "f5b_im2col7: ld=%d must be >= 7*n_mels and a multiple of 8", ld);
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, (double)B * T * (4.0 * n_mels + 2.0 * ld));
  im2col7_kernel<<<B * T, 256, 0, ST(stream)>>>(mel, reinterpret_cast<__nv_bfloat16*>(out), T, n_mels, ld);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
