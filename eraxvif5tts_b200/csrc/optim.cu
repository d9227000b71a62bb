// Optimizer step of the training loop (/root/reference/src/f5_tts/model/trainer.py:1280-1287, 1321): global-norm gradient clipping
// (accelerator.clip_grad_norm_), torch.optim.AdamW (betas 0.9/0.98, eps 1e-8, decoupled weight decay; trainer.py:316-323) and
// the rank-0 EMA lerp (ema_pytorch) fused into ONE pass over flat fp32 buffers; the bf16 working copy that the GEMM engine
// reads is refreshed in the same pass.  HBM-bound: 20 B read + 18 B written per parameter.
#include <algorithm>

#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (int64_t j = i; j < n; ++j) s += g[j] * g[j];
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void sumsq_final_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;  // fixed order: deterministic
    for (int i = 0; i < nblk; ++i) a += partial[i];
    out[0] = (float)a;
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, wd, bc1, bc2, max_norm, grad_scale, ema_decay;
};

__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ ema,
                                                        __nv_bfloat16* __restrict__ pbf, int64_t n, AdamArgs a,
                                                        const float* __restrict__ sumsq) {
  float clip = a.grad_scale;
  if (sumsq != nullptr && a.max_norm > 0.f) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = sqrtf(__ldg(sumsq)) * a.grad_scale;
    clip *= fminf(1.0f, a.max_norm / (total + 1e-6f));
  }
  const float step_size = a.lr / a.bc1;
  const float inv_sqrt_bc2 = rsqrtf(a.bc2);
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const int cnt = (int)min((int64_t)4, n - i);
  float pv[4], gv[4], mv[4], vv[4];
  if (cnt == 4) {
    *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(p + i);
    *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(g + i);
    *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + i);
    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i);
  } else {
    for (int j = 0; j < cnt; ++j) { pv[j] = p[i + j]; gv[j] = g[i + j]; mv[j] = m[i + j]; vv[j] = v[i + j]; }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < cnt) {
      const float gr = gv[j] * clip;
      pv[j] *= 1.0f - a.lr * a.wd;  // decoupled weight decay first (torch.optim.AdamW)
      mv[j] = a.beta1 * mv[j] + (1.0f - a.beta1) * gr;
      vv[j] = a.beta2 * vv[j] + (1.0f - a.beta2) * gr * gr;
      const float denom = sqrtf(vv[j]) * inv_sqrt_bc2 + a.eps;
      pv[j] -= step_size * (mv[j] / denom);
    }
  }
  if (cnt == 4) {
    *reinterpret_cast<float4*>(p + i) = *reinterpret_cast<float4*>(pv);
    *reinterpret_cast<float4*>(m + i) = *reinterpret_cast<float4*>(mv);
    *reinterpret_cast<float4*>(v + i) = *reinterpret_cast<float4*>(vv);
  } else {
    for (int j = 0; j < cnt; ++j) { p[i + j] = pv[j]; m[i + j] = mv[j]; v[i + j] = vv[j]; }
  }
  if (ema != nullptr && a.ema_decay >= 0.f) {
    for (int j = 0; j < cnt; ++j) {
      const float e = ema[i + j];
      ema[i + j] = e + (1.0f - a.ema_decay) * (pv[j] - e);  // ma.lerp_(current, 1 - decay)
    }
  }
  if (pbf != nullptr)
    for (int j = 0; j < cnt; ++j) pbf[i + j] = __float2bfloat16(pv[j]);
}

}  // namespace f5b

using namespace f5b;

extern "C" {

int f5b_grad_sumsq(const float* g, int64_t n, float* ws, float* out, f5b_stream_t stream) {
  F5B_CHECK(g && ws && out && n > 0, "f5b_grad_sumsq: bad argument");
  F5B_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, "f5b_grad_sumsq: gradient buffer must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblk = (int)std::min<int64_t>(1024, (n + 1023) / 1024);
  LaunchScope scope(K_ELEMENTWISE, s, 0, 4.0 * (double)n, 2);
  sumsq_partial_kernel<<<nblk, 256, 0, s>>>(g, n, ws);
  F5B_CUDA(cudaGetLastError());
  sumsq_final_kernel<<<1, 32, 0, s>>>(ws, nblk, out);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, const float* grad_sumsq, float max_norm, float grad_scale,
                       float ema_decay, f5b_stream_t stream) {
  F5B_CHECK(p && g && m && v && n > 0 && step >= 1, "f5b_adamw_ema_step: bad argument");
  F5B_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
              reinterpret_cast<uintptr_t>(v)) & 15) == 0, "f5b_adamw_ema_step: buffers must be 16-byte aligned");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = 1.0f - powf(beta1, (float)step);
  a.bc2 = 1.0f - powf(beta2, (float)step);
  a.max_norm = max_norm; a.grad_scale = grad_scale; a.ema_decay = ema_decay;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t nthreads = (n + 3) / 4;
  LaunchScope scope(K_ELEMENTWISE, s, 0, (double)n * (16.0 + 12.0 + (ema ? 8.0 : 0.0) + (p_bf16 ? 2.0 : 0.0)));
  adamw_ema_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, s>>>(p, g, m, v, ema, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, a,
                                                                      grad_sumsq);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
