// MelSpec ("vocos" type) and the Vocos ISTFT head: fp32 1024-point FFTs done entirely in shared memory, one CTA per frame,
// coalesced global traffic (algorithmic bytes: MelSpec 1024 B in + 400 B out per frame; iSTFT 4104 B in + 1024 B out).
//   MelSpec : /root/reference/src/f5_tts/model/modules.py:83-101 (torchaudio MelSpectrogram: reflect pad n_fft/2, periodic Hann,
//             |rFFT| (power=1), HTK filterbank without norm, log(clamp 1e-5)).
//   iSTFT   : vocos ISTFTHead (third-party; call sites infer/f5tts_wrapper.py:524, infer/utils_infer.py:488):
//             mag = clip(exp(.), 1e2), S = mag (cos p + i sin p), torch.istft(center=True, hann).
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int NFFT = 1024;
constexpr int HOP = 256;

// in-place radix-2 DIT FFT over smem (input already in bit-reversed order); 256 threads
__device__ __forceinline__ void fft1024_smem(float* re, float* im, const float2* tw) {
#pragma unroll 1
  for (int s = 0; s < 10; ++s) {
    const int half = 1 << s;
    for (int i = threadIdx.x; i < NFFT / 2; i += blockDim.x) {
      const int k = i & (half - 1);
      const int i0 = ((i >> s) << (s + 1)) + k;
      const int i1 = i0 + half;
      const float2 w = tw[k << (9 - s)];
      const float xr = re[i1], xi = im[i1];
      const float tr = w.x * xr - w.y * xi;
      const float ti = w.x * xi + w.y * xr;
      const float ur = re[i0], ui = im[i0];
      re[i0] = ur + tr;
      im[i0] = ui + ti;
      re[i1] = ur - tr;
      im[i1] = ui - ti;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void make_twiddles(float2* tw) {
  for (int j = threadIdx.x; j < NFFT / 2; j += blockDim.x) {
    float sn, cs;
    sincospif((float)j / 512.0f, &sn, &cs);  // angle 2*pi*j/1024
    tw[j] = make_float2(cs, -sn);
  }
}

__device__ __forceinline__ float hann_periodic(int i) { return 0.5f - 0.5f * cospif((float)i / 512.0f); }

__global__ void __launch_bounds__(256) melspec_kernel(const float* __restrict__ wav, const float* __restrict__ fb,
                                                      const int32_t* __restrict__ ranges, float* __restrict__ out, int L, int T,
                                                      int n_mels) {
  __shared__ float re[NFFT], im[NFFT];
  __shared__ float2 tw[NFFT / 2];
  __shared__ float mag[NFFT / 2 + 1];
  const int t = blockIdx.x, b = blockIdx.y;
  make_twiddles(tw);
  const float* w = wav + (size_t)b * L;
  for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
    int idx = t * HOP + i - NFFT / 2;
    if (idx < 0) idx = -idx;                 // reflect (no edge repeat), torch "reflect" padding
    if (idx >= L) idx = 2 * (L - 1) - idx;
    idx = max(0, min(L - 1, idx));
    const int rev = (int)(__brev((unsigned)i) >> 22);
    re[rev] = w[idx] * hann_periodic(i);
    im[rev] = 0.f;
  }
  __syncthreads();
  fft1024_smem(re, im, tw);
  for (int f = threadIdx.x; f <= NFFT / 2; f += blockDim.x) mag[f] = sqrtf(re[f] * re[f] + im[f] * im[f]);
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    const int f0 = ranges[2 * m], f1 = ranges[2 * m + 1];
    float acc = 0.f;
    for (int f = f0; f < f1; ++f) acc += mag[f] * __ldg(fb + (size_t)f * n_mels + m);
    out[((size_t)b * T + t) * n_mels + m] = logf(fmaxf(acc, 1e-5f));
  }
}

__global__ void __launch_bounds__(256) istft_frames_kernel(const float* __restrict__ head, float* __restrict__ frames, int ld) {
  __shared__ float re[NFFT], im[NFFT];
  __shared__ float2 tw[NFFT / 2];
  const size_t row = blockIdx.x;
  make_twiddles(tw);
  const float* hr = head + row * ld;
  // X[f] = mag (cos p + i sin p), f = 0..512; Hermitian extension; inverse transform = conj(FFT(conj(X))) / N
  for (int f = threadIdx.x; f <= NFFT / 2; f += blockDim.x) {
    const float mg = fminf(expf(hr[f]), 1e2f);
    float sn, cs;
    sincosf(hr[NFFT / 2 + 1 + f], &sn, &cs);
    float xr = mg * cs, xi = mg * sn;
    if (f == 0 || f == NFFT / 2) xi = 0.f;  // irfft ignores the imaginary part of DC and Nyquist
    const int rev = (int)(__brev((unsigned)f) >> 22);
    re[rev] = xr;
    im[rev] = -xi;  // conj(X[f])
    if (f > 0 && f < NFFT / 2) {
      const int rev2 = (int)(__brev((unsigned)(NFFT - f)) >> 22);
      re[rev2] = xr;
      im[rev2] = xi;  // conj(conj(X[f]))
    }
  }
  __syncthreads();
  fft1024_smem(re, im, tw);
  float* fr = frames + row * NFFT;
  for (int i = threadIdx.x; i < NFFT; i += blockDim.x) fr[i] = re[i] * (1.0f / NFFT) * hann_periodic(i);
}

__global__ void istft_ola_kernel(const float* __restrict__ frames, float* __restrict__ wav, int T, int out_len) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (s >= out_len) return;
  const int p = s + NFFT / 2;
  const int t_hi = min(T - 1, p / HOP);
  const int t_lo = max(0, (p - (NFFT - HOP)) / HOP);
  float acc = 0.f, env = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int i = p - t * HOP;
    if (i < 0 || i >= NFFT) continue;
    acc += frames[((size_t)b * T + t) * NFFT + i];
    const float hw = hann_periodic(i);
    env += hw * hw;
  }
  wav[(size_t)b * out_len + s] = acc / env;
}

}  // namespace f5b

using namespace f5b;

extern "C" {

int f5b_melspec(const float* wav, const float* fb, const int32_t* ranges, float* out, int B, int L, int n_mels,
                f5b_stream_t stream) {
  F5B_CHECK(wav && fb && ranges && out, "f5b_melspec: null pointer");
  F5B_CHECK(B > 0 && L > NFFT / 2 && n_mels > 0, "f5b_melspec: need L > 512 samples (reflect padding), got B %d L %d", B, L);
  const int T = 1 + L / HOP;
  LaunchScope scope(K_SPECTRAL, static_cast<cudaStream_t>(stream), 0, (double)B * T * (HOP * 4.0 + n_mels * 4.0));
  melspec_kernel<<<dim3(T, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(wav, fb, ranges, out, L, T, n_mels);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_istft_head(const float* head, int ld, float* frames_ws, float* wav, int B, int T, f5b_stream_t stream) {
  F5B_CHECK(head && frames_ws && wav, "f5b_istft_head: null pointer");
  F5B_CHECK(B > 0 && T > 1 && ld >= NFFT + 2, "f5b_istft_head: bad shape B %d T %d ld %d", B, T, ld);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LaunchScope scope(K_SPECTRAL, s, 0, (double)B * T * ((NFFT + 2) * 4.0 + HOP * 4.0), 2);
  istft_frames_kernel<<<B * T, 256, 0, s>>>(head, frames_ws, ld);
  F5B_CUDA(cudaGetLastError());
  const int out_len = HOP * (T - 1);
  istft_ola_kernel<<<dim3((out_len + 255) / 256, B), 256, 0, s>>>(frames_ws, wav, T, out_len);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
