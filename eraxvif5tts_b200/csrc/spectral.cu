// MelSpec ("vocos" type) and the Vocos ISTFT head: fp32 1024-point FFTs done entirely in shared memory, one CTA per frame,
// coalesced global traffic (algorithmic bytes: MelSpec 1024 B in + 400 B out per frame; iSTFT 4104 B in + 1024 B out).
//   MelSpec : /root/reference/src/f5_tts/model/modules.py:83-101 (torchaudio MelSpectrogram: reflect pad n_fft/2, periodic Hann,
//             |rFFT| (power=1), HTK filterbank without norm, log(clamp 1e-5)).
//   iSTFT   : vocos ISTFTHead (third-party; call sites infer/f5tts_wrapper.py:524, infer/utils_infer.py:488):
//             mag = clip(exp(.), 1e2), S = mag (cos p + i sin p), torch.istft(center=True, hann).
#include <math.h>

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

// Twiddles exp(-2*pi*i*j/1024) and the periodic Hann window live in a device table built once per process (double precision
// on the host); every CTA copies them into shared memory and then transforms FRAMES_PER_CTA consecutive frames.
constexpr int FRAMES_PER_CTA = 4;
__device__ float2 g_twiddle[NFFT / 2];
__device__ float g_hann[NFFT];

static int ensure_tables() {
  static bool done = false;  // one process per GPU, so a process-wide flag is enough
  if (done) return 0;
  static float2 tw[NFFT / 2];
  static float hw[NFFT];
  const double pi = 3.14159265358979323846;
  for (int j = 0; j < NFFT / 2; ++j) {
    tw[j].x = (float)cos(2.0 * pi * j / NFFT);
    tw[j].y = (float)-sin(2.0 * pi * j / NFFT);
  }
  for (int i = 0; i < NFFT; ++i) hw[i] = (float)(0.5 - 0.5 * cos(2.0 * pi * i / NFFT));
  F5B_CUDA(cudaMemcpyToSymbol(g_twiddle, tw, sizeof(tw)));
  F5B_CUDA(cudaMemcpyToSymbol(g_hann, hw, sizeof(hw)));
  done = true;
  return 0;
}

__device__ __forceinline__ void load_tables(float2* tw, float* hw) {
  for (int j = threadIdx.x; j < NFFT / 2; j += blockDim.x) tw[j] = g_twiddle[j];
  for (int j = threadIdx.x; j < NFFT; j += blockDim.x) hw[j] = g_hann[j];
}

__global__ void __launch_bounds__(256) melspec_kernel(const float* __restrict__ wav, const float* __restrict__ fb,
                                                      const int32_t* __restrict__ ranges, float* __restrict__ out, int L, int T,
                                                      int n_mels) {
  __shared__ float re[NFFT], im[NFFT];
  __shared__ float2 tw[NFFT / 2];
  __shared__ float hw[NFFT];
  __shared__ float mag[NFFT / 2 + 1];
  const int b = blockIdx.y;
  load_tables(tw, hw);
  __syncthreads();
  const float* w = wav + (size_t)b * L;
  for (int f = 0; f < FRAMES_PER_CTA; ++f) {
    const int t = blockIdx.x * FRAMES_PER_CTA + f;
    if (t >= T) break;  // block-uniform
    for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
      int idx = t * HOP + i - NFFT / 2;
      if (idx < 0) idx = -idx;                 // reflect (no edge repeat), torch "reflect" padding
      if (idx >= L) idx = 2 * (L - 1) - idx;
      idx = max(0, min(L - 1, idx));
      const int rev = (int)(__brev((unsigned)i) >> 22);
      re[rev] = w[idx] * hw[i];
      im[rev] = 0.f;
    }
    __syncthreads();
    fft1024_smem(re, im, tw);
    for (int k = threadIdx.x; k <= NFFT / 2; k += blockDim.x) mag[k] = sqrtf(re[k] * re[k] + im[k] * im[k]);
    __syncthreads();
    for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
      const int f0 = ranges[2 * m], f1 = ranges[2 * m + 1];
      float acc = 0.f;
      for (int k = f0; k < f1; ++k) acc += mag[k] * __ldg(fb + (size_t)k * n_mels + m);
      out[((size_t)b * T + t) * n_mels + m] = logf(fmaxf(acc, 1e-5f));
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) istft_frames_kernel(const float* __restrict__ head, float* __restrict__ frames, int ld,
                                                           int rows) {
  __shared__ float re[NFFT], im[NFFT];
  __shared__ float2 tw[NFFT / 2];
  __shared__ float hw[NFFT];
  load_tables(tw, hw);
  __syncthreads();
  for (int f = 0; f < FRAMES_PER_CTA; ++f) {
    const size_t row = (size_t)blockIdx.x * FRAMES_PER_CTA + f;
    if (row >= (size_t)rows) break;  // block-uniform
    const float* hr = head + row * ld;
    // X[k] = mag (cos p + i sin p), k = 0..512; Hermitian extension; inverse transform = conj(FFT(conj(X))) / N
    for (int k = threadIdx.x; k <= NFFT / 2; k += blockDim.x) {
      const float mg = fminf(expf(hr[k]), 1e2f);
      float sn, cs;
      sincosf(hr[NFFT / 2 + 1 + k], &sn, &cs);
      float xr = mg * cs, xi = mg * sn;
      if (k == 0 || k == NFFT / 2) xi = 0.f;  // irfft ignores the imaginary part of DC and Nyquist
      const int rev = (int)(__brev((unsigned)k) >> 22);
      re[rev] = xr;
      im[rev] = -xi;  // conj(X[k])
      if (k > 0 && k < NFFT / 2) {
        const int rev2 = (int)(__brev((unsigned)(NFFT - k)) >> 22);
        re[rev2] = xr;
        im[rev2] = xi;  // conj(conj(X[k]))
      }
    }
    __syncthreads();
    fft1024_smem(re, im, tw);
    float* fr = frames + row * NFFT;
    for (int i = threadIdx.x; i < NFFT; i += blockDim.x) fr[i] = re[i] * (1.0f / NFFT) * hw[i];
    __syncthreads();
  }
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
    const float hwv = g_hann[i];
    env += hwv * hwv;
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
  if (ensure_tables()) return -2;
  LaunchScope scope(K_SPECTRAL, static_cast<cudaStream_t>(stream), 0, (double)B * T * (HOP * 4.0 + n_mels * 4.0));
  melspec_kernel<<<dim3((T + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(wav, fb, ranges, out,
                                                                                                                 L, T, n_mels);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_istft_head(const float* head, int ld, float* frames_ws, float* wav, int B, int T, f5b_stream_t stream) {
  F5B_CHECK(head && frames_ws && wav, "f5b_istft_head: null pointer");
  F5B_CHECK(B > 0 && T > 1 && ld >= NFFT + 2, "f5b_istft_head: bad shape B %d T %d ld %d", B, T, ld);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ensure_tables()) return -2;
  LaunchScope scope(K_SPECTRAL, s, 0, (double)B * T * ((NFFT + 2) * 4.0 + HOP * 4.0), 2);
  istft_frames_kernel<<<(B * T + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA, 256, 0, s>>>(head, frames_ws, ld, B * T);
  F5B_CUDA(cudaGetLastError());
  const int out_len = HOP * (T - 1);
  istft_ola_kernel<<<dim3((out_len + 255) / 256, B), 256, 0, s>>>(frames_ws, wav, T, out_len);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
