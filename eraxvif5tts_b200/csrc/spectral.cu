// MelSpec ("vocos" type) and the Vocos ISTFT head: fp32 1024-point FFTs, ONE WARP PER FRAME, the transform register-resident:
// 1024 = 32 x 32 — every lane runs a fully unrolled 32-point FFT over the 32 values it holds (constant twiddles, no memory), the
// warp transposes once through padded shared memory while applying the W_1024 twiddles, and every lane runs a second 32-point FFT.
// Two shared-memory passes and two __syncwarp per frame instead of the 10 barrier-separated radix-2 stages of round 1 (which ran a
// 256-thread CTA per frame and was latency-bound at 60 / 215 GB/s of algorithmic traffic).  Global traffic is coalesced in both
// directions (algorithmic bytes: MelSpec 1024 B in + 400 B out per frame; iSTFT 4104 B in + 1024 B out).
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
constexpr int SP_WARPS = 4;           // frames in flight per CTA (one per warp)
constexpr int SP_PITCH = 33;          // padded row of the per-warp [32 x 32] transpose buffer (float2 elements)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ constexpr int brev5(int k) { return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4); }

// In-place 32-point decimation-in-frequency FFT over a thread's registers, fully unrolled (every index and twiddle is a
// compile-time constant).  Input in natural order, output bit-reversed: a[brev5(k)] = X[k].
__device__ __forceinline__ void fft32_dif(float2 (&a)[32]) {
  // exp(-2 pi i j / 32), j = 0..15
  constexpr float WR[16] = {1.f, 0.980785251f, 0.923879504f, 0.831469595f, 0.707106769f, 0.555570245f, 0.382683426f, 0.195090324f,
                            0.f, -0.195090324f, -0.382683426f, -0.555570245f, -0.707106769f, -0.831469595f, -0.923879504f, -0.980785251f};
  constexpr float WI[16] = {0.f, -0.195090324f, -0.382683426f, -0.555570245f, -0.707106769f, -0.831469595f, -0.923879504f, -0.980785251f,
                            -1.f, -0.980785251f, -0.923879504f, -0.831469595f, -0.707106769f, -0.555570245f, -0.382683426f, -0.195090324f};
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int grp = i / half, k = i % half;
      const int i0 = grp * 2 * half + k, i1 = i0 + half;
      const float2 u = a[i0], v = a[i1];
      a[i0] = make_float2(u.x + v.x, u.y + v.y);
      const float2 d = make_float2(u.x - v.x, u.y - v.y);
      const int t = k * (16 / half);  // W_{2 half}^k = W_32^t
      if (t == 0) a[i1] = d;
      else if (t == 8) a[i1] = make_float2(d.y, -d.x);  // multiply by -i
      else a[i1] = make_float2(d.x * WR[t] - d.y * WI[t], d.x * WI[t] + d.y * WR[t]);
    }
  }
}

// Forward 1024-point FFT of one frame by one warp.   in: a[n1] = x[32 n1 + lane];   out: a[brev5(k2)] = X[lane + 32 k2].
// sbuf: this warp's [32][SP_PITCH] float2 buffer; tw: exp(-2 pi i j / 1024), j = 0..1023 (shared memory).
__device__ __forceinline__ void fft1024_warp(float2 (&a)[32], float2* sbuf, const float2* tw, int lane) {
  fft32_dif(a);  // over n1: a[brev5(k1)] = Y[k1][n2 = lane]
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) sbuf[k1 * SP_PITCH + lane] = cmul(a[brev5(k1)], tw[(lane * k1) & (NFFT - 1)]);
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) a[n2] = sbuf[lane * SP_PITCH + n2];  // lane = k1 from here on
  __syncwarp();
  fft32_dif(a);  // over n2
}

// Twiddles exp(-2*pi*i*j/1024) and the periodic Hann window: device tables built once per process (double precision on the host).
__device__ float2 g_twiddle[NFFT];
__device__ float g_hann[NFFT];

static int ensure_tables() {
  static bool done = false;  // one process per GPU, so a process-wide flag is enough
  if (done) return 0;
  static float2 tw[NFFT];
  static float hw[NFFT];
  const double pi = 3.14159265358979323846;
  for (int j = 0; j < NFFT; ++j) {
    tw[j].x = (float)cos(2.0 * pi * j / NFFT);
    tw[j].y = (float)-sin(2.0 * pi * j / NFFT);
  }
  for (int i = 0; i < NFFT; ++i) hw[i] = (float)(0.5 - 0.5 * cos(2.0 * pi * i / NFFT));
  F5B_CUDA(cudaMemcpyToSymbol(g_twiddle, tw, sizeof(tw)));
  F5B_CUDA(cudaMemcpyToSymbol(g_hann, hw, sizeof(hw)));
  done = true;
  return 0;
}

__global__ void __launch_bounds__(SP_WARPS * 32) melspec_kernel(const float* __restrict__ wav, const float* __restrict__ fb,
                                                                const int32_t* __restrict__ ranges, float* __restrict__ out, int L, int T,
                                                                int n_mels, int frames) {
  __shared__ float2 tw[NFFT];
  __shared__ float2 sbuf_all[SP_WARPS][32 * SP_PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < NFFT; j += blockDim.x) tw[j] = g_twiddle[j];
  __syncthreads();
  float2* sbuf = sbuf_all[warp];
  float* mag = reinterpret_cast<float*>(sbuf);  // the transpose buffer is free again after the FFT
  for (int fr = blockIdx.x * SP_WARPS + warp; fr < frames; fr += gridDim.x * SP_WARPS) {
    const int b = fr / T, t = fr - b * T;
    const float* w = wav + (size_t)b * L;
    float2 a[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      const int i = 32 * n1 + lane;
      int idx = t * HOP + i - NFFT / 2;
      if (idx < 0) idx = -idx;                 // reflect (no edge repeat), torch "reflect" padding
      if (idx >= L) idx = 2 * (L - 1) - idx;
      idx = max(0, min(L - 1, idx));
      a[n1] = make_float2(__ldg(w + idx) * __ldg(g_hann + i), 0.f);
    }
    fft1024_warp(a, sbuf, tw, lane);
    // |X[k]|, k = lane + 32 k2 <= 512
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float2 x = a[brev5(k2)];
      mag[lane + 32 * k2] = sqrtf(x.x * x.x + x.y * x.y);
    }
    if (lane == 0) {
      const float2 x = a[brev5(16)];
      mag[NFFT / 2] = sqrtf(x.x * x.x + x.y * x.y);
    }
    __syncwarp();
    for (int m = lane; m < n_mels; m += 32) {
      const int f0 = ranges[2 * m], f1 = ranges[2 * m + 1];
      float acc = 0.f;
      for (int k = f0; k < f1; ++k) acc += mag[k] * __ldg(fb + (size_t)k * n_mels + m);
      out[(size_t)fr * n_mels + m] = logf(fmaxf(acc, 1e-5f));
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(SP_WARPS * 32) istft_frames_kernel(const float* __restrict__ head, float* __restrict__ frames, int ld,
                                                                     int rows) {
  __shared__ float2 tw[NFFT];
  __shared__ float2 sbuf_all[SP_WARPS][32 * SP_PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < NFFT; j += blockDim.x) tw[j] = g_twiddle[j];
  __syncthreads();
  float2* sbuf = sbuf_all[warp];
  for (int row = blockIdx.x * SP_WARPS + warp; row < rows; row += gridDim.x * SP_WARPS) {
    const float* hr = head + (size_t)row * ld;
    // half spectrum conj(X[k]) = mag (cos p - i sin p), k = 0..512, staged once (exp / sincos are evaluated once per bin)
    for (int k = lane; k <= NFFT / 2; k += 32) {
      const float mg = fminf(expf(hr[k]), 1e2f);
      float sn, cs;
      sincosf(hr[NFFT / 2 + 1 + k], &sn, &cs);
      float xi = mg * sn;
      if (k == 0 || k == NFFT / 2) xi = 0.f;  // irfft ignores the imaginary part of DC and Nyquist
      sbuf[k] = make_float2(mg * cs, -xi);
    }
    __syncwarp();
    // inverse transform = conj(FFT(conj(X))) / N with the Hermitian extension X[N - k] = conj(X[k])
    float2 a[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      const int k = 32 * n1 + lane;
      if (k <= NFFT / 2) a[n1] = sbuf[k];
      else {
        const float2 v = sbuf[NFFT - k];
        a[n1] = make_float2(v.x, -v.y);
      }
    }
    __syncwarp();
    fft1024_warp(a, sbuf, tw, lane);
    float* fr = frames + (size_t)row * NFFT;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
      const int i = lane + 32 * k2;
      fr[i] = a[brev5(k2)].x * (1.0f / NFFT) * __ldg(g_hann + i);
    }
    __syncwarp();
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
  const int frames = B * T;
  const int grid = min((frames + SP_WARPS - 1) / SP_WARPS, 8 * sm_count());
  melspec_kernel<<<grid, SP_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(wav, fb, ranges, out, L, T, n_mels, frames);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_istft_head(const float* head, int ld, float* frames_ws, float* wav, int B, int T, f5b_stream_t stream) {
  F5B_CHECK(head && frames_ws && wav, "f5b_istft_head: null pointer");
  F5B_CHECK(B > 0 && T > 1 && ld >= NFFT + 2, "f5b_istft_head: bad shape B %d T %d ld %d", B, T, ld);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ensure_tables()) return -2;
  LaunchScope scope(K_SPECTRAL, s, 0, (double)B * T * ((NFFT + 2) * 4.0 + HOP * 4.0), 2);
  istft_frames_kernel<<<min((B * T + SP_WARPS - 1) / SP_WARPS, 8 * sm_count()), SP_WARPS * 32, 0, s>>>(head, frames_ws, ld, B * T);
  F5B_CUDA(cudaGetLastError());
  const int out_len = HOP * (T - 1);
  istft_ola_kernel<<<dim3((out_len + 255) / 256, B), 256, 0, s>>>(frames_ws, wav, T, out_len);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
