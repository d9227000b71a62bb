// Monotonic alignment search and the duration predictor: the pieces the distillation script's duration loss is
// built from.  Reference: model/alignment_utils.py:154-212 (viterbi_vectorized_alignment), :214-257 (windowed_monotonic_alignment),
// model/duration_predictor.py:4-44 (DurationPredictor.forward, eval mode).
//
// The reference runs the Viterbi recurrence as a Python double loop over (token, frame) -- nt x mel_len tiny torch launches per
// call -- and backtracks with .item() per token.  Here: one CTA per batch item sweeps the anti-diagonals of the [nt, mel_len]
// lattice (thread = token row, one __syncthreads per diagonal), then a second kernel backtracks with block-wide reverse searches.
// Only additions and maxima of the same fp32 operands as the reference are involved, so path_prob and the alignment are bit-exact.
#include "common.cuh"
#include "f5b_internal.h"

namespace f5b {

constexpr int VT_THREADS = 1024;

// path[n, t] = sim[n, t] + max(path[n-1, t], path[n, t-1]);  first row / first column are running sums (alignment_utils.py:159-175)
__global__ void __launch_bounds__(VT_THREADS) viterbi_forward_kernel(const float* __restrict__ sim, float* __restrict__ path, int nt, int T) {
  __shared__ float buf[2][VT_THREADS];
  const int b = blockIdx.x, i = threadIdx.x;
  const float* S = sim + (size_t)b * nt * T;
  float* P = path + (size_t)b * nt * T;
  const int chunk = blockDim.x;  // rows swept together (<= VT_THREADS)
  for (int r0 = 0; r0 < nt; r0 += chunk) {
    const int rows = min(chunk, nt - r0);
    const int n = r0 + i;
    const bool mine = i < rows;
    float left = 0.f;
    float sv_next = (mine && i == 0) ? S[(size_t)n * T] : 0.f;  // thread i's first element is (n, 0) at step s = i
    const int steps = rows + T - 1;
    for (int s = 0; s < steps; ++s) {
      const int t = s - i;
      const bool active = mine && t >= 0 && t < T;
      const float sv = sv_next;
      // prefetch the next step's similarity (off the dependency chain)
      const int tn = t + 1;
      if (mine && tn >= 0 && tn < T) sv_next = S[(size_t)n * T + tn];
      if (active) {
        float val;
        if (n == 0) {
          val = t == 0 ? sv : left + sv;
        } else {
          const float up = i == 0 ? P[(size_t)(n - 1) * T + t] : buf[(s - 1) & 1][i - 1];
          val = t == 0 ? up + sv : sv + fmaxf(up, left);
        }
        P[(size_t)n * T + t] = val;
        left = val;
        buf[s & 1][i] = val;
      }
      __syncthreads();
    }
  }
}

// backtracking (alignment_utils.py:183-210): from the last token up, the segment of token n ends at curr and starts at the LAST
// index j < curr where path[n, j+1] - path[n, j] > 0 (0 if none, and for token 0).  align must be zero-filled.
__global__ void __launch_bounds__(256) viterbi_backtrack_kernel(const float* __restrict__ path, float* __restrict__ align,
                                                                int32_t* __restrict__ durations, int nt, int T) {
  __shared__ int red[8];
  __shared__ int s_best;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* P = path + (size_t)b * nt * T;
  float* A = align + (size_t)b * nt * T;
  int curr = T - 1;
  for (int n = nt - 1; n >= 0; --n) {
    int bidx = 0;
    if (n > 0) {
      const float* row = P + (size_t)n * T;
      for (int base = curr - 1; base >= 0; base -= 256) {
        const int j = base - tid;
        int cand = -1;
        if (j >= 0 && row[j + 1] - row[j] > 0.f) cand = j;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = max(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if ((tid & 31) == 0) red[tid >> 5] = cand;
        __syncthreads();
        if (tid == 0) {
          int m = -1;
          for (int w = 0; w < 8; ++w) m = max(m, red[w]);
          s_best = m;
        }
        __syncthreads();
        const int m = s_best;
        __syncthreads();
        if (m >= 0) { bidx = m; break; }
      }
    }
    for (int t = bidx + tid; t <= curr; t += 256) A[(size_t)n * T + t] = 1.f;
    if (durations && tid == 0) durations[(size_t)b * nt + n] = curr - bidx + 1;
    curr = bidx - 1;
    if (curr < 0) break;  // tokens above keep an empty segment (durations pre-zeroed)
  }
}

// windowed greedy search (alignment_utils.py:214-257): token n (all but the last) ends at the arg-max of sim[n, .] inside a window of
// +-W frames around its proportional position; the last token takes the rest.  err[b] = 1 when a window is empty (the reference's
// torch.argmax raises there).
__global__ void __launch_bounds__(256) window_align_kernel(const float* __restrict__ sim, float* __restrict__ align,
                                                           int32_t* __restrict__ durations, int32_t* __restrict__ err, int nt, int T, int W) {
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ int s_best;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* S = sim + (size_t)b * nt * T;
  float* A = align + (size_t)b * nt * T;
  const double fpp = (double)T / (double)nt;
  int start = 0;
  for (int n = 0; n < nt - 1; ++n) {
    const int expected_end = (int)((double)(n + 1) * fpp);
    const int ws = max(start, expected_end - W), we = min(T - 1, expected_end + W);
    if (ws > we) {
      if (tid == 0) err[b] = 1;
      return;
    }
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = ws + tid; t <= we; t += 256) {
      const float v = S[(size_t)n * T + t];
      if (v > bv || bi == 0x7fffffff) { bv = v; bi = t; }  // strictly greater: the first maximum wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = rv[0];
      int idx = ri[0];
      for (int w = 1; w < 8; ++w)
        if (ri[w] != 0x7fffffff && (idx == 0x7fffffff || rv[w] > v || (rv[w] == v && ri[w] < idx))) { v = rv[w]; idx = ri[w]; }
      s_best = idx;
    }
    __syncthreads();
    const int best_end = s_best;
    __syncthreads();
    for (int t = start + tid; t <= best_end; t += 256) A[(size_t)n * T + t] = 1.f;
    if (durations && tid == 0) durations[(size_t)b * nt + n] = best_end - start + 1;
    start = best_end + 1;
    if (start >= T) break;
  }
  // tokens the loop never reached keep an empty segment (durations pre-zeroed); the last token takes what is left
  if (start < T) {
    for (int t = start + tid; t < T; t += 256) A[(size_t)(nt - 1) * T + t] = 1.f;
    if (durations && tid == 0) durations[(size_t)b * nt + nt - 1] = T - start;
  }
}

// ---- DurationPredictor.forward, eval mode (model/duration_predictor.py:27-44) ----
// stage 1: h1[b, t, f] = relu(b1[f] + sum_k sum_c W1[f, c, k] * E[ids[b, t+k-pad] + 1, c] * mask[b, t+k-pad])   (warp per position)
__global__ void __launch_bounds__(256) durpred_conv1_kernel(const int64_t* __restrict__ ids, const float* __restrict__ mask,
                                                            const float* __restrict__ table, const float* __restrict__ w1,
                                                            const float* __restrict__ b1, float* __restrict__ h1, int nt, int Cin, int F,
                                                            int ksize, int id_shift) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t = blockIdx.x * 8 + warp;
  if (t >= nt) return;
  const int pad = ksize / 2;
  for (int f = 0; f < F; ++f) {
    float acc = 0.f;
    for (int k = 0; k < ksize; ++k) {
      const int tt = t + k - pad;
      if (tt < 0 || tt >= nt) continue;
      const float m = mask[(size_t)b * nt + tt];
      if (m == 0.f) continue;
      const float* e = table + (size_t)(ids[(size_t)b * nt + tt] + id_shift) * Cin;
      const float* w = w1 + (size_t)f * Cin * ksize + k;
      float a = 0.f;
      for (int c = lane; c < Cin; c += 32) a = fmaf(w[(size_t)c * ksize], e[c], a);
      acc += a * m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) h1[((size_t)b * nt + t) * F + f] = fmaxf(acc + b1[f], 0.f);
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// stage 2 (one CTA per batch item, h1 in shared memory): GroupNorm(1, F) over (F, nt) -> * mask -> conv_2 -> relu -> GroupNorm(1, F)
// -> * mask -> proj (1x1, F -> 1) -> * mask
__global__ void __launch_bounds__(256) durpred_tail_kernel(const float* __restrict__ h1, const float* __restrict__ mask,
                                                           const float* __restrict__ g1w, const float* __restrict__ g1b,
                                                           const float* __restrict__ w2, const float* __restrict__ b2,
                                                           const float* __restrict__ g2w, const float* __restrict__ g2b,
                                                           const float* __restrict__ pw, const float* __restrict__ pb,
                                                           float* __restrict__ out, int nt, int F, int ksize, float eps) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  float* a = sm;            // [nt][F]  normalised + masked conv_2 input
  float* c = sm + (size_t)nt * F;  // [nt][F]  conv_2 output after relu
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = nt * F;
  const float* H = h1 + (size_t)b * N;
  const float* M = mask + (size_t)b * nt;
  float s = 0.f;
  for (int i = tid; i < N; i += 256) { const float v = H[i]; a[i] = v; s += v; }
  const float mean1 = block_sum_256(s, red) / (float)N;
  float q = 0.f;
  for (int i = tid; i < N; i += 256) { const float d = a[i] - mean1; q += d * d; }
  const float rstd1 = rsqrtf(block_sum_256(q, red) / (float)N + eps);
  for (int i = tid; i < N; i += 256) {
    const int t = i / F, f = i - t * F;
    a[i] = ((a[i] - mean1) * rstd1 * g1w[f] + g1b[f]) * M[t];
  }
  __syncthreads();
  const int pad = ksize / 2;
  s = 0.f;
  for (int i = tid; i < N; i += 256) {
    const int t = i / F, g = i - t * F;
    float acc = b2[g];
    for (int k = 0; k < ksize; ++k) {
      const int tt = t + k - pad;
      if (tt < 0 || tt >= nt) continue;
      const float* w = w2 + (size_t)g * F * ksize + k;
      const float* x = a + (size_t)tt * F;
      for (int f = 0; f < F; ++f) acc = fmaf(w[(size_t)f * ksize], x[f], acc);
    }
    acc = fmaxf(acc, 0.f);
    c[i] = acc;
    s += acc;
  }
  const float mean2 = block_sum_256(s, red) / (float)N;
  q = 0.f;
  for (int i = tid; i < N; i += 256) { const float d = c[i] - mean2; q += d * d; }
  const float rstd2 = rsqrtf(block_sum_256(q, red) / (float)N + eps);
  for (int t = tid; t < nt; t += 256) {
    const float m = M[t];
    float acc = pb[0];
    for (int f = 0; f < F; ++f) acc = fmaf(pw[f], ((c[(size_t)t * F + f] - mean2) * rstd2 * g2w[f] + g2b[f]) * m, acc);
    out[(size_t)b * nt + t] = acc * m;
  }
}

}  // namespace f5b

using namespace f5b;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int f5b_align_viterbi(const float* sim, float* path_ws, float* align, int32_t* durations, int B, int nt, int T, f5b_stream_t stream) {
  F5B_CHECK(sim && path_ws && align && B > 0 && nt > 0 && T > 0, "f5b_align_viterbi: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 16.0 * B * nt * T);
  F5B_CUDA(cudaMemsetAsync(align, 0, sizeof(float) * (size_t)B * nt * T, ST(stream)));
  if (durations) F5B_CUDA(cudaMemsetAsync(durations, 0, sizeof(int32_t) * (size_t)B * nt, ST(stream)));
  const int threads = nt >= VT_THREADS ? VT_THREADS : (nt + 31) / 32 * 32;  // fewer warps = cheaper per-diagonal barrier
  viterbi_forward_kernel<<<B, threads, 0, ST(stream)>>>(sim, path_ws, nt, T);
  F5B_CUDA(cudaGetLastError());
  viterbi_backtrack_kernel<<<B, 256, 0, ST(stream)>>>(path_ws, align, durations, nt, T);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_align_window(const float* sim, float* align, int32_t* durations, int32_t* err, int B, int nt, int T, int window,
                     f5b_stream_t stream) {
  F5B_CHECK(sim && align && err && B > 0 && nt > 0 && T > 0 && window >= 0, "f5b_align_window: bad argument");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 0, 8.0 * B * nt * T);
  F5B_CUDA(cudaMemsetAsync(align, 0, sizeof(float) * (size_t)B * nt * T, ST(stream)));
  F5B_CUDA(cudaMemsetAsync(err, 0, sizeof(int32_t) * (size_t)B, ST(stream)));
  if (durations) F5B_CUDA(cudaMemsetAsync(durations, 0, sizeof(int32_t) * (size_t)B * nt, ST(stream)));
  window_align_kernel<<<B, 256, 0, ST(stream)>>>(sim, align, durations, err, nt, T, window);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_duration_predictor(const int64_t* ids, int id_shift, const float* mask, const float* table, int vocab_rows, const float* conv1_w,
                           const float* conv1_b, const float* norm1_w, const float* norm1_b, const float* conv2_w, const float* conv2_b,
                           const float* norm2_w, const float* norm2_b, const float* proj_w, const float* proj_b, float* h1_ws, float* out,
                           int B, int nt, int Cin, int F, int ksize, f5b_stream_t stream) {
  F5B_CHECK(ids && mask && table && conv1_w && conv1_b && norm1_w && norm1_b && conv2_w && conv2_b && norm2_w && norm2_b && proj_w &&
                proj_b && h1_ws && out && B > 0 && nt > 0 && Cin > 0 && F > 0 && (ksize & 1) && vocab_rows > 0,
            "f5b_duration_predictor: bad argument");
  const size_t smem = 2 * sizeof(float) * (size_t)nt * F;
  F5B_CHECK(smem <= 200 * 1024, "f5b_duration_predictor: nt * filter_channels too large for one CTA's shared memory");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 2.0 * B * nt * ((double)Cin * F * ksize + (double)F * F * ksize), 0);
  durpred_conv1_kernel<<<dim3((nt + 7) / 8, B), 256, 0, ST(stream)>>>(ids, mask, table, conv1_w, conv1_b, h1_ws, nt, Cin, F, ksize,
                                                                      id_shift);
  F5B_CUDA(cudaGetLastError());
  F5B_CUDA(cudaFuncSetAttribute(durpred_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  durpred_tail_kernel<<<B, 256, smem, ST(stream)>>>(h1_ws, mask, norm1_w, norm1_b, conv2_w, conv2_b, norm2_w, norm2_b, proj_w, proj_b, out,
                                                    nt, F, ksize, 1e-5f);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
