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

// dropout of the predictor's train mode (nn.Dropout(p_dropout), duration_predictor.py:16): the counter-based generator of the DiT
// training path (train_kernels.cu / oracle.dropout_multipliers with layer 0), one decision per element index, regenerated by the
// backward.  thr16 = 0: eval mode.
struct DpDrop {
  uint32_t thr16;
  float scale;
  uint64_t seed;
};
__device__ __forceinline__ float dp_mult(const DpDrop& d, int site, size_t idx) {
  if (d.thr16 == 0) return 1.f;
  const uint64_t key = d.seed * 0xD1342543DE82EF95ull + (uint64_t)(site + 1) * 0x9E3779B97F4A7C15ull;
  uint64_t z = (uint64_t)(idx >> 2) * 0x9E3779B97F4A7C15ull + key;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((uint32_t)(z >> (16 * (idx & 3))) & 0xFFFFu) >= d.thr16 ? d.scale : 0.f;
}
static DpDrop make_dpdrop(float p, uint64_t seed) {
  if (!(p > 0.f)) return DpDrop{0u, 1.f, 0ull};
  uint32_t thr = (uint32_t)(p * 65536.0f + 0.5f);
  if (thr > 65535u) thr = 65535u;
  return DpDrop{thr, 65536.0f / (65536.0f - (float)thr), seed};
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
                                                           float* __restrict__ out, int nt, int F, int ksize, float eps,
                                                           const DpDrop dr, float* __restrict__ a_out, float* __restrict__ c_out,
                                                           float* __restrict__ stats_out) {
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
    a[i] = ((a[i] - mean1) * rstd1 * g1w[f] + g1b[f]) * M[t] * dp_mult(dr, 0, (size_t)blockIdx.x * N + i);
    if (a_out) a_out[(size_t)blockIdx.x * N + i] = a[i];
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
    if (c_out) c_out[(size_t)blockIdx.x * N + i] = acc;
    s += acc;
  }
  const float mean2 = block_sum_256(s, red) / (float)N;
  q = 0.f;
  for (int i = tid; i < N; i += 256) { const float d = c[i] - mean2; q += d * d; }
  const float rstd2 = rsqrtf(block_sum_256(q, red) / (float)N + eps);
  for (int t = tid; t < nt; t += 256) {
    const float m = M[t];
    float acc = pb[0];
    for (int f = 0; f < F; ++f)
      acc = fmaf(pw[f], ((c[(size_t)t * F + f] - mean2) * rstd2 * g2w[f] + g2b[f]) * m * dp_mult(dr, 1, (size_t)b * N + (size_t)t * F + f), acc);
    out[(size_t)b * nt + t] = acc * m;
  }
  if (stats_out && tid == 0) {
    stats_out[b * 4 + 0] = mean1; stats_out[b * 4 + 1] = rstd1; stats_out[b * 4 + 2] = mean2; stats_out[b * 4 + 3] = rstd2;
  }
}

// ---- DurationPredictor backward (train mode).  One CTA per batch item walks back from d loss / d logw to d(conv_1 pre-activation):
//   out = m (pb + sum_f pw[f] z2[t,f]),  z2 = m d2 GN2(c),  c = relu(conv_2(a)),  a = m d1 GN1(h1),  h1 = relu(conv_1(m E[ids]))
// parameter gradients are accumulated with atomics (the caller zeroes them).  Shared memory: X (c, then d a), Y (a, then h1), G.
__global__ void __launch_bounds__(256) durpred_bwd_tail_kernel(const float* __restrict__ dlogw, const float* __restrict__ mask,
                                                               const float* __restrict__ h1, const float* __restrict__ a_sv,
                                                               const float* __restrict__ c_sv, const float* __restrict__ stats,
                                                               const float* __restrict__ g1w, const float* __restrict__ w2,
                                                               const float* __restrict__ g2w, const float* __restrict__ g2b,
                                                               const float* __restrict__ pw, float* __restrict__ d_b1,
                                                               float* __restrict__ d_g1w, float* __restrict__ d_g1b, float* __restrict__ d_w2,
                                                               float* __restrict__ d_b2, float* __restrict__ d_g2w, float* __restrict__ d_g2b,
                                                               float* __restrict__ d_pw, float* __restrict__ d_pb, float* __restrict__ dpre1,
                                                               int nt, int F, int ksize, const DpDrop dr) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = nt * F;
  float* X = sm;
  float* Y = sm + N;
  float* G = sm + 2 * (size_t)N;
  const float* M = mask + (size_t)b * nt;
  const float* DL = dlogw + (size_t)b * nt;
  const float mean1 = stats[b * 4], rstd1 = stats[b * 4 + 1], mean2 = stats[b * 4 + 2], rstd2 = stats[b * 4 + 3];
  const int pad = ksize / 2;
  // 1. through proj and the second dropout: G = d loss / d GN2 output;  X = xhat2
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid; i < N; i += 256) {
    const int t = i / F, f = i - t * F;
    const float m = M[t];
    const float xh = (c_sv[(size_t)b * N + i] - mean2) * rstd2;
    const float d2 = dp_mult(dr, 1, (size_t)b * N + i);
    const float gy = DL[t] * m * m * pw[f] * d2;
    X[i] = xh;
    G[i] = gy;
    Y[i] = DL[t] * m * m * d2 * (xh * g2w[f] + g2b[f]);  // contribution to d proj.weight[f]
    const float dxh = gy * g2w[f];
    s1 += dxh;
    s2 += dxh * xh;
  }
  s1 = block_sum_256(s1, red) / (float)N;
  s2 = block_sum_256(s2, red) / (float)N;
  if (tid < F) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int t = 0; t < nt; ++t) { a0 += Y[t * F + tid]; a1 += G[t * F + tid] * X[t * F + tid]; a2 += G[t * F + tid]; }
    atomicAdd(d_pw + tid, a0); atomicAdd(d_g2w + tid, a1); atomicAdd(d_g2b + tid, a2);
  }
  if (tid == 0) {
    float a0 = 0.f;
    for (int t = 0; t < nt; ++t) a0 += DL[t] * M[t];
    atomicAdd(d_pb, a0);
  }
  __syncthreads();
  // 2. GroupNorm 2 + relu backward: G = d loss / d conv_2 pre-activation;  Y = a (conv_2's input)
  for (int i = tid; i < N; i += 256) {
    const int f = i % F;
    const float dxh = G[i] * g2w[f];
    const float dc = rstd2 * (dxh - s1 - X[i] * s2);
    G[i] = c_sv[(size_t)b * N + i] > 0.f ? dc : 0.f;
    Y[i] = a_sv[(size_t)b * N + i];
  }
  __syncthreads();
  // 3. conv_2 backward: bias, weight, input (X = d loss / d a)
  if (tid < F) {
    float a0 = 0.f;
    for (int t = 0; t < nt; ++t) a0 += G[t * F + tid];
    atomicAdd(d_b2 + tid, a0);
  }
  for (int e = tid; e < F * F * ksize; e += 256) {
    const int g = e / (F * ksize), r = e - g * F * ksize, f = r / ksize, k = r - f * ksize;
    float acc = 0.f;
    for (int t = 0; t < nt; ++t) {
      const int tt = t + k - pad;
      if (tt >= 0 && tt < nt) acc = fmaf(G[t * F + g], Y[tt * F + f], acc);
    }
    atomicAdd(d_w2 + e, acc);
  }
  for (int i = tid; i < N; i += 256) {
    const int tt = i / F, f = i - tt * F;
    float acc = 0.f;
    for (int k = 0; k < ksize; ++k) {
      const int t = tt - k + pad;
      if (t < 0 || t >= nt) continue;
      const float* w = w2 + (size_t)f * ksize + k;
      const float* gr = G + (size_t)t * F;
      for (int g = 0; g < F; ++g) acc = fmaf(w[(size_t)g * F * ksize], gr[g], acc);
    }
    X[i] = acc;
  }
  __syncthreads();
  // 4. first dropout, mask, GroupNorm 1, relu: G = d loss / d GN1 output, Y = xhat1
  s1 = 0.f; s2 = 0.f;
  for (int i = tid; i < N; i += 256) {
    const int t = i / F, f = i - t * F;
    const float hv = h1[(size_t)b * N + i];
    const float xh = (hv - mean1) * rstd1;
    const float gy = X[i] * M[t] * dp_mult(dr, 0, (size_t)b * N + i);
    Y[i] = xh;
    G[i] = gy;
    const float dxh = gy * g1w[f];
    s1 += dxh;
    s2 += dxh * xh;
  }
  s1 = block_sum_256(s1, red) / (float)N;
  s2 = block_sum_256(s2, red) / (float)N;
  if (tid < F) {
    float a1 = 0.f, a2 = 0.f;
    for (int t = 0; t < nt; ++t) { a1 += G[t * F + tid] * Y[t * F + tid]; a2 += G[t * F + tid]; }
    atomicAdd(d_g1w + tid, a1); atomicAdd(d_g1b + tid, a2);
  }
  __syncthreads();
  for (int i = tid; i < N; i += 256) {
    const int f = i % F;
    const float dxh = G[i] * g1w[f];
    const float dh = rstd1 * (dxh - s1 - Y[i] * s2);
    const float v = h1[(size_t)b * N + i] > 0.f ? dh : 0.f;
    X[i] = v;
    dpre1[(size_t)b * N + i] = v;
  }
  __syncthreads();
  if (tid < F) {
    float a0 = 0.f;
    for (int t = 0; t < nt; ++t) a0 += X[t * F + tid];
    atomicAdd(d_b1 + tid, a0);
  }
}

// d conv_1.weight[f, c, k] += sum_{b,t} dpre1[b,t,f] * m[b,tt] * E[ids[b,tt] + shift, c],  tt = t + k - pad.
// thread = (c, k); blockIdx.y = a slice of the (b, t) range; F <= 64 accumulators in registers.
template <int FMAX>
__global__ void __launch_bounds__(256) durpred_bwd_w1_kernel(const float* __restrict__ dpre1, const int64_t* __restrict__ ids,
                                                             const float* __restrict__ mask, const float* __restrict__ table,
                                                             float* __restrict__ d_w1, int B, int nt, int Cin, int F, int ksize, int id_shift) {
  const int e = blockIdx.x * 256 + threadIdx.x;  // c * ksize + k
  if (e >= Cin * ksize) return;
  const int c = e / ksize, k = e - c * ksize;
  const int pad = ksize / 2;
  float acc[FMAX];
#pragma unroll
  for (int f = 0; f < FMAX; ++f) acc[f] = 0.f;
  const int total = B * nt;
  const int per = (total + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * per, hi = min(total, lo + per);
  for (int r = lo; r < hi; ++r) {
    const int b = r / nt, t = r - b * nt;
    const int tt = t + k - pad;
    if (tt < 0 || tt >= nt) continue;
    const float m = mask[(size_t)b * nt + tt];
    if (m == 0.f) continue;
    const float ev = table[(size_t)(ids[(size_t)b * nt + tt] + id_shift) * Cin + c] * m;
    const float* g = dpre1 + (size_t)r * F;
#pragma unroll
    for (int f = 0; f < FMAX; ++f)
      if (f < F) acc[f] = fmaf(g[f], ev, acc[f]);
  }
#pragma unroll
  for (int f = 0; f < FMAX; ++f)
    if (f < F) atomicAdd(d_w1 + (size_t)f * Cin * ksize + e, acc[f]);
}

// d text_embed.weight[ids[b,tt] + shift, c] += m[b,tt] * sum_f sum_k W1[f, c, k] * dpre1[b, tt - k + pad, f]   (warp per position)
__global__ void __launch_bounds__(256) durpred_bwd_embed_kernel(const float* __restrict__ dpre1, const int64_t* __restrict__ ids,
                                                                const float* __restrict__ mask, const float* __restrict__ w1,
                                                                float* __restrict__ d_table, int nt, int Cin, int F, int ksize, int id_shift) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int tt = blockIdx.x * 8 + warp;
  if (tt >= nt) return;
  const float m = mask[(size_t)b * nt + tt];
  if (m == 0.f) return;
  const int pad = ksize / 2;
  float* row = d_table + (size_t)(ids[(size_t)b * nt + tt] + id_shift) * Cin;
  for (int c = lane; c < Cin; c += 32) {
    float acc = 0.f;
    for (int k = 0; k < ksize; ++k) {
      const int t = tt - k + pad;
      if (t < 0 || t >= nt) continue;
      const float* g = dpre1 + ((size_t)b * nt + t) * F;
      for (int f = 0; f < F; ++f) acc = fmaf(w1[((size_t)f * Cin + c) * ksize + k], g[f], acc);
    }
    atomicAdd(row + c, acc * m);
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
                                                    nt, F, ksize, 1e-5f, make_dpdrop(0.f, 0), nullptr, nullptr, nullptr);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_duration_predictor_train_forward(const int64_t* ids, int id_shift, const float* mask, const F5bDurPredParams* p, float p_dropout,
                                         uint64_t seed, float* h1_ws, float* a_ws, float* c_ws, float* stats_ws, float* out, int B, int nt,
                                         int Cin, int F, int ksize, f5b_stream_t stream) {
  F5B_CHECK(ids && mask && p && h1_ws && a_ws && c_ws && stats_ws && out && B > 0 && nt > 0 && Cin > 0 && F > 0 && (ksize & 1) &&
                p_dropout >= 0.f && p_dropout < 1.f, "f5b_duration_predictor_train_forward: bad argument");
  F5B_CHECK(((size_t)B * nt * F) % 4 == 0, "f5b_duration_predictor_train_forward: B * nt * filter_channels must be a multiple of 4");
  const size_t smem = 2 * sizeof(float) * (size_t)nt * F;
  F5B_CHECK(smem <= 200 * 1024, "f5b_duration_predictor_train_forward: nt * filter_channels too large for one CTA's shared memory");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 2.0 * B * nt * ((double)Cin * F * ksize + (double)F * F * ksize), 0);
  durpred_conv1_kernel<<<dim3((nt + 7) / 8, B), 256, 0, ST(stream)>>>(ids, mask, p->table, p->conv1_w, p->conv1_b, h1_ws, nt, Cin, F, ksize,
                                                                      id_shift);
  F5B_CUDA(cudaGetLastError());
  F5B_CUDA(cudaFuncSetAttribute(durpred_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  durpred_tail_kernel<<<B, 256, smem, ST(stream)>>>(h1_ws, mask, p->norm1_w, p->norm1_b, p->conv2_w, p->conv2_b, p->norm2_w, p->norm2_b,
                                                    p->proj_w, p->proj_b, out, nt, F, ksize, 1e-5f, make_dpdrop(p_dropout, seed), a_ws, c_ws,
                                                    stats_ws);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

int f5b_duration_predictor_backward(const float* dlogw, const int64_t* ids, int id_shift, const float* mask, const F5bDurPredParams* p,
                                    const F5bDurPredParams* grads, float p_dropout, uint64_t seed, const float* h1_ws, const float* a_ws,
                                    const float* c_ws, const float* stats_ws, float* dpre1_ws, int B, int nt, int Cin, int F, int ksize,
                                    f5b_stream_t stream) {
  F5B_CHECK(dlogw && ids && mask && p && grads && h1_ws && a_ws && c_ws && stats_ws && dpre1_ws && B > 0 && nt > 0 && Cin > 0 && F > 0 &&
                F <= 64 && (ksize & 1), "f5b_duration_predictor_backward: bad argument (filter_channels <= 64)");
  const size_t smem = 3 * sizeof(float) * (size_t)nt * F;
  F5B_CHECK(smem <= 200 * 1024, "f5b_duration_predictor_backward: nt * filter_channels too large for one CTA's shared memory");
  LaunchScope scope(K_ELEMENTWISE, ST(stream), 4.0 * B * nt * ((double)Cin * F * ksize + (double)F * F * ksize), 0);
  auto G = [](const float* q) { return const_cast<float*>(q); };
  F5B_CUDA(cudaFuncSetAttribute(durpred_bwd_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  durpred_bwd_tail_kernel<<<B, 256, smem, ST(stream)>>>(dlogw, mask, h1_ws, a_ws, c_ws, stats_ws, p->norm1_w, p->conv2_w, p->norm2_w,
                                                        p->norm2_b, p->proj_w, G(grads->conv1_b), G(grads->norm1_w), G(grads->norm1_b),
                                                        G(grads->conv2_w), G(grads->conv2_b), G(grads->norm2_w), G(grads->norm2_b),
                                                        G(grads->proj_w), G(grads->proj_b), dpre1_ws, nt, F, ksize,
                                                        make_dpdrop(p_dropout, seed));
  F5B_CUDA(cudaGetLastError());
  const int slices = 32;
  durpred_bwd_w1_kernel<64><<<dim3((Cin * ksize + 255) / 256, slices), 256, 0, ST(stream)>>>(dpre1_ws, ids, mask, p->table, G(grads->conv1_w),
                                                                                            B, nt, Cin, F, ksize, id_shift);
  F5B_CUDA(cudaGetLastError());
  durpred_bwd_embed_kernel<<<dim3((nt + 7) / 8, B), 256, 0, ST(stream)>>>(dpre1_ws, ids, mask, p->conv1_w, G(grads->table), nt, Cin, F, ksize,
                                                                          id_shift);
  F5B_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
