// MUFU.EX2 throughput of the softmax inner loop (scale-subtract FFMA2, ex2, FADD2 row sum, bf16 pack) as a function of the number
// of warps per SM sub-partition and of the fraction of exponentials evaluated by the FMA-pipe polynomial.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f); x.y = fmaxf(x.y, -125.0f);
  const float2 xf = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));
  const float2 nn = __fadd2_rn(xf, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __ffma2_rn(nn, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(make_float2(0.0551716648f, 0.0551716648f), r, make_float2(0.2426111251f, 0.2426111251f));
  p = __ffma2_rn(p, r, make_float2(0.6932609677f, 0.6932609677f));
  p = __ffma2_rn(p, r, make_float2(0.9999280572f, 0.9999280572f));
  float2 e;
  e.x = __int_as_float(__float_as_int(xf.x) * (1 << 23) + __float_as_int(p.x));
  e.y = __int_as_float(__float_as_int(xf.y) * (1 << 23) + __float_as_int(p.y));
  return e;
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// SPIN > 0: one extra warp per sub-partition (the LAST four warps of the block, i.e. the highest warp ids) polls an mbarrier that only
// completes when the compute warps are done — the situation of the MMA / TMA warps of the attention kernel.  SPIN == 2: all 32 lanes
// poll, SPIN == 1: one lane polls.
template <int NPOLY, int MODE, int SPIN>
__global__ void __launch_bounds__(640, 1) k(const float* in, float* out, int iters, long long* clk) {
  extern __shared__ float pad[];
  __shared__ uint64_t bar;
  const int nwarps = blockDim.x / 32;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"((nwarps - 4) * 32));
  }
  __syncthreads();
  if (SPIN > 0 && threadIdx.x / 32 >= nwarps - 4) {
    if (SPIN == 2 || (threadIdx.x & 31) == 0) while (!try_wait(&bar, 0)) {}
    return;
  }
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 4095];
  float2 sum = make_float2(0.f, 0.f);
  uint32_t acc = 0;
  const float2 sc = make_float2(0.18f, 0.18f);
  float2 negm = make_float2(-1.f, -1.f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      float2 e;
      if (MODE == 0) {  // the full inner loop
        const float2 x = __ffma2_rn(make_float2(s[2 * p], s[2 * p + 1]), sc, negm);
        if (((p % 16 + 1) * NPOLY) / 16 != ((p % 16) * NPOLY) / 16) e = exp2_poly2(x);
        else { e.x = ex2(x.x); e.y = ex2(x.y); }
        sum = __fadd2_rn(sum, e);
        acc ^= pack(e.x, e.y);
      } else {  // MUFU only
        e.x = ex2(s[2 * p]); e.y = ex2(s[2 * p + 1]);
        s[2 * p] = e.x * 0.5f; s[2 * p + 1] = e.y * 0.5f;
      }
    }
    negm.x += 1e-3f; negm.y += 1e-3f;
  }
  const long long t1 = clock64();
  if (SPIN > 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
  float r = sum.x + sum.y + __uint_as_float(acc & 0x3fffffff);
  if (MODE == 1) for (int i = 0; i < 64; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NPOLY, int MODE, int SPIN = 0> void run(int warps_per_smsp, const float* in, float* out, long long* clk) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<NPOLY, MODE, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int threads = 128 * (warps_per_smsp + (SPIN > 0 ? 1 : 0));
  k<NPOLY, MODE, SPIN><<<148, threads, 200 * 1024>>>(in, out, 10, clk);
  k<NPOLY, MODE, SPIN><<<148, threads, 200 * 1024>>>(in, out, iters, clk);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  const double exps_per_warp = 64.0 * iters;  // per thread = per warp-instruction count (MUFU + poly)
  const double mufu_frac = MODE == 1 ? 1.0 : 1.0 - NPOLY / 16.0;
  printf("spin %d mode %d poly %d/16 warps/SMSP %d: %6.2f clk per exp (all warps of the SMSP), MUFU busy %5.1f %% (8 clk per warp-MUFU)\n", SPIN, MODE, NPOLY,
         warps_per_smsp, c / (exps_per_warp * warps_per_smsp), 100.0 * 8.0 * exps_per_warp * warps_per_smsp * mufu_frac / c);
}
int main() {
  float *in, *out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -3.0f * (i % 97) / 97.0f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w = 1; w <= 3; ++w) run<0, 0, 1>(w, in, out, clk);
  for (int w = 1; w <= 3; ++w) run<0, 0, 2>(w, in, out, clk);
  for (int w = 1; w <= 4; ++w) run<0, 1>(w, in, out, clk);
  for (int w = 1; w <= 4; ++w) run<0, 0>(w, in, out, clk);
  for (int w = 1; w <= 4; ++w) run<4, 0>(w, in, out, clk);
  for (int w = 1; w <= 4; ++w) run<6, 0>(w, in, out, clk);
  for (int w = 1; w <= 4; ++w) run<8, 0>(w, in, out, clk);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
