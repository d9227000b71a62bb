// tcgen05.mma issue / execution cost per instruction for small shapes (attention-sized MMAs), from one or two issuing warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../eraxvif5tts_b200/csrc -I../../include -o mma_issue mma_issue.cu
#include <cstdio>
#include "common.cuh"
using namespace f5b;
namespace f5b { void set_error(const char*, ...) {} }
// mode 0: SS, mode 1: TS (A from tensor memory).  N = UMMA N.  nw = issuing warps (1 or 2), each issues `cnt` MMAs back to back into its
// own accumulator columns, commits, waits.  Reports clocks from first issue to commit arrival, per MMA.
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(int cnt, int nw, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bars[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&slot, 512);
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
  if (warp < nw) {
    const uint32_t a_addr = smem_u32(smem) + warp * 16384, b_addr = smem_u32(smem) + 32768;
    const uint32_t idesc = idesc_bf16(128, N, 0, 0);
    const uint32_t d = tm + warp * 256;
    long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) {
        if (MODE == 0) umma_bf16(d, smem_desc_sw128(a_addr + (i & 3) * 32, 1024, 16), smem_desc_sw128(b_addr + (i & 3) * 32, 1024, 16), idesc, 1);
        else umma_bf16_ts(d, tm + 128 + warp * 256 + (i & 7) * 8, smem_desc_sw128(b_addr + (i & 3) * 32, 1024, 16), idesc, 1);
      }
      umma_commit(&bars[warp]);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bars[warp], 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
template <int N, int MODE> void run(int nw, long long* out) {
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int cnt = 256;
  k<N, MODE><<<148, 128, 100 * 1024>>>(cnt, nw, out);
  k<N, MODE><<<148, 128, 100 * 1024>>>(cnt, nw, out);
  cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  printf("%s N=%3d warps %d: issue %6.1f clk/MMA, issue->retire %6.1f clk/MMA (nominal %d)", MODE ? "TS" : "SS", N, nw, (double)h[0] / cnt, (double)h[1] / cnt, 128 * N / 256);
  if (nw == 2) printf("   warp1: %6.1f / %6.1f", (double)h[2] / cnt, (double)h[3] / cnt);
  printf("   %s\n", cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  for (int nw = 1; nw <= 2; ++nw) {
    run<64, 0>(nw, out); run<128, 0>(nw, out); run<256, 0>(nw, out);
    run<64, 1>(nw, out); run<128, 1>(nw, out);
  }
  return 0;
}
