// Developer microbenchmark (not product): issue rate of tcgen05.mma kind::f16 M=128, SS mode, for several N and
// operand-address patterns.  One CTA per SM, one warp issues `iters` groups of 4 MMAs (K=64), a single commit at the end.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pssr2_b200/csrc/common.cuh"
using namespace pssr;

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: A and B fixed; 1: A walks 128-byte row shifts (like the 9 taps), B walks stages; 2: like 1 but A step 1024 (aligned)
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int mode, int T, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u * (mode >= 10 ? 0 : 1) + (i * 2654435761u & 0x03ff03ffu);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_f16(0, N);
  const uint32_t a_base = base;              // 64 KB region for A
  const uint32_t b_base = base + 64 * 1024;  // 128 KB region for B (4 stages of 32 KB)
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        const int m = mode % 10;
        const uint32_t aoff = m == 0 ? 0u : (m == 1 ? (uint32_t)((it % 9) * 128 * 3) : (uint32_t)((it % 9) * 1024));
        const uint32_t boff = m == 0 ? 0u : (uint32_t)((it & 3) * 32 * 1024);
        const uint64_t bd = mk_desc(b_base + boff);
        for (int mt = 0; mt < T; ++mt) {
          const uint64_t ad = mk_desc(a_base + aoff + (uint32_t)mt * 16384u);
          const uint32_t d = tmem + (uint32_t)(mt * N);
          umma_f16(d, ad, bd, idesc, 1u);
          umma_f16(d, ad + 2, bd + 2, idesc, 1u);
          umma_f16(d, ad + 4, bd + 4, idesc, 1u);
          umma_f16(d, ad + 6, bd + 6, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 2000;
  for (int grid : {1, 148})
    for (int mode : {0, 1, 2, 10})
      for (int N : {64, 128, 256})
        for (int T : {1, 2}) {
          if (T * N > 512) continue;
          k<<<grid, 128, 205 * 1024>>>(N, iters, mode, T, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          k<<<grid, 128, 205 * 1024>>>(N, iters, mode, T, out);
          cudaEventRecord(e1);
          cudaDeviceSynchronize();
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          const double per = (double)mx / (iters * 4.0 * T);
          const double tf = 2.0 * 128 * N * 16 * iters * 4.0 * T * grid / (ms * 1e-3) / 1e12;
          printf("grid %3d mode %2d N %3d T %d: %7.1f cyc/MMA (floor %3d)  kernel %.3f ms  %7.1f TF/s\n", grid, mode, N, T, per, N / 2, ms, tf);
        }
  return 0;
}
