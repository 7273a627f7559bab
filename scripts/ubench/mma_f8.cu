// Developer microbenchmark (not product): kind::f8f6f4 (e5m2) MMAs of a CTA pair -- their issue rate alone, and interleaved with
// kind::f16 MMAs in the patterns the compensated Reconstruction.pre issues (is there a cost per kind switch?).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../pssr2_b200/csrc/common.cuh"
using namespace pssr;

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t sbo, uint32_t swz) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)swz << 61;
  return d;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void f16mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
  asm volatile("tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void f16mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id) {
  asm volatile("tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(d), "r"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void f8mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
  asm volatile("tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}

// pattern per iteration: nf16 x (kind::f16, N) then nf8 x (kind::f8f6f4, N8), repeated `rep` times inside the iteration
// mode 0: SS f16;  mode 1: TS f16 (A in TMEM)
template <int N, int N8, int NF16, int NF8, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_rank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u + (i * 2654435761u & 0x03ff03ffu);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t id16 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t id8 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N8 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t a16 = mk_desc(base, 1024, 2), b16 = mk_desc(base + 64 * 1024, 1024, 2);
    const uint64_t a8 = mk_desc(base + 32 * 1024, 512, 4), b8 = mk_desc(base + 128 * 1024, 512, 4);
    t0 = clock64();
    if (rank == 0) {
      for (int it = 0; it < iters; ++it) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < NF16; ++k) {
            if (MODE == 0) f16mma(tmem, a16 + (uint64_t)(2 * (k & 3)), b16 + (uint64_t)(2 * (k & 3)), id16);
            else f16mma_ts(tmem, tmem + 256 + 8 * (k & 3), b16 + (uint64_t)(2 * (k & 3)), id16);
          }
#pragma unroll
          for (int k = 0; k < NF8; ++k) f8mma(tmem, a8 + (uint64_t)(2 * (k & 1)), b8 + (uint64_t)(2 * (k & 1)), id8);
        }
        __syncwarp();
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int N, int N8, int NF16, int NF8, int MODE>
void run(long long* out, const char* what) {
  const int iters = 1000;
  cudaFuncSetAttribute(rate<N, N8, NF16, NF8, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  rate<N, N8, NF16, NF8, MODE><<<148, 128, 205 * 1024>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: error %s\n", what, cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-58s %8.1f cyc / iteration  (%d f16 N=%d %s + %d f8 N=%d)\n", what, (double)mx / iters, NF16, N, MODE ? "TS" : "SS", NF8, N8);
}

int main() {
  long long* out; cudaMalloc(&out, 256 * 8);
  run<256, 256, 12, 0, 0>(out, "12 f16 N=256 (floor 12 x 128 = 1536)");
  run<256, 256, 0, 6, 0>(out, "6 f8 N=256 (floor 6 x 128 = 768)");
  run<256, 256, 12, 6, 0>(out, "12 f16 + 6 f8 N=256 (floor 2304)");
  run<256, 256, 4, 2, 0>(out, "4 f16 + 2 f8 N=256 (floor 768)");
  run<256, 256, 36, 18, 0>(out, "36 f16 + 18 f8 N=256 (floor 6912)");
  run<32, 32, 4, 0, 1>(out, "tail: 4 TS f16 N=32");
  run<32, 16, 0, 2, 1>(out, "tail: 2 f8 N=16");
  run<32, 16, 4, 2, 1>(out, "tail: 4 TS f16 N=32 + 2 f8 N=16");
  run<16, 16, 4, 0, 1>(out, "tail: 4 TS f16 N=16");
  return 0;
}
