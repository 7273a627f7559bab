// Developer microbenchmark (not product): how fast can ONE warp issue tcgen05.mma when the loop is lean -- descriptors kept in
// uniform registers, address math outside the elected region, the stage-ready mbarrier poll software-pipelined one stage ahead.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__device__ __forceinline__ void mma1(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

// variant 4: like 2 but the ready poll is a volatile shared-memory flag load (no SYNCS in the issuing warp); 5: like 2 with the
// try_wait for the next stage issued AFTER this stage's MMAs + commit
// variant 0: T tiles x 4 K-steps per iteration, nothing else;  1: + commit per iteration;  2: + commit + pipelined try_wait on a
// completed barrier + fence;  3: like 2 but 3 taps per iteration (G=3: 3 x T x 4 MMAs between barrier operations)
template <int N, int T, int VAR>
__global__ void __launch_bounds__(384, 1) lean(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[10];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile uint32_t flag;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u + (i * 2654435761u & 0x03ff03ffu);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 10; ++i) mbar_init(smem_u32(&bars[i]), 1); mbar_fence_init(); flag = 1u; }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_f16(0, N);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    constexpr int G = VAR == 3 ? 3 : 1;
    const uint64_t a0 = mk_desc(base);
    const uint64_t b0 = mk_desc(base + 64 * 1024);
    const uint32_t bar_ready = smem_u32(&bars[9]);
    uint32_t stage = 0;
    bool ready = (VAR == 2 || VAR == 3 || VAR == 5) ? mbar_try_wait(bar_ready, 1u) : true;
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // uniform address math (outside the elected region)
      const uint64_t bd = b0 + (uint64_t)(stage * (32 * 1024 / 16));
      const uint64_t ad = a0 + (uint64_t)(stage * 24);          // 3 rows further per stage
      const uint32_t cbar = smem_u32(&bars[stage]);
      if (VAR == 2 || VAR == 3) {
        if (!ready) mbar_wait(bar_ready, 1u);
        tc_fence_after();
        ready = mbar_try_wait(bar_ready, 1u);                   // poll for the NEXT stage, consumed next iteration
      }
      if (VAR == 4) {
        while (flag < 1u) {}
        tc_fence_after();
      }
      if (VAR == 5) {
        if (!ready) mbar_wait(bar_ready, 1u);
        tc_fence_after();
      }
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
          for (int mt = 0; mt < T; ++mt) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma1(tmem + mt * N, ad + (uint64_t)(g * 8 + mt * 1024 + 2 * k), bd + (uint64_t)(g * (N * 128 / 16) + 2 * k), idesc);
          }
        }
        if (VAR >= 1) umma_commit(cbar);
      }
      __syncwarp();
      if (VAR == 5) ready = mbar_try_wait(bar_ready, 1u);
      stage = stage == 3 ? 0 : stage + 1;
    }
    if (elect_one()) umma_commit(smem_u32(&bars[8]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[8]), 0);
    t1 = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int N, int T, int VAR>
void run(long long* out) {
  const int iters = 2000;
  cudaFuncSetAttribute(lean<N, T, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  lean<N, T, VAR><<<148, 384, 205 * 1024>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N %d T %d var %d error %s\n", N, T, VAR, cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const int per_it = (VAR == 3 ? 3 : 1) * T * 4;
  printf("lean N %3d T %d var %d: %6.1f cyc/MMA (floor %3d), %6.0f cyc/iteration of %d MMAs\n", N, T, VAR, (double)mx / ((double)iters * per_it), N / 2,
         (double)mx / iters, per_it);
}

int main() {
  long long* out; cudaMalloc(&out, 256 * 8);
  run<64, 2, 2>(out); run<64, 2, 4>(out); run<64, 2, 5>(out);
  run<128, 2, 2>(out); run<128, 2, 4>(out); run<128, 2, 5>(out);
  run<192, 1, 0>(out); run<192, 2, 0>(out);
  run<256, 1, 0>(out); run<256, 1, 1>(out); run<256, 1, 2>(out); run<256, 1, 4>(out); run<256, 1, 5>(out);
  return 0;
}
