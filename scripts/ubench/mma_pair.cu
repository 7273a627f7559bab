// Developer microbenchmark (not product): tcgen05.mma.cta_group::2 (CTA pair, M = 256) -- operand conventions, the multicast
// commit, the TS-mode tail MMA with N = 16, and the issue rate for N = 64 / 128 / 256.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../pssr2_b200/csrc/common.cuh"
using namespace pssr;

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc2(int n) {   // kind::f16, fp16 inputs, fp32 accumulate, M = 256 (pair)
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void alloc2(uint32_t dst, uint32_t n) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(n) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc2(uint32_t t, uint32_t n) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(n) : "memory");
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__host__ __device__ inline float aval(int m, int k) { return (float)((m * 7 + k * 3) % 17 - 8); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 5 + k) % 13 - 6); }

// mode 0: SS (A in smem of each CTA, its own 128 rows).  mode 1: TS (A in each CTA's TMEM).  B: CTA r holds rows [r*N/2, (r+1)*N/2).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_check(int N, int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_rank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A tile [128 x 64] at offset 0, B half [N/2 x 64] at offset 32 KB, both K-major SWIZZLE_128B
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int m = i / 64, k = i % 64;
    reinterpret_cast<__half*>(sm + m * 128 + (((k / 8) ^ (m & 7)) * 16))[k % 8] = __float2half(aval(m + 128 * rank, k));
  }
  for (int i = threadIdx.x; i < (N / 2) * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    reinterpret_cast<__half*>(sm + 32768 + n * 128 + (((k / 8) ^ (n & 7)) * 16))[k % 8] = __float2half(bval(n + (N / 2) * rank, k));
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) alloc2(smem_u32(&tmem_base_s), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int m = warp * 32 + lane;
  if (mode == 1) {
    for (int c8 = 0; c8 < 4; ++c8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const int k = (c8 * 8 + j) * 2;
        __half2 h = __floats2half2_rn(aval(m + 128 * rank, k), aval(m + 128 * rank, k + 1));
        v[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c8 * 8, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
  }
  if (rank == 0 && warp == 0) {
    if (elect_one()) {
      const uint32_t id = idesc2(N);
      for (int ks = 0; ks < 4; ++ks) {
        if (mode == 0) mma2(tmem, mk_desc(base) + 2 * ks, mk_desc(base + 32768) + 2 * ks, id, ks ? 1u : 0u);
        else mma2_ts(tmem, tmem + 256 + ks * 8, mk_desc(base + 32768) + 2 * ks, id, ks ? 1u : 0u);
      }
      commit2(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < N; ++j) out[(m + 128 * rank) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) { tc_fence_after(); dealloc2(tmem, 512); }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[256 * 256] = __uint_as_float(tmem);
}

// issue rate: the leader issues `iters` groups of T x 4 MMAs (M = 256 per instruction), one commit at the end
template <int N, int T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_rate(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_rank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u + (i * 2654435761u & 0x03ff03ffu);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) alloc2(smem_u32(&tmem_base_s), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t id = idesc2(N);
    const uint64_t a0 = mk_desc(base), b0 = mk_desc(base + 64 * 1024);
    t0 = clock64();
    if (rank == 0) {
      uint32_t stage = 0;
      for (int it = 0; it < iters; ++it) {
        const uint64_t bd = b0 + (uint64_t)(stage * (16 * 1024 / 16));
        const uint64_t ad = a0 + (uint64_t)(stage * 24);
        if (elect_one()) {
#pragma unroll
          for (int mt = 0; mt < T; ++mt)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              asm volatile("tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem + mt * N), "l"(ad + (uint64_t)(mt * 1024 + 2 * k)),
                           "l"(bd + (uint64_t)(2 * k)), "r"(id)
                           : "memory");
        }
        __syncwarp();
        stage = stage == 3 ? 0 : stage + 1;
      }
      if (elect_one()) commit2(smem_u32(&bar));
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) { tc_fence_after(); dealloc2(tmem, 512); }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int N, int T>
void run_rate(long long* out) {
  const int iters = 2000;
  cudaFuncSetAttribute(pair_rate<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  pair_rate<N, T><<<148, 128, 205 * 1024>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("pair_rate N %d T %d error %s\n", N, T, cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("pair rate N %3d T %d: %6.1f cyc per M=256 MMA (1-CTA floor for the same per-SM work: %d)\n", N, T, (double)mx / ((double)iters * 4 * T), N / 2);
}

int main() {
  float* dout; cudaMalloc(&dout, (256 * 256 + 4) * 4);
  cudaFuncSetAttribute(pair_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int mode : {0, 1})
    for (int N : {16, 64, 128, 256}) {
      cudaMemset(dout, 0, (256 * 256 + 4) * 4);
      pair_check<<<2, 128, 70 * 1024>>>(N, mode, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pair_check mode %d N=%d error %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      float* h = (float*)malloc((256 * 256 + 4) * 4);
      cudaMemcpy(h, dout, (256 * 256 + 4) * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 256; ++m)
        for (int n = 0; n < N; ++n) {
          float ref = 0; for (int k = 0; k < 64; ++k) ref += aval(m, k) * bval(n, k);
          if (fabs(ref - h[m * N + n]) > 1e-3) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, h[m * N + n], ref); ++bad; }
        }
      printf("pair check %s N=%3d: %s (mismatches %d)\n", mode ? "TS" : "SS", N, bad ? "FAIL" : "ok", bad);
      free(h);
    }
  long long* out; cudaMalloc(&out, 256 * 8);
  run_rate<64, 1>(out); run_rate<64, 2>(out); run_rate<128, 1>(out); run_rate<128, 2>(out); run_rate<256, 1>(out);
  return 0;
}
