// Developer microbenchmark (not product): (1) correctness + rate of tcgen05.mma with the A operand in TMEM (TS mode);
// (2) effect of a tcgen05.commit after every 4 MMAs; (3) effect of concurrent tcgen05.ld traffic on the MMA rate.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline float aval(int m, int k) { return (float)((m * 7 + k * 3) % 17 - 8); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 5 + k) % 13 - 6); }

// ---- (1) TS-mode correctness: D[128 x N] = A[128 x 64] (TMEM, fp16 pairs) * B[N x 64]^T (smem, K-major SW128)
__global__ void __launch_bounds__(128, 1) ts_check(int N, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < N * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const int chunk = (k / 8) ^ (n & 7);
    reinterpret_cast<__half*>(bptr + n * 128 + chunk * 16)[k % 8] = __float2half(bval(n, k));
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_col = 256;    // A lives at columns 256..287 (64 fp16 = 32 columns)
  const int m = warp * 32 + lane;
  for (int c8 = 0; c8 < 4; ++c8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      const int k = (c8 * 8 + j) * 2;
      __half2 h = __floats2half2_rn(aval(m, k), aval(m, k + 1));
      v[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + a_col + c8 * 8, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(0, N);
      const uint64_t bd = mk_desc(base);
      for (int ks = 0; ks < 4; ++ks) umma_ts(tmem, tmem + a_col + ks * 8, bd + 2 * ks, idesc, ks ? 1u : 0u);
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < N; ++j) out[m * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- (2)/(3) rates.  mode 0: SS, one commit at the end; 1: SS, commit after every 4 MMAs (ring of 8 barriers, never waited);
// 2: SS + 4 extra warps looping tcgen05.ld on another TMEM region; 3: TS (A in TMEM), one commit at the end
__global__ void __launch_bounds__(384, 1) rate(int N, int iters, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[9];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u + (i * 2654435761u & 0x03ff03ffu);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 9; ++i) mbar_init(smem_u32(&bars[i]), 1); mbar_fence_init(); stop = 0; }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_f16(0, N);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 8 || mode == 9) mbar_wait(smem_u32(&bars[7]), 1u);   // fresh barrier, parity 1: returns at once
      if (mode == 7 || mode == 8) tc_fence_after();
      if (elect_one()) {
        const uint32_t aoff = (uint32_t)((it % 9) * 128 * 3);
        const uint64_t bd = mk_desc(base + 64 * 1024 + (uint32_t)((it & 3) * 32 * 1024));
        const uint64_t ad = mk_desc(base + aoff);
        const uint32_t d = tmem + (uint32_t)((it & 1) * (N <= 128 ? N : 0));
        if (mode == 3) {
          const uint32_t ta = tmem + 384 + (uint32_t)((it & 3) * 32);
          umma_ts(d, ta, bd, idesc, 1u);
          umma_ts(d, ta + 8, bd + 2, idesc, 1u);
          umma_ts(d, ta + 16, bd + 4, idesc, 1u);
          umma_ts(d, ta + 24, bd + 6, idesc, 1u);
        } else {
          umma_f16(d, ad, bd, idesc, 1u);
          umma_f16(d, ad + 2, bd + 2, idesc, 1u);
          umma_f16(d, ad + 4, bd + 4, idesc, 1u);
          umma_f16(d, ad + 6, bd + 6, idesc, 1u);
        }
        if (mode == 1) umma_commit(smem_u32(&bars[1 + (it & 7)]));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bars[0]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);
    t1 = clock64();
    stop = 1;
  } else if (warp >= 4 && mode == 2) {
    uint32_t acc = 0;
    while (!stop) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (acc & 64), v);
      tmem_ld_wait();
      acc += v[0] + 32;
    }
    if (acc == 0x12345678u) out[200] = acc;
  } else if (warp >= 2 && mode >= 4 && mode <= 6) {
    // spin on an mbarrier that completes only at the end (bars[8] is never arrived): mode 4 tight try_wait loop (as mbar_wait),
    // mode 5 with __nanosleep(128) between polls, mode 6 try_wait with a suspend-time hint
    const uint32_t bar = smem_u32(&bars[8]);
    uint32_t polls = 0;
    while (!stop) {
      uint32_t ok;
      if (mode == 6) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0u), "r"(2000u) : "memory");
      } else {
        ok = mbar_try_wait(bar, 0) ? 1u : 0u;
        if (mode == 5) __nanosleep(128);
      }
      polls += ok + 1;
    }
    if (polls == 0x12345678u) out[201] = polls;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  float* dout; cudaMalloc(&dout, 128 * 256 * 4);
  cudaFuncSetAttribute(ts_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {16, 64, 256}) {
    cudaMemset(dout, 0, 128 * 256 * 4);
    ts_check<<<1, 128, 40 * 1024>>>(N, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ts_check N=%d error %s\n", N, cudaGetErrorString(e)); return 1; }
    float* h = (float*)malloc(128 * N * 4);
    cudaMemcpy(h, dout, 128 * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0; for (int k = 0; k < 64; ++k) ref += aval(m, k) * bval(n, k);
        const double err = fabs(ref - h[m * N + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, h[m * N + n], ref); ++bad; }
      }
    printf("TS-mode check N=%3d: %s (mismatches %d, max err %g)\n", N, bad ? "FAIL" : "ok", bad, maxerr);
    free(h);
  }
  long long* out; cudaMalloc(&out, 256 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 2000;
  for (int mode : {0, 7, 8, 9})
    for (int N : {16, 64, 128, 256}) {
      if (N == 16) continue;
      rate<<<148, 384, 205 * 1024>>>(N, iters, mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rate mode %d N %d error %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("rate mode %d N %3d: %7.1f cyc/MMA (floor %3d)\n", mode, N, (double)mx / (iters * 4.0), N / 2);
    }
  return 0;
}
