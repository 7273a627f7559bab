// RDNet encoder companions of the tensor-core convolution (config 3, RDResUNet): everything in
// pssr/models/_rdnet.py that is not a dense GEMM.  All are small, HBM/L2-bound CUDA-core kernels on NHWC
// 16-bit activations with fp32 math:
//   stem    : x/128-1 -> BatchNorm(eval) -> PatchifyStem conv (k = stride = patch) -> LayerNorm2d   :106-116
//   ln      : LayerNorm2d of the transition layers (optionally space-to-depth 2x2 so that the 2x2 stride-2
//             transition conv becomes a 1x1 GEMM for the tcgen05 kernel)                             :57-62
//   dwln    : depthwise 7x7 conv + bias + LayerNorm2d (first two layers of Block / BlockESE)          :181-183
//   ese     : EffectiveSEModule (global mean -> 1x1 fc -> hard-sigmoid gate) + layer-scale gamma     :172-174,200-202
// One warp owns one pixel, lanes stride over 8-channel (16-byte) groups, LayerNorm statistics by warp shuffles.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "plan.h"

namespace pssr {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8], int fp16) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = unpack1((uint16_t)(w[k] & 0xFFFFu), fp16);
    f[2 * k + 1] = unpack1((uint16_t)(w[k] >> 16), fp16);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int fp16) {
  return make_uint4(pack2(f[0], f[1], fp16), pack2(f[2], f[3], fp16), pack2(f[4], f[5], fp16), pack2(f[6], f[7], fp16));
}

// compensated precision: a tensor travels as a (hi, lo) pair of 16-bit buffers of identical layout, value = hi + lo
__device__ __forceinline__ void pack8_pair(const float (&f)[8], int fp16, uint4& hi, uint4& lo) {
  hi = pack8(f, fp16);
  float h[8], r[8];
  unpack8(hi, h, fp16);
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f[j] - h[j];
  lo = pack8(r, fp16);
}
__device__ __forceinline__ void add8(float (&f)[8], const uint4& v, int fp16) {
  float t[8];
  unpack8(v, t, fp16);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] += t[j];
}

static constexpr int kMaxGroupsPerLane = 6;   // channels <= 32 lanes * 6 groups * 8 = 1536

// ------------------------------------------------------------------------------- stem
__global__ void __launch_bounds__(256) stem_kernel(pssr_stem_desc_t d, int fp16) {
  extern __shared__ float stem_sm[];            // [Cout][K] weights, bias, ln_w, ln_b
  const int Ho = d.H / d.patch, Wo = d.W / d.patch;
  const long long total = (long long)d.B * Ho * Wo;
  const int lane = threadIdx.x & 31;
  const int K = d.C * d.patch * d.patch;
  float* w_s = stem_sm;
  float* b_s = w_s + (size_t)d.Cout * K;
  float* lw_s = b_s + d.Cout;
  float* lb_s = lw_s + d.Cout;
  for (int i = threadIdx.x; i < d.Cout * K; i += blockDim.x) w_s[(i % K) * d.Cout + i / K] = d.weight[i];   // [k][c]: 16-byte loads per lane
  for (int i = threadIdx.x; i < d.Cout; i += blockDim.x) { b_s[i] = d.bias[i]; lw_s[i] = d.ln_w[i]; lb_s[i] = d.ln_b[i]; }
  __syncthreads();
  for (long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; pix < total; pix += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    // the K = C*patch*patch normalised inputs of this patch: lane k loads element k, shuffles broadcast them
    float vin = 0.f;
    if (lane < K) {
      const int ci = lane / (d.patch * d.patch), rem = lane % (d.patch * d.patch);
      const int yy = y * d.patch + rem / d.patch, xx = x * d.patch + rem % d.patch;
      const size_t idx = (((size_t)n * d.C + ci) * d.H + yy) * d.W + xx;
      const float raw = d.x_u8 ? (float)reinterpret_cast<const uint8_t*>(d.x)[idx] : reinterpret_cast<const float*>(d.x)[idx];
      vin = __fadd_rn(__fmul_rn(__fsub_rn(__fdiv_rn(raw, 128.f), 1.f), d.in_scale[ci]), d.in_shift[ci]);
    }
    // channel groups unrolled with compile-time indices: vals[] stays in registers (a runtime group counter put it in local
    // memory); groups beyond Cout are predicated off
    float vals[kMaxGroupsPerLane * 8];
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < kMaxGroupsPerLane; ++g) {
      const int c0 = lane * 8 + g * 256;
#pragma unroll
      for (int j = 0; j < 8; ++j) vals[g * 8 + j] = c0 < d.Cout ? b_s[c0 + j] : 0.f;
    }
    for (int k = 0; k < K; ++k) {
      const float v = __shfl_sync(0xffffffffu, vin, k);
#pragma unroll
      for (int g = 0; g < kMaxGroupsPerLane; ++g) {
        const int c0 = lane * 8 + g * 256;
        if (c0 < d.Cout) {
          const float4 w0 = *reinterpret_cast<const float4*>(w_s + (size_t)k * d.Cout + c0), w1 = *reinterpret_cast<const float4*>(w_s + (size_t)k * d.Cout + c0 + 4);
          vals[g * 8 + 0] = fmaf(v, w0.x, vals[g * 8 + 0]); vals[g * 8 + 1] = fmaf(v, w0.y, vals[g * 8 + 1]);
          vals[g * 8 + 2] = fmaf(v, w0.z, vals[g * 8 + 2]); vals[g * 8 + 3] = fmaf(v, w0.w, vals[g * 8 + 3]);
          vals[g * 8 + 4] = fmaf(v, w1.x, vals[g * 8 + 4]); vals[g * 8 + 5] = fmaf(v, w1.y, vals[g * 8 + 5]);
          vals[g * 8 + 6] = fmaf(v, w1.z, vals[g * 8 + 6]); vals[g * 8 + 7] = fmaf(v, w1.w, vals[g * 8 + 7]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kMaxGroupsPerLane; ++g)
      if (lane * 8 + g * 256 < d.Cout) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s += vals[g * 8 + j];
      }
    const float mean = warp_sum_f(s) / d.Cout;
    float q = 0.f;
#pragma unroll
    for (int g = 0; g < kMaxGroupsPerLane; ++g)
      if (lane * 8 + g * 256 < d.Cout) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = vals[g * 8 + j] - mean; q += t * t; }
      }
    const float rstd = rsqrtf(warp_sum_f(q) / d.Cout + d.eps);
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + (size_t)pix * d.out_cstride + d.out_choff;
    uint16_t* out_lo = d.out_lo != nullptr ? reinterpret_cast<uint16_t*>(d.out_lo) + (size_t)pix * d.out_cstride + d.out_choff : nullptr;
#pragma unroll
    for (int g = 0; g < kMaxGroupsPerLane; ++g) {
      const int c0 = lane * 8 + g * 256;
      if (c0 < d.Cout) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (vals[g * 8 + j] - mean) * rstd * lw_s[c0 + j] + lb_s[c0 + j];
        if (out_lo != nullptr) {
          uint4 hi, lo;
          pack8_pair(f, fp16, hi, lo);
          *reinterpret_cast<uint4*>(out + c0) = hi;
          *reinterpret_cast<uint4*>(out_lo + c0) = lo;
        } else {
          *reinterpret_cast<uint4*>(out + c0) = pack8(f, fp16);
        }
      }
    }
  }
}

// Fast path (patch 2 on a single-channel input: K = 4): LANES = Cout / 8 lanes own one pixel (8 channels each, the 4 x 8 filter
// taps, bias and LayerNorm parameters live in registers), so a warp works on 32 / LANES pixels at once and PX of those groups are
// in flight per iteration; the statistics reduce by xor-shuffles inside the LANES-wide group.  The generic kernel below spends a
// whole warp (half of it idle for Cout = 128) and two full shuffle reductions per pixel: 270 us for 52 MB; this one is write-bound.
template <int LANES, int PX>
__global__ void __launch_bounds__(256) stem4_kernel(pssr_stem_desc_t d, int fp16) {
  constexpr int PPW = 32 / LANES;               // pixels per warp and step
  const int Ho = d.H / 2, Wo = d.W / 2;
  const long long total = (long long)d.B * Ho * Wo;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, c0 = (lane % LANES) * 8;
  float w[4][8], bs[8], lw[8], lb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k][j] = __ldg(d.weight + (size_t)(c0 + j) * 4 + k);
    bs[j] = __ldg(d.bias + c0 + j); lw[j] = __ldg(d.ln_w + c0 + j); lb[j] = __ldg(d.ln_b + c0 + j);
  }
  const float sc = __ldg(d.in_scale), sh = __ldg(d.in_shift);
  const float inv_c = 1.0f / (float)(LANES * 8);
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p0 = ((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5) * (PPW * PX); p0 < total; p0 += warps * (PPW * PX)) {
    float raw[PX][4];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      long long pix = p0 + px * PPW + sub;
      if (pix >= total) pix = total - 1;
      // 32-bit index arithmetic (stem_launch checks the pixel count)
      const uint32_t pi = (uint32_t)pix, rr = pi / (uint32_t)Wo;
      const int x = (int)(pi - rr * (uint32_t)Wo), n = (int)(rr / (uint32_t)Ho), y = (int)(rr - (uint32_t)n * (uint32_t)Ho);
      const size_t idx = ((size_t)n * d.H + 2 * y) * d.W + 2 * x;
      if (d.x_u8) {
        const uint8_t* xp = reinterpret_cast<const uint8_t*>(d.x) + idx;
        raw[px][0] = xp[0]; raw[px][1] = xp[1]; raw[px][2] = xp[d.W]; raw[px][3] = xp[d.W + 1];
      } else {
        const float* xp = reinterpret_cast<const float*>(d.x) + idx;
        raw[px][0] = __ldg(xp); raw[px][1] = __ldg(xp + 1); raw[px][2] = __ldg(xp + d.W); raw[px][3] = __ldg(xp + d.W + 1);
      }
    }
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const long long pix = p0 + px * PPW + sub;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = bs[j];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = __fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(raw[px][k], 0.0078125f), 1.f), sc), sh);     // x / 128 == x * 2^-7 exactly
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(a, w[k][j], v[j]);
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[j] -= mean; q = fmaf(v[j], v[j], q); }
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * inv_c + d.eps);
      if (pix < total) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = v[j] * rstd * lw[j] + lb[j];
        uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + (size_t)pix * d.out_cstride + d.out_choff + c0;
        if (d.out_lo != nullptr) {
          uint4 hi, lo;
          pack8_pair(f, fp16, hi, lo);
          *reinterpret_cast<uint4*>(out) = hi;
          *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.out_lo) + (size_t)pix * d.out_cstride + d.out_choff + c0) = lo;
        } else {
          *reinterpret_cast<uint4*>(out) = pack8(f, fp16);
        }
      }
    }
  }
}

int stem_launch(const pssr_stem_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.Cout % 8 == 0 && d.Cout <= 256 * kMaxGroupsPerLane, PSSR_EUNSUP, "stem: Cout=%d unsupported", d.Cout);
  PSSR_REQUIRE(d.patch >= 1 && d.H % d.patch == 0 && d.W % d.patch == 0, PSSR_EUNSUP, "stem: size not divisible by the patch");
  PSSR_REQUIRE(d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "stem: output alignment");
  const long long total = (long long)d.B * (d.H / d.patch) * (d.W / d.patch);
  PSSR_REQUIRE(total < (1ll << 31), PSSR_EUNSUP, "stem: more than 2^31 output pixels");
  long long blocks = (total + 7) / 8;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  const int K = d.C * d.patch * d.patch;
  PSSR_REQUIRE(K <= 32, PSSR_EUNSUP, "stem: C*patch^2 = %d must be <= 32", K);
  if (d.C == 1 && d.patch == 2 && (d.Cout == 64 || d.Cout == 128 || d.Cout == 256) && getenv("PSSR_STEM_GENERIC") == nullptr) {
    const int f16 = dtype == PSSR_DT_FP16;
    const int ppw = 256 / d.Cout * 4;                       // pixels per warp and iteration (PX = 4)
    long long nb = (total + 8LL * ppw - 1) / (8LL * ppw);
    const long long cap4 = (long long)device_sm_count() * 2;      // two resident CTAs per SM (56 parameter registers per thread are loaded once)
    if (nb > cap4) nb = cap4;
    if (d.Cout == 64) stem4_kernel<8, 4><<<(int)nb, 256, 0, stream>>>(d, f16);
    else if (d.Cout == 128) stem4_kernel<16, 4><<<(int)nb, 256, 0, stream>>>(d, f16);
    else stem4_kernel<32, 4><<<(int)nb, 256, 0, stream>>>(d, f16);
    count_launch();
    PSSR_CHECK_CUDA(cudaGetLastError());
    return PSSR_OK;
  }
  const size_t smem = ((size_t)d.Cout * K + 3 * (size_t)d.Cout) * sizeof(float);
  PSSR_REQUIRE(smem <= 48 * 1024, PSSR_EUNSUP, "stem: weights do not fit in shared memory");
  stem_kernel<<<(int)blocks, 256, smem, stream>>>(d, dtype == PSSR_DT_FP16);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

// --------------------------------------------------------------------------------- ln
// G = 256-channel groups per lane (compile time: the per-pixel values stay in registers), PX = pixels a warp handles per
// iteration -- all PX * G 16-byte loads are issued before the first reduction, so a warp keeps PX * C * 2 bytes in flight
// instead of one pixel's (the one-pixel version ran at 0.75 TB/s: latency-bound).  In-place use is safe: a warp reads its
// pixels completely before it writes them.
// MINB = CTAs per SM the register allocation is held to (ncu of the <2, 4> variant at 114 registers: 25 % occupancy, 28 % issue
// slots busy, 13 cycles per issued instruction -- latency-bound; fewer pixels per warp and twice the warps hide it better).
template <int G, int PX, int MINB>
__global__ void __launch_bounds__(256, MINB) ln_kernel(pssr_ln_desc_t d, int fp16) {
  const long long total = (long long)d.B * d.H * d.W;
  const int lane = threadIdx.x & 31;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff;
  const uint16_t* in_lo = d.in_lo != nullptr ? reinterpret_cast<const uint16_t*>(d.in_lo) + d.in_choff : nullptr;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p0 = ((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5) * PX; p0 < total; p0 += warps * PX) {
    float vals[PX][G * 8];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const long long pix = p0 + px;
      const uint16_t* src = in + (size_t)(pix < total ? pix : p0) * d.in_cstride;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c0 = lane * 8 + g * 256;
        if (c0 < d.C) {
          unpack8(__ldg(reinterpret_cast<const uint4*>(src + c0)), *reinterpret_cast<float(*)[8]>(&vals[px][g * 8]), fp16);
          if (in_lo != nullptr)
            add8(*reinterpret_cast<float(*)[8]>(&vals[px][g * 8]), __ldg(reinterpret_cast<const uint4*>(in_lo + (src - in) + c0)), fp16);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) vals[px][g * 8 + j] = 0.f;
        }
      }
    }
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const long long pix = p0 + px;
      if (pix >= total) break;                      // warp-uniform
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < G * 8; ++i) s += vals[px][i];
      const float mean = warp_sum_f(s) / d.C;
      float q = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (lane * 8 + g * 256 < d.C) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float t = vals[px][g * 8 + j] - mean; q += t * t; }
        }
      const float rstd = rsqrtf(warp_sum_f(q) / d.C + d.eps);
      size_t opix = (size_t)pix;
      int coff = 0;
      if (d.s2d == 2) {
        // 32-bit index arithmetic (ln_launch checks B*H*W < 2^31): three 64-bit divisions per pixel cost more than the LayerNorm
        const uint32_t pi = (uint32_t)pix, row = pi / (uint32_t)d.W;
        const int x = (int)(pi - row * (uint32_t)d.W), n = (int)(row / (uint32_t)d.H), y = (int)(row - (uint32_t)n * (uint32_t)d.H);
        opix = ((size_t)n * (d.H / 2) + y / 2) * (d.W / 2) + x / 2;
        coff = ((y & 1) * 2 + (x & 1)) * d.C;
      }
      uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + opix * d.out_cstride + d.out_choff + coff;
      uint16_t* out_lo = d.out_lo != nullptr ? reinterpret_cast<uint16_t*>(d.out_lo) + opix * d.out_cstride + d.out_choff + coff : nullptr;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c0 = lane * 8 + g * 256;
        if (c0 < d.C) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(d.w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(d.w + c0 + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(d.b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(d.b + c0 + 4));
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = (vals[px][g * 8 + j] - mean) * rstd * wv[j] + bv[j];
          if (out_lo != nullptr) {
            uint4 hi, lo;
            pack8_pair(f, fp16, hi, lo);
            *reinterpret_cast<uint4*>(out + c0) = hi;
            *reinterpret_cast<uint4*>(out_lo + c0) = lo;
          } else {
            *reinterpret_cast<uint4*>(out + c0) = pack8(f, fp16);
          }
        }
      }
    }
  }
}

int ln_launch(const pssr_ln_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.C % 8 == 0 && d.C <= 256 * kMaxGroupsPerLane, PSSR_EUNSUP, "layernorm: C=%d unsupported", d.C);
  PSSR_REQUIRE(d.in_cstride % 8 == 0 && d.in_choff % 8 == 0 && d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "layernorm: alignment");
  PSSR_REQUIRE(d.s2d == 1 || (d.s2d == 2 && d.H % 2 == 0 && d.W % 2 == 0), PSSR_EUNSUP, "layernorm: bad space-to-depth factor");
  const long long total = (long long)d.B * d.H * d.W;
  PSSR_REQUIRE(total < (1ll << 31), PSSR_EUNSUP, "layernorm: more than 2^31 pixels");
  long long blocks = (total + 7) / 8;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  const int G = (d.C + 255) / 256;
  const int f16 = dtype == PSSR_DT_FP16;
  switch (G) {
    case 1: ln_kernel<1, 4, 4><<<(int)blocks, 256, 0, stream>>>(d, f16); break;
    case 2: ln_kernel<2, 2, 4><<<(int)blocks, 256, 0, stream>>>(d, f16); break;
    case 3: ln_kernel<3, 2, 3><<<(int)blocks, 256, 0, stream>>>(d, f16); break;
    case 4: ln_kernel<4, 1, 4><<<(int)blocks, 256, 0, stream>>>(d, f16); break;
    case 5: ln_kernel<5, 1, 3><<<(int)blocks, 256, 0, stream>>>(d, f16); break;
    default: ln_kernel<6, 1, 3><<<(int)blocks, 256, 0, stream>>>(d, f16); break;
  }
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

// ------------------------------------------------------------------------------- dwln
// (developer builds only, -DPSSR_DEV_KERNELS: the warp-per-pixel fused depthwise + LayerNorm kernel that measured slower)
#ifdef PSSR_DEV_KERNELS
// One warp produces TWO horizontally adjacent pixels: the 8 input columns x-3..x+4 of a filter row are loaded once and
// feed both outputs, and each weight vector is loaded once per pair.
__global__ void __launch_bounds__(256) dwln_kernel(pssr_dwln_desc_t d, int fp16) {
  const int Wp = (d.W + 1) / 2;
  const long long total = (long long)d.B * d.H * Wp;
  const int lane = threadIdx.x & 31;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff;
  for (long long pp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; pp < total; pp += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int xp = (int)(pp % Wp), y = (int)((pp / Wp) % d.H), n = (int)(pp / ((long long)Wp * d.H));
    const int x0 = 2 * xp;
    const bool has1 = x0 + 1 < d.W;
    float v0[kMaxGroupsPerLane * 8], v1[kMaxGroupsPerLane * 8];
    float s0 = 0.f, s1 = 0.f;
    int cnt = 0;
    for (int c0 = lane * 8; c0 < d.C; c0 += 256, ++cnt) {
      float a0[8], a1[8];
      {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(d.dw_b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(d.dw_b + c0 + 4));
        a0[0] = b0.x; a0[1] = b0.y; a0[2] = b0.z; a0[3] = b0.w; a0[4] = b1.x; a0[5] = b1.y; a0[6] = b1.z; a0[7] = b1.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) a1[j] = a0[j];
      }
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= d.H) continue;
        float wr[7][8];
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float* w = d.dw_w + (size_t)(ky * 7 + kx) * d.C + c0;
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w)), w1 = __ldg(reinterpret_cast<const float4*>(w + 4));
          wr[kx][0] = w0.x; wr[kx][1] = w0.y; wr[kx][2] = w0.z; wr[kx][3] = w0.w; wr[kx][4] = w1.x; wr[kx][5] = w1.y; wr[kx][6] = w1.z; wr[kx][7] = w1.w;
        }
        const uint16_t* rowp = in + (((size_t)n * d.H + yy) * d.W) * d.in_cstride + c0;
#pragma unroll
        for (int cx = 0; cx < 8; ++cx) {           // input column x0 - 3 + cx : tap cx of pixel 0, tap cx-1 of pixel 1
          const int xx = x0 - 3 + cx;
          if (xx < 0 || xx >= d.W) continue;
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(rowp + (size_t)xx * d.in_cstride)), f, fp16);
          if (cx < 7) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a0[j] = fmaf(f[j], wr[cx][j], a0[j]);
          }
          if (cx >= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a1[j] = fmaf(f[j], wr[cx - 1][j], a1[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { v0[cnt * 8 + j] = a0[j]; v1[cnt * 8 + j] = a1[j]; s0 += a0[j]; s1 += a1[j]; }
    }
    const float m0 = warp_sum_f(s0) / d.C, m1 = warp_sum_f(s1) / d.C;
    float q0 = 0.f, q1 = 0.f;
    for (int i = 0; i < cnt * 8; ++i) { const float t0 = v0[i] - m0, t1 = v1[i] - m1; q0 += t0 * t0; q1 += t1 * t1; }
    const float r0 = rsqrtf(warp_sum_f(q0) / d.C + d.eps), r1 = rsqrtf(warp_sum_f(q1) / d.C + d.eps);
    const size_t pix0 = ((size_t)n * d.H + y) * d.W + x0;
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + pix0 * d.out_cstride + d.out_choff;
    cnt = 0;
    for (int c0 = lane * 8; c0 < d.C; c0 += 256, ++cnt) {
      float f[8], g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float lw = d.ln_w[c0 + j], lb = d.ln_b[c0 + j];
        f[j] = (v0[cnt * 8 + j] - m0) * r0 * lw + lb;
        g[j] = (v1[cnt * 8 + j] - m1) * r1 * lw + lb;
      }
      *reinterpret_cast<uint4*>(out + c0) = pack8(f, fp16);
      if (has1) *reinterpret_cast<uint4*>(out + d.out_cstride + c0) = pack8(g, fp16);
    }
  }
}
#endif  // PSSR_DEV_KERNELS

// Tiled depthwise 7x7: a CTA owns an 8x16 pixel tile x 64 channels.  The (14 x 22) halo tile is converted to fp32
// ONCE while it is staged in shared memory ([pixel][64 channels], 256 B per pixel; LO: hi + lo summed there, exact in fp32), so the
// FMA loop has no conversions: a thread computes NPX = 8 consecutive pixels of one row for 4 channels and one filter row costs 14
// input + 7 weight 16-byte shared loads for 224 FMAs (91 % of the issued instructions are FMAs; the first version -- a 16-bit tile,
// 8 channels x 4 pixels per thread -- spent 80 conversions per 224 FMAs and sat at 35 % issue slots busy, latency-bound).
// 256 threads x 2 CTAs per SM.  Output: pre-LayerNorm values, 16-bit NHWC (LO: as a hi + lo pair).
static constexpr int kDwTH = 8, kDwTW = 16, kDwC = 64;
// NPX = pixels of a row per thread (2048 / NPX threads)
static constexpr int kDwTileF32Bytes = (kDwTH + 6) * (kDwTW + 6) * kDwC * 4;
static constexpr int kDwSmemF32 = kDwTileF32Bytes + (49 * kDwC + kDwC) * 4;
template <bool LO, int NPX>
__global__ void __launch_bounds__(2048 / NPX, 2) dwconv7_kernel(pssr_dwln_desc_t d, int fp16) {
  constexpr int kDwThreads = 2048 / NPX;
  constexpr int XQ = kDwTW / NPX;              // threads along a 16-pixel row
  extern __shared__ __align__(16) uint8_t dw_sm[];
  float4* tile = reinterpret_cast<float4*>(dw_sm);                      // [row][col][16 channel quads]
  float* wsm = reinterpret_cast<float*>(dw_sm + kDwTileF32Bytes);       // [49][64]
  float* bsm = wsm + 49 * kDwC;
  const int tiles_x = (d.W + kDwTW - 1) / kDwTW, tiles_y = (d.H + kDwTH - 1) / kDwTH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y; bid /= tiles_y;
  const int slabs = (d.C + kDwC - 1) / kDwC;
  const int sl = bid % slabs;
  const int n = bid / slabs;
  const int c_base = sl * kDwC;
  const int cw = d.C - c_base < kDwC ? d.C - c_base : kDwC;   // channels in this slab (multiple of 8)
  const int x0 = tx * kDwTW, y0 = ty * kDwTH;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff + c_base;
  const uint16_t* in_lo = LO ? reinterpret_cast<const uint16_t*>(d.in_lo) + d.in_choff + c_base : nullptr;
  // staging: every global load of the thread (halo vectors, filter taps) is issued before the first conversion, so the CTA pays
  // one memory latency instead of one per loop iteration (ncu: long-scoreboard stalls led the first version)
  constexpr int kItems = (kDwTH + 6) * (kDwTW + 6) * 8, kIter = (kItems + kDwThreads - 1) / kDwThreads;
  constexpr int kWIter = (49 * kDwC + kDwThreads - 1) / kDwThreads;
  uint4 vh[kIter], vl[kIter];
  float wv[kWIter];
#pragma unroll
  for (int it = 0; it < kIter; ++it) {
    const int i = threadIdx.x + it * kDwThreads;
    const int g = i & 7, pp = i >> 3;
    const int yy = y0 + pp / (kDwTW + 6) - 3, xx = x0 + pp % (kDwTW + 6) - 3;
    vh[it] = make_uint4(0, 0, 0, 0);
    if (LO) vl[it] = make_uint4(0, 0, 0, 0);
    if (i < kItems && yy >= 0 && yy < d.H && xx >= 0 && xx < d.W && g * 8 < cw) {
      const size_t off = (((size_t)n * d.H + yy) * d.W + xx) * d.in_cstride + g * 8;
      vh[it] = __ldg(reinterpret_cast<const uint4*>(in + off));
      if (LO) vl[it] = __ldg(reinterpret_cast<const uint4*>(in_lo + off));
    }
  }
#pragma unroll
  for (int k = 0; k < kWIter; ++k) {
    const int i = threadIdx.x + k * kDwThreads;
    const int t = i / kDwC, c = i % kDwC;
    wv[k] = (i < 49 * kDwC && c < cw) ? __ldg(d.dw_w + (size_t)t * d.C + c_base + c) : 0.f;
  }
  if (threadIdx.x < kDwC) bsm[threadIdx.x] = threadIdx.x < cw ? d.dw_b[c_base + threadIdx.x] : 0.f;
#pragma unroll
  for (int it = 0; it < kIter; ++it) {
    const int i = threadIdx.x + it * kDwThreads;
    if (i < kItems) {
      const int g = i & 7, pp = i >> 3;
      float f[8];
      unpack8(vh[it], f, fp16);
      if (LO) add8(f, vl[it], fp16);
      // the two 16-byte halves of the group go out in swapped order for g >= 4: the eight threads of a quarter-warp then hit
      // eight different 16-byte bank groups in each store (ncu: 3.7-way conflicts with the straight order)
      const int sw = (g >> 2) & 1;
      const float4 a = make_float4(f[0], f[1], f[2], f[3]), b = make_float4(f[4], f[5], f[6], f[7]);
      tile[pp * 16 + g * 2 + sw] = sw ? b : a;
      tile[pp * 16 + g * 2 + 1 - sw] = sw ? a : b;
    }
  }
#pragma unroll
  for (int k = 0; k < kWIter; ++k) {
    const int i = threadIdx.x + k * kDwThreads;
    if (i < 49 * kDwC) wsm[i] = wv[k];
  }
  __syncthreads();
  const int q = threadIdx.x & 15;            // channel quad
  const int xq = (threadIdx.x >> 4) & (XQ - 1);     // which NPX-pixel run of the 16-wide row
  const int row = threadIdx.x / (16 * XQ);          // 0..7
  float acc[NPX][4];
  {
    const float4 bq = *reinterpret_cast<const float4*>(bsm + q * 4);
#pragma unroll
    for (int p4 = 0; p4 < NPX; ++p4) { acc[p4][0] = bq.x; acc[p4][1] = bq.y; acc[p4][2] = bq.z; acc[p4][3] = bq.w; }
  }
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    float4 wr[7];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) wr[kx] = *reinterpret_cast<const float4*>(wsm + (ky * 7 + kx) * kDwC + q * 4);
    const float4* trow = tile + ((row + ky) * (kDwTW + 6) + xq * NPX) * 16 + q;
#pragma unroll
    for (int cx = 0; cx < NPX + 6; ++cx) {   // input column (xq*NPX - 3 + cx): tap (cx - p4) of output pixel p4
      const float4 f = trow[cx * 16];
#pragma unroll
      for (int p4 = 0; p4 < NPX; ++p4) {
        const int kx = cx - p4;
        if (kx >= 0 && kx < 7) {
          acc[p4][0] = fmaf(f.x, wr[kx].x, acc[p4][0]);
          acc[p4][1] = fmaf(f.y, wr[kx].y, acc[p4][1]);
          acc[p4][2] = fmaf(f.z, wr[kx].z, acc[p4][2]);
          acc[p4][3] = fmaf(f.w, wr[kx].w, acc[p4][3]);
        }
      }
    }
  }
  const int y = y0 + row;
  if (y < d.H && q * 4 < cw) {
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + d.out_choff + c_base + q * 4;
#pragma unroll
    for (int p4 = 0; p4 < NPX; ++p4) {
      const int x = x0 + xq * NPX + p4;
      if (x < d.W) {
        const size_t off = (((size_t)n * d.H + y) * d.W + x) * d.out_cstride;
        const uint2 hi = make_uint2(pack2(acc[p4][0], acc[p4][1], fp16), pack2(acc[p4][2], acc[p4][3], fp16));
        *reinterpret_cast<uint2*>(out + off) = hi;
        if (LO) {
          const float h0 = unpack1((uint16_t)(hi.x & 0xFFFFu), fp16), h1 = unpack1((uint16_t)(hi.x >> 16), fp16);
          const float h2 = unpack1((uint16_t)(hi.y & 0xFFFFu), fp16), h3 = unpack1((uint16_t)(hi.y >> 16), fp16);
          *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(d.out_lo) + d.out_choff + c_base + q * 4 + off) =
              make_uint2(pack2(acc[p4][0] - h0, acc[p4][1] - h1, fp16), pack2(acc[p4][2] - h2, acc[p4][3] - h3, fp16));
        }
      }
    }
  }
}

// Pipelined variant for large maps: a CTA owns one 8x16 spatial tile and walks ALL 64-channel slabs of it; the halo tile and the
// filter of slab s+1 arrive by cp.async (zero-filled outside the image / beyond C) while slab s is computed, so the staging
// latency that the one-slab kernel exposes on every CTA is hidden behind the FMAs.
__device__ __forceinline__ void dw_cp16(uint32_t dst, const void* src, bool valid) {
  const int nbytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
static constexpr int kDwTileBytes = (kDwTH + 6) * (kDwTW + 6) * (kDwC / 8) * 16;
static constexpr int kDwBufBytes = kDwTileBytes + (49 * kDwC + kDwC) * 4;

__global__ void __launch_bounds__(256, 2) dwconv7_pipe_kernel(pssr_dwln_desc_t d, int fp16) {
  extern __shared__ __align__(16) uint8_t dw_sm[];
  const int tiles_x = (d.W + kDwTW - 1) / kDwTW, tiles_y = (d.H + kDwTH - 1) / kDwTH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int n = bid / tiles_y;
  const int slabs = (d.C + kDwC - 1) / kDwC;
  const int x0 = tx * kDwTW, y0 = ty * kDwTH;
  const uint16_t* in0 = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff;
  const uint32_t sm0 = smem_u32(dw_sm);

  auto stage = [&](int sl, int buf) {
    const int c_base = sl * kDwC;
    const int cw = d.C - c_base < kDwC ? d.C - c_base : kDwC;
    const uint32_t tile_s = sm0 + (uint32_t)buf * kDwBufBytes;
    const uint32_t w_s = tile_s + kDwTileBytes;
    for (int i = threadIdx.x; i < (kDwTH + 6) * (kDwTW + 6) * 8; i += 256) {
      const int g = i & 7, pp = i >> 3;
      const int yy = y0 + pp / (kDwTW + 6) - 3, xx = x0 + pp % (kDwTW + 6) - 3;
      const bool ok = yy >= 0 && yy < d.H && xx >= 0 && xx < d.W && g * 8 < cw;
      const uint16_t* src = ok ? in0 + (((size_t)n * d.H + yy) * d.W + xx) * d.in_cstride + c_base + g * 8 : in0;
      dw_cp16(tile_s + (uint32_t)i * 16u, src, ok);
    }
    for (int i = threadIdx.x; i < 49 * (kDwC / 4); i += 256) {          // filter taps: 16 chunks of 4 floats per tap
      const int t = i / (kDwC / 4), c4 = (i % (kDwC / 4)) * 4;
      const bool ok = c4 < cw;
      dw_cp16(w_s + (uint32_t)(t * kDwC + c4) * 4u, ok ? d.dw_w + (size_t)t * d.C + c_base + c4 : d.dw_w, ok);
    }
    if (threadIdx.x < kDwC / 4) {
      const int c4 = threadIdx.x * 4;
      const bool ok = c4 < cw;
      dw_cp16(w_s + (uint32_t)(49 * kDwC + c4) * 4u, ok ? d.dw_b + c_base + c4 : d.dw_b, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int g = threadIdx.x & 7;             // channel group
  const int xq = (threadIdx.x >> 3) & 3;     // which 4-pixel quad of the 16-wide row
  const int row = threadIdx.x >> 5;          // 0..7
  stage(0, 0);
  for (int sl = 0; sl < slabs; ++sl) {
    const int buf = sl & 1;
    if (sl + 1 < slabs) {
      stage(sl + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint4* tile = reinterpret_cast<const uint4*>(dw_sm + (size_t)buf * kDwBufBytes);
    const float* wsm = reinterpret_cast<const float*>(dw_sm + (size_t)buf * kDwBufBytes + kDwTileBytes);
    const float* bsm = wsm + 49 * kDwC;
    const int c_base = sl * kDwC;
    const int cw = d.C - c_base < kDwC ? d.C - c_base : kDwC;
    float acc[4][8];
#pragma unroll
    for (int p4 = 0; p4 < 4; ++p4)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p4][j] = bsm[g * 8 + j];
    for (int ky = 0; ky < 7; ++ky) {
      float wr[7][8];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + (ky * 7 + kx) * kDwC + g * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + (ky * 7 + kx) * kDwC + g * 8 + 4);
        wr[kx][0] = w0.x; wr[kx][1] = w0.y; wr[kx][2] = w0.z; wr[kx][3] = w0.w; wr[kx][4] = w1.x; wr[kx][5] = w1.y; wr[kx][6] = w1.z; wr[kx][7] = w1.w;
      }
      const uint4* trow = tile + ((row + ky) * (kDwTW + 6) + xq * 4) * 8 + g;
#pragma unroll
      for (int cx = 0; cx < 10; ++cx) {
        float f[8];
        unpack8(trow[cx * 8], f, fp16);
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const int kx = cx - p4;
          if (kx >= 0 && kx < 7) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p4][j] = fmaf(f[j], wr[kx][j], acc[p4][j]);
          }
        }
      }
    }
    const int y = y0 + row;
    if (y < d.H && g * 8 < cw) {
      uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + d.out_choff + c_base + g * 8;
#pragma unroll
      for (int p4 = 0; p4 < 4; ++p4) {
        const int x = x0 + xq * 4 + p4;
        if (x < d.W) *reinterpret_cast<uint4*>(out + (((size_t)n * d.H + y) * d.W + x) * d.out_cstride) = pack8(acc[p4], fp16);
      }
    }
    __syncthreads();           // every thread is done with buffer `buf` before the stage after next overwrites it
  }
}

int ln_launch(const pssr_ln_desc_t& d, int dtype, cudaStream_t stream);

int dwln_launch(const pssr_dwln_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.C % 8 == 0 && d.C <= 256 * kMaxGroupsPerLane, PSSR_EUNSUP, "dwconv: C=%d unsupported", d.C);
  PSSR_REQUIRE(d.in_cstride % 8 == 0 && d.in_choff % 8 == 0 && d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "dwconv: alignment");
  PSSR_REQUIRE(((uintptr_t)d.dw_w & 15) == 0 && ((uintptr_t)d.dw_b & 15) == 0, PSSR_EINVAL, "dwconv: weights misaligned");
#ifdef PSSR_DEV_KERNELS
  if (getenv("PSSR_DWLN_FUSED") != nullptr) {      // single fused kernel (one warp per pixel pair), kept for comparison
    const long long total = (long long)d.B * d.H * ((d.W + 1) / 2);
    long long blocks = (total + 7) / 8;
    const long long cap = (long long)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    dwln_kernel<<<(int)blocks, 256, 0, stream>>>(d, dtype == PSSR_DT_FP16);
    count_launch();
    PSSR_CHECK_CUDA(cudaGetLastError());
    return PSSR_OK;
  }
#endif
  const long long sp_tiles = (long long)d.B * ((d.H + kDwTH - 1) / kDwTH) * ((d.W + kDwTW - 1) / kDwTW);
  const int slabs_all = (d.C + kDwC - 1) / kDwC;
  // enough spatial tiles to fill the machine twice over and at least two slabs to pipeline: the slab-walking kernel
  const bool lo = d.in_lo != nullptr || d.out_lo != nullptr;
  PSSR_REQUIRE(!lo || (d.in_lo != nullptr && d.out_lo != nullptr), PSSR_EINVAL, "dwconv: in_lo and out_lo come together");
  if (!lo && sp_tiles >= 4LL * device_sm_count() && slabs_all >= 2 && ((uintptr_t)d.dw_w & 15) == 0 && d.C % 4 == 0 && getenv("PSSR_DW_NOPIPE") == nullptr) {
    static PerDeviceOnce attr_pipe;
    if (attr_pipe.first()) PSSR_CHECK_CUDA(cudaFuncSetAttribute(dwconv7_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kDwBufBytes));
    PSSR_REQUIRE(sp_tiles < (1ll << 31), PSSR_EUNSUP, "dwconv: too many blocks");
    dwconv7_pipe_kernel<<<(unsigned)sp_tiles, 256, 2 * kDwBufBytes, stream>>>(d, dtype == PSSR_DT_FP16);
    count_launch();
    PSSR_CHECK_CUDA(cudaGetLastError());
    pssr_ln_desc_t lnp;
    memset(&lnp, 0, sizeof(lnp));
    lnp.in = d.out; lnp.in_cstride = d.out_cstride; lnp.in_choff = d.out_choff; lnp.C = d.C; lnp.B = d.B; lnp.H = d.H; lnp.W = d.W; lnp.s2d = 1;
    lnp.w = d.ln_w; lnp.b = d.ln_b; lnp.eps = d.eps; lnp.out = d.out; lnp.out_cstride = d.out_cstride; lnp.out_choff = d.out_choff;
    return ln_launch(lnp, dtype, stream);
  }
  const long long blocks = (long long)d.B * ((d.C + kDwC - 1) / kDwC) * ((d.H + kDwTH - 1) / kDwTH) * ((d.W + kDwTW - 1) / kDwTW);
  PSSR_REQUIRE(blocks < (1ll << 31), PSSR_EUNSUP, "dwconv: too many blocks");
  const size_t smem = kDwSmemF32;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    PSSR_CHECK_CUDA(cudaFuncSetAttribute(dwconv7_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemF32));
    PSSR_CHECK_CUDA(cudaFuncSetAttribute(dwconv7_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemF32));
  }
  // 8 pixels per thread (256 threads): measured against 4 pixels x 512 threads on the RDResUNet plan, depthwise + LayerNorm ops
  // 1.88 -> 1.71 ms (the 4-pixel kernel stalled on shared-memory issue: MIO throttle + short scoreboard led its ncu stall list)
  const int f16 = dtype == PSSR_DT_FP16;
  if (lo) dwconv7_kernel<true, 8><<<(unsigned)blocks, 256, smem, stream>>>(d, f16);
  else dwconv7_kernel<false, 8><<<(unsigned)blocks, 256, smem, stream>>>(d, f16);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  pssr_ln_desc_t ln;
  memset(&ln, 0, sizeof(ln));
  ln.in = d.out; ln.in_cstride = d.out_cstride; ln.in_choff = d.out_choff; ln.C = d.C; ln.B = d.B; ln.H = d.H; ln.W = d.W; ln.s2d = 1;
  ln.w = d.ln_w; ln.b = d.ln_b; ln.eps = d.eps; ln.out = d.out; ln.out_cstride = d.out_cstride; ln.out_choff = d.out_choff;
  if (lo) { ln.in_lo = d.out_lo; ln.out_lo = d.out_lo; }
  return ln_launch(ln, dtype, stream);     // in place: a warp reads its whole pixel before writing it
}

// -------------------------------------------------------------------------------- ese
// gate[b][c] = relu6(fc(mean_yx in[b]) + 3) / 6 * gamma[c]; one CTA per image.
// One CTA per image.  Phase 1: global mean per channel -- thread = (8-channel group, pixel subset), 16-byte loads, partial sums
// reduced through shared memory in a fixed order (deterministic).  Phase 2: the C x C squeeze-excite fc as one warp per output
// channel (coalesced weight rows, shuffle reduction) -> hard-sigmoid gate x layer scale.
__global__ void __launch_bounds__(256) ese_gate_kernel(pssr_ese_desc_t d, int fp16) {
  extern __shared__ float ese_sm[];          // [C] means, then [nsets][C] partial sums
  float* ese_mean = ese_sm;
  float* part = ese_sm + d.C;
  const int b = blockIdx.x;
  const int HW = d.H * d.W;
  const int groups = d.C / 8;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + (size_t)b * HW * d.in_cstride;
  if (groups <= 256 && 256 % groups == 0) {
    const int nsets = 256 / groups;
    const int g = threadIdx.x % groups, ps = threadIdx.x / groups;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = ps; p < HW; p += nsets) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(in + (size_t)p * d.in_cstride + g * 8)), f, fp16);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[(size_t)ps * d.C + g * 8 + j] = acc[j];
    __syncthreads();
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < nsets; ++k) s += part[(size_t)k * d.C + c];
      ese_mean[c] = s / HW;
    }
  } else {
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
      float s = 0.f;
      for (int p = 0; p < HW; ++p) s += unpack1(in[(size_t)p * d.in_cstride + c], fp16);
      ese_mean[c] = s / HW;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < d.C; c += 8) {
    const float* w = d.fc_w + (size_t)c * d.C;
    float a = 0.f;
    for (int k = lane; k < d.C; k += 32) a = fmaf(__ldg(w + k), ese_mean[k], a);
    a = warp_sum_f(a);
    if (lane == 0) {
      a += d.fc_b[c];
      const float gte = fminf(fmaxf(a + 3.f, 0.f), 6.f) / 6.f;
      d.gate_ws[(size_t)b * d.C + c] = gte * (d.gamma != nullptr ? d.gamma[c] : 1.f);
    }
  }
}

// Small maps (RDNet's eSE blocks sit on the 16^2 / 8^2 stages): ONE launch, one CTA of 512 threads per image.  Every thread keeps
// its PXT 16-byte vectors of the image in registers (phase 1: per-channel sums through shared memory, fixed order), the C x C fc
// runs as TPC = 2^k threads per output channel reading interleaved float4 columns of the weight row (coalesced, every load of the
// phase in flight at once -- the two-kernel version walked C / 8 dependent warp-rounds), and the gated image is written straight
// from the registers: no second pass over the image, no gate buffer.
static constexpr int kEseT = 512;
template <int PXT>
__global__ void __launch_bounds__(kEseT) ese_fused_kernel(pssr_ese_desc_t d, int fp16) {
  extern __shared__ __align__(16) float ese_sm[];          // [C] means, [C] gates, [nsets][C] partial sums
  float* ese_mean = ese_sm;
  float* gate = ese_sm + d.C;
  float* part = ese_sm + 2 * d.C;
  const int b = blockIdx.x;
  const int HW = d.H * d.W;
  const int groups = d.C / 8;
  const int nsets = kEseT / groups;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + (size_t)b * HW * d.in_cstride;
  const int g = threadIdx.x % groups, ps = threadIdx.x / groups;
  const bool own = ps < nsets;
  uint4 v[PXT];
#pragma unroll
  for (int i = 0; i < PXT; ++i) {
    const int p = ps + i * nsets;
    v[i] = own && p < HW ? __ldg(reinterpret_cast<const uint4*>(in + (size_t)p * d.in_cstride + g * 8)) : make_uint4(0, 0, 0, 0);
  }
  if (own) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < PXT; ++i) add8(acc, v[i], fp16);
#pragma unroll
    for (int j = 0; j < 8; ++j) part[(size_t)ps * d.C + g * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d.C; c += kEseT) {
    float s = 0.f;
    for (int k = 0; k < nsets; ++k) s += part[(size_t)k * d.C + c];
    ese_mean[c] = s / HW;
  }
  __syncthreads();
  int tpc = 1;
  while (tpc * 2 * d.C <= kEseT && tpc < 32) tpc *= 2;
  const int oc = threadIdx.x / tpc, ks = threadIdx.x % tpc;      // oc >= C: the thread only takes part in the shuffles
  float a0 = 0.f, a1 = 0.f;
  if (oc < d.C) {
    const float4* wrow = reinterpret_cast<const float4*>(d.fc_w + (size_t)oc * d.C);
    const float4* m4 = reinterpret_cast<const float4*>(ese_mean);
    const int n4 = d.C / 4;
    int j = ks;
    for (; j + tpc < n4; j += 2 * tpc) {
      const float4 w0 = __ldg(wrow + j), w1 = __ldg(wrow + j + tpc);
      const float4 m0 = m4[j], m1 = m4[j + tpc];
      a0 = fmaf(w0.x, m0.x, fmaf(w0.y, m0.y, fmaf(w0.z, m0.z, fmaf(w0.w, m0.w, a0))));
      a1 = fmaf(w1.x, m1.x, fmaf(w1.y, m1.y, fmaf(w1.z, m1.z, fmaf(w1.w, m1.w, a1))));
    }
    if (j < n4) {
      const float4 w0 = __ldg(wrow + j);
      const float4 m0 = m4[j];
      a0 = fmaf(w0.x, m0.x, fmaf(w0.y, m0.y, fmaf(w0.z, m0.z, fmaf(w0.w, m0.w, a0))));
    }
  }
  float a = a0 + a1;
  for (int o = tpc >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (oc < d.C && ks == 0) gate[oc] = fminf(fmaxf(a + d.fc_b[oc] + 3.f, 0.f), 6.f) / 6.f * (d.gamma != nullptr ? d.gamma[oc] : 1.f);
  __syncthreads();
  if (!own) return;
  float gt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gt[j] = gate[g * 8 + j];
  uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + (size_t)b * HW * d.out_cstride + d.out_choff + g * 8;
#pragma unroll
  for (int i = 0; i < PXT; ++i) {
    const int p = ps + i * nsets;
    if (p < HW) {
      float f[8];
      unpack8(v[i], f, fp16);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= gt[j];
      *reinterpret_cast<uint4*>(out + (size_t)p * d.out_cstride) = pack8(f, fp16);
    }
  }
}

__global__ void __launch_bounds__(256) ese_apply_kernel(pssr_ese_desc_t d, int fp16) {
  const int groups = d.C / 8;
  const long long total = (long long)d.B * d.H * d.W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long pix = i / groups;
    const int b = (int)(pix / ((long long)d.H * d.W));
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(d.in) + (size_t)pix * d.in_cstride + g * 8)), f, fp16);
    const float* gate = d.gate_ws + (size_t)b * d.C + g * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= gate[j];
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.out) + (size_t)pix * d.out_cstride + d.out_choff + g * 8) = pack8(f, fp16);
  }
}

int ese_launch(const pssr_ese_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.C % 8 == 0 && d.C <= 4096, PSSR_EUNSUP, "ese: C=%d unsupported", d.C);
  PSSR_REQUIRE(d.in_cstride % 8 == 0 && d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "ese: alignment");
  PSSR_REQUIRE(d.gate_ws != nullptr && d.fc_w != nullptr && d.fc_b != nullptr, PSSR_EINVAL, "ese: null pointer");
  const int groups = d.C / 8;
  const int nsets = (groups > 0 && groups <= 256 && 256 % groups == 0) ? 256 / groups : 0;
  const size_t ese_smem = ((size_t)d.C + (size_t)nsets * d.C) * sizeof(float);
  PSSR_REQUIRE(d.C % 8 == 0 && ese_smem <= 48 * 1024, PSSR_EUNSUP, "ese: C=%d unsupported", d.C);
  // small maps: the single-launch kernel (the image of a CTA fits its threads' registers)
  if (groups <= kEseT && d.C <= kEseT && ((uintptr_t)d.fc_w & 15) == 0 && getenv("PSSR_ESE_TWOPASS") == nullptr) {
    const int fsets = kEseT / groups;
    const int need = (d.H * d.W + fsets - 1) / fsets;
    const size_t sm = ((size_t)2 * d.C + (size_t)fsets * d.C) * sizeof(float);
    if (need <= 8 && sm <= 48 * 1024) {
      const int f16 = dtype == PSSR_DT_FP16;
      if (need <= 1) ese_fused_kernel<1><<<d.B, kEseT, sm, stream>>>(d, f16);
      else if (need <= 2) ese_fused_kernel<2><<<d.B, kEseT, sm, stream>>>(d, f16);
      else if (need <= 4) ese_fused_kernel<4><<<d.B, kEseT, sm, stream>>>(d, f16);
      else ese_fused_kernel<8><<<d.B, kEseT, sm, stream>>>(d, f16);
      count_launch();
      PSSR_CHECK_CUDA(cudaGetLastError());
      return PSSR_OK;
    }
  }
  ese_gate_kernel<<<d.B, 256, ese_smem, stream>>>(d, dtype == PSSR_DT_FP16);
  const long long total = (long long)d.B * d.H * d.W * (d.C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ese_apply_kernel<<<(int)blocks, 256, 0, stream>>>(d, dtype == PSSR_DT_FP16);
  count_launch(2);
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

}  // namespace pssr
