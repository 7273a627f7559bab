// RDNet encoder companions of the tensor-core convolution (config 3, RDResUNet): everything in
// pssr/models/_rdnet.py that is not a dense GEMM.  All are small, HBM/L2-bound CUDA-core kernels on NHWC
// 16-bit activations with fp32 math:
//   stem    : x/128-1 -> BatchNorm(eval) -> PatchifyStem conv (k = stride = patch) -> LayerNorm2d   :106-116
//   ln      : LayerNorm2d of the transition layers (optionally space-to-depth 2x2 so that the 2x2 stride-2
//             transition conv becomes a 1x1 GEMM for the tcgen05 kernel)                             :57-62
//   dwln    : depthwise 7x7 conv + bias + LayerNorm2d (first two layers of Block / BlockESE)          :181-183
//   ese     : EffectiveSEModule (global mean -> 1x1 fc -> hard-sigmoid gate) + layer-scale gamma     :172-174,200-202
// One warp owns one pixel, lanes stride over 8-channel (16-byte) groups, LayerNorm statistics by warp shuffles.
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

static constexpr int kMaxGroupsPerLane = 6;   // channels <= 32 lanes * 6 groups * 8 = 1536

// ------------------------------------------------------------------------------- stem
__global__ void __launch_bounds__(256) stem_kernel(pssr_stem_desc_t d, int fp16) {
  const int Ho = d.H / d.patch, Wo = d.W / d.patch;
  const long long total = (long long)d.B * Ho * Wo;
  const int lane = threadIdx.x & 31;
  const int K = d.C * d.patch * d.patch;
  for (long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; pix < total; pix += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float vals[kMaxGroupsPerLane * 8];
    float s = 0.f;
    int cnt = 0;
    for (int c0 = lane * 8; c0 < d.Cout; c0 += 256, ++cnt) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int co = c0 + j;
        float acc = d.bias[co];
        for (int k = 0; k < K; ++k) {
          const int ci = k / (d.patch * d.patch), rem = k % (d.patch * d.patch);
          const int yy = y * d.patch + rem / d.patch, xx = x * d.patch + rem % d.patch;
          const size_t idx = (((size_t)n * d.C + ci) * d.H + yy) * d.W + xx;
          const float raw = d.x_u8 ? (float)reinterpret_cast<const uint8_t*>(d.x)[idx] : reinterpret_cast<const float*>(d.x)[idx];
          const float v = __fadd_rn(__fmul_rn(__fsub_rn(__fdiv_rn(raw, 128.f), 1.f), d.in_scale[ci]), d.in_shift[ci]);
          acc = fmaf(v, d.weight[(size_t)co * K + k], acc);
        }
        vals[cnt * 8 + j] = acc;
        s += acc;
      }
    }
    const float mean = warp_sum_f(s) / d.Cout;
    float q = 0.f;
    for (int i = 0; i < cnt * 8; ++i) { const float t = vals[i] - mean; q += t * t; }
    const float rstd = rsqrtf(warp_sum_f(q) / d.Cout + d.eps);
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + (size_t)pix * d.out_cstride + d.out_choff;
    cnt = 0;
    for (int c0 = lane * 8; c0 < d.Cout; c0 += 256, ++cnt) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (vals[cnt * 8 + j] - mean) * rstd * d.ln_w[c0 + j] + d.ln_b[c0 + j];
      *reinterpret_cast<uint4*>(out + c0) = pack8(f, fp16);
    }
  }
}

int stem_launch(const pssr_stem_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.Cout % 8 == 0 && d.Cout <= 256 * kMaxGroupsPerLane, PSSR_EUNSUP, "stem: Cout=%d unsupported", d.Cout);
  PSSR_REQUIRE(d.patch >= 1 && d.H % d.patch == 0 && d.W % d.patch == 0, PSSR_EUNSUP, "stem: size not divisible by the patch");
  PSSR_REQUIRE(d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "stem: output alignment");
  const long long total = (long long)d.B * (d.H / d.patch) * (d.W / d.patch);
  long long blocks = (total + 7) / 8;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  stem_kernel<<<(int)blocks, 256, 0, stream>>>(d, dtype == PSSR_DT_FP16);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

// --------------------------------------------------------------------------------- ln
__global__ void __launch_bounds__(256) ln_kernel(pssr_ln_desc_t d, int fp16) {
  const long long total = (long long)d.B * d.H * d.W;
  const int lane = threadIdx.x & 31;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff;
  for (long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; pix < total; pix += ((long long)gridDim.x * blockDim.x) >> 5) {
    const uint16_t* src = in + (size_t)pix * d.in_cstride;
    float vals[kMaxGroupsPerLane * 8];
    float s = 0.f;
    int cnt = 0;
    for (int c0 = lane * 8; c0 < d.C; c0 += 256, ++cnt) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + c0)), f, fp16);
#pragma unroll
      for (int j = 0; j < 8; ++j) { vals[cnt * 8 + j] = f[j]; s += f[j]; }
    }
    const float mean = warp_sum_f(s) / d.C;
    float q = 0.f;
    for (int i = 0; i < cnt * 8; ++i) { const float t = vals[i] - mean; q += t * t; }
    const float rstd = rsqrtf(warp_sum_f(q) / d.C + d.eps);
    size_t opix = (size_t)pix;
    int coff = 0;
    if (d.s2d == 2) {
      const int x = (int)(pix % d.W), y = (int)((pix / d.W) % d.H), n = (int)(pix / ((long long)d.W * d.H));
      opix = ((size_t)n * (d.H / 2) + y / 2) * (d.W / 2) + x / 2;
      coff = ((y & 1) * 2 + (x & 1)) * d.C;
    }
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + opix * d.out_cstride + d.out_choff + coff;
    cnt = 0;
    for (int c0 = lane * 8; c0 < d.C; c0 += 256, ++cnt) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (vals[cnt * 8 + j] - mean) * rstd * d.w[c0 + j] + d.b[c0 + j];
      *reinterpret_cast<uint4*>(out + c0) = pack8(f, fp16);
    }
  }
}

int ln_launch(const pssr_ln_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.C % 8 == 0 && d.C <= 256 * kMaxGroupsPerLane, PSSR_EUNSUP, "layernorm: C=%d unsupported", d.C);
  PSSR_REQUIRE(d.in_cstride % 8 == 0 && d.in_choff % 8 == 0 && d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "layernorm: alignment");
  PSSR_REQUIRE(d.s2d == 1 || (d.s2d == 2 && d.H % 2 == 0 && d.W % 2 == 0), PSSR_EUNSUP, "layernorm: bad space-to-depth factor");
  const long long total = (long long)d.B * d.H * d.W;
  long long blocks = (total + 7) / 8;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ln_kernel<<<(int)blocks, 256, 0, stream>>>(d, dtype == PSSR_DT_FP16);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

// ------------------------------------------------------------------------------- dwln
__global__ void __launch_bounds__(256) dwln_kernel(pssr_dwln_desc_t d, int fp16) {
  const long long total = (long long)d.B * d.H * d.W;
  const int lane = threadIdx.x & 31;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff;
  for (long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; pix < total; pix += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int x = (int)(pix % d.W), y = (int)((pix / d.W) % d.H), n = (int)(pix / ((long long)d.W * d.H));
    float vals[kMaxGroupsPerLane * 8];
    float s = 0.f;
    int cnt = 0;
    for (int c0 = lane * 8; c0 < d.C; c0 += 256, ++cnt) {
      float acc[8];
      {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(d.dw_b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(d.dw_b + c0 + 4));
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
      }
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= d.H) continue;
        for (int kx = 0; kx < 7; ++kx) {
          const int xx = x + kx - 3;
          if (xx < 0 || xx >= d.W) continue;
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(in + (((size_t)n * d.H + yy) * d.W + xx) * d.in_cstride + c0)), f, fp16);
          const float* w = d.dw_w + (size_t)(ky * 7 + kx) * d.C + c0;
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w)), w1 = __ldg(reinterpret_cast<const float4*>(w + 4));
          acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]); acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
          acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]); acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { vals[cnt * 8 + j] = acc[j]; s += acc[j]; }
    }
    const float mean = warp_sum_f(s) / d.C;
    float q = 0.f;
    for (int i = 0; i < cnt * 8; ++i) { const float t = vals[i] - mean; q += t * t; }
    const float rstd = rsqrtf(warp_sum_f(q) / d.C + d.eps);
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + (size_t)pix * d.out_cstride + d.out_choff;
    cnt = 0;
    for (int c0 = lane * 8; c0 < d.C; c0 += 256, ++cnt) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (vals[cnt * 8 + j] - mean) * rstd * d.ln_w[c0 + j] + d.ln_b[c0 + j];
      *reinterpret_cast<uint4*>(out + c0) = pack8(f, fp16);
    }
  }
}

int dwln_launch(const pssr_dwln_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.C % 8 == 0 && d.C <= 256 * kMaxGroupsPerLane, PSSR_EUNSUP, "dwconv: C=%d unsupported", d.C);
  PSSR_REQUIRE(d.in_cstride % 8 == 0 && d.in_choff % 8 == 0 && d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "dwconv: alignment");
  PSSR_REQUIRE(((uintptr_t)d.dw_w & 15) == 0 && ((uintptr_t)d.dw_b & 15) == 0, PSSR_EINVAL, "dwconv: weights misaligned");
  const long long total = (long long)d.B * d.H * d.W;
  long long blocks = (total + 7) / 8;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  dwln_kernel<<<(int)blocks, 256, 0, stream>>>(d, dtype == PSSR_DT_FP16);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

// -------------------------------------------------------------------------------- ese
// gate[b][c] = relu6(fc(mean_yx in[b]) + 3) / 6 * gamma[c]; one CTA per image.
__global__ void __launch_bounds__(256) ese_gate_kernel(pssr_ese_desc_t d, int fp16) {
  extern __shared__ float ese_mean[];
  const int b = blockIdx.x;
  const int HW = d.H * d.W;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + (size_t)b * HW * d.in_cstride;
  for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += unpack1(in[(size_t)p * d.in_cstride + c], fp16);
    ese_mean[c] = s / HW;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
    float a = d.fc_b[c];
    const float* w = d.fc_w + (size_t)c * d.C;
    for (int k = 0; k < d.C; ++k) a = fmaf(w[k], ese_mean[k], a);
    const float g = fminf(fmaxf(a + 3.f, 0.f), 6.f) / 6.f;
    d.gate_ws[(size_t)b * d.C + c] = g * (d.gamma != nullptr ? d.gamma[c] : 1.f);
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
  ese_gate_kernel<<<d.B, 256, d.C * sizeof(float), stream>>>(d, dtype == PSSR_DT_FP16);
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
