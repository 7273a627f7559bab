// HBM-bound companions of the tensor-core convolution (family 2):
//   prep  : x/128-1 -> BatchNorm2d(eval) -> 3x3 im2col of the (few-channel) network input
//           pssr/models/resunet.py:66-70 (the normalised input is also the last skip, :90)
//   pool  : F.max_pool2d(x, 2) on NHWC 16-bit                 pssr/models/resunet.py:76
//   tail  : Reconstruction.conv (3x3, hidden -> out channels) + x*128+128, with the fused
//           clip -> uint8 truncation -> centre channel of `_pred_array`
//           pssr/models/_blocks.py:17, resunet.py:95, pssr/predict.py:245-246
#include <stdlib.h>
#include <cuda_fp8.h>
#include "common.cuh"
#include "plan.h"

namespace pssr {

// ------------------------------------------------------------------------------ prep
// One thread per pixel builds its 64-channel (128 B) im2col row in registers; the warp then transposes its 32 rows through
// shared memory (16-byte chunks XOR-swizzled by pixel: conflict-free both ways) so that every store instruction writes
// 512 contiguous bytes (four whole pixel rows) instead of 32 scattered 16-byte pieces.  Zero padding is applied AFTER
// normalisation (reference quirk: BN acts before the conv's padding), so out-of-image taps are 0.
static constexpr int kPrepThreads = 128;
__global__ void __launch_bounds__(kPrepThreads) prep_im2col_kernel(pssr_prep_desc_t d, int fp16) {
  __shared__ uint4 tr[kPrepThreads / 32][32 * 8];
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)d.B * d.H * d.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool cols16 = d.cols == 16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = blockIdx.x * (long long)blockDim.x + warp * 32; base < total; base += stride) {   // warp-uniform
    const long long i = base + lane;
    uint32_t row[32];
    uint32_t rowl[32];      // compensated precision: rn16(v - rn16(v)), same layout (dead code without im2col_lo)
#pragma unroll
    for (int k = 0; k < 32; ++k) row[k] = rowl[k] = 0u;
    if (i < total) {
      // 32-bit index arithmetic (prep_launch checks the pixel count): three 64-bit divisions cost ~300 instructions per pixel
      const uint32_t ii = (uint32_t)i, rr = ii / (uint32_t)d.W;
      const int x = (int)(ii - rr * (uint32_t)d.W);
      const int n = (int)(rr / (uint32_t)d.H);
      const int y = (int)(rr - (uint32_t)n * (uint32_t)d.H);
      // channel / tap loops fully unrolled: every index into row[] is a compile-time constant, so the im2col row lives in
      // registers (a runtime index put it in local memory: 128 B of local stores + loads per pixel, 59 us for 1 M pixels)
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        if (c < d.C) {
          const float s = d.scale[c], t = d.shift[c];
          const size_t plane = ((size_t)n * d.C + c) * d.H * (size_t)d.W;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
            float v = 0.f;
            if (yy >= 0 && yy < d.H && xx >= 0 && xx < d.W) {
              const size_t idx = plane + (size_t)yy * d.W + xx;
              const float raw = d.x_u8 ? (float)reinterpret_cast<const uint8_t*>(d.x)[idx]
                                       : reinterpret_cast<const float*>(d.x)[idx];
              // no FMA contraction: same bits as the unfused ops (x / 128 == x * 2^-7 exactly)
              v = __fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(raw, 0.0078125f), 1.f), s), t);
              if (tap == 4 && d.xnorm_f32 != nullptr) d.xnorm_f32[idx] = v;
            }
            const uint16_t hi16 = pack1(v, fp16);
            row[(c * 9 + tap) >> 1] |= (uint32_t)hi16 << (16 * ((c * 9 + tap) & 1));
            if (d.im2col_lo != nullptr) rowl[(c * 9 + tap) >> 1] |= (uint32_t)pack1(v - unpack1(hi16, fp16), fp16) << (16 * ((c * 9 + tap) & 1));
          }
        }
      }
    }
    if (cols16) {
      // 32 bytes per pixel: consecutive lanes write consecutive pixels, already coalesced
      if (i < total) {
        uint4* dst16 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.im2col) + (size_t)i * 16);
        dst16[0] = make_uint4(row[0], row[1], row[2], row[3]);
        dst16[1] = make_uint4(row[4], row[5], row[6], row[7]);
        if (d.im2col_lo != nullptr) {
          uint4* dl = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.im2col_lo) + (size_t)i * 16);
          dl[0] = make_uint4(rowl[0], rowl[1], rowl[2], rowl[3]);
          dl[1] = make_uint4(rowl[4], rowl[5], rowl[6], rowl[7]);
        }
      }
      continue;
    }
    uint4* my = tr[warp];
#pragma unroll
    for (int k = 0; k < 8; ++k) my[lane * 8 + (k ^ (lane & 7))] = make_uint4(row[4 * k], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]);
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.im2col) + (size_t)base * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int pix = k * 4 + (lane >> 3), chunk = lane & 7;
      if (base + pix < total) dst[pix * 8 + chunk] = my[pix * 8 + (chunk ^ (pix & 7))];
    }
    __syncwarp();
    if (d.im2col_lo != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) my[lane * 8 + (k ^ (lane & 7))] = make_uint4(rowl[4 * k], rowl[4 * k + 1], rowl[4 * k + 2], rowl[4 * k + 3]);
      __syncwarp();
      uint4* dl = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.im2col_lo) + (size_t)base * 64);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int pix = k * 4 + (lane >> 3), chunk = lane & 7;
        if (base + pix < total) dl[pix * 8 + chunk] = my[pix * 8 + (chunk ^ (pix & 7))];
      }
      __syncwarp();
    }
  }
}

// Inputs with more than 7 channels: the normalised input itself as an NHWC 16-bit tensor (one thread per pixel and 8-channel group)
__global__ void prep_nhwc_kernel(pssr_prep_desc_t d, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = d.cols / 8;
  const long long total = (long long)d.B * d.H * d.W * groups;
  const size_t hw = (size_t)d.H * d.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long pix = i / groups;
    const int n = (int)(pix / (long long)hw);
    const size_t rem = (size_t)(pix - (long long)n * (long long)hw);
    uint32_t o[4] = {0, 0, 0, 0};
    uint32_t ol[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float v = 0.f;
      if (c < d.C) {
        const size_t idx = ((size_t)n * d.C + c) * hw + rem;
        const float raw = d.x_u8 ? (float)reinterpret_cast<const uint8_t*>(d.x)[idx] : reinterpret_cast<const float*>(d.x)[idx];
        v = __fadd_rn(__fmul_rn(__fsub_rn(__fdiv_rn(raw, 128.f), 1.f), d.scale[c]), d.shift[c]);
        if (d.xnorm_f32 != nullptr) d.xnorm_f32[idx] = v;
      }
      const uint16_t hi16 = pack1(v, fp16);
      o[j >> 1] |= (uint32_t)hi16 << (16 * (j & 1));
      ol[j >> 1] |= (uint32_t)pack1(v - unpack1(hi16, fp16), fp16) << (16 * (j & 1));
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.im2col) + (size_t)pix * d.cols + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    if (d.im2col_lo != nullptr)
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.im2col_lo) + (size_t)pix * d.cols + g * 8) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

int prep_launch(const pssr_prep_desc_t& d, int dtype, cudaStream_t stream) {
  if (d.centre_only) {
    PSSR_REQUIRE(d.x && d.im2col && d.scale && d.shift, PSSR_EINVAL, "prep: null pointer");
    PSSR_REQUIRE(d.C >= 1 && d.cols % 8 == 0 && d.cols >= d.C && d.cols <= 64, PSSR_EUNSUP, "prep: NHWC mode needs C <= cols <= 64, cols %% 8 == 0");
    const long long total = (long long)d.B * d.H * d.W * (d.cols / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    PSSR_CHECK_CUDA(launch_pdl(prep_nhwc_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, d, (int)(dtype == PSSR_DT_FP16)));
    count_launch();
    return PSSR_OK;
  }
  PSSR_REQUIRE(d.C >= 1 && d.C * 9 <= 64, PSSR_EUNSUP, "prep: %d input channels unsupported (9*C must be <= 64)", d.C);
  PSSR_REQUIRE(d.x && d.im2col && d.scale && d.shift, PSSR_EINVAL, "prep: null pointer");
  PSSR_REQUIRE(d.cols == 0 || d.cols == 64 || (d.cols == 16 && d.C * 9 <= 16), PSSR_EUNSUP, "prep: im2col width %d unsupported", d.cols);
  const long long total = (long long)d.B * d.H * d.W;
  PSSR_REQUIRE(total < (1ll << 31), PSSR_EUNSUP, "prep: more than 2^31 pixels");
  const int threads = kPrepThreads;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 32;
  if (blocks > cap) blocks = cap;
  PSSR_CHECK_CUDA(launch_pdl(prep_im2col_kernel, dim3((unsigned)blocks), dim3(threads), 0, stream, d, (int)(dtype == PSSR_DT_FP16)));
  count_launch();
  return PSSR_OK;
}

// ------------------------------------------------------------------------------ cast8
// 16-bit NHWC view * scale -> e5m2 NHWC: the operand of the e5m2 correction segments of the compensated precision
// (include/pssr_b200.h PSSR_SEG_E5M2).  One thread per 16 channels: 32 bytes in, 16 bytes out.
__global__ void cast8_kernel(pssr_cast8_desc_t d, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = d.C / 16;
  const long long total = (long long)d.B * d.H * d.W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long pix = i / groups;
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(d.in) + (size_t)pix * d.in_cstride + d.in_choff + g * 16);
    const uint4 a = src[0], b = src[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float f0 = unpack1((uint16_t)(w[2 * j] & 0xffffu), fp16) * d.scale, f1 = unpack1((uint16_t)(w[2 * j] >> 16), fp16) * d.scale;
      const float f2 = unpack1((uint16_t)(w[2 * j + 1] & 0xffffu), fp16) * d.scale, f3 = unpack1((uint16_t)(w[2 * j + 1] >> 16), fp16) * d.scale;
      const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(f0, f1), __NV_SATFINITE, __NV_E5M2);
      const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(f2, f3), __NV_SATFINITE, __NV_E5M2);
      o[j] = lo | (hi << 16);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(d.out) + (size_t)pix * d.out_cstride + d.out_choff + g * 16) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int cast8_launch(const pssr_cast8_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.in && d.out, PSSR_EINVAL, "cast8: null pointer");
  PSSR_REQUIRE(d.C >= 16 && d.C % 16 == 0 && d.in_cstride % 8 == 0 && d.in_choff % 8 == 0 && d.out_cstride % 16 == 0 && d.out_choff % 16 == 0 &&
               ((uintptr_t)d.in & 15) == 0 && ((uintptr_t)d.out & 15) == 0, PSSR_EUNSUP, "cast8: channels must come in aligned groups of 16");
  const long long total = (long long)d.B * d.H * d.W * (d.C / 16);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 32;
  if (blocks > cap) blocks = cap;
  PSSR_CHECK_CUDA(launch_pdl(cast8_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, d, (int)(dtype == PSSR_DT_FP16)));
  count_launch();
  return PSSR_OK;
}

// --------------------------------------------------------------------------- resample
// CUDA-core pieces of the atrous / PSP variants (include/pssr_b200.h PSSR_OP_RESAMPLE): one thread per (output pixel, 8 channels),
// 16-byte loads and stores, fp32 math.
__device__ __forceinline__ void rs_unpack8(const uint4& v, float (&f)[8], int fp16) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = unpack1((uint16_t)(w[k] & 0xFFFFu), fp16);
    f[2 * k + 1] = unpack1((uint16_t)(w[k] >> 16), fp16);
  }
}
__global__ void resample_kernel(pssr_resample_desc_t d, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = d.C / 8;
  const int Ho = (d.mode == 0 || d.mode == 3) ? d.H : d.Ho, Wo = (d.mode == 0 || d.mode == 3) ? d.W : d.Wo;
  const long long total = (long long)d.B * Ho * Wo * groups;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long pix = i / groups;
    const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float f[8];
    if (d.mode == 0) {
      rs_unpack8(__ldg(reinterpret_cast<const uint4*>(in + (size_t)pix * d.in_cstride + g * 8)), f, fp16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (d.scale != nullptr) f[j] = fmaf(f[j], __ldg(d.scale + g * 8 + j), __ldg(d.shift + g * 8 + j));
        if (d.relu == 1) f[j] = fmaxf(f[j], 0.f);
        else if (d.relu == 2) f[j] = f[j] > 0.f ? f[j] : 0.01f * f[j];
      }
    } else if (d.mode == 3) {
      // channel gather: k valid channels from an arbitrarily aligned channel offset, zero-filled up to C (scalar loads)
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g * 8 + j < d.k ? unpack1(in[(size_t)pix * d.in_cstride + g * 8 + j], fp16) : 0.f;
    } else if (d.mode == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = -INFINITY;
      for (int dy = 0; dy < d.k; ++dy)
        for (int dx = 0; dx < d.k; ++dx) {
          float t[8];
          rs_unpack8(__ldg(reinterpret_cast<const uint4*>(in + (((size_t)n * d.H + y * d.k + dy) * d.W + x * d.k + dx) * d.in_cstride + g * 8)), t, fp16);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], t[j]);
        }
    } else {
      // torch upsample_bilinear2d, align_corners=False: scale = in / out, src = max(scale * (dst + 0.5) - 0.5, 0)
      const float sy = (float)d.H / (float)Ho, sx = (float)d.W / (float)Wo;
      const float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.f);
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < d.H - 1 ? 1 : 0), x1 = x0 + (x0 < d.W - 1 ? 1 : 0);
      const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
      float a[8], b[8], c[8], e[8];
      const size_t base = (size_t)n * d.H;
      rs_unpack8(__ldg(reinterpret_cast<const uint4*>(in + ((base + y0) * d.W + x0) * d.in_cstride + g * 8)), a, fp16);
      rs_unpack8(__ldg(reinterpret_cast<const uint4*>(in + ((base + y0) * d.W + x1) * d.in_cstride + g * 8)), b, fp16);
      rs_unpack8(__ldg(reinterpret_cast<const uint4*>(in + ((base + y1) * d.W + x0) * d.in_cstride + g * 8)), c, fp16);
      rs_unpack8(__ldg(reinterpret_cast<const uint4*>(in + ((base + y1) * d.W + x1) * d.in_cstride + g * 8)), e, fp16);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = hy * (hx * a[j] + lx * b[j]) + ly * (hx * c[j] + lx * e[j]);
    }
    const uint4 o = make_uint4(pack2(f[0], f[1], fp16), pack2(f[2], f[3], fp16), pack2(f[4], f[5], fp16), pack2(f[6], f[7], fp16));
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.out) + (size_t)pix * d.out_cstride + d.out_choff + g * 8) = o;
  }
}

int resample_launch(const pssr_resample_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.in != nullptr && d.out != nullptr, PSSR_EINVAL, "resample: null pointer");
  PSSR_REQUIRE(d.C > 0 && d.C % 8 == 0 && (d.mode == 3 || (d.in_cstride % 8 == 0 && d.in_choff % 8 == 0)) && d.out_cstride % 8 == 0 && d.out_choff % 8 == 0,
               PSSR_EUNSUP, "resample: channel counts/strides/offsets must be multiples of 8");
  PSSR_REQUIRE(d.mode >= 0 && d.mode <= 3, PSSR_EINVAL, "resample: mode %d", d.mode);
  PSSR_REQUIRE(d.mode != 3 || (d.k >= 1 && d.k <= d.C), PSSR_EINVAL, "resample: gather of %d channels into %d", d.k, d.C);
  PSSR_REQUIRE(d.mode != 0 || (d.scale == nullptr) == (d.shift == nullptr), PSSR_EINVAL, "resample: scale and shift come together");
  PSSR_REQUIRE(d.mode != 1 || (d.k >= 1 && d.Ho == d.H / d.k && d.Wo == d.W / d.k && d.Ho >= 1 && d.Wo >= 1), PSSR_EINVAL, "resample: pool geometry");
  PSSR_REQUIRE(d.mode != 2 || (d.Ho >= 1 && d.Wo >= 1), PSSR_EINVAL, "resample: output size");
  const bool same = d.mode == 0 || d.mode == 3;
  const long long total = (long long)d.B * (same ? d.H : d.Ho) * (same ? d.W : d.Wo) * (d.C / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  PSSR_CHECK_CUDA(launch_pdl(resample_kernel, dim3((unsigned)blocks), dim3(threads), 0, stream, d, (int)(dtype == PSSR_DT_FP16)));
  count_launch();
  return PSSR_OK;
}

// ------------------------------------------------------------------------------ pool
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b, int fp16) {
  if (fp16) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 max4(uint4 a, uint4 b, int fp16) {
  return make_uint4(max2(a.x, b.x, fp16), max2(a.y, b.y, fp16), max2(a.z, b.z, fp16), max2(a.w, b.w, fp16));
}

// One thread per (output pixel, 8-channel group): 4 x 16-byte loads, 1 x 16-byte store.
__global__ void maxpool2_kernel(pssr_pool_desc_t d, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = d.C / 8;
  const int Ho = d.H / 2, Wo = d.W / 2;
  const long long total = (long long)d.B * Ho * Wo * groups;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in) + d.in_choff;
  uint16_t* out = reinterpret_cast<uint16_t*>(d.out) + d.out_choff;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const uint32_t ii = (uint32_t)i, pix = ii / (uint32_t)groups;      // pool_launch checks total < 2^31
    const int g = (int)(ii - pix * (uint32_t)groups);
    const uint32_t rr = pix / (uint32_t)Wo;
    const int x = (int)(pix - rr * (uint32_t)Wo);
    const int n = (int)(rr / (uint32_t)Ho);
    const int y = (int)(rr - (uint32_t)n * (uint32_t)Ho);
    const size_t p00 = (((size_t)n * d.H + 2 * y) * d.W + 2 * x) * d.in_cstride + g * 8;
    const size_t rowstride = (size_t)d.W * d.in_cstride;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(in + p00));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(in + p00 + d.in_cstride));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(in + p00 + rowstride));
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(in + p00 + rowstride + d.in_cstride));
    const uint4 m = max4(max4(a, b, fp16), max4(c, e, fp16), fp16);
    *reinterpret_cast<uint4*>(out + (size_t)pix * d.out_cstride + g * 8) = m;
  }
}

int pool_launch(const pssr_pool_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.C % 8 == 0 && d.in_cstride % 8 == 0 && d.out_cstride % 8 == 0 && d.in_choff % 8 == 0 &&
                   d.out_choff % 8 == 0,
               PSSR_EUNSUP, "pool: channel counts/strides/offsets must be multiples of 8");
  PSSR_REQUIRE(d.H % 2 == 0 && d.W % 2 == 0, PSSR_EUNSUP, "pool: odd spatial size %dx%d", d.H, d.W);
  const long long total = (long long)d.B * (d.H / 2) * (d.W / 2) * (d.C / 8);
  PSSR_REQUIRE(total < (1ll << 31), PSSR_EUNSUP, "pool: more than 2^31 output vectors");
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  PSSR_CHECK_CUDA(launch_pdl(maxpool2_kernel, dim3((unsigned)blocks), dim3(threads), 0, stream, d, (int)(dtype == PSSR_DT_FP16)));
  count_launch();
  return PSSR_OK;
}

// ------------------------------------------------------------------------------ tail
// 3x3 conv from C (<=128) NHWC 16-bit channels to Cout (<=4) fp32 channels on CUDA cores: the
// layer has N = 1..4 outputs, far below a tensor-core tile, and is bound by reading its input.
// A CTA computes a 32x8 pixel tile from a (34x10) halo tile staged in shared memory; 16-byte
// channel chunks are XOR-swizzled by pixel so that the 8 threads of a quarter-warp hit 8
// different bank groups.
static constexpr int kTailTW = 32, kTailTH = 8, kTailMaxCout = 4;

__global__ void __launch_bounds__(kTailTW* kTailTH) tail_conv_kernel(pssr_tail_desc_t d, int fp16) {
  extern __shared__ uint8_t tail_smem[];
  const int C = d.C;
  const int chunks = C / 8;                 // 16-byte chunks per pixel
  const int HW = kTailTW + 2, HH = kTailTH + 2;
  uint4* tile = reinterpret_cast<uint4*>(tail_smem);                      // [HH*HW][chunks] swizzled
  float* wsm = reinterpret_cast<float*>(tail_smem + (size_t)HH * HW * chunks * 16);  // [Cout][9][C]

  const int tiles_x = (d.W + kTailTW - 1) / kTailTW;
  const int tiles_y = (d.H + kTailTH - 1) / kTailTH;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int trem = blockIdx.x % (tiles_x * tiles_y);
  const int x0 = (trem % tiles_x) * kTailTW, y0 = (trem / tiles_x) * kTailTH;

  for (int i = threadIdx.x; i < d.Cout * 9 * C; i += blockDim.x) wsm[i] = d.weight[i];
  const uint16_t* in = reinterpret_cast<const uint16_t*>(d.in);
  const int mask = chunks >= 8 ? 7 : (chunks - 1);  // chunks is 1,2,4 or a multiple of 8 (checked on host)
  for (int i = threadIdx.x; i < HH * HW * chunks; i += blockDim.x) {
    const int ch = i % chunks;
    const int pp = i / chunks;
    const int yy = y0 + pp / HW - 1, xx = x0 + pp % HW - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < d.H && xx >= 0 && xx < d.W)
      v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)n * d.H + yy) * d.W + xx) * d.cstride + ch * 8));
    tile[pp * chunks + (ch ^ (pp & mask))] = v;
  }
  __syncthreads();

  const int lx = threadIdx.x % kTailTW, ly = threadIdx.x / kTailTW;
  float acc[kTailMaxCout];
#pragma unroll
  for (int o = 0; o < kTailMaxCout; ++o) acc[o] = 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    const int pp = (ly + tap / 3) * HW + lx + tap % 3;
    const uint4* prow = tile + pp * chunks;
    const int sw = pp & mask;
    for (int ch = 0; ch < chunks; ++ch) {
      const uint4 v = prow[ch ^ sw];
      float f[8];
      const uint32_t w32[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f[2 * k] = unpack1((uint16_t)(w32[k] & 0xFFFFu), fp16);
        f[2 * k + 1] = unpack1((uint16_t)(w32[k] >> 16), fp16);
      }
#pragma unroll
      for (int o = 0; o < kTailMaxCout; ++o) {
        if (o < d.Cout) {
          const float* w = wsm + ((size_t)o * 9 + tap) * C + ch * 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[o] = fmaf(f[k], w[k], acc[o]);
        }
      }
    }
  }
  const int x = x0 + lx, y = y0 + ly;
  if (x < d.W && y < d.H) {
#pragma unroll
    for (int o = 0; o < kTailMaxCout; ++o) {
      if (o < d.Cout) {
        const float yv = (acc[o] + d.bias[o]) * d.mul + d.add;
        if (d.out_f32 != nullptr) d.out_f32[(((size_t)n * d.Cout + o) * d.H + y) * d.W + x] = yv;
        if (d.out_u8 != nullptr && o == d.Cout / 2) {
          // np.clip(., 0, 255).astype(np.uint8): truncation toward zero (predict.py:246)
          const float cl = fminf(fmaxf(yv, 0.f), 255.f);
          d.out_u8[((size_t)n * d.H + y) * d.W + x] = (uint8_t)(int)cl;
        }
      }
    }
  }
}

int tail_launch(const pssr_tail_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.Cout >= 1 && d.Cout <= kTailMaxCout, PSSR_EUNSUP, "tail: Cout=%d unsupported (1..4)", d.Cout);
  const int chunks = d.C / 8;
  PSSR_REQUIRE(d.C % 8 == 0 && d.C <= 128 && (chunks == 1 || chunks == 2 || chunks == 4 || chunks % 8 == 0),
               PSSR_EUNSUP, "tail: C=%d unsupported", d.C);
  PSSR_REQUIRE(d.cstride % 8 == 0 && d.cstride >= d.C, PSSR_EUNSUP, "tail: bad channel stride");
  const size_t smem = (size_t)(kTailTW + 2) * (kTailTH + 2) * chunks * 16 + (size_t)d.Cout * 9 * d.C * 4;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    PSSR_CHECK_CUDA(cudaFuncSetAttribute(tail_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  const int tiles = ((d.W + kTailTW - 1) / kTailTW) * ((d.H + kTailTH - 1) / kTailTH);
  tail_conv_kernel<<<d.B * tiles, kTailTW * kTailTH, smem, stream>>>(d, dtype == PSSR_DT_FP16);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

// --------------------------------------------------------------------------- tailsum
// Second half of the fused Reconstruction tail: gather the nine per-tap projections the conv epilogue
// left at LR resolution (z[b][y][s*9+t][x]: the r*r*9 plane rows of one LR row are contiguous, 74 KB at r = 4, W = 128) at their shifted HR positions, add the bias, apply
// x*128+128 (resunet.py:95) and emit fp32 + `_pred_array` uint8.  Pure streaming: ~r^2*9*4 B read and 5 B
// written per LR pixel.
template <bool POW2>
__global__ void __launch_bounds__(256) tailsum_kernel(pssr_tailsum_desc_t d, int lg) {
  const int Hh = d.H * d.r, Wh = d.W * d.r;
  const long long total = (long long)d.B * Hh * Wh;
  const size_t plane = (size_t)d.H * d.W;
  const int planes = d.r * d.r * 9;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % Wh);
    const int Y = (int)((i / Wh) % Hh);
    const int n = (int)(i / ((long long)Wh * Hh));
    const float* zb = d.z + (size_t)n * planes * plane;
    float acc = d.bias;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int Yt = Y + t / 3 - 1, Xt = X + t % 3 - 1;
      if (Yt >= 0 && Yt < Hh && Xt >= 0 && Xt < Wh) {
        const int y = POW2 ? (Yt >> lg) : Yt / d.r, x = POW2 ? (Xt >> lg) : Xt / d.r;
        const int s = (Yt - y * d.r) * d.r + (Xt - x * d.r);
        acc += __ldg(zb + ((size_t)y * planes + (size_t)(s * 9 + t)) * d.W + x);       // z[b][y][s*9+t][x]
      }
    }
    const float yv = acc * d.mul + d.add;
    if (d.out_f32 != nullptr) d.out_f32[i] = yv;
    if (d.out_u8 != nullptr) d.out_u8[i] = (uint8_t)(int)fminf(fmaxf(yv, 0.f), 255.f);
  }
}

// Row-coalesced variant (r = 2, 4, 8): a thread owns the r outputs (Y = r*y + si, X = r*x .. r*x + r-1) of one LR pixel and one
// sub-row; every one of its 9r loads is coalesced along the LR x axis (consecutive threads = consecutive x in one z plane),
// the r outputs leave as one vector store.  Each z value is read exactly once: HBM traffic = |z| + |out|.
template <int R>
__global__ void __launch_bounds__(256) tailsum_rows_kernel(pssr_tailsum_desc_t d) {
  const int Hh = d.H * R, Wh = d.W * R;
  const size_t plane = (size_t)d.H * d.W;
  const long long total = (long long)d.B * d.H * R * d.W;       // (n, y, si, x), x fastest
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % d.W);
    long long rest = i / d.W;
    const int si = (int)(rest % R);
    rest /= R;
    const int y = (int)(rest % d.H);
    const int n = (int)(rest / d.H);
    const float* zb = d.z + (size_t)n * (R * R * 9) * plane;
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = d.bias;
#pragma unroll
    for (int ty = -1; ty <= 1; ++ty) {
      const int sy = si + ty;                                   // source sub-row relative to LR row y
      const int ys = y + (sy < 0 ? -1 : (sy >= R ? 1 : 0));
      const int sis = sy < 0 ? sy + R : (sy >= R ? sy - R : sy);
      if (ys < 0 || ys >= d.H) continue;
#pragma unroll
      for (int c = -1; c <= R; ++c) {                           // source HR column relative to R*x
        const int xs = x + (c < 0 ? -1 : (c >= R ? 1 : 0));
        const int sjs = c < 0 ? c + R : (c >= R ? c - R : c);
        if (xs < 0 || xs >= d.W) continue;
        const float* zp = zb + ((size_t)ys * (R * R * 9) + (size_t)(sis * R + sjs) * 9 + (size_t)(ty + 1) * 3) * d.W + xs;
#pragma unroll
        for (int tx = -1; tx <= 1; ++tx) {
          const int sj = c - tx;                                // the output column this (source, tap) pair belongs to
          if (sj >= 0 && sj < R) acc[sj] += __ldg(zp + (size_t)(tx + 1) * d.W);
        }
      }
    }
    const size_t o = ((size_t)n * Hh + (size_t)(y * R + si)) * Wh + (size_t)x * R;
    float yv[R];
#pragma unroll
    for (int j = 0; j < R; ++j) yv[j] = acc[j] * d.mul + d.add;
    if (d.out_f32 != nullptr) {
      if (R % 4 == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 4) *reinterpret_cast<float4*>(d.out_f32 + o + j) = make_float4(yv[j], yv[j + 1], yv[j + 2], yv[j + 3]);
      } else {
        *reinterpret_cast<float2*>(d.out_f32 + o) = make_float2(yv[0], yv[1]);
      }
    }
    if (d.out_u8 != nullptr) {
      uint8_t b[R];
#pragma unroll
      for (int j = 0; j < R; ++j) b[j] = (uint8_t)(int)fminf(fmaxf(yv[j], 0.f), 255.f);
      if (R % 4 == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 4) *reinterpret_cast<uchar4*>(d.out_u8 + o + j) = make_uchar4(b[j], b[j + 1], b[j + 2], b[j + 3]);
      } else {
        *reinterpret_cast<uchar2*>(d.out_u8 + o) = make_uchar2(b[0], b[1]);
      }
    }
  }
}

// PSSR_TAIL_WINDOW48 (r = 4): z[b][y][e*24 + oi*4 + ojl][x] holds, per LR pixel, the sums of its own projections by HR output
// position: window row oi = 0..5 is HR row 4y + oi - 1, window column ojg = ojl + 2e = 0..5 is HR column 4x + ojg - 1.  A
// thread owns the 4 outputs (4y + si, 4x .. 4x+3): from its own pixel the window row si + 1 (columns 1..4: e = 0 holds
// 0..3, e = 1 holds 2..5), from the pixel left / right the columns 5 / 0, and for si = 0 / 3 the rows 5 / 0 of the pixel
// above / below -- 8 or 16 coalesced loads per thread instead of 36, every z value read exactly once.
__global__ void __launch_bounds__(256) tailsum_win48_kernel(pssr_tailsum_desc_t d) {
  pdl_launch_dependents();
  pdl_wait();
  const int Hh = d.H * 4, Wh = d.W * 4;
  const long long total = (long long)d.B * d.H * 4 * d.W;       // (n, y, si, x), x fastest
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % d.W);
    long long rest = i / d.W;
    const int si = (int)(rest & 3);
    rest >>= 2;
    const int y = (int)(rest % d.H);
    const int n = (int)(rest / d.H);
    float a0 = d.bias, a1 = d.bias, a2 = d.bias, a3 = d.bias;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      // k = 0: own LR row, window row si + 1;  k = 1: the neighbouring LR row whose window reaches this HR row
      int ys = y, oi = si + 1;
      if (k == 1) {
        if (si == 0) { ys = y - 1; oi = 5; }
        else if (si == 3) { ys = y + 1; oi = 0; }
        else continue;
        if (ys < 0 || ys >= d.H) continue;
      }
      const float* zr = d.z + (((size_t)n * d.H + ys) * 48 + (size_t)oi * 4) * d.W + x;     // e = 0, ojl = 0
      const size_t e1 = (size_t)24 * d.W;
      a0 += __ldg(zr + 1 * d.W);
      a1 += __ldg(zr + 2 * d.W) + __ldg(zr + e1);
      a2 += __ldg(zr + 3 * d.W) + __ldg(zr + e1 + d.W);
      a3 += __ldg(zr + e1 + 2 * d.W);
      if (x > 0) a0 += __ldg(zr + e1 + 3 * d.W - 1);          // left pixel, window column 5
      if (x + 1 < d.W) a3 += __ldg(zr + 1);                  // right pixel, window column 0
    }
    const size_t o = ((size_t)n * Hh + (size_t)(y * 4 + si)) * Wh + (size_t)x * 4;
    const float y0 = a0 * d.mul + d.add, y1 = a1 * d.mul + d.add, y2 = a2 * d.mul + d.add, y3 = a3 * d.mul + d.add;
    if (d.out_f32 != nullptr) *reinterpret_cast<float4*>(d.out_f32 + o) = make_float4(y0, y1, y2, y3);
    if (d.out_u8 != nullptr)
      *reinterpret_cast<uchar4*>(d.out_u8 + o) = make_uchar4((uint8_t)(int)fminf(fmaxf(y0, 0.f), 255.f), (uint8_t)(int)fminf(fmaxf(y1, 0.f), 255.f),
                                                             (uint8_t)(int)fminf(fmaxf(y2, 0.f), 255.f), (uint8_t)(int)fminf(fmaxf(y3, 0.f), 255.f));
  }
}

int tailsum_launch(const pssr_tailsum_desc_t& d, cudaStream_t stream) {
  PSSR_REQUIRE(d.z != nullptr && d.B >= 1 && d.H >= 1 && d.W >= 1 && d.r >= 1, PSSR_EINVAL, "tailsum: bad arguments");
  const long long cap = (long long)device_sm_count() * 32;
  const bool aligned = (d.out_f32 == nullptr || ((uintptr_t)d.out_f32 & 15) == 0) && (d.out_u8 == nullptr || ((uintptr_t)d.out_u8 & 3) == 0);
  if (d.layout == PSSR_TAIL_WINDOW48) {
    PSSR_REQUIRE(d.r == 4 && aligned, PSSR_EUNSUP, "tailsum: the window layout needs r = 4 and aligned outputs");
    const long long total = (long long)d.B * d.H * 4 * d.W;
    long long blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    PSSR_CHECK_CUDA(launch_pdl(tailsum_win48_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, d));
    count_launch();
    return PSSR_OK;
  }
  if ((d.r == 2 || d.r == 4 || d.r == 8) && aligned && getenv("PSSR_TAILSUM_V1") == nullptr) {
    const long long total = (long long)d.B * d.H * d.r * d.W;
    long long blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    if (d.r == 2) tailsum_rows_kernel<2><<<(int)blocks, 256, 0, stream>>>(d);
    else if (d.r == 4) tailsum_rows_kernel<4><<<(int)blocks, 256, 0, stream>>>(d);
    else tailsum_rows_kernel<8><<<(int)blocks, 256, 0, stream>>>(d);
    count_launch();
    PSSR_CHECK_CUDA(cudaGetLastError());
    return PSSR_OK;
  }
  const long long total = (long long)d.B * d.H * d.r * d.W * d.r;
  long long blocks = (total + 255) / 256;
  if (blocks > cap) blocks = cap;
  int lg = 0;
  while ((1 << lg) < d.r) ++lg;
  if ((1 << lg) == d.r) tailsum_kernel<true><<<(int)blocks, 256, 0, stream>>>(d, lg);
  else tailsum_kernel<false><<<(int)blocks, 256, 0, stream>>>(d, lg);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

}  // namespace pssr
