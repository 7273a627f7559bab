// SwinIR's window attention (pssr/models/swinir.py:335-373 SwinTransformerBlock.forward, :563-592 WindowAttention.forward) on the
// NHWC token map the 1x1-GEMM qkv projection leaves: cyclic shift, window partition, (q * scale) k^T + relative position bias
// (+ the shifted-window mask, computed from the three row / column regions of calculate_mask :320-341), softmax, @ v, window
// merge and reverse shift -- one CTA per window, fp32 math on CUDA cores.  The projections around it (qkv, proj, fc1, fc2) and
// every convolution of the model run on the tcgen05 kernels.
#include "common.cuh"
#include "plan.h"

namespace pssr {

static constexpr int kWaMaxTokens = 64;   // window_size <= 8
static constexpr int kWaMaxHd = 32;       // head_dim <= 32, even

__global__ void __launch_bounds__(256) winattn_kernel(pssr_winattn_desc_t d, int fp16) {
  extern __shared__ __align__(16) uint8_t wa_sm[];
  uint16_t* qkv_s = reinterpret_cast<uint16_t*>(wa_sm);       // [token][3C]
  const int ws = d.ws, N = ws * ws, C = d.C, C3 = 3 * C;
  const int nwx = d.W / ws, nwy = d.H / ws;
  int wid = blockIdx.x;
  const int wx = wid % nwx; wid /= nwx;
  const int wy = wid % nwy;
  const int b = wid / nwy;
  const uint16_t* qkv = reinterpret_cast<const uint16_t*>(d.qkv);
  // ---- window partition of the cyclically shifted map: token (iy, ix) comes from ((y' + shift) % H, (x' + shift) % W)
  const int vec_per_tok = C3 / 8;
  for (int i = threadIdx.x; i < N * vec_per_tok; i += blockDim.x) {
    const int t = i / vec_per_tok, v = i - t * vec_per_tok;
    const int yo = (wy * ws + t / ws + d.shift) % d.H, xo = (wx * ws + t % ws + d.shift) % d.W;
    reinterpret_cast<uint4*>(qkv_s)[i] = __ldg(reinterpret_cast<const uint4*>(qkv + (((size_t)b * d.H + yo) * d.W + xo) * d.cstride) + v);
  }
  __syncthreads();
  const int hd = C / d.heads;
  for (int item = threadIdx.x; item < d.heads * N; item += blockDim.x) {
    const int h = item / N, i = item - h * N;
    float q[kWaMaxHd];
#pragma unroll
    for (int e = 0; e < kWaMaxHd; ++e) q[e] = e < hd ? unpack1(qkv_s[i * C3 + h * hd + e], fp16) * d.scale : 0.f;
    // region of token i in the shifted map (calculate_mask: three slices per axis)
    int reg_i = 0;
    if (d.shift > 0) {
      const int ys = wy * ws + i / ws, xs = wx * ws + i % ws;
      reg_i = (ys < d.H - ws ? 0 : (ys < d.H - d.shift ? 1 : 2)) * 3 + (xs < d.W - ws ? 0 : (xs < d.W - d.shift ? 1 : 2));
    }
    float s[kWaMaxTokens];
    float mx = -INFINITY;
    const float* bias = d.biasT + ((size_t)h * N) * N + i;
#pragma unroll
    for (int j = 0; j < kWaMaxTokens; ++j) {
      if (j < N) {
        const uint16_t* kr = qkv_s + j * C3 + C + h * hd;
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < kWaMaxHd; e += 2) {
          if (e < hd) {
            const uint32_t kk = *reinterpret_cast<const uint32_t*>(kr + e);
            acc = fmaf(q[e], unpack1((uint16_t)(kk & 0xffffu), fp16), acc);
            acc = fmaf(q[e + 1], unpack1((uint16_t)(kk >> 16), fp16), acc);
          }
        }
        acc += __ldg(bias + (size_t)j * N);
        if (d.shift > 0) {
          const int ys = wy * ws + j / ws, xs = wx * ws + j % ws;
          const int reg_j = (ys < d.H - ws ? 0 : (ys < d.H - d.shift ? 1 : 2)) * 3 + (xs < d.W - ws ? 0 : (xs < d.W - d.shift ? 1 : 2));
          if (reg_j != reg_i) acc += -100.0f;
        }
        s[j] = acc;
        mx = fmaxf(mx, acc);
      } else {
        s[j] = -INFINITY;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kWaMaxTokens; ++j) {
      s[j] = j < N ? __expf(s[j] - mx) : 0.f;
      sum += s[j];
    }
    const float inv = 1.0f / sum;
    float o[kWaMaxHd];
#pragma unroll
    for (int e = 0; e < kWaMaxHd; ++e) o[e] = 0.f;
#pragma unroll
    for (int j = 0; j < kWaMaxTokens; ++j) {
      if (j < N) {
        const uint16_t* vr = qkv_s + j * C3 + 2 * C + h * hd;
        const float pj = s[j] * inv;
#pragma unroll
        for (int e = 0; e < kWaMaxHd; e += 2) {
          if (e < hd) {
            const uint32_t vv = *reinterpret_cast<const uint32_t*>(vr + e);
            o[e] = fmaf(pj, unpack1((uint16_t)(vv & 0xffffu), fp16), o[e]);
            o[e + 1] = fmaf(pj, unpack1((uint16_t)(vv >> 16), fp16), o[e + 1]);
          }
        }
      }
    }
    // window merge + reverse shift: back to where the token came from
    const int yo = (wy * ws + i / ws + d.shift) % d.H, xo = (wx * ws + i % ws + d.shift) % d.W;
    uint16_t* op = reinterpret_cast<uint16_t*>(d.out) + (((size_t)b * d.H + yo) * d.W + xo) * d.out_cstride + d.out_choff + h * hd;
#pragma unroll
    for (int e = 0; e < kWaMaxHd; e += 2)
      if (e < hd) *reinterpret_cast<uint32_t*>(op + e) = pack2(o[e], o[e + 1], fp16);
  }
}

int winattn_launch(const pssr_winattn_desc_t& d, int dtype, cudaStream_t stream) {
  PSSR_REQUIRE(d.qkv != nullptr && d.out != nullptr && d.biasT != nullptr, PSSR_EINVAL, "winattn: null pointer");
  PSSR_REQUIRE(d.ws >= 1 && d.ws * d.ws <= kWaMaxTokens && d.H % d.ws == 0 && d.W % d.ws == 0, PSSR_EUNSUP,
               "winattn: window %d on a %dx%d map (windows of at most 64 tokens that tile the map)", d.ws, d.H, d.W);
  PSSR_REQUIRE(d.heads >= 1 && d.C % d.heads == 0 && (d.C / d.heads) % 2 == 0 && d.C / d.heads <= kWaMaxHd, PSSR_EUNSUP,
               "winattn: head_dim %d must be even and <= 32", d.heads ? d.C / d.heads : 0);
  PSSR_REQUIRE((3 * d.C) % 8 == 0 && d.cstride % 8 == 0 && d.cstride >= 3 * d.C && ((uintptr_t)d.qkv & 15) == 0, PSSR_EUNSUP, "winattn: qkv layout");
  PSSR_REQUIRE(d.out_cstride % 2 == 0 && d.out_choff % 2 == 0 && ((uintptr_t)d.out & 3) == 0, PSSR_EUNSUP, "winattn: output layout");
  PSSR_REQUIRE(d.shift >= 0 && d.shift < d.ws, PSSR_EINVAL, "winattn: shift must lie in [0, window)");
  const size_t smem = (size_t)d.ws * d.ws * 3 * d.C * 2;
  PSSR_REQUIRE(smem <= 160 * 1024, PSSR_EUNSUP, "winattn: window does not fit in shared memory");
  static PerDeviceOnce attr_once;
  if (attr_once.first()) PSSR_CHECK_CUDA(cudaFuncSetAttribute(winattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const long long blocks = (long long)d.B * (d.H / d.ws) * (d.W / d.ws);
  PSSR_REQUIRE(blocks > 0 && blocks < (1ll << 31), PSSR_EUNSUP, "winattn: grid size");
  winattn_kernel<<<(unsigned)blocks, 256, smem, stream>>>(d, dtype == PSSR_DT_FP16);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

}  // namespace pssr
