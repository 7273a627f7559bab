// Implicit-GEMM convolution on Blackwell tensor cores (tcgen05.mma, accumulators in TMEM,
// operands fetched by TMA).  Replaces every F.conv2d on the reference's network forward:
//   ResBlock convs + BatchNorm(eval) + ReLU + 1x1 respass + residual add   pssr/models/_blocks.py:23-41
//   Reconstruction.pre + ReLU + F.pixel_shuffle                             pssr/models/_blocks.py:15-17
//   torch.cat of decoder inputs / F.pixel_shuffle(x, 2)                     pssr/models/resunet.py:81-85
//   RDNet 1x1 / 2x2-stride-2 convs                                          pssr/models/_rdnet.py:59-62,110-113,183-187
//
// GEMM view: D[M = 128 output pixels, N = output channels] = sum over K blocks of
//   A_kb[128 pixels, 64 channels] * W_kb[N, 64]^T
// A K block is one filter tap (dy,dx) of one 64-channel slice of one NHWC source tensor.  Its A tile
// is ONE 4-D TMA box {64 ch, BW, BH, BNI} (BW*BH*BNI = 128) at (c0, x0+dx, y0+dy, n0): the TMA unit
// zero-fills out-of-image coordinates, which IS the convolution's zero padding, and writes the
// 128-byte-swizzled K-major layout tcgen05.mma reads.  torch.cat never materialises: a concat is
// two sources in the K schedule; the residual 1x1 "respass" is just more K blocks accumulated
// into the same TMEM tile; BatchNorm is folded into weights/bias on the host; pixel_shuffle is
// the epilogue's store address (the packed weights order N as (i*r+j)*C' + c').
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..5 = epilogue (TMEM -> registers -> bias/act -> 16-bit NHWC stores).  Two TMEM
// accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <string.h>
#include "common.cuh"
#include "plan.h"

namespace pssr {

static constexpr int kThreads = 192;
static constexpr int kATileBytes = 128 * 128;  // 128 pixels x 64 ch x 2 B
static constexpr int kMaxStages = 8;

struct ConvKParams {
  const CUtensorMap* tmaps;  // device: [0..2] = sources, [3] = weights
  int n_segs;
  int seg_src[6], seg_taps[6], seg_cblocks[6], seg_dil[6];
  int num_kb;
  int Ho, Wo, B;
  int BW, BH, BNI;
  int tiles_x, tiles_y, tiles_n;
  int n_tiles, total_tiles;
  int block_n, n_valid;
  int num_stages;
  uint32_t stage_bytes;
  const float* bias;
  const float* out_scale;
  uint16_t* out;
  float* out_f32;
  uint16_t* out_lo;
  int lo_cstride, lo_choff;
  const uint16_t* resid;      // y = act(acc + bias + resid_scale * resid[pixel][n]) (shuffle == 1)
  int resid_cstride, resid_choff;
  float resid_scale;
  int out_cstride, out_choff, shuffle, cps, act, fp16;
  int Hout, Wout;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * kMaxStages + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * kMaxStages + 2 + b); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int block_n = p.block_n;

  if (warp == 0) {
    // ================================ TMA producer ==================================
    if (lane == 0) {
      const CUtensorMap* tmB = p.tmaps + kTmW;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const int tx = m_tile % p.tiles_x;
        m_tile /= p.tiles_x;
        const int ty = m_tile % p.tiles_y;
        const int tn = m_tile / p.tiles_y;
        const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = tn * p.BNI;
        int kb = 0;
        for (int sg = 0; sg < p.n_segs; ++sg) {
          const CUtensorMap* tmA = p.tmaps + p.seg_src[sg];
          const int taps = p.seg_taps[sg];
          const int cblocks = p.seg_cblocks[sg];
          for (int t = 0; t < taps; ++t) {
            int dy = 0, dx = 0, cs = 1;
            if (taps == 9) {
              dy = (t / 3 - 1) * p.seg_dil[sg];      // atrous taps: the box moves by the dilation, TMA zero-fills outside the image
              dx = (t % 3 - 1) * p.seg_dil[sg];
            } else if (taps == 4) {
              dy = t >> 1;
              dx = t & 1;
              cs = 2;
            }
            for (int cb = 0; cb < cblocks; ++cb, ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
              mbar_arrive_expect_tx(full_bar(stage), p.stage_bytes);
              tma_load_4d(sa, tmA, full_bar(stage), cb * 64, x0 * cs + dx, y0 * cs + dy, n0);
              tma_load_2d(sa + kATileBytes, tmB, full_bar(stage), kb * 64, n_tile * block_n);
              if (++stage == p.num_stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer ===================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(p.fp16 ? 0 : 1, block_n);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * block_n);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + kATileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in the >>4 field
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                     (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else {
    // ================================== epilogue ====================================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int bw = p.BW, bh = p.BH;
    const int lx = row % bw;
    const int ly = (row / bw) % bh;
    const int ln = row / (bw * bh);
    const int r = p.shuffle;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int n_tile = tile % p.n_tiles;
      int m_tile = tile / p.n_tiles;
      const int tx = m_tile % p.tiles_x;
      m_tile /= p.tiles_x;
      const int ty = m_tile % p.tiles_y;
      const int tn = m_tile / p.tiles_y;
      const int x = tx * bw + lx, y = ty * bh + ly, n = tn * p.BNI + ln;
      const bool valid = (x < p.Wo) && (y < p.Ho) && (n < p.B);
      mbar_wait(tfull_bar(buf), use & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * block_n);
      for (int c0 = 0; c0 < block_n; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) {
          const int nbase = n_tile * block_n + c0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int nn = nbase + g * 8;
            if (nn < p.n_valid) {
              float f[8];
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + nn));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + nn + 4));
              f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x;
              f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
              f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z;
              f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
              f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x;
              f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
              f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z;
              f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
              if (p.resid != nullptr) {
                const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.resid + (((size_t)n * p.Ho + y) * p.Wo + x) * p.resid_cstride + p.resid_choff + nn));
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  f[2 * k] = fmaf(p.resid_scale, unpack1((uint16_t)(rw[k] & 0xffffu), p.fp16), f[2 * k]);
                  f[2 * k + 1] = fmaf(p.resid_scale, unpack1((uint16_t)(rw[k] >> 16), p.fp16), f[2 * k + 1]);
                }
              }
              if (p.act == PSSR_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
              } else if (p.act == PSSR_ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
              }
              if (p.out_scale != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] *= __ldg(p.out_scale + nn + j);
              }
              const int sub = nn / p.cps;
              const int cc = nn - sub * p.cps;
              const int si = sub / r, sj = sub - si * r;
              const size_t pix = ((size_t)n * p.Hout + (size_t)(y * r + si)) * p.Wout + (size_t)(x * r + sj);
              if (p.out != nullptr) {
                uint4 o;
                o.x = pack2(f[0], f[1], p.fp16);
                o.y = pack2(f[2], f[3], p.fp16);
                o.z = pack2(f[4], f[5], p.fp16);
                o.w = pack2(f[6], f[7], p.fp16);
                *reinterpret_cast<uint4*>(p.out + pix * p.out_cstride + p.out_choff + cc) = o;
                if (p.out_lo != nullptr) {      // compensated precision: what the 16-bit rounding dropped
                  uint4 l;
                  l.x = pack2(f[0] - unpack1((uint16_t)(o.x & 0xffffu), p.fp16), f[1] - unpack1((uint16_t)(o.x >> 16), p.fp16), p.fp16);
                  l.y = pack2(f[2] - unpack1((uint16_t)(o.y & 0xffffu), p.fp16), f[3] - unpack1((uint16_t)(o.y >> 16), p.fp16), p.fp16);
                  l.z = pack2(f[4] - unpack1((uint16_t)(o.z & 0xffffu), p.fp16), f[5] - unpack1((uint16_t)(o.z >> 16), p.fp16), p.fp16);
                  l.w = pack2(f[6] - unpack1((uint16_t)(o.w & 0xffffu), p.fp16), f[7] - unpack1((uint16_t)(o.w >> 16), p.fp16), p.fp16);
                  *reinterpret_cast<uint4*>(p.out_lo + pix * p.lo_cstride + p.lo_choff + cc) = l;
                }
              }
              if (p.out_f32 != nullptr) {
                float4* d = reinterpret_cast<float4*>(p.out_f32 + pix * p.out_cstride + p.out_choff + cc);
                d[0] = make_float4(f[0], f[1], f[2], f[3]);
                d[1] = make_float4(f[4], f[5], f[6], f[7]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// --------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

static int floor_pow2(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

int conv_prepare(const pssr_conv_desc_t& d, int dtype, ConvOp& op) {
  EncodeTiledFn enc = get_encode_fn();
  PSSR_REQUIRE(enc != nullptr, PSSR_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  PSSR_REQUIRE(d.n_srcs >= 1 && d.n_srcs <= 4 && d.n_segs >= 1 && d.n_segs <= 6, PSSR_EINVAL,
               "conv: n_srcs/n_segs out of range");
  PSSR_REQUIRE(d.n >= 32 && d.n % 32 == 0, PSSR_EUNSUP, "conv: n=%d must be a positive multiple of 32", d.n);
  PSSR_REQUIRE(d.n_valid > 0 && d.n_valid <= d.n && d.n_valid % 8 == 0, PSSR_EUNSUP,
               "conv: n_valid=%d must be a multiple of 8 and <= n", d.n_valid);
  PSSR_REQUIRE(d.shuffle >= 1, PSSR_EINVAL, "conv: shuffle must be >= 1");
  PSSR_REQUIRE(d.n_valid % (d.shuffle * d.shuffle) == 0, PSSR_EUNSUP, "conv: n_valid %% shuffle^2 != 0");
  const int cps = d.n_valid / (d.shuffle * d.shuffle);
  PSSR_REQUIRE(cps % 8 == 0, PSSR_EUNSUP, "conv: channels after pixel shuffle (%d) must be a multiple of 8", cps);
  PSSR_REQUIRE(d.shuffle == 1 || d.n == d.n_valid, PSSR_EUNSUP, "conv: padded N with pixel shuffle unsupported");
  PSSR_REQUIRE(d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP,
               "conv: output channel stride/offset must be multiples of 8");
  PSSR_REQUIRE(d.tail_z == nullptr, PSSR_EUNSUP, "conv: the fused Reconstruction tail needs the v3 kernel (64 channels per sub-position, N %% 256 == 0)");
  PSSR_REQUIRE(d.out != nullptr || d.out_f32 != nullptr, PSSR_EINVAL, "conv: no output buffer");
  // (with a pixel shuffle the residual is indexed by GEMM column, i.e. it must come from a GEMM packed with the same N permutation)
  PSSR_REQUIRE(d.resid == nullptr || (d.resid_cstride % 8 == 0 && d.resid_choff % 8 == 0 && ((uintptr_t)d.resid & 15) == 0),
               PSSR_EUNSUP, "conv: the epilogue residual needs 16-byte aligned channel slices");

  ConvKParams& p = *reinterpret_cast<ConvKParams*>(op.kparams);
  static_assert(sizeof(ConvKParams) <= sizeof(op.kparams), "ConvOp::kparams too small");
  memset(&p, 0, sizeof(p));
  memset(op.tmaps, 0, sizeof(op.tmaps));

  // block_n: largest of 256/128/64/32 dividing n
  int block_n = 256;
  while (d.n % block_n != 0) block_n >>= 1;
  p.block_n = block_n;
  p.n_tiles = d.n / block_n;
  p.n_valid = d.n_valid;

  // pixel tile
  p.Ho = d.Ho;
  p.Wo = d.Wo;
  p.B = d.B;
  p.BW = floor_pow2(d.Wo < 16 ? d.Wo : 16);
  int bh_max = 128 / p.BW;
  p.BH = floor_pow2(d.Ho < bh_max ? d.Ho : bh_max);
  p.BNI = 128 / (p.BW * p.BH);
  p.tiles_x = (d.Wo + p.BW - 1) / p.BW;
  p.tiles_y = (d.Ho + p.BH - 1) / p.BH;
  p.tiles_n = (d.B + p.BNI - 1) / p.BNI;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;

  const CUtensorMapDataType tdt = dtype == PSSR_DT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int num_kb = 0;
  bool src_stride2[4] = {false, false, false, false};
  p.n_segs = d.n_segs;
  for (int s = 0; s < d.n_segs; ++s) {
    const pssr_kseg_t& sg = d.segs[s];
    PSSR_REQUIRE(sg.src >= 0 && sg.src < d.n_srcs, PSSR_EINVAL, "conv: segment source index out of range");
    PSSR_REQUIRE(sg.taps == 1 || sg.taps == 9 || sg.taps == 4, PSSR_EUNSUP, "conv: taps must be 1, 4 or 9");
    PSSR_REQUIRE(sg.fmt == PSSR_SEG_F16, PSSR_EUNSUP, "conv: e5m2 segments need the rows-mode kernel (3x3, width %% 128 == 0)");
    PSSR_REQUIRE(sg.cblocks >= 1, PSSR_EINVAL, "conv: cblocks must be >= 1");
    p.seg_src[s] = sg.src;
    p.seg_taps[s] = sg.taps;
    p.seg_cblocks[s] = sg.cblocks;
    p.seg_dil[s] = sg.dilation > 1 ? sg.dilation : 1;
    PSSR_REQUIRE(sg.dilation <= 1 || sg.taps == 9, PSSR_EINVAL, "conv: dilation applies to 3x3 segments");
    num_kb += sg.taps * sg.cblocks;
    if (sg.taps == 4) src_stride2[sg.src] = true;
  }
  p.num_kb = num_kb;

  for (int s = 0; s < d.n_srcs; ++s) {
    const pssr_src_t& src = d.srcs[s];
    PSSR_REQUIRE(src.base != nullptr && ((uintptr_t)src.base & 15) == 0, PSSR_EINVAL,
                 "conv: source %d base must be 16-byte aligned", s);
    PSSR_REQUIRE(src.cstride % 8 == 0 && src.channels >= 1 && src.channels <= src.cstride, PSSR_EUNSUP,
                 "conv: source %d channel stride must be a multiple of 8 and >= channels", s);
    const int st = src_stride2[s] ? 2 : 1;
    PSSR_REQUIRE(src.H == d.Ho * st && src.W == d.Wo * st && src.B == d.B, PSSR_EINVAL,
                 "conv: source %d geometry %dx%dx%d does not match output %dx%dx%d (stride %d)", s, src.B,
                 src.H, src.W, d.B, d.Ho, d.Wo, st);
    cuuint64_t gdim[4] = {(cuuint64_t)src.channels, (cuuint64_t)src.W, (cuuint64_t)src.H, (cuuint64_t)src.B};
    cuuint64_t gstr[3] = {(cuuint64_t)src.cstride * 2, (cuuint64_t)src.cstride * 2 * src.W,
                          (cuuint64_t)src.cstride * 2 * src.W * src.H};
    // with element strides the box is given in traversed elements; the tile holds ceil(box/stride)
    cuuint32_t box[4] = {64, (cuuint32_t)(p.BW * st - (st - 1)), (cuuint32_t)(p.BH * st - (st - 1)), (cuuint32_t)p.BNI};
    cuuint32_t estr[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
    CUresult r = enc(&op.tmaps[s], tdt, 4, const_cast<void*>(src.base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(source %d) failed with %d", s, (int)r);
  }
  {
    PSSR_REQUIRE(d.weights != nullptr && ((uintptr_t)d.weights & 15) == 0, PSSR_EINVAL, "conv: weights misaligned");
    const cuuint64_t ktot = (cuuint64_t)num_kb * 64;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)d.n};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&op.tmaps[kTmW], tdt, 2, const_cast<void*>(d.weights), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  }

  p.stage_bytes = (uint32_t)(kATileBytes + block_n * 128);
  const int smem_budget = 200 * 1024;
  int stages = smem_budget / (int)p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > num_kb) stages = num_kb < 2 ? 2 : num_kb;
  p.num_stages = stages;
  op.smem_bytes = stages * (int)p.stage_bytes + 1024;

  p.bias = d.bias;
  PSSR_REQUIRE(d.bias != nullptr && ((uintptr_t)d.bias & 15) == 0, PSSR_EINVAL, "conv: bias missing or misaligned");
  p.out_scale = d.out_scale;
  p.out = reinterpret_cast<uint16_t*>(d.out);
  p.out_f32 = d.out_f32;
  p.out_lo = reinterpret_cast<uint16_t*>(d.out_lo);
  p.lo_cstride = d.out_lo_cstride;
  p.lo_choff = d.out_lo_choff;
  PSSR_REQUIRE(d.out_lo == nullptr || (d.out != nullptr && d.out_lo_cstride % 8 == 0 && d.out_lo_choff % 8 == 0), PSSR_EUNSUP, "conv: out_lo needs a 16-bit primary output");
  p.resid = reinterpret_cast<const uint16_t*>(d.resid);
  p.resid_cstride = d.resid_cstride;
  p.resid_choff = d.resid_choff;
  p.resid_scale = d.resid_scale;
  p.out_cstride = d.out_cstride;
  p.out_choff = d.out_choff;
  p.shuffle = d.shuffle;
  p.cps = cps;
  p.act = d.act;
  p.fp16 = dtype == PSSR_DT_FP16 ? 1 : 0;
  p.Hout = d.Ho * d.shuffle;
  p.Wout = d.Wo * d.shuffle;
  const int sms = device_sm_count();
  op.grid = p.total_tiles < sms ? p.total_tiles : sms;

  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    PSSR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  }
  return PSSR_OK;
}

int conv_launch(const ConvOp& op, const void* tmaps_dev, cudaStream_t stream) {
  ConvKParams p = *reinterpret_cast<const ConvKParams*>(op.kparams);
  p.tmaps = reinterpret_cast<const CUtensorMap*>(tmaps_dev);
  conv_igemm_kernel<<<op.grid, kThreads, op.smem_bytes, stream>>>(p);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

}  // namespace pssr
