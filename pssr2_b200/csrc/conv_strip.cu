// Strip implicit-GEMM convolution (v2 of the tcgen05 path) -- the kernel that runs every stride-1
// 3x3 / 1x1 convolution of the network forward (pssr/models/_blocks.py:15-41, resunet.py:65-96).
//
// Why a second kernel: conv_igemm.cu fetches one [128 px x 64 ch] box per filter tap, i.e. every input
// pixel crosses L2->SMEM nine times and every weight block once per 128-pixel tile; on B200 that is
// L2-bandwidth bound at ~20 % of the tensor peak (profiles/r01_*).  Here
//   * the batch is viewed as ONE flat sequence of zero-padded pixels  q = (n*(H+2) + y+1)*(W+2) + x+1,
//     so a filter tap (dy,dx) is the pure linear shift  q + dy*(W+2) + dx;
//   * a work unit = T consecutive 128-pixel M tiles (x one N tile).  The padded input rows the unit
//     touches (incl. one halo row above/below) are staged ONCE per 64-channel block by per-row TMA boxes
//     (out-of-image rows/columns are zero-filled by the TMA unit = the conv's zero padding), and the nine
//     shifted A operands are nine UMMA shared-memory descriptors into that one buffer (start address
//     advanced by whole 128-byte rows; the descriptor's base-offset field carries the swizzle phase);
//   * every weight block [N x 64] is fetched once per unit and multiplies all T tiles (T accumulators in
//     TMEM), cutting weight traffic by T.
// Outputs computed at padding positions (2/(W+2) of the columns, 2/(H+2) of the rows) are discarded by
// the epilogue.
//
// Warps: 0 = A producer (TMA rows), 1 = MMA issuer + TMEM owner, 2 = B producer (TMA weights), 3 idle,
// 4..11 = epilogue, two warps per TMEM lane quarter splitting the columns (TMEM -> registers -> bias / act ->
// 16-bit NHWC stores with pixel-shuffle as addressing, or the fused Reconstruction tail).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "plan.h"

namespace pssr {

static constexpr int kSThreads = 384;   // 4 control warps + 8 epilogue warps (two per TMEM lane quarter)
static constexpr int kSMaxB = 8;

struct StripKParams {
  const CUtensorMap* tmaps;   // device: [0..2] sources (row boxes), [3] weights
  int n_segs;
  int seg_src[4], seg_taps[4], seg_cblocks[4], seg_kb0[4];
  int num_kb;
  int H, W, B, P, IP;         // P = Wb + 2*pad (row pitch of a virtual image), IP = (H + 2*pad) * P
  int pad;                    // 1 when a 3x3 segment needs the zero halo, 0 for pure 1x1 layers (exact pixels, no waste)
  int HP;                     // H + 2*pad
  int NJ, Wb;                 // images wider than 128 are split into NJ column blocks of Wb pixels: each block is a
                              // virtual image of pitch P = Wb + 2 whose halo columns are the neighbouring block's pixels
  int q_begin, q_end;         // first / one-past-last real pixel in q space (host checks it fits 31 bits)
  int T;                      // M tiles per unit
  int units_m, n_tiles, total_units;
  int block_n, n_valid, n_total, wide_store;
  int rmax;                   // padded rows per A buffer
  uint32_t a_bytes, b_bytes;  // bytes per A buffer / per B stage
  int b_stages, tmem_bufs;
  int G;                      // filter taps fetched per B stage for 3x3 segments (1, 3 or 9): fewer barrier round trips
  uint32_t tap_bytes;         // bytes of one tap's weight block = block_n * 128
  int desc_mode;              // 0 (default): base_offset field = 0 -- measured on B200: the UMMA swizzle is a function of the
                              // absolute shared-memory address, a non-zero base offset double-applies the phase
  int dbg;                    // developer experiments (PSSR_DBG): 1 no stores, 2 no MMA, 4 no A loads, 8 no B loads
  const float* bias;
  const float* out_scale;
  uint16_t* out;
  float* out_f32;
  int out_cstride, out_choff, shuffle, cps, act, fp16;
  int Hout, Wout;
  const float* tail_w;        // fused Reconstruction tail: fp32 [9][cps] or nullptr
  float* tail_z;              // fp32 [B][H][r*r*9][W]
};

// developer timeline (PSSR_DBG bit 16): per CTA 128 clock64 stamps -- [0] entry, [1] setup done, [2+2u] unit u first MMA issued,
// [3+2u] unit u committed, [64+2u] unit u accumulators ready (epilogue side), [65+2u] unit u epilogue done, [127] exit
__device__ long long g_strip_trace[148 * 256];
#define STRIP_TRACE(slot)                                                                                   \
  do {                                                                                                      \
    if ((p.dbg & 16) && lane == 0 && (slot) < 127) g_strip_trace[(blockIdx.x % 148) * 256 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ uint64_t strip_desc(uint32_t addr, int mode) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (mode) d |= (uint64_t)((addr >> 7) & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t (&o)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
               : "memory");
}

// exact-erf GELU (nn.GELU(), _rdnet.py:186) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the 16-bit
// rounding of the stored activation) -- erff() costs ~3x more instructions and made the RDNet 1x1 expansions epilogue-bound
__device__ __forceinline__ float gelu_erf2(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);      // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

__global__ void __launch_bounds__(kSThreads, 1) conv_strip_kernel(const __grid_constant__ StripKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 + 2 + 2 * kSMaxB + 4];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 1) STRIP_TRACE(0);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + 2u * p.a_bytes;
  // bias (and optional per-channel scale) of the whole layer, staged once per CTA
  float* bias_s = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + 2u * p.a_bytes + (uint32_t)p.b_stages * p.b_bytes);
  float* scale_s = bias_s + p.n_total;
  float* tailw_s = scale_s + (p.out_scale != nullptr ? p.n_total : 0);
  for (int i = threadIdx.x; i < p.n_total; i += kSThreads) {
    bias_s[i] = p.bias[i];
    if (p.out_scale != nullptr) scale_s[i] = p.out_scale[i];
  }
  if (p.tail_w != nullptr)   // global [tap][c]  ->  shared [c/4][tap][c%4]
    for (int i = threadIdx.x; i < 9 * p.cps; i += kSThreads) {
      const int t = i / p.cps, c = i - t * p.cps;
      tailw_s[((c >> 2) * 9 + t) * 4 + (c & 3)] = p.tail_w[i];
    }
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (4 + kSMaxB + s); };
  auto t_full = [&](int b) { return bar0 + 8u * (4 + 2 * kSMaxB + b); };
  auto t_empty = [&](int b) { return bar0 + 8u * (4 + 2 * kSMaxB + 2 + b); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(t_full(b), 1); mbar_init(t_empty(b), 8); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp == 1) STRIP_TRACE(1);
  const int block_n = p.block_n;
  const int T = p.T;
  const int unit_q = 128 * T;

  if (warp == 0) {
    // ============================ A producer: padded rows via TMA ===========================
    if (lane == 0) {
      int as = 0;
      uint32_t aphase = 0;
      const int rows_total = p.B * p.NJ * p.HP;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        const int um = unit / p.n_tiles;
        const int qa = p.q_begin + um * unit_q;
        int qb = qa + unit_q;
        if (qb > p.q_end) qb = p.q_end;
        const int halo = p.pad * (p.P + 1);
        const int r0 = (qa - halo) / p.P;
        int r1 = (qb - 1 + halo) / p.P;
        if (r1 > rows_total - 1) r1 = rows_total - 1;
        const int nrows = r1 - r0 + 1;
        for (int sg = 0; sg < p.n_segs; ++sg) {
          const CUtensorMap* tm = p.tmaps + p.seg_src[sg];
          for (int cb = 0; cb < p.seg_cblocks[sg]; ++cb) {
            mbar_wait(a_empty(as), aphase ^ 1u);
            if (p.dbg & 4) { mbar_arrive(a_full(as)); if (++as == 2) { as = 0; aphase ^= 1u; } continue; }
            if (!p.pad) {
              // pure 1x1 layer: pixel space is the flat NHWC pixel index, the whole A tile is ONE 2-D box [128T px x 64 ch]
              mbar_arrive_expect_tx(a_full(as), (uint32_t)unit_q * 128u);
              tma_load_2d(a_base + (uint32_t)as * p.a_bytes, tm, a_full(as), cb * 64, qa);
              if (++as == 2) { as = 0; aphase ^= 1u; }
              continue;
            }
            mbar_arrive_expect_tx(a_full(as), (uint32_t)nrows * (uint32_t)p.P * 128u);
            const uint32_t dst0 = a_base + (uint32_t)as * p.a_bytes;
            for (int r = 0; r < nrows; ++r) {
              const int rho = r0 + r;
              const int v = rho / p.HP;
              const int py = rho - v * p.HP;
              const int n = v / p.NJ;
              const int j = v - n * p.NJ;
              const uint32_t dst = dst0 + (uint32_t)r * (uint32_t)p.P * 128u;
              tma_load_4d(dst, tm, a_full(as), cb * 64, j * p.Wb - p.pad, py - p.pad, n);
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ B producer: weights ===================================
    if (lane == 0) {
      const CUtensorMap* tmB = p.tmaps + 3;
      int bs = 0;
      uint32_t bphase = 0;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        const int n_tile = unit % p.n_tiles;
        for (int sg = 0; sg < p.n_segs; ++sg) {
          const int taps = p.seg_taps[sg], cbs = p.seg_cblocks[sg];
          for (int cb = 0; cb < cbs; ++cb) {
            const int gs = taps == 9 ? p.G : 1;
            for (int t0 = 0; t0 < taps; t0 += gs) {
              mbar_wait(b_empty(bs), bphase ^ 1u);
              if (p.dbg & 8) mbar_arrive(b_full(bs));
              else {
                mbar_arrive_expect_tx(b_full(bs), (uint32_t)gs * p.tap_bytes);
                for (int t = 0; t < gs; ++t) {
                  const int kb = p.seg_kb0[sg] + (t0 + t) * cbs + cb;   // weights are packed tap-major, then channel block
                  tma_load_2d(b_base + (uint32_t)bs * p.b_bytes + (uint32_t)t * p.tap_bytes, tmB, b_full(bs), kb * 64, n_tile * block_n);
                }
              }
              if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================== MMA issuer ==========================================
    // The whole warp walks the loops (warp-uniform control flow keeps descriptors in uniform registers);
    // one elected lane issues the tcgen05 instructions.
    const uint32_t idesc = umma_idesc_f16(p.fp16 ? 0 : 1, block_n);
    int as = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    int it = 0;
    int trace_stage = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++it) {
      const int um = unit / p.n_tiles;
      const int qa = p.q_begin + um * (128 * T);
      const int r0 = (qa - p.pad * (p.P + 1)) / p.P;
      const int row_off0 = p.pad ? qa - r0 * p.P : 0;     // smem row of the unit's first pixel
      int tv = (p.q_end - qa + 127) / 128;
      if (tv > T) tv = T;
      if (p.dbg & 2) tv = 0;
      const int buf = p.tmem_bufs == 2 ? (it & 1) : 0;
      const uint32_t use = p.tmem_bufs == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
      mbar_wait(t_empty(buf), (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(buf * T * block_n);
      uint32_t first = 1;
      for (int sg = 0; sg < p.n_segs; ++sg) {
        const int taps = p.seg_taps[sg];
        for (int cb = 0; cb < p.seg_cblocks[sg]; ++cb) {
          mbar_wait(a_full(as), aphase);
          tc_fence_after();
          if (first) STRIP_TRACE(2 + 2 * it);
          // descriptor of the unit's first pixel row in this A buffer; +8 per 128-byte row, +2 per 16 K elements
          const uint64_t adesc0 = strip_desc(a_base + (uint32_t)as * p.a_bytes + (uint32_t)row_off0 * 128u, 0);
          const int gs = taps == 9 ? p.G : 1;
          for (int t0 = 0; t0 < taps; t0 += gs) {
            mbar_wait(b_full(bs), bphase);
            tc_fence_after();
            if ((p.dbg & 16) && it == 1 && lane == 0 && trace_stage < 128) g_strip_trace[(blockIdx.x % 148) * 256 + 128 + trace_stage++] = clock64();
            if (elect_one()) {
              for (int tt = 0; tt < gs; ++tt) {
                const int t = t0 + tt;
                const int shift = taps == 9 ? (t / 3 - 1) * p.P + (t % 3 - 1) : 0;
                const uint64_t bdesc = strip_desc(b_base + (uint32_t)bs * p.b_bytes + (uint32_t)tt * p.tap_bytes, 0);
                const uint64_t adesc_t = adesc0 + (uint64_t)(long long)(shift * 8);
                for (int mt = 0; mt < tv; ++mt) {
                  const uint64_t ad = adesc_t + (uint64_t)(mt * 128 * 8);
                  const uint32_t dcol = d0 + (uint32_t)(mt * block_n);
                  umma_f16(dcol, ad, bdesc, idesc, first ? 0u : 1u);
                  umma_f16(dcol, ad + 2, bdesc + 2, idesc, 1u);
                  umma_f16(dcol, ad + 4, bdesc + 4, idesc, 1u);
                  umma_f16(dcol, ad + 6, bdesc + 6, idesc, 1u);
                }
                first = 0;
              }
              umma_commit(b_empty(bs));
            }
            __syncwarp();
            first = 0;
            if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
          }
          if (elect_one()) umma_commit(a_empty(as));
          __syncwarp();
          if (++as == 2) { as = 0; aphase ^= 1u; }
        }
      }
      if (elect_one()) umma_commit(t_full(buf));
      __syncwarp();
      STRIP_TRACE(3 + 2 * it);
    }
  } else if (warp >= 4) {
    // ==================================== epilogue ==========================================
    const int q4 = warp & 3;
    const int eg = (warp - 4) >> 2;           // epilogue group 0 / 1
    const int row = q4 * 32 + lane;
    const int r = p.shuffle;
    const int pw = p.tail_z != nullptr ? p.cps : 64;        // columns per work item (a whole sub-pixel in tail mode)
    const int npairs = (block_n + pw - 1) / pw;
    int it = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++it) {
      const int n_tile = unit % p.n_tiles;
      const int um = unit / p.n_tiles;
      const int qa = p.q_begin + um * unit_q;
      int tv = (p.q_end - qa + 127) / 128;
      if (tv > T) tv = T;
      const int buf = p.tmem_bufs == 2 ? (it & 1) : 0;
      const uint32_t use = p.tmem_bufs == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
      mbar_wait(t_full(buf), use & 1u);
      tc_fence_after();
      if (warp == 4) STRIP_TRACE(64 + 2 * it);
      for (int item = eg; item < tv * npairs; item += 2) {
        const int mt = item / npairs;
        const int pi = item - mt * npairs;
        const int c_lo = pi * pw;
        const int c_hi = c_lo + pw < block_n ? c_lo + pw : block_n;
        const int q = qa + mt * 128 + row;
        const int vimg = q / p.IP;
        const int rem = q - vimg * p.IP;
        const int py = rem / p.P;
        const int px = rem - py * p.P;
        const int n = vimg / p.NJ;
        const int x = (vimg - n * p.NJ) * p.Wb + px - p.pad, y = py - p.pad;
        const bool valid = (q < p.q_end) && px >= p.pad && px < p.Wb + p.pad && x < p.W && py >= p.pad && py < p.H + p.pad;
        // sub-pixel / channel position of the unit's first output column, advanced incrementally (no divisions per chunk)
        int sub = (n_tile * block_n + c_lo) / p.cps;
        int cc = n_tile * block_n + c_lo - sub * p.cps;
        int si = sub / r, sj = sub - si * r;
        const size_t pix00 = ((size_t)n * p.Hout + (size_t)(y * r)) * p.Wout + (size_t)(x * r);
        float zacc[9];
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * T * block_n + mt * block_n);
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (p.tail_z != nullptr) {
            // fused Reconstruction tail: per-tap projection of this sub-pixel's channels, fp32 on CUDA cores
            if (valid && !(p.dbg & 1)) {
              const int nbase = n_tile * block_n + c0;
              const int sub_t = nbase / p.cps;
              const int ccb = nbase - sub_t * p.cps;
              if (ccb == 0) {
#pragma unroll
                for (int t = 0; t < 9; ++t) zacc[t] = 0.f;
              }
              float rr[32];
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 bb = *reinterpret_cast<const float4*>(bias_s + nbase + 4 * j4);
                rr[4 * j4 + 0] = fmaxf(__uint_as_float(v[4 * j4 + 0]) + bb.x, 0.f);
                rr[4 * j4 + 1] = fmaxf(__uint_as_float(v[4 * j4 + 1]) + bb.y, 0.f);
                rr[4 * j4 + 2] = fmaxf(__uint_as_float(v[4 * j4 + 2]) + bb.z, 0.f);
                rr[4 * j4 + 3] = fmaxf(__uint_as_float(v[4 * j4 + 3]) + bb.w, 0.f);
              }
              // 9 independent accumulation chains interleaved (one per tap): weights are staged as [c/4][tap][4]
              const float4* w4p = reinterpret_cast<const float4*>(tailw_s) + (ccb >> 2) * 9;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                  const float4 w4 = w4p[j4 * 9 + t];
                  zacc[t] = fmaf(rr[4 * j4 + 0], w4.x, zacc[t]);
                  zacc[t] = fmaf(rr[4 * j4 + 1], w4.y, zacc[t]);
                  zacc[t] = fmaf(rr[4 * j4 + 2], w4.z, zacc[t]);
                  zacc[t] = fmaf(rr[4 * j4 + 3], w4.w, zacc[t]);
                }
              }
              if (ccb + 32 == p.cps) {
                const int planes = r * r * 9;
                float* zp = p.tail_z + (((size_t)n * p.H + y) * planes + (size_t)sub_t * 9) * p.W + x;
#pragma unroll
                for (int t = 0; t < 9; ++t) zp[(size_t)t * p.W] = zacc[t];
              }
            }
          } else {
            const int nbase = n_tile * block_n + c0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int nn = nbase + h * 16;
              if (h == 1 || c0 > c_lo) {   // advance the (sub-pixel, channel) cursor by 16 columns
                cc += 16;
                if (cc >= p.cps) { cc -= p.cps; if (++sj == r) { sj = 0; ++si; } }
              }
              if (valid && !(p.dbg & 1) && nn < p.n_valid) {
                float f[16];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 bb = *reinterpret_cast<const float4*>(bias_s + nn + 4 * j4);   // LDS.128, warp-uniform (broadcast)
                  f[4 * j4 + 0] = __uint_as_float(v[h * 16 + 4 * j4 + 0]) + bb.x;
                  f[4 * j4 + 1] = __uint_as_float(v[h * 16 + 4 * j4 + 1]) + bb.y;
                  f[4 * j4 + 2] = __uint_as_float(v[h * 16 + 4 * j4 + 2]) + bb.z;
                  f[4 * j4 + 3] = __uint_as_float(v[h * 16 + 4 * j4 + 3]) + bb.w;
                }
                if (p.act == PSSR_ACT_RELU) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                } else if (p.act == PSSR_ACT_GELU) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] = gelu_erf2(f[j]);
                }
                if (p.out_scale != nullptr) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] *= scale_s[nn + j];
                }
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = pack2(f[2 * j], f[2 * j + 1], p.fp16);
                if (p.wide_store) {
                  // 16 channels = one full 32-byte sector per thread and instruction
                  const size_t pix = pix00 + (size_t)si * p.Wout + (size_t)sj;
                  st_global_v8(p.out + pix * p.out_cstride + p.out_choff + cc, o);
                } else {
#pragma unroll
                  for (int g = 0; g < 2; ++g) {
                    const int n8 = nn + g * 8;
                    if (n8 < p.n_valid) {
                      const int sub = n8 / p.cps;
                      const int cc = n8 - sub * p.cps;
                      const int si = sub / r, sj = sub - si * r;
                      const size_t pix = ((size_t)n * p.Hout + (size_t)(y * r + si)) * p.Wout + (size_t)(x * r + sj);
                      if (p.out != nullptr)
                        *reinterpret_cast<uint4*>(p.out + pix * p.out_cstride + p.out_choff + cc) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                      if (p.out_f32 != nullptr) {
                        float4* d = reinterpret_cast<float4*>(p.out_f32 + pix * p.out_cstride + p.out_choff + cc);
                        d[0] = make_float4(f[8 * g + 0], f[8 * g + 1], f[8 * g + 2], f[8 * g + 3]);
                        d[1] = make_float4(f[8 * g + 4], f[8 * g + 5], f[8 * g + 6], f[8 * g + 7]);
                      }
                    }
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(buf));
      if (warp == 4) STRIP_TRACE(65 + 2 * it);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
    if ((p.dbg & 16) && lane == 0) g_strip_trace[(blockIdx.x % 148) * 256 + 127] = clock64();
  }
}

int strip_trace_fetch(long long* host, int n) {
  if (n > 148 * 256) n = 148 * 256;
  PSSR_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_strip_trace, sizeof(long long) * (size_t)n));
  return PSSR_OK;
}

// --------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn strip_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

bool strip_supported(const pssr_conv_desc_t& d) {
  if (getenv("PSSR_CONV_V1") != nullptr) return false;
  for (int s = 0; s < d.n_segs; ++s)
    if (d.segs[s].taps != 1 && d.segs[s].taps != 9) return false;
  if (d.n % 32 != 0 || d.n < 32) return false;
  // small feature maps: the padded pixel space wastes (1 - HW/((H+2)(W+2))) of the MMAs (36 % at 8x8, 21 % at 16x16)
  // and the exact-tile kernel (conv_igemm.cu) is faster there
  bool any9 = false;
  for (int s = 0; s < d.n_segs; ++s) any9 = any9 || d.segs[s].taps == 9;
  if (any9 && d.Wo < 32 && d.tail_z == nullptr && getenv("PSSR_STRIP_ALWAYS") == nullptr) return false;
  return true;
}

int strip_prepare(const pssr_conv_desc_t& d, int dtype, ConvOp& op) {
  EncodeTiledFn enc = strip_encode_fn();
  PSSR_REQUIRE(enc != nullptr, PSSR_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  PSSR_REQUIRE(d.n_srcs >= 1 && d.n_srcs <= 3 && d.n_segs >= 1 && d.n_segs <= 4, PSSR_EINVAL, "conv: n_srcs/n_segs out of range");
  PSSR_REQUIRE(d.n_valid > 0 && d.n_valid <= d.n && d.n_valid % 8 == 0, PSSR_EUNSUP, "conv: n_valid=%d must be a multiple of 8 and <= n", d.n_valid);
  PSSR_REQUIRE(d.shuffle >= 1 && d.n_valid % (d.shuffle * d.shuffle) == 0, PSSR_EUNSUP, "conv: n_valid %% shuffle^2 != 0");
  const int cps = d.n_valid / (d.shuffle * d.shuffle);
  PSSR_REQUIRE(cps % 8 == 0, PSSR_EUNSUP, "conv: channels after pixel shuffle (%d) must be a multiple of 8", cps);
  PSSR_REQUIRE(d.shuffle == 1 || d.n == d.n_valid, PSSR_EUNSUP, "conv: padded N with pixel shuffle unsupported");
  PSSR_REQUIRE(d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "conv: output channel stride/offset must be multiples of 8");
  PSSR_REQUIRE(d.out != nullptr || d.out_f32 != nullptr || d.tail_z != nullptr, PSSR_EINVAL, "conv: no output buffer");
  PSSR_REQUIRE(d.bias != nullptr && ((uintptr_t)d.bias & 15) == 0, PSSR_EINVAL, "conv: bias missing or misaligned");
  if (d.tail_z != nullptr) {
    PSSR_REQUIRE(d.tail_weight != nullptr && ((uintptr_t)d.tail_weight & 15) == 0, PSSR_EINVAL, "conv: tail_weight missing or misaligned");
    PSSR_REQUIRE(cps % 32 == 0 && d.act == PSSR_ACT_RELU && d.n == d.n_valid, PSSR_EUNSUP, "conv: fused tail needs C' %% 32 == 0, ReLU and unpadded N");
  }

  StripKParams& p = *reinterpret_cast<StripKParams*>(op.kparams);
  static_assert(sizeof(StripKParams) <= sizeof(op.kparams), "ConvOp::kparams too small");
  memset(&p, 0, sizeof(p));
  memset(op.tmaps, 0, sizeof(op.tmaps));
  op.variant = 2;

  int block_n = 256;
  while (d.n % block_n != 0) block_n >>= 1;
  p.block_n = block_n;
  p.n_tiles = d.n / block_n;
  p.n_valid = d.n_valid;
  p.n_total = d.n;
  p.H = d.Ho; p.W = d.Wo; p.B = d.B;
  p.NJ = (d.Wo + 127) / 128;
  p.Wb = (d.Wo + p.NJ - 1) / p.NJ;
  p.pad = 0;
  for (int s2 = 0; s2 < d.n_segs; ++s2)
    if (d.segs[s2].taps == 9) p.pad = 1;
  if (!p.pad) { p.NJ = 1; p.Wb = d.Wo; }
  p.P = p.Wb + 2 * p.pad;
  p.HP = d.Ho + 2 * p.pad;
  p.IP = p.HP * p.P;
  p.q_begin = p.pad * (p.P + 1);
  PSSR_REQUIRE((long long)d.B * p.NJ * p.IP < (1ll << 30), PSSR_EUNSUP, "conv: batch x padded image exceeds the 30-bit pixel index");
  p.q_end = d.B * p.NJ * p.IP - p.pad * (p.P + 1);

  int num_kb = 0;
  p.n_segs = d.n_segs;
  for (int s = 0; s < d.n_segs; ++s) {
    const pssr_kseg_t& sg = d.segs[s];
    PSSR_REQUIRE(sg.src >= 0 && sg.src < d.n_srcs && sg.cblocks >= 1, PSSR_EINVAL, "conv: bad K segment");
    p.seg_src[s] = sg.src; p.seg_taps[s] = sg.taps; p.seg_cblocks[s] = sg.cblocks; p.seg_kb0[s] = num_kb;
    num_kb += sg.taps * sg.cblocks;
  }
  p.num_kb = num_kb;

  // T: as many M tiles per unit as TMEM (512 columns) and shared memory allow, capped at 2 so the A halo
  // buffers can be double-buffered; TMEM is double-buffered when T * block_n <= 256.
  const int tailw_bytes = d.tail_z != nullptr ? 9 * cps * 4 : 0;
  const int vec_bytes = 4 * d.n * (d.out_scale != nullptr ? 2 : 1);   // staged bias (+ scale) vectors
  const int smem_cap = 226 * 1024 - 1024 - vec_bytes - tailw_bytes;
  // measured on B200 (scripts/dev_time_layer.py): keeping T * block_n <= 256 so that TMEM is double-buffered and the
  // epilogue of unit i overlaps the MMAs of unit i+1 beats the halved weight traffic of a larger T.
  int T = 256 / block_n;
  if (T > 2) T = 2;
  if (T < 1) T = 1;
  const char* envT = getenv("PSSR_STRIP_T");
  if (envT) { int t = atoi(envT); if (t >= 1 && t * block_n <= 512) T = t; }
  int rmax = 0, b_stages = 0, G = 1;
  bool has9 = false;
  for (int s2 = 0; s2 < d.n_segs; ++s2) has9 = has9 || d.segs[s2].taps == 9;
  const char* envG = getenv("PSSR_STRIP_G");
  for (;; --T) {
    rmax = (128 * T + 1) / p.P + 4;   // rows [floor((qa-P-1)/P), floor((qb+P)/P)] plus the overhang a partial last tile may read
    const long long a_bytes = p.pad ? (long long)rmax * p.P * 128 : 128LL * T * 128;
    const long long a_total = ((2 * a_bytes + 1023) / 1024) * 1024;
    const long long left = smem_cap - a_total;
    // taps per B stage: as many as still leave two stages (a stage holds G weight blocks on one barrier)
    G = 1;
    if (has9) {
      for (int g : {9, 3}) {
        if (envG && atoi(envG) != g) continue;
        if (left >= 2LL * g * block_n * 128) { G = g; break; }
      }
      if (envG && atoi(envG) == 1) G = 1;
    }
    b_stages = (int)(left / ((long long)G * block_n * 128));
    if (b_stages > kSMaxB) b_stages = kSMaxB;
    if (b_stages >= 2 || T == 1) break;
  }
  PSSR_REQUIRE(b_stages >= 2, PSSR_EUNSUP, "conv: image width %d needs more shared memory than available for the strip kernel", d.Wo);
  p.G = G;
  p.tap_bytes = (uint32_t)(block_n * 128);
  p.T = T;
  p.rmax = rmax;
  p.a_bytes = (uint32_t)((((p.pad ? (long long)rmax * p.P * 128 : 128LL * T * 128) + 1023) / 1024) * 1024);
  p.b_bytes = (uint32_t)(G * block_n * 128);
  p.b_stages = b_stages;
  p.tmem_bufs = (T * block_n <= 256) ? 2 : 1;
  const char* envm = getenv("PSSR_DESC_MODE");
  p.desc_mode = envm ? atoi(envm) : 0;
  const char* envd = getenv("PSSR_DBG");
  p.dbg = envd ? atoi(envd) : 0;
  const int total_q = p.q_end - p.q_begin;
  p.units_m = (total_q + 128 * T - 1) / (128 * T);
  p.total_units = p.units_m * p.n_tiles;

  const CUtensorMapDataType tdt = dtype == PSSR_DT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  for (int s = 0; s < d.n_srcs; ++s) {
    const pssr_src_t& src = d.srcs[s];
    PSSR_REQUIRE(src.base != nullptr && ((uintptr_t)src.base & 15) == 0, PSSR_EINVAL, "conv: source %d base must be 16-byte aligned", s);
    PSSR_REQUIRE(src.cstride % 8 == 0 && src.channels >= 1 && src.channels <= src.cstride, PSSR_EUNSUP, "conv: source %d bad channel stride", s);
    PSSR_REQUIRE(src.H == d.Ho && src.W == d.Wo && src.B == d.B, PSSR_EINVAL, "conv: source %d geometry does not match the output", s);
    cuuint64_t gdim[4] = {(cuuint64_t)src.channels, (cuuint64_t)src.W, (cuuint64_t)src.H, (cuuint64_t)src.B};
    cuuint64_t gstr[3] = {(cuuint64_t)src.cstride * 2, (cuuint64_t)src.cstride * 2 * src.W, (cuuint64_t)src.cstride * 2 * src.W * src.H};
    cuuint32_t box[4] = {64, (cuuint32_t)p.P, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (p.pad) {
      r = enc(&op.tmaps[s], tdt, 4, const_cast<void*>(src.base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gdim2[2] = {(cuuint64_t)src.channels, (cuuint64_t)src.W * src.H * src.B};
      cuuint64_t gstr2[1] = {(cuuint64_t)src.cstride * 2};
      cuuint32_t box2[2] = {64, (cuuint32_t)(128 * T)};
      cuuint32_t estr2[2] = {1, 1};
      r = enc(&op.tmaps[s], tdt, 2, const_cast<void*>(src.base), gdim2, gstr2, box2, estr2, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(source %d) failed with %d", s, (int)r);
  }
  {
    PSSR_REQUIRE(d.weights != nullptr && ((uintptr_t)d.weights & 15) == 0, PSSR_EINVAL, "conv: weights misaligned");
    const cuuint64_t ktot = (cuuint64_t)num_kb * 64;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)d.n};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&op.tmaps[3], tdt, 2, const_cast<void*>(d.weights), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  }
  op.smem_bytes = (int)(2 * p.a_bytes + (uint32_t)b_stages * p.b_bytes + 1024 + (uint32_t)vec_bytes + (uint32_t)tailw_bytes);

  p.bias = d.bias;
  p.out_scale = d.out_scale;
  p.out = reinterpret_cast<uint16_t*>(d.out);
  p.out_f32 = d.out_f32;
  p.tail_w = d.tail_weight;
  p.tail_z = d.tail_z;
  PSSR_REQUIRE(d.tail_z == nullptr || d.tail_layout == PSSR_TAIL_TAPS, PSSR_EUNSUP, "conv: this kernel only writes the per-tap tail layout");
  p.out_cstride = d.out_cstride;
  p.out_choff = d.out_choff;
  p.shuffle = d.shuffle;
  p.cps = cps;
  p.act = d.act;
  p.fp16 = dtype == PSSR_DT_FP16 ? 1 : 0;
  p.Hout = d.Ho * d.shuffle;
  p.Wout = d.Wo * d.shuffle;
  p.wide_store = (d.out != nullptr && d.out_f32 == nullptr && cps % 16 == 0 && d.out_choff % 16 == 0 && d.out_cstride % 16 == 0 &&
                  d.n_valid % 16 == 0 && ((uintptr_t)d.out & 31) == 0 && getenv("PSSR_NO_WIDE_STORE") == nullptr) ? 1 : 0;
  const int sms = device_sm_count();
  op.grid = p.total_units < sms ? p.total_units : sms;
  static bool attr_set = false;
  if (!attr_set) {
    PSSR_CHECK_CUDA(cudaFuncSetAttribute(conv_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    attr_set = true;
  }
  return PSSR_OK;
}

int strip_launch(const ConvOp& op, const void* tmaps_dev, cudaStream_t stream) {
  StripKParams p = *reinterpret_cast<const StripKParams*>(op.kparams);
  p.tmaps = reinterpret_cast<const CUtensorMap*>(tmaps_dev);
  conv_strip_kernel<<<op.grid, kSThreads, op.smem_bytes, stream>>>(p);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

}  // namespace pssr
