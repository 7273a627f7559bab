// Lean-issue implicit-GEMM convolution (v3 of the tcgen05 path) -- runs every stride-1 3x3 / 1x1 convolution of the
// network forward whose output is at least 32 pixels wide (pssr/models/_blocks.py:15-41, resunet.py:65-96,
// rdresunet.py:104-130).
//
// What the B200 measurements behind this kernel say (scripts/ubench/*.cu, profiles/r01_ubench_mma.txt):
//   * one warp issues tcgen05.mma into a SHALLOW queue: every instruction the issuing warp executes between two MMAs is
//     on the critical path (~6 cycles each), and one mbarrier poll costs ~100 unhidden cycles unless the MMA is 128 cycles
//     long.  The v2 loop (runtime tap arithmetic, ~100 instructions per 4 MMAs) ran N=256 MMAs at 163 cycles instead of
//     128 and N=64 MMAs at 75-89 instead of 48.  Here the loop is a template: taps, tiles and K steps are unrolled,
//     descriptors are base + compile-time constant, a stage holds G taps (up to the whole layer, "resident" weights).
//   * M=128 x N=64 x K=16 costs 48 cycles (shared-memory operand reads), N=128 64, N=192 96, N=256 128.
//   * the A operand in TMEM (TS mode) works with lane = row, column = k/2, low half = even k; the fused Reconstruction
//     tail uses it: the epilogue writes relu(acc + bias) back to TMEM as 16-bit and the tensor core projects it on the
//     nine tail taps (N = 16), instead of 9216 fp32 FMAs per pixel on the CUDA cores.
//
// Pixel space.  "flat" mode = v2's: the batch is one sequence of zero-padded pixels q = (v*(H+2) + y+1)*(W+2) + x+1 and a
// unit is T consecutive 128-pixel tiles; its input pixels [q0-P-1, q0+128T+P+1) are staged by per-row TMA boxes at a FIXED
// offset inside the stage buffer, so every MMA descriptor is a kernel-lifetime constant.  "rows" mode (W % 128 == 0: the
// 128^2 / 256^2 layers, where the flat halo is 5x the tile) = a tile is one 128-pixel image row segment, a unit is T rows
// of one image, staged with their two halo rows by ONE 4-D TMA box; no MMA is spent on padding columns.
//
// Warps: 0 = A producer, 1 = MMA issuer + TMEM owner, 2 = B producer, 3 idle, 4..11 = epilogue (two per TMEM lane quarter).
#include <cuda.h>
#include <cuda_fp8.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "plan.h"

namespace pssr {

static constexpr int kV3Threads = 384;
static constexpr int kV3MaxB = 8;
static constexpr int kV3MaxR = 12;   // ring slots (row groups) in rows mode

struct V3Params {
  const CUtensorMap* tmaps;   // device: [0..3] sources, [kTmW] weights, [kTmOut] output, [kTmW8] e5m2 weights
  int n_segs;
  int seg_src[6], seg_taps[6], seg_cblocks[6], seg_kb0[6];
  int num_kb;
  int H, W, B, P, IP, pad, HP, NJ, Wb;
  int rows_mode;              // 1: tile = one 128-pixel row segment (W % 128 == 0)
  int total_vrows;            // rows mode: B * NJ * H
  int ring_R;                 // rows mode: row groups in the ring
  int row_major;              // rows mode: filter-row-major K order with early release / late acquire of ring rows (shallow ring)
  uint32_t slot_bytes;        // rows mode: one row of one 64-channel plane = P * 128
  uint32_t slot16_bytes;      // rows mode: one row of a narrow (16-channel, SWIZZLE_32B) plane = P * 32 rounded up to 128
  uint32_t group_bytes;       // rows mode: all planes of one row (layout stride of a ring slot)
  uint32_t group_tx;          // rows mode: bytes the TMA unit writes per ring slot
  int seg_k16[6];             // rows mode: the segment's source is a 16-channel tensor staged as a narrow plane (one K = 16 MMA)
  int seg_f8[6];              // rows mode: e5m2 source (64-byte pixels, SWIZZLE_64B) x e5m2 weights, kind::f8f6f4 (K = 32)
  uint32_t slot8_bytes;       // rows mode: one row of an e5m2 plane = P * 64 rounded up to 128
  int seg_load[6];            // rows mode: 0 = the segment reads the planes an earlier segment over the same source staged
  uint32_t seg_plane[6];      // rows mode: byte offset of the segment's first plane inside a ring slot
  uint32_t ring_bytes;        // rows mode: ring_R * group_bytes rounded up to 1024
  int q_begin, q_end;         // flat mode
  // cols mode (small feature maps, W % 8 == 0, W < 32): a tile = 16 groups of 8 pixels = cR image rows x 8 columns of cG images
  // (16x16 maps: 16 rows of one image half; 8x8 maps: 8 rows of two images, rows interleaved).  One TMA box {64 ch, 10 px, cG,
  // cR+2 rows} stages the tile with its halo; the 8-pixel groups sit 10 pixels apart, so the UMMA descriptors use a 1280-byte
  // group stride and NO padded pixel is multiplied (flat mode wastes 21 % / 36 % of the MMAs at 16^2 / 8^2).
  int cols_mode, cR, cG, c_strips, c_yblocks, tiles_total;
  uint32_t tile_bytes;        // cols mode: staged bytes per tile (box bytes rounded up to 1024)
  uint32_t box_bytes;         // cols mode: bytes one tile box transfers
  uint32_t dy_units;          // descriptor units (16 B) between consecutive filter rows of the staged A operand
  uint32_t a_sbo;             // byte stride between 8-row groups of the A operand (1024; 1280 in cols mode)
  int off_px;                 // pixel offset of a unit's first pixel inside an A stage buffer
  uint32_t tile_step;         // descriptor units (16 B) between the A operands of consecutive tiles of a unit
  uint32_t a_bytes, a_tx_bytes, b_bytes, tap_bytes;
  int b_stages;
  int units_m, n_tiles, total_units;
  int block_n, n_valid, n_total, wide_store;
  int dbg;
  const float* bias;
  const float* out_scale;
  uint16_t* out;
  float* out_f32;
  int out_cstride, out_choff, shuffle, cps, act, fp16;
  int Hout, Wout;
  int tma_store;              // 1: the epilogue stages 16-bit tiles in shared memory and writes them with TMA (tmaps[kTmOut])
  uint32_t stage_off;         // byte offset of the 8 x 4 KB store staging tiles inside the aligned dynamic shared memory
  const float* tail_w;        // fused Reconstruction tail: fp32 [9][64]
  float* tail_z;              // fp32 [B][H][r*r*9][W]: the r*r*9 plane rows of one LR row are contiguous
  int tail_win48;             // PSSR_TAIL_WINDOW48: [B][H][48][W], the projections pre-summed by HR output position (r = 4, rows mode)
  // compensated precision (pssr_conv_desc_t::out_lo / PSSR_TAIL_COMP)
  uint16_t* out_lo;           // second 16-bit output: rn16(y - rn16(y))
  int lo_cstride, lo_choff;
  int tail_comp;              // tail: B = [W_hi ; W_lo] (N = 32) against the 16-bit activation in TMEM, plus an e5m2 pass
                              // (kind::f8f6f4, A = 64 * (y - rn16(y)) staged in shared memory, B = W / 64) into the same accumulator
  const uint16_t* resid;      // epilogue residual: y = act(acc + bias + resid_scale * resid[pixel][n]) (shuffle == 1)
  int resid_cstride, resid_choff;
  float resid_scale;
  uint32_t tailw_bytes;       // bytes of the tail operand region: 16-bit weight tile [+ e5m2 weight tile + 4 x 8 KB e5m2 A tiles]
};

// developer timeline (PSSR_DBG bit 16): per CTA 256 clock64 stamps -- [0] entry, [1] setup done, [2+2u] unit u: accumulator buffer
// free (MMA warp), [3+2u] unit u committed, [64+2u] unit u accumulators ready (epilogue warp 4), [65+2u] unit u epilogue done,
// [127] exit, [128+2u] unit u: first A stage landed, [129+2u] TAIL: MMA warp starts waiting for unit u's 16-bit activations,
// [192+u] TAIL: tail MMAs of unit u issued, [232+2u] / [233+2u] (u < 8): epilogue warp 4 handed the activations over / saw the projections
__device__ long long g_v3_trace[148 * 256];
#define V3_TRACE(slot)                                                                                   \
  do {                                                                                                   \
    if ((p.dbg & 16) && lane == 0 && (slot) >= 0 && (slot) < 256) g_v3_trace[(blockIdx.x % 148) * 256 + (slot)] = clock64(); \
  } while (0)

// developer build (-DPSSR_V3_WAITPROF): cycles the MMA warp spends in each kind of barrier wait, summed over the launch, in trace
// slots [240] accumulator buffer, [241] input rows, [242] weight stages, [243] tail activations, [244] whole issue loop
#ifdef PSSR_V3_WAITPROF
#define V3_WAIT(cat, call)            \
  do {                                \
    const long long _t = clock64();   \
    call;                             \
    wprof[cat] += clock64() - _t;     \
  } while (0)
#else
#define V3_WAIT(cat, call) call
#endif

__device__ __forceinline__ uint64_t v3_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D (+)= A * B^T with the accumulate flag as an immediate / a register.  PAIR: cta_group::2 -- one instruction drives the
// tensor cores of both SMs of the CTA pair (M = 256: each CTA's own 128 A rows, each CTA holds half of B's N rows).
template <bool PAIR>
__device__ __forceinline__ void v3_mma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (PAIR) asm volatile("tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void v3_mma_p(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// A operand in TMEM (16-bit pairs: lane = row, column = k/2)
template <bool PAIR>
__device__ __forceinline__ void v3_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// e5m2 operands (kind::f8f6f4, K = 32 per instruction), fp32 accumulate into the same TMEM columns as the f16 MMAs
template <bool PAIR>
__device__ __forceinline__ void v3_mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// K-major operand tile with 64-byte rows, SWIZZLE_64B: 8-row groups 512 B apart
__device__ __forceinline__ uint64_t v3_desc64(uint32_t addr) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// four floats -> four e5m2 bytes (byte i = value i)
__device__ __forceinline__ uint32_t v3_e5m2x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E5M2);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E5M2);
  return lo | (hi << 16);
}
__device__ __forceinline__ float2 v3_unpack2(uint32_t v, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<__half2*>(&v));
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ void v3_arrive_cluster_release(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// wait that pairs with a cluster-scope release from the peer CTA (its shared-memory writes must be visible to the MMA)
__device__ __forceinline__ void v3_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// completion of all MMAs issued so far -> one arrival on the barrier (PAIR: on the barrier at this offset in BOTH CTAs)
template <bool PAIR>
__device__ __forceinline__ void v3_commit(uint32_t bar) {
  if (PAIR)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
  else
    umma_commit(bar);
}
__device__ __forceinline__ uint32_t v3_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void v3_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (an address in this CTA's shared memory) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t v3_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void v3_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA loads of a CTA pair: the data lands in this CTA's shared memory, the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void v3_tma_2d_pair(uint32_t dst, const void* tmap, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void v3_tma_4d_pair(uint32_t dst, const void* tmap, uint32_t cluster_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void v3_tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void v3_tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void v3_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void v3_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void v3_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void v3_st_global_v8(void* ptr, const uint32_t (&o)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
               : "memory");
}
// exact-erf GELU (nn.GELU(), _rdnet.py:186), erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 << 16-bit output rounding)
__device__ __forceinline__ float v3_gelu(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// TMA stores of the epilogue: shared memory tile -> global, out-of-range coordinates are clipped by the TMA unit
__device__ __forceinline__ void v3_tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void v3_tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void v3_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 16-bit pack with saturation to the finite range (and optional ReLU) in one F2FP instruction
__device__ __forceinline__ uint32_t v3_pack2(float lo, float hi, int fp16, bool relu) {
  uint32_t d;
  if (fp16) {
    if (relu) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  } else {
    if (relu) asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  }
  return d;
}

// PSSR_TAIL_WINDOW48: add the nine projections of source sub-position (i' = I, j' = 2*eg + SL) to the 6 x 4 window of HR output
// positions this epilogue group owns: tap (dy, dx) feeds (oi, oj) = (I - dy, j' - dx); window row = oi + 1, column = oj + 1 - 2*eg
// = SL - dx + 1.  All indices are compile-time constants: the window lives in registers.
template <int I, int SL>
__device__ __forceinline__ void v3_tail_window_add(float (&win)[24], const uint32_t (&zv)[16]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int dy = t / 3 - 1, dx = t % 3 - 1;
    win[(I - dy + 1) * 4 + (SL - dx + 1)] += __uint_as_float(zv[t]);
  }
}

// T tiles per CTA and unit, G filter taps per weight stage (3x3 segments), RES: the whole layer's weights stay in shared memory,
// TAIL: fused Reconstruction tail (block_n = 256, T = 1, 64 channels per pixel-shuffle sub-position), PAIR: the kernel runs
// as clusters of two CTAs; a unit is 2T tiles, the leader (cluster rank 0) issues cta_group::2 MMAs for both, each CTA stages
// its own A tiles and half of the weight rows, and drains its own TMEM.  ROWS: rows mode -- a tile is a 128-pixel row segment,
// every worker walks a CONTIGUOUS range of row groups and keeps the input rows in a shared-memory ring (each row is fetched
// once and prefetched several units ahead; all N tiles of a row group run against the same staged rows).
// X: the compensated-precision extras (e5m2 segments, epilogue residual, second output, compensated tail) are compiled in; the
// plain instantiation is what every other layer runs (the extras cost the lean issue loop and the epilogue ~25 % on the
// 64-channel layers when they are merely present as untaken branches).
template <int T, int G, bool RES, bool TAIL, bool PAIR, bool ROWS, bool X>
__global__ void __launch_bounds__(kV3Threads, 1) conv_v3_kernel(const __grid_constant__ V3Params p) {
  const bool x_tail_comp = X && p.tail_comp != 0;
  const uint16_t* const x_resid = X ? p.resid : nullptr;
  uint16_t* const x_out_lo = X ? p.out_lo : nullptr;
  constexpr int C = PAIR ? 2 : 1;
  const uint32_t rank = PAIR ? v3_cluster_rank() : 0u;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // persistent worker (CTA or CTA pair)
  const int workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[4 + 2 * kV3MaxB + 8 + 2 * kV3MaxR];
  __shared__ uint64_t rbars[8];      // X: one per epilogue warp -- its residual tile has landed in shared memory
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 1) V3_TRACE(0);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;                                       // A stages (flat) or the row ring (ROWS)
  const uint32_t a_total = ROWS ? p.ring_bytes : 2u * p.a_bytes;
  const uint32_t b_base = smem_base + a_total;
  const uint32_t b_total = RES ? (uint32_t)p.num_kb * p.tap_bytes : (uint32_t)p.b_stages * p.b_bytes;
  const uint32_t tailw_off = a_total + b_total;                            // tail operands, 1024-aligned
  const uint32_t vec_off = tailw_off + (TAIL ? p.tailw_bytes : 0u);
  // compensated tail: [0, 4096) 16-bit weight tile (N = 32), [4096, 5120) e5m2 weight tile, [6144, 6144 + 32 KB) e5m2 A tiles
  const uint32_t w8_off = tailw_off + 4096u;
  const uint32_t lo8_off = tailw_off + 6144u;
  float* bias_s = reinterpret_cast<float*>(smem_al + vec_off);
  float* scale_s = bias_s + p.n_total;
  for (int i = threadIdx.x; i < p.n_total; i += kV3Threads) {
    bias_s[i] = p.bias[i];
    if (p.out_scale != nullptr) scale_s[i] = p.out_scale[i];
  }
  if (TAIL) {
    // tail weights [9][64] fp32 -> 16-bit K-major SWIZZLE_128B operand tile [16 taps x 64 channels] (rows 9..15 zero);
    // a CTA pair splits the N rows: local row t of rank r is row r*NT/2 + t.  Compensated: N = 32, rows 16..31 hold
    // rn16(w - rn16(w)) of the same taps (the second half of the accumulator columns).
    const int NT = x_tail_comp ? 32 : 16;
    for (int i = threadIdx.x; i < (NT / C) * 64; i += kV3Threads) {
      const int t = i >> 6, c = i & 63;
      const int grow = t + (int)rank * (NT / C);
      const int tap = grow & 15;
      const float w = tap < 9 ? p.tail_w[tap * 64 + c] : 0.f;
      uint16_t val = pack1(w, p.fp16);
      if (grow >= 16) val = pack1(w - unpack1(val, p.fp16), p.fp16);
      const int chunk = (c >> 3) ^ (t & 7);
      reinterpret_cast<uint16_t*>(smem_al + tailw_off + t * 128 + chunk * 16)[c & 7] = val;
    }
    if (x_tail_comp) {
      // e5m2 tile [16 taps x 64 channels] of w / 64, 64-byte rows, SWIZZLE_64B (16-byte chunk index ^ ((row >> 1) & 3))
      for (int i = threadIdx.x; i < (16 / C) * 64; i += kV3Threads) {
        const int t = i >> 6, c = i & 63;
        const int tap = t + (int)rank * (16 / C);
        const float w = tap < 9 ? p.tail_w[tap * 64 + c] * 0.015625f : 0.f;
        const int chunk = (c >> 4) ^ ((t >> 1) & 3);
        (smem_al + w8_off + t * 64 + chunk * 16)[c & 15] = (uint8_t)__nv_cvt_float_to_fp8(w, __NV_SATFINITE, __NV_E5M2);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (4 + kV3MaxB + s); };
  auto t_full = [&](int b) { return bar0 + 8u * (4 + 2 * kV3MaxB + b); };
  auto t_empty = [&](int b) { return bar0 + 8u * (4 + 2 * kV3MaxB + 2 + b); };
  auto p_full = [&](int b) { return bar0 + 8u * (4 + 2 * kV3MaxB + 4 + b); };
  auto z_full = [&](int b) { return bar0 + 8u * (4 + 2 * kV3MaxB + 6 + b); };
  auto r_full = [&](int s) { return bar0 + 8u * (4 + 2 * kV3MaxB + 8 + s); };
  auto r_empty = [&](int s) { return bar0 + 8u * (4 + 2 * kV3MaxB + 8 + kV3MaxR + s); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < kV3MaxB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < kV3MaxR; ++s) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(t_full(b), 1); mbar_init(t_empty(b), 8 * C); mbar_init(p_full(b), 8 * C); mbar_init(z_full(b), 1); }
    if (X) for (int w8 = 0; w8 < 8; ++w8) mbar_init(smem_u32(&rbars[w8]), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) v3_tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
    else tmem_alloc(smem_u32(&tmem_base_smem), 512);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) v3_cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int block_n = p.block_n;
  if (warp == 1) V3_TRACE(1);
  // Programmatic dependent launch: the next kernel of the stream may start its own prologue (barriers, TMEM, bias, weight
  // prefetch) on SMs this grid has left; this grid's own prologue above ran the same way while its predecessor drained.
  // Everything that reads the predecessor's output (A loads) or overwrites buffers it may still read (epilogue stores)
  // first waits for the predecessor to finish (griddepcontrol.wait); weights / bias are constants and do not wait.
  if (threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- ROWS: this worker's contiguous range of row groups (a group = T rows of one virtual image) -----------------------
  // groups [g_lo, g_hi); a pair splits the range in two halves walked in lockstep (rank r: g_lo + r*msteps + ms).  The ring
  // advances identically in both CTAs of a pair (one MMA descriptor addresses both): a step is "fresh" (all T+2 rows
  // fetched) when either CTA starts a new image, else only the T new rows are fetched.
  const int groups = ROWS ? p.total_vrows / T : 0;
  const int g_lo = ROWS ? (int)((long long)worker * groups / workers) : 0;
  const int g_hi = ROWS ? (int)((long long)(worker + 1) * groups / workers) : 0;
  const int msteps = ROWS ? (PAIR ? (g_hi - g_lo + 1) / 2 : g_hi - g_lo) : 0;
  auto group_of = [&](int ms, int rk, bool& valid) {
    int g = g_lo + (PAIR ? rk * msteps : 0) + ms;
    valid = g < g_hi;
    if (!valid) g = g_hi - 1;
    return g;
  };
  auto step_fresh = [&](int ms) {
    if (ms == 0) return true;
    bool fresh = false;
#pragma unroll
    for (int rk = 0; rk < C; ++rk) {
      bool valid;
      const int g = group_of(ms, rk, valid);
      fresh = fresh || !valid || ((g * T) % p.H == 0);
    }
    return fresh;
  };

  if (warp == 0) {
    // ============================ A producer ================================================
    // every CTA stages its own tiles; in a pair the bytes of both CTAs are counted on the leader's barrier
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      if (ROWS) {
        int slot = 0;
        uint32_t phase = 0;
        for (int ms = 0; ms < msteps; ++ms) {
          bool valid;
          const int g = group_of(ms, (int)rank, valid);
          const int vr0 = g * T;
          const int v = vr0 / p.H;
          const int y0 = vr0 - v * p.H;
          const int n = v / p.NJ;
          const int x0 = (v - n * p.NJ) * 128 - 1;
          const bool fresh = step_fresh(ms);
          const int first_row = fresh ? y0 - 1 : y0 + 1;
          const int count = fresh ? T + 2 : T;
          for (int i = 0; i < count; ++i) {
            mbar_wait(r_empty(slot), phase ^ 1u);
            const uint32_t fbar = PAIR ? v3_mapa(r_full(slot), 0) : r_full(slot);
            if (rank == 0) mbar_arrive_expect_tx(r_full(slot), p.group_tx * C);
            for (int sg = 0; sg < p.n_segs; ++sg) {
              if (!p.seg_load[sg]) continue;
              uint32_t dst = a_base + (uint32_t)slot * p.group_bytes + p.seg_plane[sg];
              const CUtensorMap* tm = p.tmaps + p.seg_src[sg];
              for (int cb = 0; cb < p.seg_cblocks[sg]; ++cb) {
                if (PAIR) v3_tma_4d_pair(dst, tm, fbar, cb * 64, x0, first_row + i, n);
                else tma_load_4d(dst, tm, fbar, cb * 64, x0, first_row + i, n);
                dst += p.seg_k16[sg] ? p.slot16_bytes : (X && p.seg_f8[sg]) ? p.slot8_bytes : p.slot_bytes;
              }
            }
            if (++slot == p.ring_R) { slot = 0; phase ^= 1u; }
          }
        }
      } else {
        int as = 0;
        uint32_t aphase = 0;
        const int rows_total = p.B * p.NJ * p.HP;
        for (int unit = worker; unit < p.total_units; unit += workers) {
          const int um = unit / p.n_tiles;
          int qa = 0, r0 = 0, nrows = 0;
          uint32_t tx = 0;
          if (p.cols_mode) {
            tx = (uint32_t)T * p.box_bytes * C;
          } else if (p.pad) {
#pragma unroll
            for (int rk = 0; rk < C; ++rk) {        // row range of each CTA of the pair (the leader needs both byte counts)
              const int qa_r = p.q_begin + (um * C + rk) * (128 * T);
              const int r0_r = (qa_r - p.P - 1) / p.P;
              int r1_r = (qa_r + 128 * T + p.P) / p.P;
              if (r1_r > rows_total - 1) r1_r = rows_total - 1;
              int nr = r1_r - r0_r + 1;
              if (nr < 0) nr = 0;
              tx += (uint32_t)nr * (uint32_t)p.P * 128u;
              if (rk == (int)rank) { qa = qa_r; r0 = r0_r; nrows = nr; }
            }
          } else {
            qa = (um * C + (int)rank) * (128 * T);
            tx = (uint32_t)(128 * T) * 128u * C;
          }
          for (int sg = 0; sg < p.n_segs; ++sg) {
            const CUtensorMap* tm = p.tmaps + p.seg_src[sg];
            for (int cb = 0; cb < p.seg_cblocks[sg]; ++cb) {
              mbar_wait(a_empty(as), aphase ^ 1u);
              const uint32_t dst0 = a_base + (uint32_t)as * p.a_bytes;
              const uint32_t fbar = PAIR ? v3_mapa(a_full(as), 0) : a_full(as);
              if (rank == 0) mbar_arrive_expect_tx(a_full(as), tx);
              if (p.cols_mode) {
                for (int mt = 0; mt < T; ++mt) {
                  int t = (um * C + (int)rank) * T + mt;
                  if (t >= p.tiles_total) t = p.tiles_total - 1;       // overhanging tiles re-read the last one; the epilogue drops them
                  const int sx = t % p.c_strips;
                  const int t2 = t / p.c_strips;
                  const int yb = t2 % p.c_yblocks;
                  const int ng = t2 / p.c_yblocks;
                  const uint32_t dst = dst0 + (uint32_t)mt * p.tile_bytes;
                  // tensor map dims: (channel, x, image, y)
                  if (PAIR) v3_tma_4d_pair(dst, tm, fbar, cb * 64, sx * 8 - 1, ng * p.cG, yb * p.cR - 1);
                  else tma_load_4d(dst, tm, fbar, cb * 64, sx * 8 - 1, ng * p.cG, yb * p.cR - 1);
                }
              } else if (p.pad) {
                for (int r = 0; r < nrows; ++r) {
                  const int rho = r0 + r;
                  const int v = rho / p.HP;
                  const int py = rho - v * p.HP;
                  const int n = v / p.NJ;
                  const int j = v - n * p.NJ;
                  const uint32_t dst = dst0 + (uint32_t)(rho * p.P - qa + p.off_px) * 128u;
                  if (PAIR) v3_tma_4d_pair(dst, tm, fbar, cb * 64, j * p.Wb - 1, py - 1, n);
                  else tma_load_4d(dst, tm, fbar, cb * 64, j * p.Wb - 1, py - 1, n);
                }
              } else {
                if (PAIR) v3_tma_2d_pair(dst0, tm, fbar, cb * 64, qa);
                else tma_load_2d(dst0, tm, fbar, cb * 64, qa);
              }
              as ^= 1;
              if (as == 0) aphase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ B producer: weights ===================================
    // p.tap_bytes = bytes of one tap's weight block held by ONE CTA (a pair splits the N rows); bytes counted on the leader
    if (lane == 0) {
      const CUtensorMap* tmB16 = p.tmaps + kTmW;
      const CUtensorMap* tmB = tmB16;
      const int nrow0 = (int)rank * (block_n / C);
      if (RES) {
        const uint32_t fbar = PAIR ? v3_mapa(b_full(0), 0) : b_full(0);
        if (rank == 0) mbar_arrive_expect_tx(b_full(0), (uint32_t)p.num_kb * p.tap_bytes * C);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (PAIR) v3_tma_2d_pair(b_base + (uint32_t)kb * p.tap_bytes, tmB, fbar, kb * 64, nrow0);
          else tma_load_2d(b_base + (uint32_t)kb * p.tap_bytes, tmB, fbar, kb * 64, nrow0);
        }
      } else {
        int bs = 0;
        uint32_t bphase = 0;
        const int steps = ROWS ? msteps * p.n_tiles : 0;
        int k = 0;
        for (int unit = worker; ROWS ? k < steps : unit < p.total_units; unit += workers, ++k) {
          const int n_tile = ROWS ? k % p.n_tiles : unit % p.n_tiles;
          // ROWS: filter-row-major order (dy = -1, 0, +1; 1x1 segments ride with dy = 0), the order the MMA issuer walks so that
          // the top input row can be released after the first third of a unit.  Flat: channel-block-major (A is staged per block).
          const bool rm = ROWS && p.row_major;
          for (int g = 0; g < (rm ? 3 : 1); ++g)
          for (int sg = 0; sg < p.n_segs; ++sg) {
            const int taps = p.seg_taps[sg], cbs = p.seg_cblocks[sg];
            const int gs = taps == 9 ? G : 1;
            // e5m2 segments: 64-byte weight rows from their own matrix, half a 16-bit tap block per tap
            const bool f8 = X && ROWS && p.seg_f8[sg] != 0;
            tmB = f8 ? p.tmaps + kTmW8 : tmB16;
            const uint32_t tapb = f8 ? (p.tap_bytes >> 1) : p.tap_bytes;
            if (rm && taps != 9 && g != 1) continue;
            const int t_lo = (rm && taps == 9) ? 3 * g : 0, t_hi = (rm && taps == 9) ? 3 * g + 3 : taps;
            const int gsf = f8 ? 2 * gs : gs;       // e5m2 taps are half as large: two per 16-bit tap slot of a stage
            for (int cb = 0; cb < cbs; ++cb) {
              for (int t0 = t_lo; t0 < t_hi; t0 += gsf) {
                const int cnt = t_hi - t0 < gsf ? t_hi - t0 : gsf;
                mbar_wait(b_empty(bs), bphase ^ 1u);
                const uint32_t fbar = PAIR ? v3_mapa(b_full(bs), 0) : b_full(bs);
                if ((p.dbg & 32) && k > 0) {            // developer ablation: stale weights, no shared-memory fill traffic
                  if (rank == 0) mbar_arrive(b_full(bs));
                  if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
                  continue;
                }
                if (rank == 0) mbar_arrive_expect_tx(b_full(bs), (uint32_t)cnt * tapb * C);
                for (int t = 0; t < cnt; ++t) {
                  const int kb = p.seg_kb0[sg] + (t0 + t) * cbs + cb;   // weights are packed tap-major, then channel block
                  const uint32_t dst = b_base + (uint32_t)bs * p.b_bytes + (uint32_t)t * tapb;
                  if (PAIR) v3_tma_2d_pair(dst, tmB, fbar, kb * 64, n_tile * block_n + nrow0);
                  else tma_load_2d(dst, tmB, fbar, kb * 64, n_tile * block_n + nrow0);
                }
                if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================================== MMA issuer ==========================================
    // The whole warp walks the loops (warp-uniform control flow); one elected lane issues the tcgen05 instructions.
    // PAIR: M = 256 in the instruction descriptor (bits 24..28 hold M >> 4)
    const uint32_t idesc = umma_idesc_f16(p.fp16 ? 0 : 1, block_n) + (PAIR ? (8u << 24) : 0u);
    const uint32_t tail_n = x_tail_comp ? 32u : 16u;
    const uint32_t idesc_tail = umma_idesc_f16(p.fp16 ? 0 : 1, (int)tail_n) + (PAIR ? (8u << 24) : 0u);
    // kind::f8f6f4: fp32 accumulate, A and B e5m2 (format 1), K-major, N = 16
    const uint32_t idesc_f8 = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24) + (PAIR ? (8u << 24) : 0u);
    const uint32_t idesc_m8 = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)block_n >> 3) << 17) | ((128u >> 4) << 24) + (PAIR ? (8u << 24) : 0u);
    // e5m2 planes / weight tiles: same start-address field, SWIZZLE_64B layout (type 4), 512 bytes between 8-row groups
    const uint64_t f8_fix = (((uint64_t)(512u >> 4) << 32) | ((uint64_t)4 << 61)) - (((uint64_t)(1024u >> 4) << 32) | ((uint64_t)2 << 61));
    const uint64_t slot8_desc = (uint64_t)(p.slot8_bytes >> 4);
    const uint64_t tdesc = v3_desc(smem_base + tailw_off);
    const uint64_t w8desc = v3_desc64(smem_base + w8_off);
    const uint64_t lo8desc = v3_desc64(smem_base + lo8_off);
    const uint64_t adesc_s0 = v3_desc(a_base + (uint32_t)p.off_px * 128u) + ((uint64_t)((p.a_sbo - 1024u) >> 4) << 32);
    const uint64_t adesc_s1 = adesc_s0 + (uint64_t)(p.a_bytes >> 4);
    const uint64_t bdesc0 = v3_desc(b_base);
    const uint32_t bstep = p.b_bytes >> 4, tapstep = p.tap_bytes >> 4;
    const int64_t P8 = (int64_t)p.dy_units;
    const uint64_t tile_step = p.tile_step;
    const uint64_t slot_desc = (uint64_t)(p.slot_bytes >> 4);
    const uint64_t slot16_desc = (uint64_t)(p.slot16_bytes >> 4);
    // narrow plane: same start-address field, SWIZZLE_32B layout (type 6) and 256-byte stride between 8-row groups
    const uint64_t k16_fix = (((uint64_t)(256u >> 4) << 32) | ((uint64_t)6 << 61)) - (((uint64_t)(1024u >> 4) << 32) | ((uint64_t)2 << 61));
    int as = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    int it = 0;
#ifdef PSSR_V3_WAITPROF
    long long wprof[5] = {0, 0, 0, 0, 0};
    const long long wprof_t0 = clock64();
#endif
    // K segments packed into registers (bit 0: 3x3, bits 1..8: channel blocks, bits 9..: first K block): an indexed load from
    // the constant bank per segment would sit on the issue path of every unit
    uint32_t segw[6];
    uint32_t segp[6];             // rows mode: descriptor units (16 B) from a ring slot to the segment's first plane
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      segw[i] = i < p.n_segs ? ((p.seg_taps[i] == 9 ? 1u : 0u) | ((uint32_t)p.seg_cblocks[i] << 1) | ((uint32_t)p.seg_kb0[i] << 9) |
                                (p.seg_k16[i] ? 0x80000000u : 0u) | (p.seg_f8[i] ? 0x40000000u : 0u)) : 0u;
      segp[i] = i < p.n_segs ? (p.seg_plane[i] >> 4) : 0u;
    }
    const int n_segs = p.n_segs;

    auto issue_tail = [&](int pit) {
      const int pbuf = pit & 1;
      if (pit < 31) V3_TRACE(129 + 2 * pit);
      if (PAIR && x_tail_comp) V3_WAIT(3, v3_mbar_wait_cluster(p_full(pbuf), (uint32_t)(pit >> 1) & 1u));
      else V3_WAIT(3, mbar_wait(p_full(pbuf), (uint32_t)(pit >> 1) & 1u));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t cbase = tmem_base + (uint32_t)(pbuf * 256);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint32_t region = (uint32_t)((s >> 1) * 128);
          const uint32_t a_col = cbase + region + (uint32_t)((s & 1) * 32);
          const uint32_t d_col = cbase + region + 64u + (uint32_t)(s & 1) * tail_n;
#pragma unroll
          for (int k = 0; k < 4; ++k) v3_mma_ts<PAIR>(d_col, a_col + 8u * k, tdesc + 2u * k, idesc_tail, k ? 1u : 0u);
          if (x_tail_comp) {
            // + (64 * lo) x (w / 64), both e5m2, K = 32 per instruction: accumulates on the hi x W_hi columns
            const uint64_t ad = lo8desc + (uint64_t)(s * (8192 >> 4));
            v3_mma_f8<PAIR>(d_col, ad, w8desc, idesc_f8, 1u);
            v3_mma_f8<PAIR>(d_col, ad + 2u, w8desc + 2u, idesc_f8, 1u);
          }
        }
        v3_commit<PAIR>(z_full(pbuf));
      }
      __syncwarp();
      if (pit < 60) V3_TRACE(192 + pit);
    };

    // one accumulator tile set: all K segments of (n_tile) against the staged A.  ea[i] (ROWS): descriptor of ring row i of
    // the current group (row y0-1+i, pixel x0 = -1), plane 0.
    auto run_k = [&](const uint64_t (&ea)[T + 2]) {
      const int buf = it & 1;
      V3_WAIT(0, mbar_wait(t_empty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u));
      tc_fence_after();
      if (it < 31) V3_TRACE(2 + 2 * it);
      const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
      uint32_t acc = 0;
      bool tail_done = false;
      for (int sg = 0; sg < n_segs; ++sg) {
        const uint32_t sw = sg == 0 ? segw[0] : sg == 1 ? segw[1] : sg == 2 ? segw[2] : sg == 3 ? segw[3] : sg == 4 ? segw[4] : segw[5];
        uint64_t plane_desc = sg == 0 ? segp[0] : sg == 1 ? segp[1] : sg == 2 ? segp[2] : sg == 3 ? segp[3] : sg == 4 ? segp[4] : segp[5];
        const bool nine = (sw & 1u) != 0;
        const bool k16 = ROWS && (sw >> 31) != 0;
        const bool f8 = X && ROWS && ((sw >> 30) & 1u) != 0;
        const int cbs = (int)((sw >> 1) & 0xffu);
        const int kb0 = (int)((sw >> 9) & 0x1fffffu);
        for (int cb = 0; cb < cbs; ++cb, plane_desc += (k16 ? slot16_desc : f8 ? slot8_desc : slot_desc)) {
          uint64_t ad0 = 0;
          if (!ROWS) {
            V3_WAIT(1, mbar_wait(a_full(as), aphase));
            tc_fence_after();
            ad0 = as ? adesc_s1 : adesc_s0;
          }
          if (acc == 0 && it < 31) V3_TRACE(128 + 2 * it);
          if (ROWS && f8) {
            // e5m2 3x3 segment (streamed weights): two K = 32 MMAs per tap and 64-channel block, 2G taps per weight stage
#pragma unroll
            for (int t0 = 0; t0 < 9; t0 += 2 * G) {
              V3_WAIT(2, mbar_wait(b_full(bs), bphase));
              tc_fence_after();
              const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)bs * bstep) + f8_fix;
              if (elect_one()) {
#pragma unroll
                for (int tt = 0; tt < 2 * G; ++tt) {
                  const int t = t0 + tt;
                  if (t >= 9) break;
                  const int dy = t / 3 - 1, dx = t % 3 - 1;
                  const uint64_t bdt = bd + (uint64_t)tt * (uint64_t)(tapstep >> 1);
#pragma unroll
                  for (int mt = 0; mt < T; ++mt) {
                    const uint64_t adm = ea[mt + 1 + dy] + plane_desc + (uint64_t)((1 + dx) * 4) + f8_fix;
                    const uint32_t dcol = d0 + (uint32_t)(mt * block_n);
                    v3_mma_f8<PAIR>(dcol, adm, bdt, idesc_m8, (tt == 0) ? acc : 1u);
                    v3_mma_f8<PAIR>(dcol, adm + 2, bdt + 2, idesc_m8, 1u);
                  }
                }
                v3_commit<PAIR>(b_empty(bs));
              }
              __syncwarp();
              acc = 1;
              if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
            }
          } else if (nine) {
#pragma unroll
            for (int t0 = 0; t0 < 9; t0 += G) {
              uint64_t bd;
              if (RES) {
                bd = bdesc0 + (uint64_t)((uint32_t)(kb0 + t0 * cbs + cb) * tapstep);
              } else {
                V3_WAIT(2, mbar_wait(b_full(bs), bphase));
                tc_fence_after();
                bd = bdesc0 + (uint64_t)((uint32_t)bs * bstep);
              }
              const uint64_t btap = RES ? (uint64_t)((uint32_t)cbs * tapstep) : (uint64_t)tapstep;
              if (elect_one()) {
#pragma unroll
                for (int tt = 0; tt < G; ++tt) {
                  const int t = t0 + tt;
                  const int dy = t / 3 - 1, dx = t % 3 - 1;
                  const uint64_t bdt = bd + (uint64_t)tt * btap;
#pragma unroll
                  for (int mt = 0; mt < T; ++mt) {
                    const uint64_t adm = ROWS ? ea[mt + 1 + dy] + plane_desc + (uint64_t)((1 + dx) * 8)
                                              : ad0 + (uint64_t)(dy * P8 + dx * 8) + (uint64_t)mt * tile_step;
                    const uint32_t dcol = d0 + (uint32_t)(mt * block_n);
                    if (tt == 0) v3_mma_p<PAIR>(dcol, adm, bdt, idesc, acc);
                    else v3_mma_acc<PAIR>(dcol, adm, bdt, idesc);
                    v3_mma_acc<PAIR>(dcol, adm + 2, bdt + 2, idesc);
                    v3_mma_acc<PAIR>(dcol, adm + 4, bdt + 4, idesc);
                    v3_mma_acc<PAIR>(dcol, adm + 6, bdt + 6, idesc);
                  }
                }
                if (!RES) v3_commit<PAIR>(b_empty(bs));
              }
              __syncwarp();
              acc = 1;
              if (!RES) { if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; } }
              if (TAIL && t0 == 6 && !tail_done) {
                // the previous unit's 16-bit activations are in TMEM by now: project them on the nine tail taps
                if (it > 0) issue_tail(it - 1);
                tail_done = true;
              }
            }
          } else {
            uint64_t bd;
            if (RES) {
              bd = bdesc0 + (uint64_t)((uint32_t)(kb0 + cb) * tapstep);
            } else {
              V3_WAIT(2, mbar_wait(b_full(bs), bphase));
              tc_fence_after();
              bd = bdesc0 + (uint64_t)((uint32_t)bs * bstep);
            }
            if (elect_one()) {
#pragma unroll
              for (int mt = 0; mt < T; ++mt) {
                const uint32_t dcol = d0 + (uint32_t)(mt * block_n);
                if (k16) {     // one K = 16 step: pixel x0 + 1 of the narrow plane (32 bytes per pixel)
                  v3_mma_p<PAIR>(dcol, ea[mt + 1] + plane_desc + 2u + k16_fix, bd, idesc, acc);
                } else {
                  const uint64_t adm = ROWS ? ea[mt + 1] + plane_desc + 8u : ad0 + (uint64_t)mt * tile_step;
                  v3_mma_p<PAIR>(dcol, adm, bd, idesc, acc);
                  v3_mma_acc<PAIR>(dcol, adm + 2, bd + 2, idesc);
                  v3_mma_acc<PAIR>(dcol, adm + 4, bd + 4, idesc);
                  v3_mma_acc<PAIR>(dcol, adm + 6, bd + 6, idesc);
                }
              }
              if (!RES) v3_commit<PAIR>(b_empty(bs));
            }
            __syncwarp();
            acc = 1;
            if (!RES) { if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; } }
          }
          if (!ROWS) {
            if (elect_one()) v3_commit<PAIR>(a_empty(as));
            __syncwarp();
            as ^= 1;
            if (as == 0) aphase ^= 1u;
          }
        }
      }
      if (TAIL && !tail_done && it > 0) issue_tail(it - 1);
      if (elect_one()) v3_commit<PAIR>(t_full(buf));
      __syncwarp();
      if (it < 31) V3_TRACE(3 + 2 * it);
      ++it;
    };

    // ROWS: one accumulator tile set in filter-row-major order.  rows[i] = ring row y0-1+i of the current group: descriptor ea[i],
    // slot sl[i], barrier parity ph[i].  Row i is first read by filter row dy = max(-1, i-T) and last by dy = min(1, i-1), so it
    // is awaited just before its first use (bit i of newmask: the row was fetched for this step) and handed back to the producer
    // right after its last (bit i of relmask) -- with a ring of only T+2 rows the next rows load while this unit computes.
    auto run_k_rows = [&](const uint64_t (&ea)[T + 2], const int (&sl)[T + 2], const uint32_t (&ph)[T + 2], uint32_t newmask, uint32_t relmask) {
      const int buf = it & 1;
      V3_WAIT(0, mbar_wait(t_empty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u));
      tc_fence_after();
      if (it < 31) V3_TRACE(2 + 2 * it);
      const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
      uint32_t acc = 0;
      bool tail_done = false;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int dy = g - 1;
#pragma unroll
        for (int i = 0; i < T + 2; ++i) {
          const int need = (i - T) > -1 ? (i - T) : -1;
          if (need == dy && ((newmask >> i) & 1u)) V3_WAIT(1, mbar_wait(r_full(sl[i]), ph[i]));
        }
        tc_fence_after();
        if (g == 0 && it < 31) V3_TRACE(128 + 2 * it);
        for (int sg = 0; sg < n_segs; ++sg) {
          const uint32_t sw = sg == 0 ? segw[0] : sg == 1 ? segw[1] : sg == 2 ? segw[2] : sg == 3 ? segw[3] : sg == 4 ? segw[4] : segw[5];
          uint64_t plane_desc = sg == 0 ? segp[0] : sg == 1 ? segp[1] : sg == 2 ? segp[2] : sg == 3 ? segp[3] : sg == 4 ? segp[4] : segp[5];
          const bool nine = (sw & 1u) != 0;
          const bool k16 = (sw >> 31) != 0;
          const bool f8 = X && ((sw >> 30) & 1u) != 0;
          const int cbs = (int)((sw >> 1) & 0xffu);
          const int kb0 = (int)((sw >> 9) & 0x1fffffu);
          if (!nine && g != 1) continue;
          for (int cb = 0; cb < cbs; ++cb, plane_desc += (k16 ? slot16_desc : f8 ? slot8_desc : slot_desc)) {
            if (f8) {
              constexpr int GS = G == 9 ? 3 : G;
#pragma unroll
              for (int x0 = 0; x0 < 3; x0 += 2 * GS) {
                V3_WAIT(2, mbar_wait(b_full(bs), bphase));
                tc_fence_after();
                const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)bs * bstep) + f8_fix;
                if (elect_one()) {
#pragma unroll
                  for (int tt = 0; tt < 2 * GS; ++tt) {
                    if (x0 + tt >= 3) break;
                    const int dx = x0 + tt - 1;
                    const uint64_t bdt = bd + (uint64_t)tt * (uint64_t)(tapstep >> 1);
#pragma unroll
                    for (int mt = 0; mt < T; ++mt) {
                      const uint64_t adm = ea[mt + 1 + dy] + plane_desc + (uint64_t)((1 + dx) * 4) + f8_fix;
                      const uint32_t dcol = d0 + (uint32_t)(mt * block_n);
                      v3_mma_f8<PAIR>(dcol, adm, bdt, idesc_m8, (tt == 0) ? acc : 1u);
                      v3_mma_f8<PAIR>(dcol, adm + 2, bdt + 2, idesc_m8, 1u);
                    }
                  }
                  v3_commit<PAIR>(b_empty(bs));
                }
                __syncwarp();
                acc = 1;
                if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
              }
            } else if (nine) {
              constexpr int GS = G == 9 ? 3 : G;          // taps per weight stage inside one filter row
#pragma unroll
              for (int x0 = 0; x0 < 3; x0 += GS) {
                uint64_t bd;
                if (RES) {
                  bd = bdesc0 + (uint64_t)((uint32_t)(kb0 + (3 * g + x0) * cbs + cb) * tapstep);
                } else {
                  V3_WAIT(2, mbar_wait(b_full(bs), bphase));
                  tc_fence_after();
                  bd = bdesc0 + (uint64_t)((uint32_t)bs * bstep);
                }
                const uint64_t btap = RES ? (uint64_t)((uint32_t)cbs * tapstep) : (uint64_t)tapstep;
                if (elect_one()) {
#pragma unroll
                  for (int tt = 0; tt < GS; ++tt) {
                    const int dx = x0 + tt - 1;
                    const uint64_t bdt = bd + (uint64_t)tt * btap;
#pragma unroll
                    for (int mt = 0; mt < T; ++mt) {
                      const uint64_t adm = ea[mt + 1 + dy] + plane_desc + (uint64_t)((1 + dx) * 8);
                      const uint32_t dcol = d0 + (uint32_t)(mt * block_n);
                      if (tt == 0) v3_mma_p<PAIR>(dcol, adm, bdt, idesc, acc);
                      else v3_mma_acc<PAIR>(dcol, adm, bdt, idesc);
                      v3_mma_acc<PAIR>(dcol, adm + 2, bdt + 2, idesc);
                      v3_mma_acc<PAIR>(dcol, adm + 4, bdt + 4, idesc);
                      v3_mma_acc<PAIR>(dcol, adm + 6, bdt + 6, idesc);
                    }
                  }
                  if (!RES) v3_commit<PAIR>(b_empty(bs));
                }
                __syncwarp();
                acc = 1;
                if (!RES) { if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; } }
                if (TAIL && g == 2 && !tail_done) {
                  // the previous unit's 16-bit activations are in TMEM by now: project them on the nine tail taps
                  if (it > 0) issue_tail(it - 1);
                  tail_done = true;
                }
              }
            } else {
              uint64_t bd;
              if (RES) {
                bd = bdesc0 + (uint64_t)((uint32_t)(kb0 + cb) * tapstep);
              } else {
                V3_WAIT(2, mbar_wait(b_full(bs), bphase));
                tc_fence_after();
                bd = bdesc0 + (uint64_t)((uint32_t)bs * bstep);
              }
              if (elect_one()) {
#pragma unroll
                for (int mt = 0; mt < T; ++mt) {
                  const uint32_t dcol = d0 + (uint32_t)(mt * block_n);
                  if (k16) {
                    v3_mma_p<PAIR>(dcol, ea[mt + 1] + plane_desc + 2u + k16_fix, bd, idesc, acc);
                  } else {
                    const uint64_t adm = ea[mt + 1] + plane_desc + 8u;
                    v3_mma_p<PAIR>(dcol, adm, bd, idesc, acc);
                    v3_mma_acc<PAIR>(dcol, adm + 2, bd + 2, idesc);
                    v3_mma_acc<PAIR>(dcol, adm + 4, bd + 4, idesc);
                    v3_mma_acc<PAIR>(dcol, adm + 6, bd + 6, idesc);
                  }
                }
                if (!RES) v3_commit<PAIR>(b_empty(bs));
              }
              __syncwarp();
              acc = 1;
              if (!RES) { if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; } }
            }
          }
        }
        if (relmask != 0u) {
          if (elect_one()) {
#pragma unroll
            for (int i = 0; i < T + 2; ++i) {
              const int last_use = (i - 1) < 1 ? (i - 1) : 1;
              if (last_use == dy && ((relmask >> i) & 1u)) v3_commit<PAIR>(r_empty(sl[i]));
            }
          }
          __syncwarp();
        }
      }
      if (TAIL && !tail_done && it > 0) issue_tail(it - 1);
      if (elect_one()) v3_commit<PAIR>(t_full(buf));
      __syncwarp();
      if (it < 31) V3_TRACE(3 + 2 * it);
      ++it;
    };

    if (RES) {
      mbar_wait(b_full(0), 0);
      tc_fence_after();
    }
    if (ROWS) {
      const uint64_t ring_desc0 = v3_desc(a_base);
      const uint64_t group_desc = (uint64_t)(p.group_bytes >> 4);
      int cslot = 0;                // next ring slot to be consumed
      uint32_t cphase = 0;
      bool fresh = true;
      // first row of each CTA's current group, advanced by T per step (no divisions on the issue path); must reproduce
      // step_fresh(): a step is fresh when a CTA enters a new image or the pair's second half has run out of groups
      int y0r[C];
#pragma unroll
      for (int rk = 0; rk < C; ++rk) {
        bool v;
        y0r[rk] = (group_of(0, rk, v) * T) % p.H;
      }
      const bool odd_tail = PAIR && (((g_hi - g_lo) & 1) != 0);
      for (int ms = 0; ms < msteps; ++ms) {
        // ring slots of the T+2 rows of this group; the rows fetched for this step (all of them after a fresh start, else the
        // last T) are consumed in slot order, each with the barrier parity of its own lap around the ring
        int first = fresh ? cslot : cslot - 2;
        if (first < 0) first += p.ring_R;
        uint64_t ea[T + 2];
        int sl[T + 2];
        uint32_t ph[T + 2];
#pragma unroll
        for (int i = 0; i < T + 2; ++i) {
          int s_i = first + i;
          if (s_i >= p.ring_R) s_i -= p.ring_R;
          sl[i] = s_i;
          ea[i] = ring_desc0 + (uint64_t)s_i * group_desc;
          ph[i] = 0;
          if (fresh || i >= 2) {
            ph[i] = cphase;
            if (++cslot == p.ring_R) { cslot = 0; cphase ^= 1u; }
          }
        }
        const uint32_t newmask = fresh ? ((1u << (T + 2)) - 1u) : (((1u << (T + 2)) - 1u) & ~3u);
        // rows that no later group of this worker needs go back to the producer
        const bool last = ms + 1 == msteps;
        bool next_fresh = odd_tail && ms + 2 == msteps;
#pragma unroll
        for (int rk = 0; rk < C; ++rk) {
          y0r[rk] += T;
          if (y0r[rk] >= p.H) y0r[rk] -= p.H;
          next_fresh = next_fresh || y0r[rk] == 0;
        }
        const uint32_t relmask = last ? 0u : (next_fresh ? ((1u << (T + 2)) - 1u) : ((1u << T) - 1u));
        if (p.row_major) {
          for (int nt = 0; nt < p.n_tiles; ++nt)
            run_k_rows(ea, sl, ph, nt == 0 ? newmask : 0u, nt + 1 == p.n_tiles ? relmask : 0u);
        } else {
#pragma unroll
          for (int i = 0; i < T + 2; ++i)
            if ((newmask >> i) & 1u) V3_WAIT(1, mbar_wait(r_full(sl[i]), ph[i]));
          tc_fence_after();
          for (int nt = 0; nt < p.n_tiles; ++nt) run_k(ea);
          if (relmask != 0u) {
            if (elect_one()) {
#pragma unroll
              for (int i = 0; i < T + 2; ++i)
                if ((relmask >> i) & 1u) v3_commit<PAIR>(r_empty(sl[i]));
            }
            __syncwarp();
          }
        }
        fresh = next_fresh;
      }
    } else {
      const uint64_t ea[T + 2] = {};
      for (int unit = worker; unit < p.total_units; unit += workers) run_k(ea);
    }
    if (TAIL && it > 0) issue_tail(it - 1);
#ifdef PSSR_V3_WAITPROF
    wprof[4] = clock64() - wprof_t0;
    if ((p.dbg & 16) && lane == 0)
      for (int k = 0; k < 5; ++k) g_v3_trace[(blockIdx.x % 148) * 256 + 240 + k] = wprof[k];
#endif
  } else if (warp >= 4) {
    // ==================================== epilogue ==========================================
    const int q4 = warp & 3;
    const int eg = (warp - 4) >> 2;           // epilogue group 0 / 1
    const int row = q4 * 32 + lane;
    const int r = p.shuffle;
    const int npairs = (block_n + 63) / 64;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const bool relu = p.act == PSSR_ACT_RELU;
    const int steps = ROWS ? msteps * p.n_tiles : 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int it = 0;
    uint32_t rphase = 0;            // X: parity of this warp's residual-tile barrier
    float win[24];                  // TAIL + window layout: this thread's pixel, persistent over the four N tiles of a row group
#pragma unroll
    for (int k = 0; k < 24; ++k) win[k] = 0.f;
    for (int unit = worker; ROWS ? it < steps : unit < p.total_units; unit += workers, ++it) {
      const int n_tile = ROWS ? it % p.n_tiles : unit % p.n_tiles;
      const int um = ROWS ? 0 : (unit / p.n_tiles) * C + (int)rank;     // flat: this CTA's group of T tiles
      // ROWS: this CTA's row group
      int gn = 0, gy0 = 0, gx0 = 0;
      bool gvalid = true;
      if (ROWS) {
        const int g = group_of(it / p.n_tiles, (int)rank, gvalid);
        const int vr0 = g * T;
        const int v = vr0 / p.H;
        gy0 = vr0 - v * p.H;
        gn = v / p.NJ;
        gx0 = (v - gn * p.NJ) * 128;
      }
      const int buf = it & 1;
      const uint32_t par = (uint32_t)(it >> 1) & 1u;
      // X, rows mode with a TMA-store epilogue and at most one (tile, 64-column) item per warp and unit: the warp's residual tile
      // [32 pixels x 64 channels] is fetched by TMA into its second staging tile while the MMAs of the unit are still running
      // (loaded from global memory in the epilogue proper, the 128 bytes per thread cost the unit ~0.6 us of exposed latency)
      const bool resid_tma = X && ROWS && !TAIL && x_resid != nullptr && p.tma_store && T * npairs <= 2;
      const int my_item = T * npairs == 1 ? ((it & 1) == eg ? 0 : -1) : (eg < T * npairs ? eg : -1);
      if (resid_tma && my_item >= 0 && gvalid && lane == 0) {
        const int mt = my_item / npairs, pi = my_item - mt * npairs;
        const uint32_t rb = smem_u32(&rbars[warp - 4]);
        mbar_arrive_expect_tx(rb, 4096u);
        tma_load_4d(smem_base + p.stage_off + (uint32_t)(8 + warp - 4) * 4096u, p.tmaps + kTmOutLo, rb, n_tile * block_n + pi * 64, gx0 + q4 * 32,
                    gy0 + mt, gn);
      }
      mbar_wait(t_full(buf), par);
      tc_fence_after();
      if (warp == 4 && it < 31) V3_TRACE(64 + 2 * it);
      if (TAIL) {
        // ---- phase 1: relu(acc + bias) -> 16-bit, written back over the accumulator's own columns ------------------
        const uint32_t region = lane_addr + (uint32_t)(buf * 256 + eg * 128);
        const int nb0 = n_tile * 256 + eg * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(region + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          uint32_t o[16];
          if (x_tail_comp) {
            // hi = rn16(y) goes back to TMEM; 64 * (y - hi) as e5m2 goes to this pixel's row of the sub-position's A tile in
            // shared memory (64-byte rows, SWIZZLE_64B): 32 channels = two 16-byte chunks
            uint32_t l8[8];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bb = *reinterpret_cast<const float4*>(bias_s + nb0 + c * 32 + 4 * j4);
              const float y0 = fmaxf(__uint_as_float(v[4 * j4 + 0]) + bb.x, 0.f), y1 = fmaxf(__uint_as_float(v[4 * j4 + 1]) + bb.y, 0.f);
              const float y2 = fmaxf(__uint_as_float(v[4 * j4 + 2]) + bb.z, 0.f), y3 = fmaxf(__uint_as_float(v[4 * j4 + 3]) + bb.w, 0.f);
              o[2 * j4 + 0] = v3_pack2(y0, y1, p.fp16, false);
              o[2 * j4 + 1] = v3_pack2(y2, y3, p.fp16, false);
              const float2 h01 = v3_unpack2(o[2 * j4 + 0], p.fp16), h23 = v3_unpack2(o[2 * j4 + 1], p.fp16);
              l8[j4] = v3_e5m2x4((y0 - h01.x) * 64.f, (y1 - h01.y) * 64.f, (y2 - h23.x) * 64.f, (y3 - h23.y) * 64.f);
            }
            const uint32_t rowaddr = smem_base + lo8_off + (uint32_t)((eg * 2 + (c >> 1)) * 8192 + row * 64);
            const uint32_t sw = (uint32_t)(row >> 1) & 3u;
            const uint32_t j0 = (uint32_t)(c & 1) * 2u;
            v3_st_shared_v4(rowaddr + ((j0 ^ sw) << 4), l8[0], l8[1], l8[2], l8[3]);
            v3_st_shared_v4(rowaddr + (((j0 + 1u) ^ sw) << 4), l8[4], l8[5], l8[6], l8[7]);
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bb = *reinterpret_cast<const float4*>(bias_s + nb0 + c * 32 + 4 * j4);
              o[2 * j4 + 0] = v3_pack2(__uint_as_float(v[4 * j4 + 0]) + bb.x, __uint_as_float(v[4 * j4 + 1]) + bb.y, p.fp16, true);
              o[2 * j4 + 1] = v3_pack2(__uint_as_float(v[4 * j4 + 2]) + bb.z, __uint_as_float(v[4 * j4 + 3]) + bb.w, p.fp16, true);
            }
          }
          v3_tmem_st16(region + (uint32_t)(c * 16), o);
        }
        if (x_tail_comp) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the e5m2 tile is read by the tensor core
        v3_tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(p_full(buf));
          else if (x_tail_comp) v3_arrive_cluster_release(v3_mapa(p_full(buf), 0));
          else v3_arrive_cluster(v3_mapa(p_full(buf), 0));
        }
        if (warp == 4 && it < 8) V3_TRACE(232 + 2 * it);
        // ---- phase 2: the nine per-tap projections of this thread's pixel, two sub-positions per warp -------------
        int n, y, x;
        bool valid;
        if (ROWS) {
          n = gn; y = gy0; x = gx0 + row; valid = gvalid;      // T == 1
        } else {
          const int q = p.q_begin + um * 128 + row;
          const int vimg = q / p.IP;
          const int rem = q - vimg * p.IP;
          const int py = rem / p.P;
          const int px = rem - py * p.P;
          n = vimg / p.NJ;
          x = (vimg - n * p.NJ) * p.Wb + px - 1;
          y = py - 1;
          valid = (q < p.q_end) && px >= 1 && px < p.Wb + 1 && x < p.W && py >= 1 && py < p.H + 1;
        }
        mbar_wait(z_full(buf), par);
        tc_fence_after();
        if (warp == 4 && it < 8) V3_TRACE(233 + 2 * it);
        const int planes = r * r * 9;
        if (ROWS && p.tail_win48) {
          uint32_t zv0[16], zv1[16];
          if (x_tail_comp) {
            // accumulator columns per sub-position: [0, 16) hi x W_hi + lo x W, [16, 32) hi x W_lo
            uint32_t zl0[16], zl1[16];
            v3_tmem_ld16(region + 64u, zv0);
            v3_tmem_ld16(region + 64u + 16u, zl0);
            v3_tmem_ld16(region + 64u + 32u, zv1);
            v3_tmem_ld16(region + 64u + 48u, zl1);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              zv0[t] = __float_as_uint(__uint_as_float(zv0[t]) + __uint_as_float(zl0[t]));
              zv1[t] = __float_as_uint(__uint_as_float(zv1[t]) + __uint_as_float(zl1[t]));
            }
          } else {
            v3_tmem_ld16(region + 64u, zv0);
            v3_tmem_ld16(region + 64u + 16u, zv1);
            tmem_ld_wait();
          }
          switch (n_tile) {
            case 0: v3_tail_window_add<0, 0>(win, zv0); v3_tail_window_add<0, 1>(win, zv1); break;
            case 1: v3_tail_window_add<1, 0>(win, zv0); v3_tail_window_add<1, 1>(win, zv1); break;
            case 2: v3_tail_window_add<2, 0>(win, zv0); v3_tail_window_add<2, 1>(win, zv1); break;
            default: v3_tail_window_add<3, 0>(win, zv0); v3_tail_window_add<3, 1>(win, zv1); break;
          }
          if (n_tile == 3) {
            if (valid && !(p.dbg & 1)) {
              float* zp = p.tail_z + (((size_t)n * p.H + y) * 48 + (size_t)(eg * 24)) * p.W + x;
#pragma unroll
              for (int k = 0; k < 24; ++k) zp[(size_t)k * p.W] = win[k];
            }
#pragma unroll
            for (int k = 0; k < 24; ++k) win[k] = 0.f;
          }
        } else
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          uint32_t zv[16];
          if (x_tail_comp) {
            uint32_t zl[16];
            v3_tmem_ld16(region + 64u + (uint32_t)(sl * 32), zv);
            v3_tmem_ld16(region + 64u + (uint32_t)(sl * 32 + 16), zl);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 9; ++t) zv[t] = __float_as_uint(__uint_as_float(zv[t]) + __uint_as_float(zl[t]));
          } else {
            v3_tmem_ld16(region + 64u + (uint32_t)(sl * 16), zv);
            tmem_ld_wait();
          }
          if (valid && !(p.dbg & 1)) {
            const int sub = n_tile * 4 + eg * 2 + sl;
            float* zp = p.tail_z + (((size_t)n * p.H + y) * planes + (size_t)sub * 9) * p.W + x;
#pragma unroll
            for (int t = 0; t < 9; ++t) zp[(size_t)t * p.W] = __uint_as_float(zv[t]);
          }
        }
      } else {
        // two epilogue groups share the unit's (tile, 64-column) items; a unit with a single item goes to the groups in turn, so
        // that the epilogues of consecutive units overlap instead of leaving four warps idle
        for (int item = (T * npairs == 1 ? ((it & 1) == eg ? 0 : 1) : eg); item < T * npairs; item += 2) {
          const int mt = item / npairs;
          const int pi = item - mt * npairs;
          const int c_lo = pi * 64;
          const int c_hi = c_lo + 64 < block_n ? c_lo + 64 : block_n;
          int n, y, x;
          bool valid;
          if (ROWS) {
            n = gn; y = gy0 + mt; x = gx0 + row; valid = gvalid;
          } else if (p.cols_mode) {
            const int t = um * T + mt;
            const int sx = t % p.c_strips;
            const int t2 = t / p.c_strips;
            const int yb = t2 % p.c_yblocks;
            const int ng = t2 / p.c_yblocks;
            const int g = row >> 3;
            const int rr = p.cG == 2 ? (g >> 1) : g, img = p.cG == 2 ? (g & 1) : 0;
            n = ng * p.cG + img;
            y = yb * p.cR + rr;
            x = sx * 8 + (row & 7);
            valid = t < p.tiles_total && n < p.B;
          } else if (p.pad) {
            const int q = p.q_begin + (um * T + mt) * 128 + row;
            const int vimg = q / p.IP;
            const int rem = q - vimg * p.IP;
            const int py = rem / p.P;
            const int px = rem - py * p.P;
            n = vimg / p.NJ;
            x = (vimg - n * p.NJ) * p.Wb + px - 1;
            y = py - 1;
            valid = (q < p.q_end) && px >= 1 && px < p.Wb + 1 && x < p.W && py >= 1 && py < p.H + 1;
          } else {
            const int q = (um * T + mt) * 128 + row;
            n = q / p.IP;
            const int rem = q - n * p.IP;
            y = rem / p.P;
            x = rem - y * p.P;
            valid = q < p.q_end;
          }
          if (p.tma_store) {
            // ---- 16-bit tile [32 pixels x 64 channels] -> swizzled shared memory -> one TMA store (two when the 32 flat
            // pixels wrap to the next image row); padding / out-of-range positions are clipped by the TMA unit ----------
            const uint32_t stg = smem_base + p.stage_off + (uint32_t)(warp - 4) * 4096u;
            const uint32_t stg_hi = stg;
            const uint32_t taddr2 = lane_addr + (uint32_t)(buf * 256 + mt * block_n + c_lo);
            const int nb = n_tile * block_n + c_lo;
            // second output (out_lo): what the 16-bit rounding dropped, staged in a second tile of this warp and stored by a
            // second TMA store of the same box
            const uint32_t stg_lo = stg + 8u * 4096u;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile(s) have left the buffers
            __syncwarp();
#pragma unroll
            for (int cq = 0; cq < 2; ++cq) {
              uint32_t v[32];
              tmem_ld_32x32(taddr2 + (uint32_t)(cq * 32), v);
              tmem_ld_wait();
              float f[32];
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 bb = *reinterpret_cast<const float4*>(bias_s + nb + cq * 32 + 4 * j4);
                f[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + bb.x;
                f[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + bb.y;
                f[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + bb.z;
                f[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + bb.w;
              }
              if (X && resid_tma) {
                if (cq == 0 && gvalid) mbar_wait(smem_u32(&rbars[warp - 4]), rphase);
                if (gvalid) {
#pragma unroll
                  for (int j8 = 0; j8 < 4; ++j8) {
                    uint4 rv;
                    const uint32_t chunk = (uint32_t)(cq * 4 + j8) ^ (uint32_t)(lane & 7);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w)
                                 : "r"(stg_lo + (uint32_t)lane * 128u + chunk * 16u));
                    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const float2 r2 = v3_unpack2(rw[k], p.fp16);
                      f[8 * j8 + 2 * k] = fmaf(p.resid_scale, r2.x, f[8 * j8 + 2 * k]);
                      f[8 * j8 + 2 * k + 1] = fmaf(p.resid_scale, r2.y, f[8 * j8 + 2 * k + 1]);
                    }
                  }
                }
                if (cq == 1 && gvalid) rphase ^= 1u;
              } else if (x_resid != nullptr && valid) {
                const uint4* rp = reinterpret_cast<const uint4*>(x_resid + (((size_t)n * p.H + y) * p.W + x) * p.resid_cstride + p.resid_choff + nb + cq * 32);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  const uint4 rv = rp[j8];
                  const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 r2 = v3_unpack2(rw[k], p.fp16);
                    f[8 * j8 + 2 * k] = fmaf(p.resid_scale, r2.x, f[8 * j8 + 2 * k]);
                    f[8 * j8 + 2 * k + 1] = fmaf(p.resid_scale, r2.y, f[8 * j8 + 2 * k + 1]);
                  }
                }
              }
              if (p.act == PSSR_ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = v3_gelu(f[j]);
              }
              if (p.out_scale != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = (relu ? fmaxf(f[j], 0.f) : f[j]) * scale_s[nb + cq * 32 + j];
              }
              const bool relu_pack = relu && p.out_scale == nullptr;
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const uint32_t chunk = (uint32_t)(cq * 4 + j4) ^ (uint32_t)(lane & 7);
                uint32_t w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = v3_pack2(f[8 * j4 + 2 * k], f[8 * j4 + 2 * k + 1], p.fp16, relu_pack);
                v3_st_shared_v4(stg + (uint32_t)lane * 128u + chunk * 16u, w[0], w[1], w[2], w[3]);
                if (X && x_out_lo != nullptr) {
                  uint32_t wl[4];
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 hq = v3_unpack2(w[k], p.fp16);
                    const float a0 = relu_pack ? fmaxf(f[8 * j4 + 2 * k], 0.f) : f[8 * j4 + 2 * k];
                    const float a1 = relu_pack ? fmaxf(f[8 * j4 + 2 * k + 1], 0.f) : f[8 * j4 + 2 * k + 1];
                    wl[k] = v3_pack2(a0 - hq.x, a1 - hq.y, p.fp16, false);
                  }
                  v3_st_shared_v4(stg_lo + (uint32_t)lane * 128u + chunk * 16u, wl[0], wl[1], wl[2], wl[3]);
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const int n_pass = (X && x_out_lo != nullptr) ? 2 : 1;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass)
            if (pass < n_pass && lane == 0 && !(p.dbg & 1)) {
              const CUtensorMap* tmo = p.tmaps + (pass == 0 ? kTmOut : kTmOutLo);
              const uint32_t stg = pass == 0 ? stg_hi : stg_lo;
              if (ROWS) {
                if (gvalid) v3_tma_store_4d(tmo, stg, nb, gx0 + q4 * 32, gy0 + mt, gn);
              } else if (p.pad) {
                const int q0 = p.q_begin + (um * T + mt) * 128 + q4 * 32;
                const int vimg = q0 / p.IP;
                const int rem = q0 - vimg * p.IP;
                const int py = rem / p.P;
                const int px = rem - py * p.P;
                // a box is issued only when it covers real pixels (rows of the zero halo / beyond the batch are skipped);
                // its columns left of x = 0 or right of x = W - 1 are clipped by the TMA unit
                if (py >= 1 && py <= p.H && vimg < p.B && px <= p.W) v3_tma_store_4d(tmo, stg, nb, px - 1, py - 1, vimg);
                if (px + 32 > p.P) {              // the tail of the 32 pixels lies in the next padded row (maybe the next image)
                  const int q1 = q0 + (p.P - px);
                  const int vimg1 = q1 / p.IP;
                  const int py1 = (q1 - vimg1 * p.IP) / p.P;
                  if (py1 >= 1 && py1 <= p.H && vimg1 < p.B) v3_tma_store_4d(tmo, stg, nb, -1 - (p.P - px), py1 - 1, vimg1);
                }
              } else {
                if ((um * T + mt) * 128 + q4 * 32 < p.q_end) v3_tma_store_2d(tmo, stg, nb, (um * T + mt) * 128 + q4 * 32);
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            continue;
          }
          // sub-pixel / channel position of the item's first output column, advanced incrementally
          int sub = (n_tile * block_n + c_lo) / p.cps;
          int cc = n_tile * block_n + c_lo - sub * p.cps;
          int si = sub / r, sj = sub - si * r;
          const size_t pix00 = ((size_t)n * p.Hout + (size_t)(y * r)) * p.Wout + (size_t)(x * r);
          const uint32_t taddr = lane_addr + (uint32_t)(buf * 256 + mt * block_n);
          if (warp == 4 && it == 2) V3_TRACE(224);
          for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (warp == 4 && it == 2) V3_TRACE(225 + 2 * ((c0 - c_lo) >> 5));
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
                  const float4 bb = *reinterpret_cast<const float4*>(bias_s + nn + 4 * j4);
                  f[4 * j4 + 0] = __uint_as_float(v[h * 16 + 4 * j4 + 0]) + bb.x;
                  f[4 * j4 + 1] = __uint_as_float(v[h * 16 + 4 * j4 + 1]) + bb.y;
                  f[4 * j4 + 2] = __uint_as_float(v[h * 16 + 4 * j4 + 2]) + bb.z;
                  f[4 * j4 + 3] = __uint_as_float(v[h * 16 + 4 * j4 + 3]) + bb.w;
                }
                if (x_resid != nullptr) {
                  const uint4* rp = reinterpret_cast<const uint4*>(x_resid + (((size_t)n * p.H + y) * p.W + x) * p.resid_cstride + p.resid_choff + nn);
#pragma unroll
                  for (int j8 = 0; j8 < 2; ++j8) {
                    const uint4 rv = rp[j8];
                    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const float2 r2 = v3_unpack2(rw[k], p.fp16);
                      f[8 * j8 + 2 * k] = fmaf(p.resid_scale, r2.x, f[8 * j8 + 2 * k]);
                      f[8 * j8 + 2 * k + 1] = fmaf(p.resid_scale, r2.y, f[8 * j8 + 2 * k + 1]);
                    }
                  }
                }
                if (p.act == PSSR_ACT_GELU) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] = v3_gelu(f[j]);
                }
                if (p.out_scale != nullptr) {
                  if (relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                  }
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] *= scale_s[nn + j];
                }
                const bool relu_pack = relu && p.out_scale == nullptr;      // ReLU folded into the 16-bit conversion
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = v3_pack2(f[2 * j], f[2 * j + 1], p.fp16, relu_pack);
                uint32_t ol[8];
                if (x_out_lo != nullptr) {          // second output: what the 16-bit rounding of the activated value dropped
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float2 h = v3_unpack2(o[j], p.fp16);
                    const float a0 = relu_pack ? fmaxf(f[2 * j], 0.f) : f[2 * j], a1 = relu_pack ? fmaxf(f[2 * j + 1], 0.f) : f[2 * j + 1];
                    ol[j] = v3_pack2(a0 - h.x, a1 - h.y, p.fp16, false);
                  }
                }
                if (p.wide_store) {
                  const size_t pix = pix00 + (size_t)si * p.Wout + (size_t)sj;
                  v3_st_global_v8(p.out + pix * p.out_cstride + p.out_choff + cc, o);
                  if (x_out_lo != nullptr) v3_st_global_v8(x_out_lo + pix * p.lo_cstride + p.lo_choff + cc, ol);
                  if (warp == 4 && it == 2 && h == 1) V3_TRACE(226 + 2 * ((c0 - c_lo) >> 5));
                } else {
                  if (relu_pack && p.out_f32 != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                  }
#pragma unroll
                  for (int g = 0; g < 2; ++g) {
                    const int n8 = nn + g * 8;
                    if (n8 < p.n_valid) {
                      const int sub8 = n8 / p.cps;
                      const int cc8 = n8 - sub8 * p.cps;
                      const int si8 = sub8 / r, sj8 = sub8 - si8 * r;
                      const size_t pix = ((size_t)n * p.Hout + (size_t)(y * r + si8)) * p.Wout + (size_t)(x * r + sj8);
                      if (p.out != nullptr)
                        *reinterpret_cast<uint4*>(p.out + pix * p.out_cstride + p.out_choff + cc8) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                      if (x_out_lo != nullptr)
                        *reinterpret_cast<uint4*>(x_out_lo + pix * p.lo_cstride + p.lo_choff + cc8) = make_uint4(ol[4 * g], ol[4 * g + 1], ol[4 * g + 2], ol[4 * g + 3]);
                      if (p.out_f32 != nullptr) {
                        float4* d = reinterpret_cast<float4*>(p.out_f32 + pix * p.out_cstride + p.out_choff + cc8);
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
      // the accumulator buffer goes back to the MMA issuer.  Only the TMEM reads must be ordered before this arrival (they
      // are: tcgen05.wait::ld + the fence); a cluster-scope RELEASE would also wait for this thread's global stores.
      if (warp == 4 && it == 2) V3_TRACE(230);
      tc_fence_before();
      if (warp == 4 && it == 2) V3_TRACE(231);
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(t_empty(buf));
        else v3_arrive_cluster(v3_mapa(t_empty(buf), 0));
      }
      if (warp == 4 && it < 31) V3_TRACE(65 + 2 * it);
    }
  }

  if (warp >= 4 && lane == 0 && p.tma_store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging tiles fully written out
  tc_fence_before();
  __syncthreads();
  if (PAIR) v3_cluster_sync();     // the peer's shared memory and barriers stay alive until the leader's MMAs are done
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) v3_tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
    V3_TRACE(127);
  }
}

int v3_trace_fetch(long long* host, int n) {
  if (n > 148 * 256) n = 148 * 256;
  PSSR_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_v3_trace, sizeof(long long) * (size_t)n));
  return PSSR_OK;
}

// --------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn v3_encode_fn() {
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

typedef void (*V3Kernel)(const V3Params);
struct V3Variant { int T, G, RES, TAIL, PAIR, ROWS, X; V3Kernel fn; };
#define V3_VARIANT(T, G, RES, TAIL, PAIR, ROWS) {T, G, RES, TAIL, PAIR, ROWS, 0, conv_v3_kernel<T, G, RES != 0, TAIL != 0, PAIR != 0, ROWS != 0, false>}, \
  {T, G, RES, TAIL, PAIR, ROWS, 1, conv_v3_kernel<T, G, RES != 0, TAIL != 0, PAIR != 0, ROWS != 0, true>}
#define V3_VARIANTS_PR(PAIR, ROWS)                                                                                             \
  V3_VARIANT(1, 1, 0, 0, PAIR, ROWS), V3_VARIANT(1, 3, 0, 0, PAIR, ROWS), V3_VARIANT(1, 9, 1, 0, PAIR, ROWS),                  \
  V3_VARIANT(2, 1, 0, 0, PAIR, ROWS), V3_VARIANT(2, 3, 0, 0, PAIR, ROWS), V3_VARIANT(2, 9, 1, 0, PAIR, ROWS),                  \
  V3_VARIANT(1, 1, 0, 1, PAIR, ROWS), V3_VARIANT(1, 3, 0, 1, PAIR, ROWS)
static const V3Variant kV3Variants[] = {V3_VARIANTS_PR(0, 0), V3_VARIANTS_PR(1, 0), V3_VARIANTS_PR(0, 1), V3_VARIANTS_PR(1, 1)};
static const int kV3NumVariants = (int)(sizeof(kV3Variants) / sizeof(kV3Variants[0]));

// cols mode applies to 3x3 layers on maps narrower than 128 pixels that split into 16 groups of 8 pixels: H % 16 == 0 (16 rows x
// 8 columns of one image) or H == 8 (8 rows x 8 columns of two images).  Measured against flat mode (ResUNet, batch 64): equal at
// 32x32, 10-15 % faster at 64x64 (1.4x instead of 2x halo over-fetch, no padded columns), 30 % faster than the per-tap kernel at
// 16x16 / 8x8.
static bool v3_cols_geometry(const pssr_conv_desc_t& d, int* cR, int* cG) {
  if (getenv("PSSR_V3_NO_COLS") != nullptr) return false;
  bool any9 = false;
  for (int s = 0; s < d.n_segs; ++s) any9 = any9 || d.segs[s].taps == 9;
  const char* envw = getenv("PSSR_V3_COLS_MAXW");
  int maxw = envw ? atoi(envw) : 128;             // exclusive bound on the map width (W % 128 == 0 runs rows mode)
  if (maxw > 128) maxw = 128;
  if (!any9 || d.tail_z != nullptr || d.Wo >= maxw || d.Wo % 8 != 0) return false;
  int r, g;
  if (d.Ho % 16 == 0) { r = 16; g = 1; }
  else if (d.Ho == 8) { r = 8; g = 2; }
  else return false;
  if (cR) *cR = r;
  if (cG) *cG = g;
  return true;
}

bool v3_supported(const pssr_conv_desc_t& d) {
  if (getenv("PSSR_CONV_V1") != nullptr) return false;
  for (int s = 0; s < d.n_segs; ++s)
    if ((d.segs[s].taps != 1 && d.segs[s].taps != 9) || d.segs[s].dilation > 1) return false;     // atrous taps: box-per-tap kernel
  if (d.n % 32 != 0 || d.n < 32) return false;
  if (d.resid != nullptr && (d.shuffle != 1 || d.n != d.n_valid || d.n_valid % 16 != 0)) return false;   // chained partial sums of the atrous blocks
  bool any9 = false;
  for (int s = 0; s < d.n_segs; ++s) any9 = any9 || d.segs[s].taps == 9;
  // small feature maps in flat mode: the padded pixel space wastes (1 - HW/((H+2)(W+2))) of the MMAs (36 % at 8x8, 21 % at
  // 16x16) and the exact-tile kernel (conv_igemm.cu) is faster there
  // ... unless the map tiles exactly into 8-pixel column groups (cols mode, v3_cols_geometry)
  if (any9 && d.Wo < 32 && d.tail_z == nullptr && getenv("PSSR_V3_SMALL") == nullptr && !v3_cols_geometry(d, nullptr, nullptr)) return false;
  if (d.tail_z != nullptr) {
    const int cps = d.n_valid / (d.shuffle * d.shuffle);
    if (cps != 64 || d.n % 256 != 0 || !any9) return false;   // other tail shapes run unfused (PSSR_OP_TAIL)
  }
  return true;
}

static int v3_prepare_impl(const pssr_conv_desc_t& d, int dtype, ConvOp& op, bool force_flat);

int v3_prepare(const pssr_conv_desc_t& d, int dtype, ConvOp& op) {
  // rows mode keeps ALL source planes of T+2 image rows in shared memory; layers with many planes (e.g. a 3x3 over 128 channels
  // plus a 1x1 over 192 at width 128: 5 planes, 81 KB per row) do not fit and run in flat mode, which stages one plane at a time
  int rc = v3_prepare_impl(d, dtype, op, false);
  if (rc == PSSR_EUNSUP && d.Wo % 128 == 0 && d.tail_layout != PSSR_TAIL_WINDOW48) rc = v3_prepare_impl(d, dtype, op, true);
  return rc;
}

static int v3_prepare_impl(const pssr_conv_desc_t& d, int dtype, ConvOp& op, bool force_flat) {
  EncodeTiledFn enc = v3_encode_fn();
  PSSR_REQUIRE(enc != nullptr, PSSR_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  PSSR_REQUIRE(d.n_srcs >= 1 && d.n_srcs <= 4 && d.n_segs >= 1 && d.n_segs <= 6, PSSR_EINVAL, "conv: n_srcs/n_segs out of range");
  PSSR_REQUIRE(d.n_valid > 0 && d.n_valid <= d.n && d.n_valid % 8 == 0, PSSR_EUNSUP, "conv: n_valid=%d must be a multiple of 8 and <= n", d.n_valid);
  PSSR_REQUIRE(d.shuffle >= 1 && d.n_valid % (d.shuffle * d.shuffle) == 0, PSSR_EUNSUP, "conv: n_valid %% shuffle^2 != 0");
  const int cps = d.n_valid / (d.shuffle * d.shuffle);
  PSSR_REQUIRE(cps % 8 == 0, PSSR_EUNSUP, "conv: channels after pixel shuffle (%d) must be a multiple of 8", cps);
  PSSR_REQUIRE(d.shuffle == 1 || d.n == d.n_valid, PSSR_EUNSUP, "conv: padded N with pixel shuffle unsupported");
  PSSR_REQUIRE(d.out_cstride % 8 == 0 && d.out_choff % 8 == 0, PSSR_EUNSUP, "conv: output channel stride/offset must be multiples of 8");
  PSSR_REQUIRE(d.out != nullptr || d.out_f32 != nullptr || d.tail_z != nullptr, PSSR_EINVAL, "conv: no output buffer");
  PSSR_REQUIRE(d.bias != nullptr && ((uintptr_t)d.bias & 15) == 0, PSSR_EINVAL, "conv: bias missing or misaligned");
  const bool tail = d.tail_z != nullptr;
  if (tail) {
    PSSR_REQUIRE(d.tail_weight != nullptr && ((uintptr_t)d.tail_weight & 15) == 0, PSSR_EINVAL, "conv: tail_weight missing or misaligned");
    PSSR_REQUIRE(cps == 64 && d.act == PSSR_ACT_RELU && d.n == d.n_valid && d.n % 256 == 0, PSSR_EUNSUP,
                 "conv: the tensor-core tail needs C' = 64, ReLU and N %% 256 == 0");
  }

  V3Params& p = *reinterpret_cast<V3Params*>(op.kparams);
  static_assert(sizeof(V3Params) <= sizeof(op.kparams), "ConvOp::kparams too small");
  memset(&p, 0, sizeof(p));
  memset(op.tmaps, 0, sizeof(op.tmaps));
  op.variant = 3;

  int block_n = 256;
  while (d.n % block_n != 0) block_n >>= 1;
  p.block_n = block_n;
  p.n_tiles = d.n / block_n;
  p.n_valid = d.n_valid;
  p.n_total = d.n;
  p.H = d.Ho; p.W = d.Wo; p.B = d.B;
  p.pad = 0;
  for (int s2 = 0; s2 < d.n_segs; ++s2)
    if (d.segs[s2].taps == 9) p.pad = 1;
  p.NJ = (d.Wo + 127) / 128;
  p.Wb = (d.Wo + p.NJ - 1) / p.NJ;
  p.rows_mode = (p.pad && d.Wo % 128 == 0 && !force_flat && getenv("PSSR_V3_FLAT") == nullptr) ? 1 : 0;
  p.cols_mode = (p.pad && v3_cols_geometry(d, &p.cR, &p.cG)) ? 1 : 0;
  if (p.cols_mode) {
    p.c_strips = d.Wo / 8;
    p.c_yblocks = d.Ho / p.cR;
    p.tiles_total = ((d.B + p.cG - 1) / p.cG) * p.c_yblocks * p.c_strips;
    p.box_bytes = (uint32_t)(128 * 10 * p.cG * (p.cR + 2));
    p.tile_bytes = ((p.box_bytes + 1023u) / 1024u) * 1024u;
  }
  if (!p.pad) { p.NJ = 1; p.Wb = d.Wo; }
  p.P = p.Wb + 2 * p.pad;
  p.HP = d.Ho + 2 * p.pad;
  p.IP = p.HP * p.P;
  p.q_begin = p.pad * (p.P + 1);
  PSSR_REQUIRE((long long)d.B * p.NJ * p.IP < (1ll << 30), PSSR_EUNSUP, "conv: batch x padded image exceeds the 30-bit pixel index");
  p.q_end = d.B * p.NJ * p.IP - p.pad * (p.P + 1);
  p.total_vrows = d.B * p.NJ * d.Ho;

  int num_kb = 0, num_kb8 = 0;
  bool src_f8[4] = {false, false, false, false};
  p.n_segs = d.n_segs;
  for (int s = 0; s < d.n_segs; ++s) {
    const pssr_kseg_t& sg = d.segs[s];
    PSSR_REQUIRE(sg.src >= 0 && sg.src < d.n_srcs && sg.cblocks >= 1, PSSR_EINVAL, "conv: bad K segment");
    p.seg_src[s] = sg.src; p.seg_taps[s] = sg.taps; p.seg_cblocks[s] = sg.cblocks;
    if (sg.fmt == PSSR_SEG_E5M2) {
      // low-order correction terms as e5m2 x e5m2 (kind::f8f6f4): their own weight matrix and K-block numbering
      PSSR_REQUIRE(p.rows_mode && sg.taps == 9 && d.weights8 != nullptr && ((uintptr_t)d.weights8 & 15) == 0, PSSR_EUNSUP,
                   "conv: e5m2 segments need rows mode (width %% 128 == 0), 3x3 taps and weights8");
      p.seg_f8[s] = 1;
      src_f8[sg.src] = true;
      p.seg_kb0[s] = num_kb8;
      num_kb8 += sg.taps * sg.cblocks;
    } else {
      PSSR_REQUIRE(sg.fmt == PSSR_SEG_F16, PSSR_EINVAL, "conv: unknown segment format %d", sg.fmt);
      p.seg_kb0[s] = num_kb;
      num_kb += sg.taps * sg.cblocks;
    }
  }
  for (int s = 0; s < d.n_segs; ++s)
    PSSR_REQUIRE((p.seg_f8[s] != 0) == src_f8[d.segs[s].src], PSSR_EINVAL, "conv: a source is read both as 16-bit and as e5m2");
  PSSR_REQUIRE(num_kb >= 1, PSSR_EINVAL, "conv: no 16-bit K segment");
  p.num_kb = num_kb;
  // CTA pairs (cta_group::2): each CTA holds half of the weight rows, one MMA instruction drives both SMs
  const int sms = device_sm_count();
  int PAIR = (getenv("PSSR_V3_NOPAIR") == nullptr && sms % 2 == 0 && block_n >= 32) ? 1 : 0;
  const int C = PAIR ? 2 : 1;
  p.tap_bytes = (uint32_t)(block_n / C * 128);

  // T tiles per CTA and unit with T * block_n <= 256 (TMEM double-buffered: the epilogue of unit i overlaps the MMAs of unit i+1)
  int T = 256 / block_n;
  if (T > 2) T = 2;
  if (tail) T = 1;
  if (p.rows_mode && d.Ho % T != 0) T = 1;
  // small maps: with two tiles per CTA a layer of few pixels occupies a fraction of the machine (the 16^2 project convolutions of
  // RDNet, 12800 pixels: 25 CTA pairs of 74) -- one tile per CTA doubles the CTAs at work
  if (T == 2 && !p.rows_mode && getenv("PSSR_V3_T2_ALWAYS") == nullptr) {
    const long long px = (long long)d.B * d.Ho * d.Wo;
    const long long units2 = ((px + 256 * C - 1) / (256 * C)) * ((d.n + block_n - 1) / block_n);
    if (units2 < sms / C) T = 1;
  }
  const char* envT = getenv("PSSR_V3_T");
  if (envT && atoi(envT) == 1) T = 1;
  const bool tail_comp = tail && (d.tail_flags & PSSR_TAIL_COMP) != 0 && getenv("PSSR_V3_NO_TAIL_COMP") == nullptr;
  const int tailw_bytes = tail ? (tail_comp ? 6144 + 4 * 8192 : 2048) : 0;
  PSSR_REQUIRE(d.out_lo == nullptr || (d.out != nullptr && !tail && d.out_lo_cstride % 8 == 0 && d.out_lo_choff % 8 == 0), PSSR_EUNSUP,
               "conv: out_lo needs a 16-bit primary output and 8-channel aligned stride / offset");
  const int vec_bytes = ((4 * d.n * (d.out_scale != nullptr ? 2 : 1) + 1023) / 1024) * 1024;
  // epilogue through shared memory + TMA stores (full 128-byte lines, asynchronous) where the output is a plain 16-bit NHWC view
  // and every 32-pixel box lies inside the image (rows mode, 1x1 layers).  Measured on B200: store boxes with out-of-range
  // coordinates on several sides raise an illegal-instruction fault, so the flat 3x3 mode keeps its direct stores.
  p.tma_store = (!tail && d.shuffle == 1 && d.out != nullptr && d.out_f32 == nullptr && d.n_valid % 64 == 0 && block_n % 64 == 0 &&
                 (d.out_lo == nullptr || getenv("PSSR_V3_NO_TMA_STORE_LO") == nullptr) &&
                 (p.rows_mode || !p.pad || getenv("PSSR_V3_TMA_STORE_FLAT") != nullptr) && getenv("PSSR_V3_NO_TMA_STORE") == nullptr) ? 1 : 0;
  const long long smem_cap0 = 226 * 1024 - 1024 - vec_bytes - tailw_bytes;
  const char* envG = getenv("PSSR_V3_G");
  int planes = 0;
  for (int s2 = 0; s2 < d.n_segs; ++s2) planes += d.segs[s2].cblocks;
  p.slot_bytes = (uint32_t)(p.P * 128);
  p.slot16_bytes = (uint32_t)(((p.P * 32 + 127) / 128) * 128);
  p.slot8_bytes = (uint32_t)(((p.P * 64 + 127) / 128) * 128);
  // narrow planes (rows mode): a 1x1 segment over a 16-channel tensor (the im2col of a 1-channel input) is staged with 32-byte
  // pixels and multiplied with ONE K = 16 MMA instead of four (3 of the 40 K steps of Reconstruction.pre were zeros)
  bool src_k16[4] = {false, false, false, false};
  p.group_bytes = 0;
  p.group_tx = 0;
  for (int s2 = 0; s2 < d.n_segs; ++s2) {
    const pssr_kseg_t& sg = d.segs[s2];
    const pssr_src_t& src = d.srcs[sg.src];
    const bool k16 = p.rows_mode && sg.taps == 1 && sg.cblocks == 1 && src.cstride == 16 && src.channels <= 16 && getenv("PSSR_V3_NO_K16") == nullptr;
    p.seg_k16[s2] = k16 ? 1 : 0;
    if (k16) src_k16[sg.src] = true;
    // segments over the same source view share its staged rows (hi x W_hi and hi x W_lo of the compensated layers, or the 3x3
    // and the 1x1 residual of a depth-0 block): only the first one is loaded
    int alias = -1;
    for (int e = 0; e < s2 && alias < 0; ++e)
      if (p.seg_load[e] && d.segs[e].src == sg.src && d.segs[e].cblocks == sg.cblocks && p.seg_k16[e] == p.seg_k16[s2] && p.seg_f8[e] == p.seg_f8[s2] &&
          getenv("PSSR_V3_NO_ALIAS") == nullptr)
        alias = e;
    if (alias >= 0) {
      p.seg_load[s2] = 0;
      p.seg_plane[s2] = p.seg_plane[alias];
      continue;
    }
    p.seg_load[s2] = 1;
    p.seg_plane[s2] = p.group_bytes;
    p.group_bytes += (uint32_t)sg.cblocks * (k16 ? p.slot16_bytes : p.seg_f8[s2] ? p.slot8_bytes : p.slot_bytes);
    p.group_tx += (uint32_t)sg.cblocks * (uint32_t)p.P * (k16 ? 32u : p.seg_f8[s2] ? 64u : 128u);
  }
  for (int s2 = 0; s2 < d.n_segs; ++s2)
    PSSR_REQUIRE(p.seg_k16[s2] || !src_k16[d.segs[s2].src], PSSR_EUNSUP, "conv: a 16-channel source must be read by 1x1 segments only");
  (void)planes;
  int G = 1, RES = 0, b_stages = 0, ring_R = 0;
  long long a_total = 0, a_bytes = 0;
  const int T_first = T;
  const int want_stage = p.tma_store;
  int stage_bytes = 0;
  long long smem_cap = smem_cap0;
  // Shared-memory budget.  Candidates in order of preference: resident weights before streamed ones (rows mode), T = 2 before
  // T = 1, a deep ring (2T+2 rows: plain K order, all nine taps of a block issued in one go) before a shallow one (T+2 rows:
  // filter-row-major order, the issuer releases the top rows early and awaits the bottom rows late so that the next rows load
  // while the unit computes), TMA-store staging before direct stores.  Whatever is left after the weights deepens the ring.
  // per-warp 4 KB tiles: one per 16-bit output, or the second one as the landing tile of the epilogue residual
  const long long stage_total = ((d.out_lo != nullptr || d.resid != nullptr) ? 16 : 8) * 4096;
  auto try_config = [&](int Tc, bool staged, bool res_only, bool deep) -> bool {
    const long long cap = smem_cap0 - (staged ? stage_total : 0);
    long long a_min, ab = 0;
    if (p.rows_mode) {
      a_min = (long long)(deep ? 2 * Tc + 2 : Tc + 2) * p.group_bytes;      // deep: the current group and one group of prefetch
    } else {
      const long long a_raw = p.cols_mode ? (long long)Tc * p.tile_bytes : (p.pad ? (long long)(128 * Tc + 4 * p.P + 1) * 128 : 128LL * Tc * 128);
      ab = ((a_raw + 1023) / 1024) * 1024;
      a_min = 2 * ab;
    }
    a_min = ((a_min + 1023) / 1024) * 1024;
    const long long left = cap - a_min;
    int g_ = 0, res_ = 0, bst_ = 0;
    if (p.n_tiles == 1 && num_kb8 == 0 && (long long)num_kb * p.tap_bytes <= left && getenv("PSSR_V3_NORES") == nullptr && !tail) {
      res_ = 1; g_ = 9; bst_ = 1;
    } else if (!res_only) {
      // small N: a stage must hold many MMAs (a barrier poll costs ~100 unhidden cycles); N = 256: finer stages, deeper prefetch
      const int cand[2] = {3, 1};
      for (int ci = (block_n >= 256 ? 1 : 0); ci < 2; ++ci) {
        const int g = cand[ci];
        if (envG && atoi(envG) != g) continue;
        if (!p.pad && g != 1) continue;
        const long long stage = (long long)g * p.tap_bytes;
        const int need = g == 1 ? 3 : 2;
        if (left >= need * stage) {
          g_ = g;
          bst_ = (int)(left / stage);
          const int want = g == 1 ? 6 : 3;                      // rows mode: the rest goes to the ring
          if (p.rows_mode && bst_ > want) bst_ = want;
          break;
        }
      }
      if (g_ == 0 && left >= 2LL * p.tap_bytes) { g_ = 1; bst_ = (int)(left / p.tap_bytes); }
      if (bst_ > kV3MaxB) bst_ = kV3MaxB;
    }
    if (g_ == 0) return false;
    T = Tc; G = g_; RES = res_; b_stages = bst_; a_bytes = ab;
    p.row_major = (p.rows_mode && !deep) ? 1 : 0;
    p.tma_store = staged ? 1 : 0;
    stage_bytes = staged ? (int)stage_total : 0;
    smem_cap = cap;
    return true;
  };
  {
    bool found = false;
    for (int pass = (p.rows_mode ? 0 : 1); pass < 2 && !found; ++pass)
      for (int Tc = T_first; Tc >= 1 && !found; --Tc)
        for (int deep = 1; deep >= (p.rows_mode ? 0 : 1) && !found; --deep)
          for (int st = want_stage; st >= 0 && !found; --st) found = try_config(Tc, st != 0, pass == 0, deep != 0);
    if (!found) G = 0;
  }
  PSSR_REQUIRE(G != 0 && b_stages >= 1, PSSR_EUNSUP, "conv: image width %d needs more shared memory than available", d.Wo);
  p.b_bytes = (uint32_t)(G * (int)p.tap_bytes);
  p.b_stages = b_stages;
  const long long b_total_ll = RES ? (long long)num_kb * p.tap_bytes : (long long)b_stages * p.b_bytes;
  if (p.rows_mode) {
    ring_R = (int)((smem_cap - b_total_ll) / p.group_bytes);
    if (ring_R > kV3MaxR) ring_R = kV3MaxR;
    PSSR_REQUIRE(ring_R >= T + 2, PSSR_EUNSUP, "conv: the row ring does not fit in shared memory");
    p.ring_R = ring_R;
    p.ring_bytes = (uint32_t)((((long long)ring_R * p.group_bytes + 1023) / 1024) * 1024);
    if ((long long)p.ring_bytes + b_total_ll > smem_cap) { p.ring_R = --ring_R; p.ring_bytes = (uint32_t)((((long long)ring_R * p.group_bytes + 1023) / 1024) * 1024); }
    PSSR_REQUIRE(ring_R >= T + 2, PSSR_EUNSUP, "conv: the row ring does not fit in shared memory");
    a_total = p.ring_bytes;
    p.a_bytes = 0;
    p.off_px = 0;
    p.tile_step = 0;
    p.units_m = p.total_vrows / T;       // row groups
  } else {
    p.a_bytes = (uint32_t)a_bytes;
    a_total = 2 * a_bytes;
    if (p.cols_mode) {
      p.off_px = p.cG * 10 + 1;                       // staged (row 0, pixel 0) of the tile: one halo row of all cG images + one halo pixel
      p.tile_step = p.tile_bytes >> 4;
      p.units_m = (p.tiles_total + T * C - 1) / (T * C);
    } else if (p.pad) {
      p.off_px = 2 * p.P + 1;
      p.tile_step = 1024;
      p.units_m = (p.q_end - p.q_begin + 128 * T * C - 1) / (128 * T * C);
    } else {
      p.off_px = 0;
      p.tile_step = 1024;
      p.units_m = (p.q_end + 128 * T * C - 1) / (128 * T * C);
    }
  }
  p.total_units = p.units_m * p.n_tiles;
  p.dy_units = p.cols_mode ? (uint32_t)(p.cG * 10 * 8) : (uint32_t)(p.P * 8);
  p.a_sbo = p.cols_mode ? 1280u : 1024u;
  const char* envd = getenv("PSSR_DBG");
  p.dbg = envd ? atoi(envd) : 0;

  const CUtensorMapDataType tdt = dtype == PSSR_DT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  for (int s = 0; s < d.n_srcs; ++s) {
    const pssr_src_t& src = d.srcs[s];
    PSSR_REQUIRE(src.base != nullptr && ((uintptr_t)src.base & 15) == 0, PSSR_EINVAL, "conv: source %d base must be 16-byte aligned", s);
    PSSR_REQUIRE(src.cstride % 8 == 0 && src.channels >= 1 && src.channels <= src.cstride, PSSR_EUNSUP, "conv: source %d bad channel stride", s);
    PSSR_REQUIRE(src.H == d.Ho && src.W == d.Wo && src.B == d.B, PSSR_EINVAL, "conv: source %d geometry does not match the output", s);
    CUresult r;
    if (src_f8[s]) {
      // e5m2 source (rows mode only): one byte per element, 64-byte pixels
      PSSR_REQUIRE(src.cstride % 16 == 0, PSSR_EUNSUP, "conv: e5m2 source %d needs a channel stride that is a multiple of 16", s);
      cuuint64_t gdim[4] = {(cuuint64_t)src.channels, (cuuint64_t)src.W, (cuuint64_t)src.H, (cuuint64_t)src.B};
      cuuint64_t gstr[3] = {(cuuint64_t)src.cstride, (cuuint64_t)src.cstride * src.W, (cuuint64_t)src.cstride * src.W * src.H};
      cuuint32_t box[4] = {64u, (cuuint32_t)p.P, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = enc(&op.tmaps[s], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(src.base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (p.cols_mode) {
      // dims (channel, x, image, y): the box lands as [y][image][x][channel], rows of the cG images interleaved
      cuuint64_t gdim[4] = {(cuuint64_t)src.channels, (cuuint64_t)src.W, (cuuint64_t)src.B, (cuuint64_t)src.H};
      cuuint64_t gstr[3] = {(cuuint64_t)src.cstride * 2, (cuuint64_t)src.cstride * 2 * src.W * src.H, (cuuint64_t)src.cstride * 2 * src.W};
      cuuint32_t box[4] = {64, 10, (cuuint32_t)p.cG, (cuuint32_t)(p.cR + 2)};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = enc(&op.tmaps[s], tdt, 4, const_cast<void*>(src.base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (p.pad) {
      cuuint64_t gdim[4] = {(cuuint64_t)src.channels, (cuuint64_t)src.W, (cuuint64_t)src.H, (cuuint64_t)src.B};
      cuuint64_t gstr[3] = {(cuuint64_t)src.cstride * 2, (cuuint64_t)src.cstride * 2 * src.W, (cuuint64_t)src.cstride * 2 * src.W * src.H};
      cuuint32_t box[4] = {src_k16[s] ? 16u : 64u, (cuuint32_t)p.P, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = enc(&op.tmaps[s], tdt, 4, const_cast<void*>(src.base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              src_k16[s] ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
    cuuint32_t box[2] = {64, (cuuint32_t)(block_n / C)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&op.tmaps[kTmW], tdt, 2, const_cast<void*>(d.weights), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  }
  if (num_kb8 > 0) {
    const cuuint64_t ktot = (cuuint64_t)num_kb8 * 64;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)d.n};
    cuuint64_t gstr[1] = {ktot};
    cuuint32_t box[2] = {64, (cuuint32_t)(block_n / C)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&op.tmaps[kTmW8], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(d.weights8), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(weights8) failed with %d", (int)r);
  }
  op.smem_bytes = (int)(a_total + b_total_ll + 1024 + vec_bytes + tailw_bytes + stage_bytes);
  p.stage_off = (uint32_t)(a_total + b_total_ll + tailw_bytes + vec_bytes);
  if (p.tma_store) {
    // output view [pixels][channels]: channel 0 of the view at out + choff, n_valid channels visible, pixel pitch out_cstride
    uint8_t* obase = reinterpret_cast<uint8_t*>(d.out) + (size_t)d.out_choff * 2;
    CUresult r;
    if (p.pad) {
      cuuint64_t gdim[4] = {(cuuint64_t)d.n_valid, (cuuint64_t)d.Wo, (cuuint64_t)d.Ho, (cuuint64_t)d.B};
      cuuint64_t gstr[3] = {(cuuint64_t)d.out_cstride * 2, (cuuint64_t)d.out_cstride * 2 * d.Wo, (cuuint64_t)d.out_cstride * 2 * d.Wo * d.Ho};
      cuuint32_t box[4] = {64, 32, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = enc(&op.tmaps[kTmOut], tdt, 4, obase, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gdim[2] = {(cuuint64_t)d.n_valid, (cuuint64_t)d.Wo * d.Ho * d.B};
      cuuint64_t gstr[1] = {(cuuint64_t)d.out_cstride * 2};
      cuuint32_t box[2] = {64, 32};
      cuuint32_t estr[2] = {1, 1};
      r = enc(&op.tmaps[kTmOut], tdt, 2, obase, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(output) failed with %d", (int)r);
    if (d.out_lo != nullptr) {        // the same view over the second output
      uint8_t* lbase = reinterpret_cast<uint8_t*>(d.out_lo) + (size_t)d.out_lo_choff * 2;
      if (p.pad) {
        cuuint64_t gdim[4] = {(cuuint64_t)d.n_valid, (cuuint64_t)d.Wo, (cuuint64_t)d.Ho, (cuuint64_t)d.B};
        cuuint64_t gstr[3] = {(cuuint64_t)d.out_lo_cstride * 2, (cuuint64_t)d.out_lo_cstride * 2 * d.Wo, (cuuint64_t)d.out_lo_cstride * 2 * d.Wo * d.Ho};
        cuuint32_t box[4] = {64, 32, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        r = enc(&op.tmaps[kTmOutLo], tdt, 4, lbase, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint64_t gdim[2] = {(cuuint64_t)d.n_valid, (cuuint64_t)d.Wo * d.Ho * d.B};
        cuuint64_t gstr[1] = {(cuuint64_t)d.out_lo_cstride * 2};
        cuuint32_t box[2] = {64, 32};
        cuuint32_t estr[2] = {1, 1};
        r = enc(&op.tmaps[kTmOutLo], tdt, 2, lbase, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(out_lo) failed with %d", (int)r);
    } else if (d.resid != nullptr && p.rows_mode) {
      // the epilogue residual as 32-pixel x 64-channel boxes (rows mode; other modes read it from global memory)
      uint8_t* rbase = reinterpret_cast<uint8_t*>(const_cast<void*>(d.resid)) + (size_t)d.resid_choff * 2;
      cuuint64_t gdim[4] = {(cuuint64_t)d.n_valid, (cuuint64_t)d.Wo, (cuuint64_t)d.Ho, (cuuint64_t)d.B};
      cuuint64_t gstr[3] = {(cuuint64_t)d.resid_cstride * 2, (cuuint64_t)d.resid_cstride * 2 * d.Wo, (cuuint64_t)d.resid_cstride * 2 * d.Wo * d.Ho};
      cuuint32_t box[4] = {64, 32, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = enc(&op.tmaps[kTmOutLo], tdt, 4, rbase, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      PSSR_REQUIRE(r == CUDA_SUCCESS, PSSR_ECUDA, "cuTensorMapEncodeTiled(resid) failed with %d", (int)r);
    }
  }

  p.bias = d.bias;
  p.out_scale = d.out_scale;
  p.out = reinterpret_cast<uint16_t*>(d.out);
  p.out_f32 = d.out_f32;
  p.tail_w = d.tail_weight;
  p.tail_z = d.tail_z;
  p.tail_win48 = (tail && d.tail_layout == PSSR_TAIL_WINDOW48) ? 1 : 0;
  p.tail_comp = tail_comp ? 1 : 0;
  p.tailw_bytes = (uint32_t)tailw_bytes;
  PSSR_REQUIRE(d.resid == nullptr || d.out_lo == nullptr, PSSR_EUNSUP, "conv: resid and out_lo share the second staging tile");
  PSSR_REQUIRE(d.resid == nullptr || (!tail && d.shuffle == 1 && d.n == d.n_valid && d.n_valid % 16 == 0 && d.resid_cstride % 8 == 0 &&
                                      d.resid_choff % 8 == 0 && ((uintptr_t)d.resid & 15) == 0),
               PSSR_EUNSUP, "conv: the epilogue residual needs shuffle == 1, unpadded N and 16-byte aligned channel slices");
  p.resid = reinterpret_cast<const uint16_t*>(d.resid);
  p.resid_cstride = d.resid_cstride;
  p.resid_choff = d.resid_choff;
  p.resid_scale = d.resid_scale;
  p.out_lo = reinterpret_cast<uint16_t*>(d.out_lo);
  p.lo_cstride = d.out_lo_cstride;
  p.lo_choff = d.out_lo_choff;
  PSSR_REQUIRE(!p.tail_win48 || (p.rows_mode && d.shuffle == 4 && p.n_tiles == 4), PSSR_EUNSUP,
               "conv: PSSR_TAIL_WINDOW48 needs scale 4, 64 channels per sub-position and Wo %% 128 == 0");
  p.out_cstride = d.out_cstride;
  p.out_choff = d.out_choff;
  p.shuffle = d.shuffle;
  p.cps = cps;
  p.act = d.act;
  p.fp16 = dtype == PSSR_DT_FP16 ? 1 : 0;
  p.Hout = d.Ho * d.shuffle;
  p.Wout = d.Wo * d.shuffle;
  p.wide_store = (d.out != nullptr && d.out_f32 == nullptr && cps % 16 == 0 && d.out_choff % 16 == 0 && d.out_cstride % 16 == 0 &&
                  d.n_valid % 16 == 0 && ((uintptr_t)d.out & 31) == 0 &&
                  (d.out_lo == nullptr || (d.out_lo_choff % 16 == 0 && d.out_lo_cstride % 16 == 0 && ((uintptr_t)d.out_lo & 31) == 0))) ? 1 : 0;
  int workers = sms / C;
  if (p.rows_mode) {
    int gmax = p.units_m / C;          // every worker gets at least one row group per CTA
    if (gmax < 1) gmax = 1;
    if (workers > gmax) workers = gmax;
  } else if (workers > p.total_units) {
    workers = p.total_units;
  }
  op.grid = workers * C;
  op.cluster = C;
  op.kernel_index = -1;
  const int need_x = (num_kb8 > 0 || d.resid != nullptr || d.out_lo != nullptr || tail_comp) ? 1 : 0;
  for (int i = 0; i < kV3NumVariants; ++i)
    if (kV3Variants[i].T == T && kV3Variants[i].G == G && kV3Variants[i].RES == RES && kV3Variants[i].TAIL == (tail ? 1 : 0) &&
        kV3Variants[i].PAIR == PAIR && kV3Variants[i].ROWS == p.rows_mode && kV3Variants[i].X == need_x)
      op.kernel_index = i;
  PSSR_REQUIRE(op.kernel_index >= 0, PSSR_EUNSUP, "conv: no kernel variant for T=%d G=%d RES=%d TAIL=%d PAIR=%d", T, G, RES, (int)tail, PAIR);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    for (int i = 0; i < kV3NumVariants; ++i)
      PSSR_CHECK_CUDA(cudaFuncSetAttribute(kV3Variants[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
  }
  if (getenv("PSSR_V3_VERBOSE") != nullptr)
    fprintf(stderr, "v3: %dx%d n=%d kb=%d block_n=%d T=%d G=%d RES=%d TAIL=%d PAIR=%d rows=%d/%d cols=%d ring=%d b_stages=%d a_bytes=%u smem=%d units=%d grid=%d\n",
            d.Ho, d.Wo, d.n, num_kb, block_n, T, G, RES, (int)tail, PAIR, p.rows_mode, p.row_major, p.cols_mode, p.ring_R, b_stages, p.a_bytes, op.smem_bytes, p.total_units,
            op.grid);
  return PSSR_OK;
}

int v3_launch(const ConvOp& op, const void* tmaps_dev, cudaStream_t stream) {
  V3Params p = *reinterpret_cast<const V3Params*>(op.kparams);
  p.tmaps = reinterpret_cast<const CUtensorMap*>(tmaps_dev);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)op.grid, 1, 1);
  cfg.blockDim = dim3(kV3Threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)op.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (op.cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)op.cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  static const bool pdl = getenv("PSSR_V3_NO_PDL") == nullptr;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  PSSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kV3Variants[op.kernel_index].fn, p));
  count_launch();
  return PSSR_OK;
}

}  // namespace pssr
