// Family 3: overlap-weighted tile stitch.  Replaces `_patch_images` (pssr/util.py:116-137) and the
// uint8 cast at pssr/util.py:100.  The reference accumulates every tile into a float64 canvas plus a
// float64 count canvas and divides; here every OUTPUT pixel gathers its contributors directly
// (tile rows/cols whose kept span covers it), so there are no atomics and no intermediate canvases:
// integer sum / integer count, truncated, equals the reference's float64 quotient truncated to uint8.
// Inner tiles drop `margin` pixels on interior edges (util.py:131); pixels nobody covers stay 0
// (count forced to 1, util.py:136).
#include <stdlib.h>
#include "common.cuh"

namespace pssr {

__global__ void stitch_kernel(const uint8_t* __restrict__ tiles, uint8_t* __restrict__ sheets, int n_rows, int n_cols,
                              int T, int step, int margin, int out_h, int out_w) {
  const int stack = blockIdx.z;
  const size_t tile_px = (size_t)T * T;
  const uint8_t* tb = tiles + (size_t)stack * n_rows * n_cols * tile_px;
  uint8_t* ob = sheets + (size_t)stack * out_h * out_w;
  const int Y = blockIdx.y;
  // tile rows whose kept span [r*step + m0, r*step + T - m1) contains Y
  int r_lo = Y - T + 1 <= 0 ? 0 : (Y - T + step) / step;  // ceil((Y-T+1)/step)
  int r_hi = Y / step;
  if (r_hi > n_rows - 1) r_hi = n_rows - 1;
  for (int X = blockIdx.x * blockDim.x + threadIdx.x; X < out_w; X += gridDim.x * blockDim.x) {
    int c_lo = X - T + 1 <= 0 ? 0 : (X - T + step) / step;
    int c_hi = X / step;
    if (c_hi > n_cols - 1) c_hi = n_cols - 1;
    int sum = 0, cnt = 0;
    for (int r = r_lo; r <= r_hi; ++r) {
      const int ly = Y - r * step;
      const int m0 = r != 0 ? margin : 0, m1 = r != n_rows - 1 ? margin : 0;
      if (ly < m0 || ly >= T - m1) continue;
      for (int c = c_lo; c <= c_hi; ++c) {
        const int lx = X - c * step;
        const int n0 = c != 0 ? margin : 0, n1 = c != n_cols - 1 ? margin : 0;
        if (lx < n0 || lx >= T - n1) continue;
        sum += tb[(size_t)(r * n_cols + c) * tile_px + (size_t)ly * T + lx];
        ++cnt;
      }
    }
    ob[(size_t)Y * out_w + X] = (uint8_t)(cnt > 0 ? sum / cnt : 0);
  }
}

// Vector path (T, step, margin, sheet width all multiples of 4 and 4-byte aligned buffers): every boundary of a kept span is a
// multiple of 4, so the four pixels of an aligned group share their contributors -- one 32-bit load per contributor, one
// 32-bit store per group, no per-pixel division (count is 1, 2 or 4 except for overlaps > T/2).
__device__ __forceinline__ uint32_t stitch_group(const uint8_t* __restrict__ tb, int n_rows, int n_cols, int T, int step, int margin,
                                                 size_t tile_px, int Y, int X, int r_lo, int r_hi) {
  const int c_lo = X - T + 1 <= 0 ? 0 : (X - T + step) / step;
  const int c_hi = min(X / step, n_cols - 1);
  int s0 = 0, s1 = 0, s2 = 0, s3 = 0, cnt = 0;
  for (int r = r_lo; r <= r_hi; ++r) {
    const int ly = Y - r * step;
    const int m0 = r != 0 ? margin : 0, m1 = r != n_rows - 1 ? margin : 0;
    if (ly < m0 || ly >= T - m1) continue;
    for (int c = c_lo; c <= c_hi; ++c) {
      const int lx = X - c * step;
      const int n0 = c != 0 ? margin : 0, n1 = c != n_cols - 1 ? margin : 0;
      if (lx < n0 || lx >= T - n1) continue;
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(tb + (size_t)(r * n_cols + c) * tile_px + (size_t)ly * T + lx));
      s0 += w & 255u; s1 += (w >> 8) & 255u; s2 += (w >> 16) & 255u; s3 += w >> 24;
      ++cnt;
    }
  }
  if (cnt == 2) { s0 >>= 1; s1 >>= 1; s2 >>= 1; s3 >>= 1; }
  else if (cnt == 4) { s0 >>= 2; s1 >>= 2; s2 >>= 2; s3 >>= 2; }
  else if (cnt > 1) { s0 /= cnt; s1 /= cnt; s2 /= cnt; s3 /= cnt; }
  return (uint32_t)s0 | ((uint32_t)s1 << 8) | ((uint32_t)s2 << 16) | ((uint32_t)s3 << 24);
}

// grid (x: 16-pixel column blocks, y: output rows, z: stacks); a thread produces 4 groups = 16 pixels (their loads are independent)
__global__ void __launch_bounds__(128) stitch_vec4_kernel(const uint8_t* __restrict__ tiles, uint8_t* __restrict__ sheets, int n_rows,
                                                          int n_cols, int T, int step, int margin, int out_h, int out_w) {
  const int stack = blockIdx.z, Y = blockIdx.y;
  const size_t tile_px = (size_t)T * T;
  const uint8_t* tb = tiles + (size_t)stack * n_rows * n_cols * tile_px;
  uint8_t* orow = sheets + ((size_t)stack * out_h + Y) * out_w;
  const int r_lo = Y - T + 1 <= 0 ? 0 : (Y - T + step) / step;
  const int r_hi = min(Y / step, n_rows - 1);
  const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (X0 >= out_w) return;
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = X0 + 4 * k < out_w ? stitch_group(tb, n_rows, n_cols, T, step, margin, tile_px, Y, X0 + 4 * k, r_lo, r_hi) : 0u;
  if (X0 + 16 <= out_w && ((reinterpret_cast<uintptr_t>(orow) + X0) & 15) == 0) {
    *reinterpret_cast<uint4*>(orow + X0) = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (X0 + 4 * k < out_w) *reinterpret_cast<uint32_t*>(orow + X0 + 4 * k) = o[k];
  }
}

// 16-pixel path (T, step, margin, sheet width all multiples of 16, 16-byte aligned buffers): every span boundary is a multiple
// of 16, so an aligned 16-pixel vector shares its contributors.  One 16-byte load per contributor; the byte sums are kept in
// 16-bit lanes (two per 32-bit word, <= 9 * 255), the division is a shift for 1 / 2 / 4 contributors.
__device__ __forceinline__ void stitch_acc16(uint32_t (&lo)[4], uint32_t (&hi)[4], const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) { lo[k] += w[k] & 0x00FF00FFu; hi[k] += (w[k] >> 8) & 0x00FF00FFu; }
}
__global__ void __launch_bounds__(128) stitch_vec16_kernel(const uint8_t* __restrict__ tiles, uint8_t* __restrict__ sheets, int n_rows,
                                                           int n_cols, int T, int step, int margin, int out_h, int out_w) {
  const int stack = blockIdx.z, Y = blockIdx.y;
  const int X = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (X >= out_w) return;
  const size_t tile_px = (size_t)T * T;
  const uint8_t* tb = tiles + (size_t)stack * n_rows * n_cols * tile_px;
  const int r_lo = Y - T + 1 <= 0 ? 0 : (Y - T + step) / step;
  const int r_hi = min(Y / step, n_rows - 1);
  const int c_lo = X - T + 1 <= 0 ? 0 : (X - T + step) / step;
  const int c_hi = min(X / step, n_cols - 1);
  uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
  int cnt = 0;
  for (int r = r_lo; r <= r_hi; ++r) {
    const int ly = Y - r * step;
    const int m0 = r != 0 ? margin : 0, m1 = r != n_rows - 1 ? margin : 0;
    if (ly < m0 || ly >= T - m1) continue;
    for (int c = c_lo; c <= c_hi; ++c) {
      const int lx = X - c * step;
      const int n0 = c != 0 ? margin : 0, n1 = c != n_cols - 1 ? margin : 0;
      if (lx < n0 || lx >= T - n1) continue;
      stitch_acc16(lo, hi, __ldg(reinterpret_cast<const uint4*>(tb + (size_t)(r * n_cols + c) * tile_px + (size_t)ly * T + lx)));
      ++cnt;
    }
  }
  uint32_t o[4];
  if (cnt <= 1 || cnt == 2 || cnt == 4) {
    const int sh = cnt == 2 ? 1 : (cnt == 4 ? 2 : 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = ((lo[k] >> sh) & 0x00FF00FFu) | (((hi[k] >> sh) & 0x00FF00FFu) << 8);
  } else {
    // floor(s / cnt) = (s * ceil(2^16 / cnt)) >> 16 exactly for s <= 9 * 255, cnt <= 9
    const uint32_t m = (65536u + (uint32_t)cnt - 1u) / (uint32_t)cnt;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a = ((lo[k] & 0xFFFFu) * m) >> 16, b = ((lo[k] >> 16) * m) >> 16;
      const uint32_t c2 = ((hi[k] & 0xFFFFu) * m) >> 16, d2 = ((hi[k] >> 16) * m) >> 16;
      o[k] = a | (c2 << 8) | (b << 16) | (d2 << 24);
    }
  }
  *reinterpret_cast<uint4*>(sheets + ((size_t)stack * out_h + Y) * out_w + X) = make_uint4(o[0], o[1], o[2], o[3]);
}

// Band kernel (same preconditions as the 16-pixel path).  The 16-pixel kernel above spends its time on the per-thread integer
// divisions and span tests (ncu round 1: ALU 57 %, 2.4 TB/s): the contributors of a pixel are the product of a per-ROW set and a
// per-COLUMN set, so a CTA of 128 threads takes a band of R output rows x 2048 columns, resolves the R row sets once (one thread
// per row, into shared memory) and every thread resolves its own column set once; the row loop is then loads, packed adds and
// one store.  Offsets are bytes inside one stack of tiles (< 2^31: checked on the host).
static constexpr int kBandRows = 16;
__global__ void __launch_bounds__(128) stitch_band16_kernel(const uint8_t* __restrict__ tiles, uint8_t* __restrict__ sheets, int n_rows,
                                                            int n_cols, int T, int step, int margin, int out_h, int out_w, int band) {
  __shared__ int s_roff[kBandRows][3];
  __shared__ int s_nr[kBandRows];
  const int stack = blockIdx.z, Y0 = blockIdx.y * band;
  const int X = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int tile_px = T * T;
  if (threadIdx.x < band) {
    const int Y = Y0 + threadIdx.x;
    int n = 0;
    if (Y < out_h) {
      const int r_lo = Y - T + 1 <= 0 ? 0 : (Y - T + step) / step;
      const int r_hi = min(Y / step, n_rows - 1);
      for (int r = r_lo; r <= r_hi && n < 3; ++r) {
        const int ly = Y - r * step;
        const int m0 = r != 0 ? margin : 0, m1 = r != n_rows - 1 ? margin : 0;
        if (ly < m0 || ly >= T - m1) continue;
        s_roff[threadIdx.x][n++] = r * n_cols * tile_px + ly * T;
      }
    }
    s_nr[threadIdx.x] = n;
  }
  int coff[3] = {0, 0, 0}, nc = 0;
  if (X < out_w) {
    const int c_lo = X - T + 1 <= 0 ? 0 : (X - T + step) / step;
    const int c_hi = min(X / step, n_cols - 1);
    for (int c = c_lo; c <= c_hi && nc < 3; ++c) {
      const int lx = X - c * step;
      const int n0 = c != 0 ? margin : 0, n1 = c != n_cols - 1 ? margin : 0;
      if (lx < n0 || lx >= T - n1) continue;
      const int off = c * tile_px + lx;
      if (nc == 0) coff[0] = off; else if (nc == 1) coff[1] = off; else coff[2] = off;
      ++nc;
    }
  }
  __syncthreads();
  if (X >= out_w) return;
  const uint8_t* tb = tiles + (size_t)stack * n_rows * n_cols * tile_px;
  uint8_t* ob = sheets + ((size_t)stack * out_h + Y0) * out_w + X;
  const int rows = min(band, out_h - Y0);
#pragma unroll 4
  for (int i = 0; i < rows; ++i) {
    const int nr = s_nr[i];
    uint4 out;
    if (nr == 1 && nc == 1) {                       // interior of a tile: a copy
      out = __ldg(reinterpret_cast<const uint4*>(tb + s_roff[i][0] + coff[0]));
    } else {
      uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
      for (int m = 0; m < nr; ++m) {
        const uint8_t* rb = tb + s_roff[i][m];
        stitch_acc16(lo, hi, nc > 0 ? __ldg(reinterpret_cast<const uint4*>(rb + coff[0])) : make_uint4(0, 0, 0, 0));
        if (nc > 1) stitch_acc16(lo, hi, __ldg(reinterpret_cast<const uint4*>(rb + coff[1])));
        if (nc > 2) stitch_acc16(lo, hi, __ldg(reinterpret_cast<const uint4*>(rb + coff[2])));
      }
      const int cnt = nr * nc;
      uint32_t o[4];
      if (cnt <= 1 || cnt == 2 || cnt == 4) {
        const int sh = cnt == 2 ? 1 : (cnt == 4 ? 2 : 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = ((lo[k] >> sh) & 0x00FF00FFu) | (((hi[k] >> sh) & 0x00FF00FFu) << 8);
      } else {
        // floor(s / cnt) = (s * ceil(2^16 / cnt)) >> 16 exactly for s <= 9 * 255, cnt <= 9
        const uint32_t m = (65536u + (uint32_t)cnt - 1u) / (uint32_t)cnt;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = ((lo[k] & 0xFFFFu) * m) >> 16, b = ((lo[k] >> 16) * m) >> 16;
          const uint32_t c2 = ((hi[k] & 0xFFFFu) * m) >> 16, d2 = ((hi[k] >> 16) * m) >> 16;
          o[k] = a | (c2 << 8) | (b << 16) | (d2 << 24);
        }
      }
      out = make_uint4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<uint4*>(ob + (size_t)i * out_w) = out;
  }
}

}  // namespace pssr

using namespace pssr;

extern "C" int pssr_stitch(const uint8_t* tiles, uint8_t* sheets, int32_t n_stacks, int32_t n_rows, int32_t n_cols,
                           int32_t tile, int32_t overlap, int32_t margin, void* stream) {
  PSSR_REQUIRE(tiles && sheets, PSSR_EINVAL, "stitch: null pointer");
  PSSR_REQUIRE(n_stacks >= 1 && n_rows >= 1 && n_cols >= 1 && tile >= 1, PSSR_EINVAL, "stitch: bad sizes");
  PSSR_REQUIRE(overlap >= 0 && overlap < tile, PSSR_EINVAL, "stitch: overlap must be in [0, tile)");
  // same contract as reassemble_sheets (util.py:76-77)
  PSSR_REQUIRE(margin >= 0 && margin <= overlap, PSSR_EINVAL, "The value of margin cannot be greater than overlap. Given %d and %d respectively.", margin, overlap);
  const int step = tile - overlap;
  const int out_h = n_rows * step + overlap, out_w = n_cols * step + overlap;
  PSSR_REQUIRE(n_stacks <= 65535 && out_h <= 65535, PSSR_EUNSUP, "stitch: sheet too large for the launch grid");
  const bool vec = tile % 4 == 0 && step % 4 == 0 && margin % 4 == 0 && out_w % 4 == 0 &&
                   (((uintptr_t)tiles | (uintptr_t)sheets) & 3) == 0 && getenv("PSSR_STITCH_SCALAR") == nullptr;
  const bool vec16 = vec && tile % 16 == 0 && step % 16 == 0 && margin % 16 == 0 && out_w % 16 == 0 &&
                     (((uintptr_t)tiles | (uintptr_t)sheets) & 15) == 0 && tile / step < 3 && getenv("PSSR_STITCH_VEC4") == nullptr;
  const bool band = vec16 && (long long)n_rows * n_cols * tile * tile < (1ll << 31) && getenv("PSSR_STITCH_NOBAND") == nullptr;
  if (band) {
    // rows per CTA: as many as leave >= ~24 CTAs per SM in the grid (a single 3968^2 sheet: 4 rows; eight sheets: 16)
    const long long xb = (out_w / 16 + 127) / 128;
    int rows_per_cta = kBandRows;
    while (rows_per_cta > 2 && xb * ((out_h + rows_per_cta - 1) / rows_per_cta) * n_stacks < 24LL * device_sm_count()) rows_per_cta >>= 1;
    const char* envb = getenv("PSSR_STITCH_BAND");
    if (envb != nullptr && atoi(envb) >= 1 && atoi(envb) <= kBandRows) rows_per_cta = atoi(envb);
    dim3 grid((unsigned)xb, (out_h + rows_per_cta - 1) / rows_per_cta, n_stacks);
    stitch_band16_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tiles, sheets, n_rows, n_cols, tile, step, margin, out_h, out_w,
                                                                                   rows_per_cta);
  } else if (vec16) {
    dim3 grid((out_w / 16 + 127) / 128, out_h, n_stacks);
    stitch_vec16_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tiles, sheets, n_rows, n_cols, tile, step, margin, out_h, out_w);
  } else if (vec) {
    dim3 grid((out_w + 16 * 128 - 1) / (16 * 128), out_h, n_stacks);
    stitch_vec4_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tiles, sheets, n_rows, n_cols, tile, step, margin, out_h, out_w);
  } else {
    dim3 grid((out_w + 255) / 256, out_h, n_stacks);
    stitch_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tiles, sheets, n_rows, n_cols, tile, step, margin, out_h, out_w);
  }
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}
