// Family 4: scoring.
//   pssr_metric_sums     : sum (a-b)^2 (-> mse, pixel, PSNR) and the SSIM map sum over the interior
//                          pssr/predict.py:193-203; skimage.metrics.peak_signal_noise_ratio /
//                          structural_similarity (7x7 uniform window, sample covariance, float64,
//                          crop 3) -- restated in oracle/thirdparty.py
//   pssr_normalize_preds : pssr/util.py:139-191 (+ _normalize_minmax :193-205).  Inputs are uint8, so
//                          every statistic the reference takes (percentiles, means, cov, var, min) is
//                          a function of two 256-bin histograms and the integer sum of products, and
//                          the per-pixel maps are two 256-entry lookup tables.
// All window / image sums are exact integers; reductions use warp shuffles; the per-block SSIM
// partials are summed in a fixed order so results are run-to-run deterministic.
#include "common.cuh"

namespace pssr {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// -------------------------------------------------------------------------- SSIM + SSE
// Row-streaming design.  A CTA of 128 threads owns a strip of 512 input columns (4 per thread, one 32-bit load per image and
// row) and a chunk of output rows.  Per input row every thread updates the VERTICAL 7-row running sums of a, b, a^2 + b^2 and
// ab of its four columns (add the new row, subtract the row seven above), publishes them to shared memory (double-buffered,
// one barrier per row), slides the HORIZONTAL 7-window over its own four columns and the six to the right, and evaluates the
// SSIM expression for four window positions.  Every input byte is read from HBM once (the row seven above comes from L1/L2).
//   S = (2 s0 s1 + c1)(2 (49 s4 - s0 s1) + c2) / ((s0^2 + s1^2 + c1)(49 s23 - s0^2 - s1^2 + c2))
// with s0, s1, s23, s4 the 49-pixel sums of a, b, a^2 + b^2, ab, c1 = 49^2 C1, c2 = 49*48 C2 (skimage structural_similarity
// with sample covariance: the 1/49^2 and 1/(49*48) factors cancel).
// Arithmetic.  Everything that cancels -- the window sums, 49 s4 - s0 s1, 49 s23 - s0^2 - s1^2 -- is EXACT int32 (all values
// < 2^31); the four factors are converted to float once (relative 6e-8, after the cancellation), the quotient takes one
// MUFU.RCP plus a Newton step, and a thread adds at most 16 rows x 4 windows in float before promoting to its double
// accumulator.  Per-window error ~2e-7 with random sign: the image mean agrees with the float64 statement to ~1e-8 (asserted
// <= 1e-6 in tests/test_gpu_ops.py; the tolerance of the path is 1e-3).  a^2 and b^2 only ever appear as a sum, so one running
// sum serves both; the squares enter as (new - old)(new + old).  Round 1 evaluated the expression in double with five running
// sums: 78 instructions per pixel, issue-bound at 0.38 TB/s; this version executes ~45.
static constexpr int kMetThreads = 128;
static constexpr int kMetCols = kMetThreads * 4;     // input columns per strip
static constexpr int kMetStep = kMetCols - 8;        // output columns a strip owns (multiple of 4; the last strip takes up to +2 more)

struct MetGeom { int strips, chunks, rows_per_chunk; };
static MetGeom metric_geom(int n, int h, int w) {
  MetGeom g;
  const int ow = w - 6 > 0 ? w - 6 : 1, oh = h - 6 > 0 ? h - 6 : 1;
  g.strips = ow <= kMetStep + 2 ? 1 : (ow - 2 + kMetStep - 1) / kMetStep;
  // enough CTAs for ~10 per SM, at least 8 output rows per chunk (6 halo rows are re-read per chunk; taller chunks with fewer
  // CTAs measured slower: 63 -> 68 us for 64 pairs of 512^2)
  long long want = (148LL * 10 + (long long)n * g.strips - 1) / ((long long)n * g.strips);
  if (want < 1) want = 1;
  int max_chunks = (oh + 7) / 8;
  if (max_chunks < 1) max_chunks = 1;
  g.chunks = (int)(want < max_chunks ? want : max_chunks);
  g.rows_per_chunk = (oh + g.chunks - 1) / g.chunks;
  g.chunks = (oh + g.rows_per_chunk - 1) / g.rows_per_chunk;
  return g;
}

template <bool VEC>
__global__ void __launch_bounds__(kMetThreads) metric_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int h, int w,
                                                             int rows_per_chunk, int want_ssim, long long* __restrict__ sse_part,
                                                             double* __restrict__ ssim_part) {
  __shared__ __align__(16) int vs[2][4][kMetCols + 8];
  __shared__ double red_d[kMetThreads / 32];
  __shared__ long long red_l[kMetThreads / 32];
  const int strip = blockIdx.x, chunk = blockIdx.y, img = blockIdx.z;
  const int strips = gridDim.x, chunks = gridDim.y;
  const int tid = threadIdx.x;
  const int ow = w - 6, oh = h - 6;
  const int x0 = strip * kMetStep + 4 * tid;                // this thread's four input columns / window left edges
  const int own_x_end = strip + 1 == strips ? w : (strip + 1) * kMetStep;      // SSE ownership of input columns
  const int out_x_end = strip + 1 == strips ? ow : (strip + 1) * kMetStep;     // window positions this strip scores
  const int oy0 = chunk * rows_per_chunk;
  const int oy1 = min(oy0 + rows_per_chunk, oh);
  const int own_y_end = chunk + 1 == chunks ? h : oy0 + rows_per_chunk;        // SSE ownership of input rows
  const int y_begin = oy0, y_end = (want_ssim && oh > 0) ? max(oy1 + 6, own_y_end) : own_y_end;
  const uint8_t* pa = a + (size_t)img * h * w;
  const uint8_t* pb = b + (size_t)img * h * w;
  if (tid < 8) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { vs[0][q][kMetCols + tid] = 0; vs[1][q][kMetCols + tid] = 0; }
  }
  auto load4 = [&](const uint8_t* p, int y) -> uint32_t {
    if (VEC) {
      if (x0 + 3 < w) return __ldg(reinterpret_cast<const uint32_t*>(p + (size_t)y * w + x0));
    }
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (x0 + j < w) v |= (uint32_t)__ldg(p + (size_t)y * w + x0 + j) << (8 * j);
    return v;
  };
  // columns this thread owns for the SSE / scores as window positions: masks instead of per-pixel compares
  bool own[4], outp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { own[j] = x0 + j < own_x_end; outp[j] = x0 + j < out_x_end; }
  int V[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[q][j] = 0;
  long long sse = 0;
  double acc = 0.0;
  float accf = 0.f;
  int rows_in_accf = 0;
  const float c1 = (float)(2401.0 * (0.01 * 255.0) * (0.01 * 255.0)), c2 = (float)(2352.0 * (0.03 * 255.0) * (0.03 * 255.0));
  for (int y = y_begin; y < y_end; ++y) {
    const uint32_t na = load4(pa, y), nb = load4(pb, y);
    uint32_t oa = 0, ob = 0;
    const bool has_old = y - 7 >= y_begin;
    if (has_old) { oa = load4(pa, y - 7); ob = load4(pb, y - 7); }
    int sq_row = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int an = (int)__byte_perm(na, 0, 0x4440 + j), bn = (int)__byte_perm(nb, 0, 0x4440 + j);
      const int ao = (int)__byte_perm(oa, 0, 0x4440 + j), bo = (int)__byte_perm(ob, 0, 0x4440 + j);
      const int da = an - ao, db = bn - bo;
      V[0][j] += da;
      V[1][j] += db;
      V[2][j] += da * (an + ao) + db * (bn + bo);       // a^2 + b^2, new minus old
      V[3][j] += an * bn - ao * bo;
      const int e = an - bn;
      if (own[j]) sq_row += e * e;
    }
    if (y < own_y_end) sse += sq_row;
    const int oy = y - 6;                       // window rows [oy, oy+6] are summed in V now
    if (!want_ssim || oy < oy0 || oy >= oy1) continue;      // uniform across the CTA
    const int pbuf = y & 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<int4*>(&vs[pbuf][q][4 * tid]) = make_int4(V[q][0], V[q][1], V[q][2], V[q][3]);
    __syncthreads();
    int H[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int4 r0 = *reinterpret_cast<const int4*>(&vs[pbuf][q][4 * tid + 4]);
      const int2 r1 = *reinterpret_cast<const int2*>(&vs[pbuf][q][4 * tid + 8]);
      const int v0 = V[q][0], v1 = V[q][1], v2 = V[q][2], v3 = V[q][3];
      const int h0 = v0 + v1 + v2 + v3 + r0.x + r0.y + r0.z;
      const int h1 = h0 - v0 + r0.w;
      const int h2 = h1 - v1 + r1.x;
      const int h3 = h2 - v2 + r1.y;
      H[q][0] = h0; H[q][1] = h1; H[q][2] = h2; H[q][3] = h3;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s0 = H[0][j], s1 = H[1][j];
      const int s01 = s0 * s1;
      const int sq = s0 * s0 + s1 * s1;
      const int cv = 49 * H[3][j] - s01;
      const int vr = 49 * H[2][j] - sq;
      const float A1 = fmaf(2.f, (float)s01, c1), A2 = fmaf(2.f, (float)cv, c2);
      const float B1 = (float)sq + c1, B2 = (float)vr + c2;
      const float N = A1 * A2, D = B1 * B2;
      float r = __frcp_rn(D);
      r = fmaf(fmaf(-D, r, 1.f), r, r);          // one Newton step: ~1 ulp
      if (outp[j]) accf = fmaf(N, r, accf);
    }
    if (++rows_in_accf == 16) { acc += (double)accf; accf = 0.f; rows_in_accf = 0; }
  }
  acc += (double)accf;
  acc = warp_sum(acc);
  sse = warp_sum(sse);
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) { red_d[warp] = acc; red_l[warp] = sse; }
  __syncthreads();
  if (tid == 0) {
    double d = 0.0; long long l = 0;
    for (int k = 0; k < kMetThreads / 32; ++k) { d += red_d[k]; l += red_l[k]; }
    const size_t part = ((size_t)img * chunks + chunk) * strips + strip;
    ssim_part[part] = d;
    sse_part[part] = l;
  }
}

__global__ void metric_finish_kernel(const long long* sse_part, const double* ssim_part, int parts, long long* sq_err, double* ssim_sum) {
  const int img = blockIdx.x;
  // one warp, fixed order: lane-strided partial sums then a shuffle tree
  double d = 0.0; long long l = 0;
  for (int k = threadIdx.x; k < parts; k += 32) { d += ssim_part[(size_t)img * parts + k]; l += sse_part[(size_t)img * parts + k]; }
  d = warp_sum(d); l = warp_sum(l);
  if (threadIdx.x == 0) { if (sq_err) sq_err[img] = l; if (ssim_sum) ssim_sum[img] = d; }
}

// ------------------------------------------------------------------------ normalize_preds
struct NormWs {                 // per image, 4096 bytes
  unsigned int hist_a[256];
  unsigned int hist_b[256];
  unsigned long long sum_ab;
  unsigned long long n_b;       // differing resolutions: pixels of hr_hat (0: same size as hr)
  double sza, sz;               // differing resolutions: sum zoom(hr_hat) * hr, sum zoom(hr_hat) over the hr grid
  unsigned long long pad[4];
  uint8_t lut_a[256];
  uint8_t lut_b[256];
  uint8_t fill[4096 - 2048 - 64 - 512];
};
static_assert(sizeof(NormWs) == 4096, "NormWs layout");

__global__ void __launch_bounds__(256) norm_stats_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n_px, NormWs* ws) {
  __shared__ unsigned int ha[256], hb[256];
  const int img = blockIdx.y;
  ha[threadIdx.x] = 0; hb[threadIdx.x] = 0;
  __syncthreads();
  const uint8_t* pa = a + (size_t)img * n_px;
  const uint8_t* pb = b + (size_t)img * n_px;
  long long sab = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_px; i += (size_t)gridDim.x * blockDim.x) {
    const int va = pa[i], vb = pb[i];
    atomicAdd(&ha[va], 1u);
    atomicAdd(&hb[vb], 1u);
    sab += va * vb;
  }
  sab = warp_sum(sab);
  if ((threadIdx.x & 31) == 0 && sab != 0) atomicAdd(&ws[img].sum_ab, (unsigned long long)sab);
  __syncthreads();
  if (ha[threadIdx.x]) atomicAdd(&ws[img].hist_a[threadIdx.x], ha[threadIdx.x]);
  if (hb[threadIdx.x]) atomicAdd(&ws[img].hist_b[threadIdx.x], hb[threadIdx.x]);
}

// 256-bin histogram of one image set (differing resolutions: hr and hr_hat are counted separately)
__global__ void __launch_bounds__(256) norm_hist_kernel(const uint8_t* __restrict__ a, size_t n_px, NormWs* ws, int which) {
  __shared__ unsigned int h[256];
  const int img = blockIdx.y;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint8_t* pa = a + (size_t)img * n_px;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_px; i += (size_t)gridDim.x * blockDim.x) atomicAdd(&h[pa[i]], 1u);
  __syncthreads();
  unsigned int* dst = which ? ws[img].hist_b : ws[img].hist_a;
  if (h[threadIdx.x]) atomicAdd(&dst[threadIdx.x], h[threadIdx.x]);
}

// `resize(hr_hat_norm, hr_norm.shape)` of pssr/util.py:179 = skimage.transform.resize(order 1, mode "reflect", no anti-aliasing when
// enlarging) = scipy.ndimage.zoom(order=1, mode="mirror", grid_mode=True): source coordinate (o + 0.5) * in / out - 0.5, reflected about
// the first / last sample, linear interpolation.  Only two sums of the enlarged image are needed (the covariance is shift-invariant):
// sum z * hr and sum z, per CTA in double, reduced in fixed order by norm_cross_finish_kernel.
__device__ __forceinline__ void zoom_coord(int o, double zoom, int len, int& i0, int& i1, double& t) {
  double cc = ((double)o + 0.5) * zoom - 0.5;
  if (cc < 0.0) cc = -cc;
  if (cc > (double)(len - 1)) cc = 2.0 * (double)(len - 1) - cc;
  if (cc < 0.0) cc = 0.0;                     // len == 1
  const double f = floor(cc);
  i0 = (int)f;
  t = cc - f;
  i1 = i0 + 1 < len ? i0 + 1 : i0;
}
__global__ void __launch_bounds__(256) norm_cross_kernel(const uint8_t* __restrict__ hr, const uint8_t* __restrict__ hat, int h, int w, int hh, int hw,
                                                         double* __restrict__ partials) {
  const int img = blockIdx.y;
  const uint8_t* pa = hr + (size_t)img * h * w;
  const uint8_t* pb = hat + (size_t)img * hh * hw;
  const double zy = (double)hh / (double)h, zx = (double)hw / (double)w;
  double sza = 0.0, sz = 0.0;
  const size_t n_px = (size_t)h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_px; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
    int y0, y1, x0, x1;
    double ty, tx;
    zoom_coord(y, zy, hh, y0, y1, ty);
    zoom_coord(x, zx, hw, x0, x1, tx);
    // scipy applies the separable weights as a product over the 2 x 2 support
    const double z = (1.0 - ty) * ((1.0 - tx) * pb[(size_t)y0 * hw + x0] + tx * pb[(size_t)y0 * hw + x1]) +
                     ty * ((1.0 - tx) * pb[(size_t)y1 * hw + x0] + tx * pb[(size_t)y1 * hw + x1]);
    sza += z * (double)pa[i];
    sz += z;
  }
  __shared__ double ra[8], rz[8];
  sza = warp_sum(sza); sz = warp_sum(sz);
  if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = sza; rz[threadIdx.x >> 5] = sz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, z = 0.0;
    for (int k = 0; k < 8; ++k) { a += ra[k]; z += rz[k]; }
    partials[((size_t)img * gridDim.x + blockIdx.x) * 2 + 0] = a;
    partials[((size_t)img * gridDim.x + blockIdx.x) * 2 + 1] = z;
  }
}
__global__ void norm_cross_finish_kernel(const double* __restrict__ partials, int parts, unsigned long long n_b, NormWs* ws) {
  const int img = blockIdx.x;
  if (threadIdx.x == 0) {
    double a = 0.0, z = 0.0;
    for (int k = 0; k < parts; ++k) { a += partials[((size_t)img * parts + k) * 2]; z += partials[((size_t)img * parts + k) * 2 + 1]; }
    ws[img].sza = a; ws[img].sz = z; ws[img].n_b = n_b;
  }
}

// value at sorted position `idx` of the multiset described by a 256-bin histogram
__device__ int hist_select(const unsigned int* hist, unsigned long long idx) {
  unsigned long long cum = 0;
  for (int v = 0; v < 256; ++v) { cum += hist[v]; if (idx < cum) return v; }
  return 255;
}
// np.percentile(x, q) with the default linear method on float32 data (numpy lib/_function_base_impl
// `_lerp`): result is float32.
__device__ float hist_percentile(const unsigned int* hist, unsigned long long n, double q) {
  const double quant = q / 100.0;
  const double vi = (double)(n - 1) * quant;
  double lo = floor(vi);
  const double g = vi - lo;
  unsigned long long ilo = (unsigned long long)lo;
  unsigned long long ihi = ilo + 1 < n ? ilo + 1 : n - 1;
  const double va = (double)hist_select(hist, ilo), vb = (double)hist_select(hist, ihi);
  const double diff = vb - va;
  double r = va + diff * g;
  if (g >= 0.5) r = vb - diff * (1.0 - g);
  return (float)r;
}

__global__ void __launch_bounds__(256) norm_lut_kernel(NormWs* ws, unsigned long long n_px, double pmin, double pmax) {
  NormWs& W = ws[blockIdx.x];
  __shared__ float s_xmin, s_den, s_mean_hn, s_mean_b, s_min_hr, s_base_max, s_base_mean, s_mean_hr2;
  __shared__ double s_amp, s_mean_hh2;
  const int v = threadIdx.x;
  const double N = (double)n_px;
  const bool resized = W.n_b != 0;                                       // hr_hat has its own resolution (util.py:179)
  const double NB = resized ? (double)W.n_b : N;
  if (v == 0) {
    double sa = 0, sb = 0, sbb = 0;
    int amin = 255;
    for (int k = 255; k >= 0; --k) {
      sa += (double)W.hist_a[k] * k; sb += (double)W.hist_b[k] * k; sbb += (double)W.hist_b[k] * k * k;
      if (W.hist_a[k]) amin = k;
    }
    const float base_max = hist_percentile(W.hist_a, n_px, pmax);      // util.py:171
    const float base_mean = (float)(sa / N);                            // util.py:172
    const float xmin = hist_percentile(W.hist_a, n_px, pmin);           // util.py:195
    const float den = (base_max - xmin) + 1e-20f;                       // util.py:203
    // mean of (x - xmin)/den over the image (util.py:177), from the histogram
    double m = 0.0;
    for (int k = 0; k < 256; ++k) m += (double)W.hist_a[k] * (double)(((float)k - xmin) / den);
    const float mean_hn = (float)(m / N);
    const float mean_b = (float)(sb / NB);                              // util.py:176
    // np.cov(hr_hat_c, hr_c)[0,1] (ddof=1, float64) / np.var(hr_hat_c) (ddof=0)  -- util.py:180
    // (differing resolutions: hr_hat enlarged to the hr grid -- sum z * hr and sum z replace sum ab and sum b)
    const double cov = resized ? (W.sza - sa * W.sz / N) / (double)den / (N - 1.0) : ((double)W.sum_ab - sa * sb / N) / (double)den / (N - 1.0);
    const double var = (double)(float)((sbb - sb * sb / NB) / NB);
    s_amp = cov / var;
    s_xmin = xmin; s_den = den; s_mean_hn = mean_hn; s_mean_b = mean_b;
    s_min_hr = (((float)amin - xmin) / den) - mean_hn;                  // hr_norm.min() after centring
    s_base_max = base_max; s_base_mean = base_mean;
  }
  __syncthreads();
  // per-value maps up to the final division (util.py:184)
  const float hr_c = (((float)v - s_xmin) / s_den) - s_mean_hn;
  const float hr2 = (hr_c - s_min_hr) * s_base_max;
  const double hh1 = s_amp * (double)((float)v - s_mean_b);             // float64 from here (amp is np.float64)
  const double hh2 = (hh1 - (double)s_min_hr) * (double)s_base_max;
  __shared__ double red_a[256], red_b[256];
  red_a[v] = (double)W.hist_a[v] * (double)hr2;
  red_b[v] = (double)W.hist_b[v] * hh2;
  __syncthreads();
  if (v == 0) {
    double ma = 0, mb = 0;
    for (int k = 0; k < 256; ++k) { ma += red_a[k]; mb += red_b[k]; }
    s_mean_hr2 = (float)(ma / N);
    s_mean_hh2 = mb / NB;
  }
  __syncthreads();
  const float hr3 = hr2 / (s_mean_hr2 / s_base_mean);                   // util.py:185
  const double hh3 = hh2 / (s_mean_hh2 / (double)s_base_mean);
  const float ca = fminf(fmaxf(hr3, 0.f), 255.f);
  const double cb = fmin(fmax(hh3, 0.0), 255.0);
  W.lut_a[v] = (uint8_t)(int)ca;   // NaN (degenerate image) -> 0, numpy's cast is undefined there
  W.lut_b[v] = (uint8_t)(int)cb;
}

__global__ void __launch_bounds__(256) norm_apply_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ oa,
                                                         uint8_t* __restrict__ ob, size_t n_px, const NormWs* ws) {
  __shared__ uint8_t la[256], lb[256];
  const int img = blockIdx.y;
  la[threadIdx.x] = ws[img].lut_a[threadIdx.x];
  lb[threadIdx.x] = ws[img].lut_b[threadIdx.x];
  __syncthreads();
  const size_t base = (size_t)img * n_px;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_px; i += (size_t)gridDim.x * blockDim.x) {
    if (oa) oa[base + i] = la[a[base + i]];
    if (ob) ob[base + i] = lb[b[base + i]];
  }
}

// ------------------------------------------------------------------- noise-profile histogram
// `_Crappifier_Objective.sample` (pssr/train.py:366-380): profile = image.astype(float32) - base.astype(float32);
// np.histogram(profile, np.arange(-256, 256)) -> 511 bins [k-256, k-255), the last one closed at 255; plus the profile's sum
// (its mean is the objective's second term).  Shared-memory histogram per CTA, one global atomic per non-empty bin.
template <typename A>
__global__ void __launch_bounds__(256) profile_hist_kernel(const A* __restrict__ a, const uint8_t* __restrict__ base, size_t n,
                                                           unsigned long long* __restrict__ hist, double* __restrict__ sum) {
  __shared__ unsigned int h[511];
  __shared__ double red[8];
  for (int i = threadIdx.x; i < 511; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = (float)a[i] - (float)base[i];
    s += (double)v;
    if (v >= -256.f && v <= 255.f) {
      int k = (int)floorf(v) + 256;
      if (k > 510) k = 510;
      atomicAdd(&h[k], 1u);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(sum, t);
  }
  for (int i = threadIdx.x; i < 511; i += blockDim.x)
    if (h[i] != 0u) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

}  // namespace pssr

using namespace pssr;

extern "C" int pssr_profile_hist(const void* a, int32_t a_kind, const uint8_t* base, int64_t n, int64_t* hist511, double* sum, void* stream) {
  PSSR_REQUIRE(a && base && hist511 && sum && n >= 0 && a_kind >= 0 && a_kind <= 2, PSSR_EINVAL, "profile_hist: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PSSR_CHECK_CUDA(cudaMemsetAsync(hist511, 0, 511 * sizeof(int64_t), st));
  PSSR_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st));
  if (n == 0) return PSSR_OK;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist511);
  if (a_kind == 0) profile_hist_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(a), base, (size_t)n, h, sum);
  else if (a_kind == 1) profile_hist_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(a), base, (size_t)n, h, sum);
  else profile_hist_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const double*>(a), base, (size_t)n, h, sum);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

extern "C" int64_t pssr_normalize_resized_workspace_bytes(int32_t n) { return n <= 0 ? 0 : (int64_t)n * (4096 + 256 * 16); }

extern "C" int pssr_normalize_preds_resized(const uint8_t* hr, const uint8_t* hr_hat, uint8_t* hr_out, uint8_t* hr_hat_out, int32_t n, int32_t h,
                                            int32_t w, int32_t hat_h, int32_t hat_w, double pmin, double pmax, void* workspace, void* stream) {
  PSSR_REQUIRE(hr && hr_hat && workspace && n >= 1 && h >= 1 && w >= 1 && hat_h >= 1 && hat_w >= 1, PSSR_EINVAL, "normalize_preds: bad arguments");
  PSSR_REQUIRE(((uintptr_t)workspace & 15) == 0, PSSR_EINVAL, "normalize_preds: workspace must be 16-byte aligned");
  PSSR_REQUIRE(n <= 65535, PSSR_EUNSUP, "normalize_preds: at most 65535 images per call");
  PSSR_REQUIRE(hat_h <= h && hat_w <= w, PSSR_EUNSUP, "normalize_preds: hr_hat larger than hr needs skimage's anti-aliasing filter (not on this path)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NormWs* ws = reinterpret_cast<NormWs*>(workspace);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + sizeof(NormWs) * (size_t)n);
  const size_t na = (size_t)h * w, nb = (size_t)hat_h * hat_w;
  PSSR_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(NormWs) * (size_t)n, st));
  auto blocks = [](size_t px, int cap) { long long b = (long long)((px + 256 * 16 - 1) / (256 * 16)); return (int)(b < 1 ? 1 : (b > cap ? cap : b)); };
  norm_hist_kernel<<<dim3(blocks(na, 1024), n), 256, 0, st>>>(hr, na, ws, 0);
  norm_hist_kernel<<<dim3(blocks(nb, 1024), n), 256, 0, st>>>(hr_hat, nb, ws, 1);
  const int parts = blocks(na, 256);
  norm_cross_kernel<<<dim3(parts, n), 256, 0, st>>>(hr, hr_hat, h, w, hat_h, hat_w, partials);
  norm_cross_finish_kernel<<<n, 32, 0, st>>>(partials, parts, (unsigned long long)nb, ws);
  norm_lut_kernel<<<n, 256, 0, st>>>(ws, (unsigned long long)na, pmin, pmax);
  if (hr_out) norm_apply_kernel<<<dim3(blocks(na, 1024), n), 256, 0, st>>>(hr, hr, hr_out, nullptr, na, ws);
  if (hr_hat_out) norm_apply_kernel<<<dim3(blocks(nb, 1024), n), 256, 0, st>>>(hr_hat, hr_hat, nullptr, hr_hat_out, nb, ws);
  count_launch(5 + (hr_out ? 1 : 0) + (hr_hat_out ? 1 : 0));
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

extern "C" int64_t pssr_metric_workspace_bytes(int32_t n, int32_t h, int32_t w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  const MetGeom g = metric_geom(n, h, w);
  return (int64_t)n * g.strips * g.chunks * 16;
}

extern "C" int pssr_metric_sums(const uint8_t* a, const uint8_t* b, int32_t n, int32_t h, int32_t w, int64_t* sq_err,
                                double* ssim_sum, void* workspace, void* stream) {
  PSSR_REQUIRE(a && b && n >= 1 && h >= 1 && w >= 1, PSSR_EINVAL, "metric_sums: bad arguments");
  PSSR_REQUIRE(ssim_sum == nullptr || (h >= 7 && w >= 7), PSSR_EINVAL, "win_size exceeds image extent.");
  PSSR_REQUIRE(n <= 65535, PSSR_EUNSUP, "metric_sums: at most 65535 images per call");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const MetGeom g = metric_geom(n, h, w);
  const int parts = g.strips * g.chunks;
  PSSR_REQUIRE(g.chunks <= 65535, PSSR_EUNSUP, "metric_sums: image too tall");
  PSSR_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 15) == 0, PSSR_EINVAL, "metric_sums: workspace missing or misaligned");
  long long* sse_part = reinterpret_cast<long long*>(workspace);
  double* ssim_part = reinterpret_cast<double*>(sse_part + (size_t)n * parts);
  const dim3 grid(g.strips, g.chunks, n);
  const bool vec = (w % 4 == 0) && (((uintptr_t)a | (uintptr_t)b) & 3) == 0;
  if (vec) metric_kernel<true><<<grid, kMetThreads, 0, st>>>(a, b, h, w, g.rows_per_chunk, ssim_sum != nullptr, sse_part, ssim_part);
  else metric_kernel<false><<<grid, kMetThreads, 0, st>>>(a, b, h, w, g.rows_per_chunk, ssim_sum != nullptr, sse_part, ssim_part);
  metric_finish_kernel<<<n, 32, 0, st>>>(sse_part, ssim_part, parts, reinterpret_cast<long long*>(sq_err), ssim_sum);
  count_launch(2);
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

extern "C" int64_t pssr_normalize_workspace_bytes(int32_t n) { return (int64_t)(n > 0 ? n : 0) * (int64_t)sizeof(NormWs); }

extern "C" int pssr_normalize_preds(const uint8_t* hr, const uint8_t* hr_hat, uint8_t* hr_out, uint8_t* hr_hat_out, int32_t n,
                                    int32_t h, int32_t w, double pmin, double pmax, void* workspace, void* stream) {
  PSSR_REQUIRE(hr && hr_hat && workspace && n >= 1 && h >= 1 && w >= 1, PSSR_EINVAL, "normalize_preds: bad arguments");
  PSSR_REQUIRE(((uintptr_t)workspace & 15) == 0, PSSR_EINVAL, "normalize_preds: workspace must be 16-byte aligned");
  PSSR_REQUIRE(n <= 65535, PSSR_EUNSUP, "normalize_preds: at most 65535 images per call");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NormWs* ws = reinterpret_cast<NormWs*>(workspace);
  const size_t n_px = (size_t)h * w;
  PSSR_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(NormWs) * (size_t)n, st));
  int bx = (int)((n_px + 256 * 16 - 1) / (256 * 16));
  if (bx < 1) bx = 1;
  if (bx > 1024) bx = 1024;
  norm_stats_kernel<<<dim3(bx, n), 256, 0, st>>>(hr, hr_hat, n_px, ws);
  norm_lut_kernel<<<n, 256, 0, st>>>(ws, (unsigned long long)n_px, pmin, pmax);
  norm_apply_kernel<<<dim3(bx, n), 256, 0, st>>>(hr, hr_hat, hr_out, hr_hat_out, n_px, ws);
  count_launch(3);
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}
