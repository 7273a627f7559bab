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
static constexpr int kSsimT = 32;            // centres per CTA edge
static constexpr int kSsimW = kSsimT + 6;    // staged pixels per edge

__global__ void __launch_bounds__(256) metric_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int h, int w,
                                                     int tiles_x, int tiles_y, long long* __restrict__ sse_part,
                                                     double* __restrict__ ssim_part) {
  __shared__ uint8_t sa[kSsimW][kSsimW + 2], sb[kSsimW][kSsimW + 2];
  __shared__ int hs[5][kSsimW][kSsimT + 1];  // horizontal 7-sums of a, b, a^2, b^2, ab
  __shared__ double red_d[8];
  __shared__ long long red_l[8];
  const int img = blockIdx.y;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int x0 = tx * kSsimT, y0 = ty * kSsimT;  // top-left of the staged window (= centre - 3)
  const uint8_t* pa = a + (size_t)img * h * w;
  const uint8_t* pb = b + (size_t)img * h * w;
  long long sse = 0;
  for (int i = threadIdx.x; i < kSsimW * kSsimW; i += blockDim.x) {
    const int r = i / kSsimW, c = i % kSsimW;
    const int y = y0 + r, x = x0 + c;
    int va = 0, vb = 0;
    if (y < h && x < w) {
      va = pa[(size_t)y * w + x];
      vb = pb[(size_t)y * w + x];
      // every pixel belongs to exactly one CTA's top-left kSsimT x kSsimT block for the SSE
      if (r < kSsimT && c < kSsimT) sse += (long long)((va - vb) * (va - vb));
    }
    sa[r][c] = (uint8_t)va;
    sb[r][c] = (uint8_t)vb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsimW * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, c = i % kSsimT;
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int va = sa[r][c + k], vb = sb[r][c + k];
      s0 += va; s1 += vb; s2 += va * va; s3 += vb * vb; s4 += va * vb;
    }
    hs[0][r][c] = s0; hs[1][r][c] = s1; hs[2][r][c] = s2; hs[3][r][c] = s3; hs[4][r][c] = s4;
  }
  __syncthreads();
  double acc = 0.0;
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  const double cov_norm = 49.0 / 48.0;
  for (int i = threadIdx.x; i < kSsimT * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, c = i % kSsimT;
    // centre (y0+r+3, x0+c+3) must satisfy 3 <= cy <= h-4  <=>  window fully inside the image
    if (y0 + r + 6 < h && x0 + c + 6 < w) {
      int s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        s0 += hs[0][r + k][c]; s1 += hs[1][r + k][c]; s2 += hs[2][r + k][c]; s3 += hs[3][r + k][c]; s4 += hs[4][r + k][c];
      }
      const double ux = s0 / 49.0, uy = s1 / 49.0, uxx = s2 / 49.0, uyy = s3 / 49.0, uxy = s4 / 49.0;
      const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
      const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
      acc += (A1 * A2) / (B1 * B2);
    }
  }
  acc = warp_sum(acc);
  sse = warp_sum(sse);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red_d[warp] = acc; red_l[warp] = sse; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0; long long l = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { d += red_d[k]; l += red_l[k]; }
    ssim_part[(size_t)img * gridDim.x + blockIdx.x] = d;
    sse_part[(size_t)img * gridDim.x + blockIdx.x] = l;
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
  unsigned long long pad[7];
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
    const float mean_b = (float)(sb / N);                               // util.py:176
    // np.cov(hr_hat_c, hr_c)[0,1] (ddof=1, float64) / np.var(hr_hat_c) (ddof=0)  -- util.py:180
    const double cov = ((double)W.sum_ab - sa * sb / N) / (double)den / (N - 1.0);
    const double var = (double)(float)((sbb - sb * sb / N) / N);
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
    s_mean_hh2 = mb / N;
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

}  // namespace pssr

using namespace pssr;

extern "C" int64_t pssr_metric_workspace_bytes(int32_t n, int32_t h, int32_t w) {
  const int64_t parts = (int64_t)((w + kSsimT - 1) / kSsimT) * ((h + kSsimT - 1) / kSsimT);
  return (int64_t)(n > 0 ? n : 0) * parts * 16;
}

extern "C" int pssr_metric_sums(const uint8_t* a, const uint8_t* b, int32_t n, int32_t h, int32_t w, int64_t* sq_err,
                                double* ssim_sum, void* workspace, void* stream) {
  PSSR_REQUIRE(a && b && n >= 1 && h >= 1 && w >= 1, PSSR_EINVAL, "metric_sums: bad arguments");
  PSSR_REQUIRE(ssim_sum == nullptr || (h >= 7 && w >= 7), PSSR_EINVAL, "win_size exceeds image extent.");
  PSSR_REQUIRE(n <= 65535, PSSR_EUNSUP, "metric_sums: at most 65535 images per call");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int tiles_x = (w + kSsimT - 1) / kSsimT, tiles_y = (h + kSsimT - 1) / kSsimT;
  const int parts = tiles_x * tiles_y;
  PSSR_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 15) == 0, PSSR_EINVAL, "metric_sums: workspace missing or misaligned");
  long long* sse_part = reinterpret_cast<long long*>(workspace);
  double* ssim_part = reinterpret_cast<double*>(sse_part + (size_t)n * parts);
  metric_kernel<<<dim3(parts, n), 256, 0, st>>>(a, b, h, w, tiles_x, tiles_y, sse_part, ssim_part);
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
