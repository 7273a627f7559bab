// Family 1: fused crappify.  One launch = tile gather from resident sheets (+ reflect pad) ->
// Pillow-exact two-pass antialiased triangle downscale -> noise stages -> round-half-even ->
// clip -> float32 LR tiles.  Replaces, per tile,
//   _sliding_window / _square_crop / _pad_image      pssr/data.py:629-638, :536-551
//   Image.fromarray(ch).resize(.., BILINEAR)          pssr/data.py:483 (Pillow Resample.c, see
//                                                      oracle/pillow_resize.py for the restatement)
//   crappifier.crappify + np.clip(lr.round(),0,255)   pssr/data.py:486-487, pssr/crappifiers.py:26-105
//   _slice_center + float32 cast                      pssr/data.py:489-495, :526-534
//
// A CTA owns a TL x TL patch of one LR frame.  It stages the HR window it needs in shared memory
// with 16-byte coalesced loads (each HR byte is read from HBM once, plus a 2*scale halo), runs the
// horizontal pass into a second shared buffer rounded to the image dtype exactly as Pillow does,
// then the vertical pass, the noise chain and the store.  uint8 uses Pillow's 22-bit fixed point;
// uint16 uses Pillow's sequential double accumulation (explicit __dmul_rn/__dadd_rn: an FMA would
// change the bits).
#include <math.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace pssr {

static constexpr int kCrapThreads = 256;
static constexpr int kPrecisionBits = 32 - 8 - 2;

struct ResampleTable {
  int in_size = 0, out_size = 0, ksize = 0;
  int2* bounds = nullptr;    // device [out] (xmin, count)
  int32_t* kq = nullptr;     // device [out][ksize] 8 bpc fixed-point coefficients
  double* kd = nullptr;      // device [out][ksize] double coefficients
  int32_t* ki = nullptr;     // device [out][ksize] k * 2^shift where every coefficient of the output is dyadic, see dshift
  int32_t* dshift = nullptr; // device [out] shift (0..14), or -1: the output needs the double path
  std::vector<int2> h_bounds;
};

// Host restatement of Resample.c precompute_coeffs / normalize_coeffs_8bpc (bilinear filter).
static void build_coeffs(int in_size, int out_size, std::vector<int2>& bounds, std::vector<double>& kk,
                         std::vector<int32_t>& kq, int& ksize) {
  double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  ksize = (int)ceil(support) * 2 + 1;
  bounds.assign(out_size, make_int2(0, 0));
  kk.assign((size_t)out_size * ksize, 0.0);
  kq.assign((size_t)out_size * ksize, 0);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double* k = &kk[(size_t)xx * ksize];
    volatile double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww = ww + w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    bounds[xx] = make_int2(xmin, xmax);
    for (int x = 0; x < ksize; ++x) {
      const double v = k[x];
      kq[(size_t)xx * ksize + x] = v < 0 ? (int32_t)(-0.5 + v * (double)(1 << kPrecisionBits))
                                         : (int32_t)(0.5 + v * (double)(1 << kPrecisionBits));
    }
  }
}

static std::mutex g_table_mu;
static std::map<std::pair<int, std::pair<int, int>>, ResampleTable> g_tables;  // (device, (in, out))

static int get_table(int in_size, int out_size, const ResampleTable** out) {
  int dev = 0;
  PSSR_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_table_mu);
  auto key = std::make_pair(dev, std::make_pair(in_size, out_size));
  auto it = g_tables.find(key);
  if (it != g_tables.end()) {
    *out = &it->second;
    return PSSR_OK;
  }
  ResampleTable t;
  std::vector<double> kk;
  std::vector<int32_t> kq;
  build_coeffs(in_size, out_size, t.h_bounds, kk, kq, t.ksize);
  t.in_size = in_size;
  t.out_size = out_size;
  PSSR_CHECK_CUDA(cudaMalloc(&t.bounds, sizeof(int2) * out_size));
  PSSR_CHECK_CUDA(cudaMalloc(&t.kq, sizeof(int32_t) * kq.size()));
  PSSR_CHECK_CUDA(cudaMalloc(&t.kd, sizeof(double) * kk.size()));
  PSSR_CHECK_CUDA(cudaMemcpy(t.bounds, t.h_bounds.data(), sizeof(int2) * out_size, cudaMemcpyHostToDevice));
  PSSR_CHECK_CUDA(cudaMemcpy(t.kq, kq.data(), sizeof(int32_t) * kq.size(), cudaMemcpyHostToDevice));
  PSSR_CHECK_CUDA(cudaMemcpy(t.kd, kk.data(), sizeof(double) * kk.size(), cudaMemcpyHostToDevice));
  // Integer fast path of the 16-bit (double) resample.  Pillow accumulates pixel * k sequentially in double and takes
  // (int)(ss + 0.5).  When every coefficient of an output is k = m / 2^s with s <= 14 (integer scales: all interior outputs,
  // e.g. [1,3,5,7,7,5,3,1]/32), every product and partial sum is an exactly representable dyadic rational (< 2^31 / 2^s,
  // far inside 53 bits), so  (sum pixel*m + 2^(s-1)) >> s  is the same number bit for bit.  Edge outputs (renormalised by a
  // non-power-of-two window weight) keep the double path.
  std::vector<int32_t> ki(kk.size(), 0), dsh(out_size, -1);
  for (int xx = 0; xx < out_size; ++xx) {
    for (int sft = 0; sft <= 14; ++sft) {
      bool ok = true;
      long long tot = 0;
      for (int x = 0; x < t.ksize && ok; ++x) {
        const double v = kk[(size_t)xx * t.ksize + x] * (double)(1 << sft);
        ok = v == floor(v) && v >= 0.0 && v <= 2147483647.0;
        tot += (long long)v;
      }
      if (ok && tot * 65535ll < (1ll << 31)) {
        dsh[xx] = sft;
        for (int x = 0; x < t.ksize; ++x) ki[(size_t)xx * t.ksize + x] = (int32_t)(kk[(size_t)xx * t.ksize + x] * (double)(1 << sft));
        break;
      }
    }
  }
  PSSR_CHECK_CUDA(cudaMalloc(&t.ki, sizeof(int32_t) * ki.size()));
  PSSR_CHECK_CUDA(cudaMalloc(&t.dshift, sizeof(int32_t) * out_size));
  PSSR_CHECK_CUDA(cudaMemcpy(t.ki, ki.data(), sizeof(int32_t) * ki.size(), cudaMemcpyHostToDevice));
  PSSR_CHECK_CUDA(cudaMemcpy(t.dshift, dsh.data(), sizeof(int32_t) * out_size, cudaMemcpyHostToDevice));
  auto ins = g_tables.emplace(key, std::move(t));
  *out = &ins.first->second;
  return PSSR_OK;
}

// ------------------------------------------------------------------------------ Philox
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u01f(uint32_t x) { return ((x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)
__device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
}

// Poisson(lam): Knuth's product method below 10, Hoermann's PTRS above (the pair NumPy's legacy
// generator uses; the stream of uniforms differs, so free-running mode is validated statistically).
// lam < 1024 (every 8-bit pipeline value) runs in float32 -- the sampler's arithmetic is ours to choose, only its
// distribution is specified, and B200's fp64 transcendental path (log, lgamma) made the fp64 version 90 % of the kernel;
// the squeeze inequality is evaluated with ~1e-4 absolute error on terms of O(1e3), i.e. it changes the acceptance of a
// vanishing fraction of borderline draws.  One Philox block feeds two PTRS attempts.
// log(k!) for integer-valued k >= 0: table below 8, Stirling series above (|error| < 1.3e-8 at k + 1 >= 9) -- lgammaf() costs
// ~100 instructions and every warp takes the squeeze branch that needs it
__constant__ float kLogFact[8] = {0.f, 0.f, 0.69314718f, 1.79175947f, 3.17805383f, 4.78749174f, 6.57925121f, 8.52516136f};
__device__ __forceinline__ float log_factorial(float k) {
  if (k < 8.f) return kLogFact[(int)k];
  const float x = k + 1.0f, rx = __frcp_rn(x);
  return fmaf(x - 0.5f, logf(x), -x) + 0.91893853f + rx * (0.083333333f - 0.0027777778f * rx * rx);
}

// Per-pixel pool of Philox words shared by all noise stages of the pixel: one Philox4x32-10 block (~70 instructions) yields two
// pairs; a PTRS attempt, a Box-Muller normal and a salt-and-pepper draw take one pair each, so the usual Poisson + Gaussian
// chain costs ONE block per pixel instead of two or three.  Counter = (pixel, block index), key = (seed, tile).
struct RngPool {
  const Philox& ph;
  uint32_t pix, blk;
  uint4 w;
  int have;
  __device__ __forceinline__ RngPool(const Philox& p, uint32_t px) : ph(p), pix(px), blk(0), w(make_uint4(0, 0, 0, 0)), have(0) {}
  __device__ __forceinline__ void next2(uint32_t& a, uint32_t& b) {
    if (have == 0) { w = ph(pix, 0x504F4F4Cu, blk++, 0x50535352u); have = 2; a = w.x; b = w.y; }
    else { have = 0; a = w.z; b = w.w; }
  }
};

// Integer rates 1..255 (every LR pixel of an 8-bit pipeline before the first noise stage) are sampled with Walker/Vose ALIAS
// tables built on the host in double precision: per rate 256 outcomes k = kmin .. kmin+255 (kmin = max(0, rate - 7.5 sqrt(rate)
// - 4): > 7.5 sigma on both sides, truncated mass < 1e-13, renormalised), thresholds quantised to 2^-32.  One pair of Philox
// words per draw, two loads, no rejection loop -- a warp running PTRS repeats the attempt until its slowest lane accepts
// (~3.5 rounds, ~500 instructions per pixel).  Other rates (non-integer after an earlier stage, > 255) keep Knuth / PTRS.
struct PtrsConst { float slam, loglam, b, a, invalpha, vr, log_invalpha, enlam; };
__device__ const uint2* g_alias_tab;      // [256][256] (threshold, alias)
__device__ const int* g_alias_kmin;       // [256]
__device__ __forceinline__ PtrsConst ptrs_const(float lam) {
  PtrsConst c;
  c.slam = sqrtf(lam);
  c.loglam = logf(lam);
  c.b = 0.931f + 2.53f * c.slam;
  c.a = -0.059f + 0.02483f * c.b;
  c.invalpha = 1.1239f + 1.1328f / (c.b - 3.4f);
  c.vr = 0.9277f - 3.6224f / (c.b - 2.0f);
  c.log_invalpha = logf(c.invalpha);
  c.enlam = expf(-lam);
  return c;
}

__device__ float poisson_sample_f32(RngPool& rng, float lam) {
  if (lam < 256.0f && lam >= 1.0f && lam == floorf(lam)) {
    uint32_t u0, u1;
    rng.next2(u0, u1);
    const int il = (int)lam, j = (int)(u0 >> 24);
    const uint2 e = __ldg(g_alias_tab + il * 256 + j);
    return (float)(__ldg(g_alias_kmin + il) + (u1 < e.x ? j : (int)e.y));
  }
  const PtrsConst c = ptrs_const(lam);
  if (lam < 10.0f) {
    float prod = 1.0f;
    int k = 0;
    while (true) {
      uint32_t u0, u1;
      rng.next2(u0, u1);
      prod *= u01f(u0);
      if (prod <= c.enlam) return (float)k;
      ++k;
      prod *= u01f(u1);
      if (prod <= c.enlam) return (float)k;
      ++k;
    }
  }
  while (true) {
    uint32_t u0, u1;
    rng.next2(u0, u1);
    const float U = u01f(u0) - 0.5f;
    const float V = u01f(u1);
    const float us = 0.5f - fabsf(U);
    const float k = floorf((2.0f * c.a / us + c.b) * U + lam + 0.43f);
    if (us >= 0.07f && V <= c.vr) return k;
    if (k < 0.0f || (us < 0.013f && V > us)) continue;
    if ((__logf(V) + c.log_invalpha - __logf(c.a / (us * us) + c.b)) <= (-lam + k * c.loglam - log_factorial(k))) return k;
  }
}

__device__ double poisson_sample(RngPool& rng, double lam) {
  if (!(lam > 0.0)) return 0.0;
  if (lam < 1024.0) return (double)poisson_sample_f32(rng, (float)lam);
  const Philox& ph = rng.ph;
  const uint32_t pix = rng.pix, stage = 0x4C415247u;
  uint32_t draw = 0;
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam;
  const double a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  const double vr = 0.9277 - 3.6224 / (b - 2.0);
  while (true) {
    const uint4 r = ph(pix, stage, draw++, 0x50545253u);
    const double U = u01d(r.x, r.y) - 0.5;
    const double V = u01d(r.z, r.w);
    const double us = 0.5 - fabs(U);
    const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0.0 || (us < 0.013 && V > us)) continue;
    if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + k * loglam - lgamma(k + 1.0))) return k;
  }
}

// one-time build + upload of the alias tables on the current device (synchronous on first use so that every stream sees them)
static int ensure_ptrs_table() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  PSSR_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 64 && done[dev]) return PSSR_OK;
  std::vector<uint2> tab(256 * 256, make_uint2(0xFFFFFFFFu, 0u));
  std::vector<int> kmin(256, 0);
  std::vector<double> q(256);
  std::vector<int> small, large;
  for (int lam = 1; lam < 256; ++lam) {
    int k0 = (int)floor((double)lam - 7.5 * sqrt((double)lam) - 4.0);
    if (k0 < 0) k0 = 0;
    kmin[lam] = k0;
    double tot = 0.0;
    for (int j = 0; j < 256; ++j) {
      const double k = (double)(k0 + j);
      q[j] = exp(-(double)lam + k * log((double)lam) - lgamma(k + 1.0));
      tot += q[j];
    }
    small.clear(); large.clear();
    for (int j = 0; j < 256; ++j) {
      q[j] = q[j] / tot * 256.0;
      (q[j] < 1.0 ? small : large).push_back(j);
    }
    uint2* row = &tab[(size_t)lam * 256];
    for (int j = 0; j < 256; ++j) row[j] = make_uint2(0xFFFFFFFFu, (unsigned)j);
    while (!small.empty() && !large.empty()) {          // Vose's alias construction
      const int sidx = small.back(); small.pop_back();
      const int lidx = large.back(); large.pop_back();
      const double thr = q[sidx] * 4294967296.0;
      row[sidx] = make_uint2(thr >= 4294967295.0 ? 0xFFFFFFFFu : (unsigned)thr, (unsigned)lidx);
      q[lidx] = (q[lidx] + q[sidx]) - 1.0;
      (q[lidx] < 1.0 ? small : large).push_back(lidx);
    }
  }
  uint2* d_tab = nullptr;
  int* d_kmin = nullptr;
  PSSR_CHECK_CUDA(cudaMalloc(&d_tab, tab.size() * sizeof(uint2)));
  PSSR_CHECK_CUDA(cudaMalloc(&d_kmin, kmin.size() * sizeof(int)));
  PSSR_CHECK_CUDA(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  PSSR_CHECK_CUDA(cudaMemcpy(d_kmin, kmin.data(), kmin.size() * sizeof(int), cudaMemcpyHostToDevice));
  PSSR_CHECK_CUDA(cudaMemcpyToSymbol(g_alias_tab, &d_tab, sizeof(d_tab)));
  PSSR_CHECK_CUDA(cudaMemcpyToSymbol(g_alias_kmin, &d_kmin, sizeof(d_kmin)));
  PSSR_CHECK_CUDA(cudaDeviceSynchronize());
  if (dev < 64) done[dev] = true;
  return PSSR_OK;
}

// ------------------------------------------------------------------------------ kernel
struct StageK {
  int kind, rng, mix_in_f32, pad;
  double intensity, gain;
  const void* injected;
};

// The crappifier chain on one value (MultiCrappifier.crappify, pssr/crappifiers.py:38-43).  `val`
// carries either a float32 or a float64 quantity exactly; casts reproduce NumPy's dtype flow.
__device__ __forceinline__ double noise_chain(double val, const StageK* stages, int n_stages, int clip_between,
                                              const Philox& ph, uint32_t pix, size_t inj) {
  RngPool rng(ph, pix);
  for (int s = 0; s < n_stages; ++s) {
    const StageK& st = stages[s];
    if (st.kind == PSSR_NOISE_POISSON) {
      // x.astype(f32) * (1 - i) + y * i + gain          (crappifiers.py:82-86)
      const float xf = (float)val;
      double y;
      if (st.rng == PSSR_RNG_INJECTED) y = (double)reinterpret_cast<const long long*>(st.injected)[inj];
      else y = poisson_sample(rng, fmax(val, 0.0));
      double t;
      if (st.mix_in_f32) t = (double)__fmul_rn(xf, (float)(1.0 - st.intensity));
      else t = __dmul_rn((double)xf, 1.0 - st.intensity);
      val = __dadd_rn(__dadd_rn(t, __dmul_rn(y, st.intensity)), st.gain);
    } else if (st.kind == PSSR_NOISE_GAUSSIAN) {
      // x.astype(f32) + normal(gain, intensity)           (crappifiers.py:62-64)
      const float xf = (float)val;
      double g;
      if (st.rng == PSSR_RNG_INJECTED) g = reinterpret_cast<const double*>(st.injected)[inj];
      else {
        uint32_t u0, u1;
        rng.next2(u0, u1);
        const float rad = sqrtf(-2.0f * logf(u01f(u0)));
        const float z = rad * cospif(2.0f * u01f(u1));
        g = __dadd_rn(st.gain, __dmul_rn(st.intensity, (double)z));
      }
      val = __dadd_rn((double)xf, g);
    } else {
      // random_noise(clip(x.astype(f32) + gain, 0, 255) / 255, "s&p", amount) * 255   (crappifiers.py:103-105)
      float v = __fadd_rn((float)val, (float)st.gain);
      v = fminf(fmaxf(v, 0.f), 255.f);
      v = __fdiv_rn(v, 255.f);
      bool flipped, salted;
      if (st.rng == PSSR_RNG_INJECTED) {
        const uint8_t m = reinterpret_cast<const uint8_t*>(st.injected)[inj];
        flipped = m & 1;
        salted = m & 2;
      } else {
        uint32_t u0, u1;
        rng.next2(u0, u1);
        flipped = (double)u01f(u0) <= st.intensity;          // 24-bit uniforms: the rates are reproduced to 6e-8
        salted = u01f(u1) <= 0.5f;
      }
      if (flipped) v = salted ? 1.f : 0.f;
      v = fminf(fmaxf(v, 0.f), 1.f);
      val = (double)__fmul_rn(v, 255.f);
    }
    if (clip_between) val = fmin(fmax(val, 0.0), 255.0);  // MultiCrappifier clip (crappifiers.py:41-42)
  }
  return val;
}

struct CrapK {
  const void* const* sheets;
  int elem_bytes, sheet_h, sheet_w;
  const int32_t *sheet_hs, *sheet_ws;      // optional per-sheet dimensions
  const int32_t *tile_sheet, *tile_frame, *tile_y, *tile_x, *tile_vh, *tile_vw;
  const int32_t* tile_xf;                  // optional per-tile rot90 / flip (training augmentation)
  int n_tiles, frames, lr_frame0, lr_frames, hr_res, lr_res;
  int TL, tiles_per_side, ksize;
  int max_rows, raw_pitch;  // shared staging geometry (bytes per raw row, multiple of 16)
  const int2* bounds;
  const int32_t* kq;
  const double* kd;
  const int32_t* ki;
  const int32_t* dshift;
  StageK stages[4];
  int n_stages, clip_between;
  uint32_t seed_lo, seed_hi;
  unsigned long long tile_index0;
  float* lr_out;
  // HR tiles emitted from the staged window (fused path: hr_res % scale == 0 and the HR frame window inside the LR one)
  float* hr_out;
  uint8_t* hr_u8;
  int hr_frame0, hr_frames, scale;
};

// (r, c) of the augmented tile -> (r, c) of the padded tile it was made from: T'[r][c] = R[r'][c'] with the flips applied to the
// indices, and np.rot90(P, axes=(1, 2))[i][j] = P[j][N-1-i]  (pssr/data.py:478-480: rot90 first, then flip)
__device__ __forceinline__ void untransform(int xf, int N, int& r, int& c) {
  if (xf & 2) r = N - 1 - r;
  if (xf & 4) c = N - 1 - c;
  if (xf & 1) { const int t = r; r = c; c = N - 1 - t; }
}

template <typename T>
__device__ __forceinline__ T load_reflect(const T* frame_base, int sheet_w, int ty, int tx, int r, int c, int vh,
                                          int vw) {
  // np.pad(mode="reflect"): period 2n-2, repeated when the pad is wider than the image (n == 1: the single value)
  int rr = r, cc = c;
  if (r >= vh) { const int per = 2 * vh - 2; rr = per > 0 ? r % per : 0; if (rr >= vh) rr = per - rr; }
  if (c >= vw) { const int per = 2 * vw - 2; cc = per > 0 ? c % per : 0; if (cc >= vw) cc = per - cc; }
  return frame_base[(size_t)(ty + rr) * sheet_w + tx + cc];
}

// KS: compile-time bound of the filter window (9: scales <= 4, 17: scales <= 8; 0: generic loops)
// MINB: CTAs per SM the register allocation aims for (3: 85 registers, no spills; 4: 64 registers, ~70 bytes of spills)
template <typename T, int KS, int MINB>
__global__ void __launch_bounds__(kCrapThreads, MINB) crappify_kernel(const CrapK p) {
  extern __shared__ __align__(16) uint8_t csm[];
  const int TL = p.TL;
  int* lead = reinterpret_cast<int*>(csm);                         // [max_rows] byte offset of each staged row
  uint8_t* raw = csm + (((size_t)p.max_rows * 4 + 15) & ~(size_t)15);  // [max_rows][raw_pitch]
  T* inter = reinterpret_cast<T*>(raw + (size_t)p.max_rows * p.raw_pitch);  // [max_rows][TL]

  int bid = blockIdx.x;
  const int txl = bid % p.tiles_per_side;
  bid /= p.tiles_per_side;
  const int tyl = bid % p.tiles_per_side;
  bid /= p.tiles_per_side;
  const int fo = bid % p.lr_frames;  // output frame index
  const int tile = bid / p.lr_frames;
  const int f = p.lr_frame0 + fo;    // frame inside the tile's window

  const int xx0 = txl * TL, xx1 = min(xx0 + TL, p.lr_res);
  const int yy0 = tyl * TL, yy1 = min(yy0 + TL, p.lr_res);
  const int2 bx0 = p.bounds[xx0], bx1 = p.bounds[xx1 - 1];
  const int2 by0 = p.bounds[yy0], by1 = p.bounds[yy1 - 1];
  const int col0 = bx0.x, ncols = bx1.x + bx1.y - col0;
  const int row0 = by0.x, nrows = by1.x + by1.y - row0;

  const int ty = p.tile_y[tile], tx = p.tile_x[tile], vh = p.tile_vh[tile], vw = p.tile_vw[tile];
  const int sheet_idx = p.tile_sheet[tile];
  const int sheet_h = p.sheet_hs != nullptr ? p.sheet_hs[sheet_idx] : p.sheet_h;
  const int sheet_w = p.sheet_ws != nullptr ? p.sheet_ws[sheet_idx] : p.sheet_w;
  const size_t frame_elems = (size_t)sheet_h * sheet_w;
  const T* sheet = reinterpret_cast<const T*>(p.sheets[sheet_idx]);
  const T* fbase = sheet + (size_t)(p.tile_frame[tile] + f) * frame_elems;

  // ---- stage 1: HR window -> shared (16-byte vectors on the aligned fast path) ----------
  const int xf = p.tile_xf != nullptr ? p.tile_xf[tile] : 0;
  const bool interior = (row0 + nrows <= vh) && (col0 + ncols <= vw);
  const bool vec_ok = interior && xf == 0 && ((reinterpret_cast<uintptr_t>(sheet) & 15) == 0);
  if (vec_ok) {
    // every 16-byte vector of the window goes global -> shared with cp.async: all ~2400 of them are in flight at once (the
    // ld.global / st.shared loop exposed one DRAM latency per iteration) and the row / vector split is a multiply-shift
    // (`i / vec_per_row` as an integer division was 17 % of the kernel's instructions)
    const int vec_per_row = p.raw_pitch / 16;
    const uint32_t magic = ((1u << 22) + (uint32_t)vec_per_row - 1u) / (uint32_t)vec_per_row;     // exact for i * vec_per_row < 2^22
    const uint8_t* g0 = reinterpret_cast<const uint8_t*>(fbase + (size_t)(ty + row0) * sheet_w + tx + col0);
    const size_t row_bytes = (size_t)sheet_w * sizeof(T);
    const uint32_t need = (uint32_t)ncols * (uint32_t)sizeof(T);
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(raw);
    for (int i = threadIdx.x; i < nrows * vec_per_row; i += kCrapThreads) {
      const int r = (int)(((uint32_t)i * magic) >> 22), v = i - r * vec_per_row;
      const uint8_t* g = g0 + (size_t)r * row_bytes;
      const uint32_t ld = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15);
      if (v == 0) lead[r] = (int)ld;
      // the vector holds at least one needed byte, lies inside the sheet's 16-byte-aligned extent
      if ((uint32_t)v * 16u < need + ld)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(raw_s + (uint32_t)r * (uint32_t)p.raw_pitch + (uint32_t)v * 16u),
                     "l"(g - ld + (size_t)v * 16)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    for (int i = threadIdx.x; i < nrows * ncols; i += kCrapThreads) {
      const int r = i / ncols, c = i - r * ncols;
      int pr = row0 + r, pc = col0 + c;
      untransform(xf, p.hr_res, pr, pc);
      reinterpret_cast<T*>(raw + (size_t)r * p.raw_pitch)[c] = load_reflect<T>(fbase, sheet_w, ty, tx, pr, pc, vh, vw);
    }
    for (int r = threadIdx.x; r < nrows; r += kCrapThreads) lead[r] = 0;
  }
  __syncthreads();

  // ---- stage 1b: the HR tile itself (dataset HR output, data.py:495; `_pred_array` view, predict.py:245-246), from the
  // staged window: this CTA owns HR rows [yy0*s, yy1*s) x cols [xx0*s, xx1*s) -- no second pass over the sheet ----------
  if (p.hr_out != nullptr || p.hr_u8 != nullptr) {
    const int hf = f - p.hr_frame0;
    const bool want_f32 = p.hr_out != nullptr && hf >= 0 && hf < p.hr_frames;
    const bool want_u8 = p.hr_u8 != nullptr && hf == p.hr_frames / 2;      // _slice_center(x, 1) keeps index shape//2
    if (want_f32 || want_u8) {
      const int s = p.scale;
      const int X0 = xx0 * s, Y0 = yy0 * s, wown = (xx1 - xx0) * s, hown = (yy1 - yy0) * s;
      const int groups = (wown + 3) >> 2;
      const size_t n_hr = (size_t)p.hr_res * p.hr_res;
      const bool vec_st = (p.hr_res & 3) == 0 && (X0 & 3) == 0;
      for (int ry = (int)(threadIdx.x >> 5); ry < hown; ry += kCrapThreads / 32)
      for (int g = (int)(threadIdx.x & 31); g < groups; g += 32) {
        const int Y = Y0 + ry, X = X0 + 4 * g;
        const int r = Y - row0;
        const uint8_t* sb = raw + (size_t)r * p.raw_pitch + lead[r] + (size_t)(X - col0) * sizeof(T);
        const T* src = reinterpret_cast<const T*>(sb);
        const int cnt = min(4, wown - 4 * g);
        int v[4] = {0, 0, 0, 0};
        if (cnt == 4 && (reinterpret_cast<uintptr_t>(sb) & 3) == 0) {          // 4 pixels in one / two 32-bit shared-memory reads
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(sb);
          if (sizeof(T) == 1) {
            const uint32_t w = s32[0];
            v[0] = w & 255u; v[1] = (w >> 8) & 255u; v[2] = (w >> 16) & 255u; v[3] = w >> 24;
          } else {
            const uint32_t w0 = s32[0], w1 = s32[1];
            v[0] = w0 & 0xffffu; v[1] = w0 >> 16; v[2] = w1 & 0xffffu; v[3] = w1 >> 16;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < cnt) v[j] = (int)src[j];
        }
        const size_t o = (size_t)Y * p.hr_res + X;
        if (want_u8) {
          uint8_t* d = p.hr_u8 + (size_t)tile * n_hr + o;
          if (vec_st && cnt == 4)
            *reinterpret_cast<uint32_t*>(d) = (uint32_t)min(v[0], 255) | ((uint32_t)min(v[1], 255) << 8) | ((uint32_t)min(v[2], 255) << 16) |
                                              ((uint32_t)min(v[3], 255) << 24);
          else
            for (int j = 0; j < cnt; ++j) d[j] = (uint8_t)min(v[j], 255);
        }
        if (want_f32) {
          float* d = p.hr_out + ((size_t)tile * p.hr_frames + hf) * n_hr + o;
          if (vec_st && cnt == 4) *reinterpret_cast<float4*>(d) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
          else
            for (int j = 0; j < cnt; ++j) d[j] = (float)v[j];
        }
      }
    }
  }

  // ---- stage 2: horizontal pass, rounded to the image dtype ------------------------------
  const int nx = xx1 - xx0, ny = yy1 - yy0;
  if (KS > 0 && (kCrapThreads % nx) == 0) {
    // a thread keeps one output column for all rows: window bounds and the (integer) coefficients stay in registers
    const int xo = threadIdx.x % nx;
    const int2 b = p.bounds[xx0 + xo];
    const int sft = sizeof(T) == 1 ? kPrecisionBits : p.dshift[xx0 + xo];
    const int32_t* kp = (sizeof(T) == 1 ? p.kq : p.ki) + (size_t)(xx0 + xo) * p.ksize;
    int32_t kr[KS > 0 ? KS : 1];
#pragma unroll
    for (int x = 0; x < KS; ++x) kr[x] = (x < b.y && sft >= 0) ? __ldg(kp + x) : 0;
    const int32_t half = sft > 0 ? (1 << (sft - 1)) : 0;
    const int coff = b.x - col0;
    for (int r = threadIdx.x / nx; r < nrows; r += kCrapThreads / nx) {
      const T* src = reinterpret_cast<const T*>(raw + (size_t)r * p.raw_pitch + lead[r]) + coff;
      if (sft >= 0) {
        int32_t ss = half;
#pragma unroll
        for (int x = 0; x < KS; ++x)
          if (x < b.y) ss += (int32_t)src[x] * kr[x];
        ss >>= sft;
        inter[r * TL + xo] = sizeof(T) == 1 ? (T)min(max(ss, 0), 255) : (T)((ss & 255) | (min(max(ss >> 8, 0), 255) << 8));
      } else {
        const double* k = p.kd + (size_t)(xx0 + xo) * p.ksize;
        double ss = 0.0;
        for (int x = 0; x < b.y; ++x) ss = __dadd_rn(ss, __dmul_rn((double)src[x], __ldg(k + x)));
        const int si = (int)(ss + 0.5);
        inter[r * TL + xo] = (T)((si & 255) | (min(max(si >> 8, 0), 255) << 8));
      }
    }
  } else {
    for (int i = threadIdx.x; i < nrows * nx; i += kCrapThreads) {
      const int r = i / nx, xo = i - r * nx;
      const int2 b = p.bounds[xx0 + xo];
      const T* src = reinterpret_cast<const T*>(raw + (size_t)r * p.raw_pitch + lead[r]) + (b.x - col0);
      if (sizeof(T) == 1) {
        const int32_t* k = p.kq + (size_t)(xx0 + xo) * p.ksize;
        int32_t ss = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < b.y; ++x) ss += (int32_t)src[x] * __ldg(k + x);
        ss >>= kPrecisionBits;
        inter[r * TL + xo] = (T)min(max(ss, 0), 255);
      } else {
        const double* k = p.kd + (size_t)(xx0 + xo) * p.ksize;
        double ss = 0.0;
        for (int x = 0; x < b.y; ++x) ss = __dadd_rn(ss, __dmul_rn((double)src[x], __ldg(k + x)));
        const int si = (int)(ss + 0.5);
        inter[r * TL + xo] = (T)((si & 255) | (min(max(si >> 8, 0), 255) << 8));
      }
    }
  }
  __syncthreads();

  // ---- stage 3: vertical pass + noise chain + store -------------------------------------
  const Philox ph{p.seed_lo ^ (uint32_t)(p.tile_index0 + tile), p.seed_hi ^ (uint32_t)((p.tile_index0 + tile) >> 32)};
  for (int yo = (int)(threadIdx.x >> 5); yo < ny; yo += kCrapThreads / 32)
  for (int xo = (int)(threadIdx.x & 31); xo < nx; xo += 32) {
    const int2 b = p.bounds[yy0 + yo];
    const T* src = inter + (size_t)(b.x - row0) * TL + xo;
    double val;  // carries either a float32 or a float64 quantity, exactly
    if (sizeof(T) == 1) {
      const int32_t* k = p.kq + (size_t)(yy0 + yo) * p.ksize;
      int32_t ss = 1 << (kPrecisionBits - 1);
      for (int y = 0; y < b.y; ++y) ss += (int32_t)src[(size_t)y * TL] * __ldg(k + y);
      ss >>= kPrecisionBits;
      val = (double)min(max(ss, 0), 255);
    } else {
      const int sft = p.dshift[yy0 + yo];
      int si;
      if (sft >= 0) {
        const int32_t* k = p.ki + (size_t)(yy0 + yo) * p.ksize;
        int32_t ss = sft > 0 ? (1 << (sft - 1)) : 0;
        for (int y = 0; y < b.y; ++y) ss += (int32_t)src[(size_t)y * TL] * __ldg(k + y);
        si = ss >> sft;
      } else {
        const double* k = p.kd + (size_t)(yy0 + yo) * p.ksize;
        double ss = 0.0;
        for (int y = 0; y < b.y; ++y) ss = __dadd_rn(ss, __dmul_rn((double)src[(size_t)y * TL], __ldg(k + y)));
        si = (int)(ss + 0.5);
      }
      val = (double)((si & 255) | (min(max(si >> 8, 0), 255) << 8));
    }
    // `.astype(np.float32)`: exact for values <= 65535 (data.py:483)
    const int yy = yy0 + yo, xx = xx0 + xo;
    const uint32_t pix = (uint32_t)((f * p.lr_res + yy) * p.lr_res + xx);
    const size_t inj = (((size_t)tile * p.frames + f) * p.lr_res + yy) * p.lr_res + xx;
    val = noise_chain(val, p.stages, p.n_stages, p.clip_between, ph, pix, inj);
    if (p.n_stages > 0) val = fmin(fmax(rint(val), 0.0), 255.0);  // np.clip(lr.round(), 0, 255), data.py:487
    p.lr_out[(((size_t)tile * p.lr_frames + fo) * p.lr_res + yy) * p.lr_res + xx] = (float)val;
  }
}

// HR tiles as the dataset returns them (float32 raw values, data.py:495) and/or as `_pred_array`
// sees them (uint8 clip+trunc of the centre frame, predict.py:245-246).  Pure gather/convert.
template <typename T>
__global__ void hr_gather_kernel(const void* const* sheets, const int32_t* tile_sheet, const int32_t* tile_frame,
                                 const int32_t* tile_y, const int32_t* tile_x, const int32_t* tile_vh,
                                 const int32_t* tile_vw, int sheet_h, int sheet_w, int hr_res, int hr_frame0,
                                 int hr_frames, float* hr_out, uint8_t* hr_u8, const int32_t* sheet_hs, const int32_t* sheet_ws,
                                 const int32_t* tile_xf) {
  const int tile = blockIdx.z;
  const int fo = blockIdx.y;
  if (sheet_hs != nullptr) { sheet_h = sheet_hs[tile_sheet[tile]]; sheet_w = sheet_ws[tile_sheet[tile]]; }
  const T* sheet = reinterpret_cast<const T*>(sheets[tile_sheet[tile]]);
  const T* fbase = sheet + (size_t)(tile_frame[tile] + hr_frame0 + fo) * sheet_h * sheet_w;
  const int ty = tile_y[tile], tx = tile_x[tile], vh = tile_vh[tile], vw = tile_vw[tile];
  const int centre = hr_frames / 2;  // _slice_center(x, 1) keeps index shape//2
  const size_t n = (size_t)hr_res * hr_res;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / hr_res), c = (int)(i - (size_t)r * hr_res);
    untransform(tile_xf != nullptr ? tile_xf[tile] : 0, hr_res, r, c);
    const T v = load_reflect<T>(fbase, sheet_w, ty, tx, r, c, vh, vw);
    if (hr_out) hr_out[((size_t)tile * hr_frames + fo) * n + i] = (float)v;
    if (hr_u8 && fo == centre) hr_u8[(size_t)tile * n + i] = (uint8_t)min((int)v, 255);
  }
}

// Crappifier.crappify(image) on an arbitrary float array (the reference operator interface,
// pssr/crappifiers.py:13-24): noise only, no downscale, no final round/clip.
struct NoiseK {
  StageK stages[4];
  int n_stages, clip_between;
  uint32_t seed_lo, seed_hi;
};
__global__ void noise_chain_kernel(const void* in, int in_f64, double* out, size_t n, NoiseK p) {
  const Philox ph{p.seed_lo, p.seed_hi};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = in_f64 ? reinterpret_cast<const double*>(in)[i] : (double)reinterpret_cast<const float*>(in)[i];
    // Philox counter: low 32 bits in `pix`, high bits folded into the key so >4G-element arrays stay distinct
    const Philox phi{ph.k0 ^ (uint32_t)(i >> 32), ph.k1};
    out[i] = noise_chain(v, p.stages, p.n_stages, p.clip_between, phi, (uint32_t)i, i);
  }
}

// Standalone resample (Pillow parity tests)
template <typename T>
__global__ void resize_plain_kernel(const T* src, T* dst, int n, int h, int w, int oh, int ow, const int2* bh,
                                    const int2* bw, const int32_t* kqh, const int32_t* kqw, const double* kdh,
                                    const double* kdw, int ksh, int ksw, T* tmp) {
  // pass 1: horizontal into tmp [n][h][ow]; pass 2 runs as a second launch (phase flag via dst==nullptr)
  const size_t total = dst == nullptr ? (size_t)n * h * ow : (size_t)n * oh * ow;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (dst == nullptr) {
      const int xo = (int)(i % ow);
      const size_t row = i / ow;
      const int2 b = bw[xo];
      const T* s = src + row * w + b.x;
      if (sizeof(T) == 1) {
        int32_t ss = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < b.y; ++x) ss += (int32_t)s[x] * kqw[(size_t)xo * ksw + x];
        tmp[i] = (T)min(max(ss >> kPrecisionBits, 0), 255);
      } else {
        double ss = 0.0;
        for (int x = 0; x < b.y; ++x) ss = __dadd_rn(ss, __dmul_rn((double)s[x], kdw[(size_t)xo * ksw + x]));
        const int si = (int)(ss + 0.5);
        tmp[i] = (T)((si & 255) | (min(max(si >> 8, 0), 255) << 8));
      }
    } else {
      const int xo = (int)(i % ow);
      const int yo = (int)((i / ow) % oh);
      const size_t img = i / ((size_t)ow * oh);
      const int2 b = bh[yo];
      const T* s = tmp + (img * h + b.x) * ow + xo;
      if (sizeof(T) == 1) {
        int32_t ss = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < b.y; ++y) ss += (int32_t)s[(size_t)y * ow] * kqh[(size_t)yo * ksh + y];
        dst[i] = (T)min(max(ss >> kPrecisionBits, 0), 255);
      } else {
        double ss = 0.0;
        for (int y = 0; y < b.y; ++y) ss = __dadd_rn(ss, __dmul_rn((double)s[(size_t)y * ow], kdh[(size_t)yo * ksh + y]));
        const int si = (int)(ss + 0.5);
        dst[i] = (T)((si & 255) | (min(max(si >> 8, 0), 255) << 8));
      }
    }
  }
}

}  // namespace pssr

using namespace pssr;

// One CTA copies the (few KB) tile table from pinned host memory, read in place over PCIe, into device memory.
__global__ void __launch_bounds__(256) table_fetch_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int64_t words) {
  for (int64_t i = threadIdx.x; i < words; i += 256) dst[i] = src[i];
}

extern "C" int pssr_table_fetch(void* dst, const void* pinned_src, int64_t bytes, void* stream) {
  PSSR_REQUIRE(dst != nullptr && pinned_src != nullptr && bytes >= 0 && bytes % 4 == 0, PSSR_EINVAL, "table_fetch: bad arguments");
  if (bytes == 0) return PSSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  void* dev_view = nullptr;
  if (cudaHostGetDevicePointer(&dev_view, const_cast<void*>(pinned_src), 0) != cudaSuccess || dev_view == nullptr) {
    (void)cudaGetLastError();          // not mapped into the device address space: the copy engine after all
    PSSR_CHECK_CUDA(cudaMemcpyAsync(dst, pinned_src, (size_t)bytes, cudaMemcpyHostToDevice, st));
    return PSSR_OK;
  }
  table_fetch_kernel<<<1, 256, 0, st>>>(reinterpret_cast<uint32_t*>(dst), reinterpret_cast<const uint32_t*>(dev_view), bytes / 4);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

extern "C" int pssr_crappify(const pssr_crappify_args_t* a, void* stream) {
  PSSR_REQUIRE(a != nullptr, PSSR_EINVAL, "crappify: null args");
  PSSR_REQUIRE(a->elem_bytes == 1 || a->elem_bytes == 2, PSSR_EUNSUP, "crappify: elem_bytes must be 1 (uint8) or 2 (uint16)");
  PSSR_REQUIRE(a->n_tiles >= 0 && a->frames >= 1 && a->hr_res >= 1 && a->lr_scale >= 1, PSSR_EINVAL, "crappify: bad sizes");
  PSSR_REQUIRE(a->lr_frames >= 1 && a->lr_frame0 >= 0 && a->lr_frame0 + a->lr_frames <= a->frames, PSSR_EINVAL,
               "crappify: LR frame window [%d,+%d) outside the %d frames of the tile", a->lr_frame0, a->lr_frames, a->frames);
  PSSR_REQUIRE(a->n_stages >= 0 && a->n_stages <= 4, PSSR_EINVAL, "crappify: at most 4 noise stages");
  PSSR_REQUIRE(a->sheets && a->tile_sheet && a->tile_frame && a->tile_y && a->tile_x && a->tile_vh && a->tile_vw,
               PSSR_EINVAL, "crappify: null tile table");
  if (a->n_tiles == 0) return PSSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int lr_res = a->hr_res / a->lr_scale;
  PSSR_REQUIRE(lr_res >= 1, PSSR_EINVAL, "crappify: lr_scale larger than hr_res");

  const bool want_hr = a->hr_out != nullptr || a->hr_u8_out != nullptr;
  if (want_hr)
    PSSR_REQUIRE(a->hr_frames >= 1 && a->hr_frame0 >= 0 && a->hr_frame0 + a->hr_frames <= a->frames, PSSR_EINVAL,
                 "crappify: HR frame window outside the tile");
  // the crappify CTAs emit the HR tile from their staged window when every wanted HR frame is one of the LR frames and the
  // HR grid is an integer multiple of the LR grid; otherwise a separate gather pass runs
  const bool hr_fused = want_hr && a->lr_out != nullptr && a->hr_res % a->lr_scale == 0 && a->hr_frame0 >= a->lr_frame0 &&
                        a->hr_frame0 + a->hr_frames <= a->lr_frame0 + a->lr_frames && getenv("PSSR_NO_HR_FUSE") == nullptr &&
                        ((uintptr_t)a->hr_out & 15) == 0 && ((uintptr_t)a->hr_u8_out & 3) == 0;
  if (a->lr_out != nullptr) {
    if (a->n_stages > 0) {
      const int rct = ensure_ptrs_table();
      if (rct != PSSR_OK) return rct;
    }
    const ResampleTable* tab = nullptr;
    int rc = get_table(a->hr_res, lr_res, &tab);
    if (rc != PSSR_OK) return rc;
    CrapK p;
    memset(&p, 0, sizeof(p));
    p.sheets = a->sheets;
    p.elem_bytes = a->elem_bytes;
    p.sheet_h = a->sheet_h;
    p.sheet_w = a->sheet_w;
    PSSR_REQUIRE((a->sheet_hs == nullptr) == (a->sheet_ws == nullptr), PSSR_EINVAL, "crappify: sheet_hs and sheet_ws go together");
    p.sheet_hs = a->sheet_hs;
    p.sheet_ws = a->sheet_ws;
    p.tile_sheet = a->tile_sheet; p.tile_frame = a->tile_frame; p.tile_y = a->tile_y; p.tile_x = a->tile_x;
    p.tile_vh = a->tile_vh; p.tile_vw = a->tile_vw;
    p.tile_xf = a->tile_xf;
    p.n_tiles = a->n_tiles; p.frames = a->frames; p.lr_frame0 = a->lr_frame0; p.lr_frames = a->lr_frames;
    p.hr_res = a->hr_res; p.lr_res = lr_res;
    int TL = 128 / a->lr_scale;
    if (TL < 4) TL = 4;
    if (TL > 64) TL = 64;
    if (TL > lr_res) TL = lr_res;
    p.TL = TL;
    p.tiles_per_side = (lr_res + TL - 1) / TL;
    p.ksize = tab->ksize;
    int max_span = 0;
    for (int t = 0; t < p.tiles_per_side; ++t) {
      const int x0 = t * TL, x1 = (x0 + TL < lr_res ? x0 + TL : lr_res) - 1;
      const int span = tab->h_bounds[x1].x + tab->h_bounds[x1].y - tab->h_bounds[x0].x;
      if (span > max_span) max_span = span;
    }
    p.max_rows = max_span;
    p.raw_pitch = ((max_span * a->elem_bytes + 15 + 15) / 16) * 16;  // + worst-case 15-byte lead
    p.bounds = tab->bounds; p.kq = tab->kq; p.kd = tab->kd; p.ki = tab->ki; p.dshift = tab->dshift;
    p.n_stages = a->n_stages;
    p.clip_between = a->clip_between;
    for (int s = 0; s < a->n_stages; ++s) {
      const pssr_noise_stage_t& ns = a->stages[s];
      PSSR_REQUIRE(ns.kind >= PSSR_NOISE_POISSON && ns.kind <= PSSR_NOISE_SALTPEPPER, PSSR_EINVAL, "crappify: bad noise kind %d", ns.kind);
      PSSR_REQUIRE(ns.rng == PSSR_RNG_PHILOX || ns.injected != nullptr, PSSR_EINVAL, "crappify: injected noise buffer missing for stage %d", s);
      p.stages[s].kind = ns.kind; p.stages[s].rng = ns.rng; p.stages[s].mix_in_f32 = ns.mix_in_f32;
      p.stages[s].intensity = ns.intensity; p.stages[s].gain = ns.gain; p.stages[s].injected = ns.injected;
    }
    p.seed_lo = (uint32_t)a->seed; p.seed_hi = (uint32_t)(a->seed >> 32);
    p.tile_index0 = a->tile_index0;
    p.lr_out = a->lr_out;
    p.scale = a->lr_scale;
    if (hr_fused) { p.hr_out = a->hr_out; p.hr_u8 = a->hr_u8_out; p.hr_frame0 = a->hr_frame0; p.hr_frames = a->hr_frames; }
    const size_t smem = (size_t)p.max_rows * p.raw_pitch + (size_t)p.max_rows * TL * a->elem_bytes + (size_t)p.max_rows * 4 + 32;
    PSSR_REQUIRE(smem <= 200 * 1024, PSSR_EUNSUP, "crappify: staging needs %zu bytes of shared memory (scale %d too large)", smem, a->lr_scale);
    const long long blocks = (long long)a->n_tiles * a->lr_frames * p.tiles_per_side * p.tiles_per_side;
    PSSR_REQUIRE(blocks < (1ll << 31), PSSR_EUNSUP, "crappify: too many blocks");
    const int ks = getenv("PSSR_CRAP_GENERIC") != nullptr ? 0 : (p.ksize <= 9 ? 9 : (p.ksize <= 17 ? 17 : 0));
    static const int minb = getenv("PSSR_CRAP_MINB") != nullptr ? atoi(getenv("PSSR_CRAP_MINB")) : 4;     // measured: 4 CTAs/SM +10 %
#define PSSR_CRAP_LAUNCH1(TT, KK, MB)                                                                                         \
  do {                                                                                                                        \
    static PerDeviceOnce attr;                                                                                                \
    if (attr.first()) PSSR_CHECK_CUDA(cudaFuncSetAttribute(crappify_kernel<TT, KK, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    crappify_kernel<TT, KK, MB><<<(unsigned)blocks, kCrapThreads, smem, st>>>(p);                                             \
  } while (0)
#define PSSR_CRAP_LAUNCH(TT, KK)                                                                                              \
  do {                                                                                                                        \
    if (minb == 4) PSSR_CRAP_LAUNCH1(TT, KK, 4); else PSSR_CRAP_LAUNCH1(TT, KK, 3);                                            \
  } while (0)
    if (a->elem_bytes == 1) {
      if (ks == 9) PSSR_CRAP_LAUNCH(uint8_t, 9); else if (ks == 17) PSSR_CRAP_LAUNCH(uint8_t, 17); else PSSR_CRAP_LAUNCH(uint8_t, 0);
    } else {
      if (ks == 9) PSSR_CRAP_LAUNCH(uint16_t, 9); else if (ks == 17) PSSR_CRAP_LAUNCH(uint16_t, 17); else PSSR_CRAP_LAUNCH(uint16_t, 0);
    }
#undef PSSR_CRAP_LAUNCH
#undef PSSR_CRAP_LAUNCH1
    count_launch();
    PSSR_CHECK_CUDA(cudaGetLastError());
  }
  if (want_hr && !hr_fused) {
    dim3 grid(64, a->hr_frames, a->n_tiles);
    if (a->elem_bytes == 1)
      hr_gather_kernel<uint8_t><<<grid, 256, 0, st>>>(a->sheets, a->tile_sheet, a->tile_frame, a->tile_y, a->tile_x, a->tile_vh,
                                                      a->tile_vw, a->sheet_h, a->sheet_w, a->hr_res, a->hr_frame0, a->hr_frames,
                                                      a->hr_out, a->hr_u8_out, a->sheet_hs, a->sheet_ws, a->tile_xf);
    else
      hr_gather_kernel<uint16_t><<<grid, 256, 0, st>>>(a->sheets, a->tile_sheet, a->tile_frame, a->tile_y, a->tile_x, a->tile_vh,
                                                       a->tile_vw, a->sheet_h, a->sheet_w, a->hr_res, a->hr_frame0, a->hr_frames,
                                                       a->hr_out, a->hr_u8_out, a->sheet_hs, a->sheet_ws, a->tile_xf);
    count_launch();
    PSSR_CHECK_CUDA(cudaGetLastError());
  }
  return PSSR_OK;
}

extern "C" int pssr_noise_chain(const void* in, int32_t in_is_f64, double* out, int64_t n, const pssr_noise_stage_t* stages,
                                int32_t n_stages, int32_t clip_between, uint64_t seed, void* stream) {
  PSSR_REQUIRE(in && out && n >= 0 && stages && n_stages >= 1 && n_stages <= 4, PSSR_EINVAL, "noise_chain: bad arguments");
  if (n == 0) return PSSR_OK;
  {
    const int rct = ensure_ptrs_table();
    if (rct != PSSR_OK) return rct;
  }
  NoiseK p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < n_stages; ++s) {
    const pssr_noise_stage_t& ns = stages[s];
    PSSR_REQUIRE(ns.kind >= PSSR_NOISE_POISSON && ns.kind <= PSSR_NOISE_SALTPEPPER, PSSR_EINVAL, "noise_chain: bad noise kind %d", ns.kind);
    PSSR_REQUIRE(ns.rng == PSSR_RNG_PHILOX || ns.injected != nullptr, PSSR_EINVAL, "noise_chain: injected buffer missing for stage %d", s);
    p.stages[s].kind = ns.kind; p.stages[s].rng = ns.rng; p.stages[s].mix_in_f32 = ns.mix_in_f32;
    p.stages[s].intensity = ns.intensity; p.stages[s].gain = ns.gain; p.stages[s].injected = ns.injected;
  }
  p.n_stages = n_stages; p.clip_between = clip_between;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  noise_chain_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, in_is_f64, out, (size_t)n, p);
  count_launch();
  PSSR_CHECK_CUDA(cudaGetLastError());
  return PSSR_OK;
}

extern "C" int pssr_resize_bilinear(const void* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t scale,
                                    int32_t elem_bytes, void* stream) {
  PSSR_REQUIRE(src && dst && n >= 1 && h >= 1 && w >= 1 && scale >= 1, PSSR_EINVAL, "resize: bad arguments");
  PSSR_REQUIRE(elem_bytes == 1 || elem_bytes == 2, PSSR_EUNSUP, "resize: elem_bytes must be 1 or 2");
  const int oh = h / scale, ow = w / scale;
  PSSR_REQUIRE(oh >= 1 && ow >= 1, PSSR_EINVAL, "resize: scale larger than the image");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const ResampleTable *th = nullptr, *tw = nullptr;
  int rc = get_table(h, oh, &th);
  if (rc != PSSR_OK) return rc;
  rc = get_table(w, ow, &tw);
  if (rc != PSSR_OK) return rc;
  void* tmp = nullptr;
  PSSR_CHECK_CUDA(cudaMallocAsync(&tmp, (size_t)n * h * ow * elem_bytes, st));
  const int blocks = device_sm_count() * 8;
  if (elem_bytes == 1) {
    resize_plain_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)src, nullptr, n, h, w, oh, ow, th->bounds, tw->bounds, th->kq,
                                                         tw->kq, th->kd, tw->kd, th->ksize, tw->ksize, (uint8_t*)tmp);
    resize_plain_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, n, h, w, oh, ow, th->bounds, tw->bounds,
                                                         th->kq, tw->kq, th->kd, tw->kd, th->ksize, tw->ksize, (uint8_t*)tmp);
  } else {
    resize_plain_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)src, nullptr, n, h, w, oh, ow, th->bounds, tw->bounds,
                                                          th->kq, tw->kq, th->kd, tw->kd, th->ksize, tw->ksize, (uint16_t*)tmp);
    resize_plain_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)src, (uint16_t*)dst, n, h, w, oh, ow, th->bounds,
                                                          tw->bounds, th->kq, tw->kq, th->kd, tw->kd, th->ksize, tw->ksize, (uint16_t*)tmp);
  }
  count_launch(2);
  PSSR_CHECK_CUDA(cudaGetLastError());
  PSSR_CHECK_CUDA(cudaFreeAsync(tmp, st));
  return PSSR_OK;
}
