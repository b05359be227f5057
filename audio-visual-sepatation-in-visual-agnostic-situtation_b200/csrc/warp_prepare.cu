// K2: log-frequency warp (utils.warpgrid + F.grid_sample, main.py:61-66 / 216-220) and the fused
// NetWrapper.prepare derivations (main.py:51-95).  HBM-bound: every input row is read once from
// DRAM (neighbouring taps hit L1), every output element is written once.
//
// The bilinear resample is separable (warp_taps.h).  A thread owns one output column and walks a
// block of output rows; the x-interpolated value of the two most recent input rows is kept in
// registers because consecutive output rows share taps on the low-frequency half of the axis.
#include "common.cuh"
#include "warp_taps.h"

namespace avsep {

constexpr int WARP_ROWS_PER_CTA = 8;

// x-interpolated sample of input row `y` (zero when the row is out of range)
__device__ __forceinline__ float row_sample(const float* __restrict__ base, int y, int H, int T, int x0, float fx, float eps) {
  if (y < 0 || y >= H) return 0.f;
  const float* r = base + (size_t)y * T;
  float a = 0.f, b = 0.f;
  if (x0 >= 0) a = __ldg(r + x0) + eps;
  if (x0 + 1 < T) b = __ldg(r + x0 + 1) + eps;
  return (1.f - fx) * a + fx * b;
}

struct RowCache {
  int y0 = -2, y1 = -2;
  float v0 = 0.f, v1 = 0.f;
  __device__ __forceinline__ float get(const float* __restrict__ base, int y, int H, int T, int x0, float fx, float eps) {
    if (y == y0) return v0;
    if (y == y1) return v1;
    return row_sample(base, y, H, T, x0, fx, eps);
  }
  __device__ __forceinline__ void put(int ya, float a, int yb, float b) {
    y0 = ya; v0 = a; y1 = yb; v1 = b;
  }
};

__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ x, int Hin, int Hout, int T, Taps ytap, Taps xtap,
                                                    float* __restrict__ y) {
  const int s = blockIdx.z;
  const int h0 = blockIdx.y * WARP_ROWS_PER_CTA;
  const float* base = x + (size_t)s * Hin * T;
  float* out = y + (size_t)s * Hout * T;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < T; w += gridDim.x * blockDim.x) {
    const int x0 = xtap.i0[w];
    const float fx = xtap.frac[w];
    RowCache rc;
    for (int h = h0; h < min(h0 + WARP_ROWS_PER_CTA, Hout); ++h) {
      const int y0 = ytap.i0[h];
      const float fy = ytap.frac[h];
      const float a = rc.get(base, y0, Hin, T, x0, fx, 0.f);
      const float b = rc.get(base, y0 + 1, Hin, T, x0, fx, 0.f);
      rc.put(y0, a, y0 + 1, b);
      out[(size_t)h * T + w] = (1.f - fy) * a + fy * b;
    }
  }
}

struct PrepareParams {
  const float* mix;                    // [B, Hin, T]
  const float* src[AVSEP_MAX_SRC];     // [B, Hin, T] each
  float* mix_w;                        // [B, Hout, T]
  float* src_w[AVSEP_MAX_SRC];
  float* log_mix;
  float* weights;
  float* gt[AVSEP_MAX_SRC];
  int B, Hin, Hout, T;
  int weighted, binary;
  Taps ytap, xtap;
};

template <int NS, bool WARP>
__global__ void __launch_bounds__(256) prepare_kernel(const PrepareParams p) {
  const int b = blockIdx.z;
  const int h0 = blockIdx.y * WARP_ROWS_PER_CTA;
  const size_t in_off = (size_t)b * p.Hin * p.T;
  const size_t out_off = (size_t)b * p.Hout * p.T;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < p.T; w += gridDim.x * blockDim.x) {
    int x0 = 0;
    float fx = 0.f;
    if (WARP) {
      x0 = p.xtap.i0[w];
      fx = p.xtap.frac[w];
    }
    RowCache rc[NS + 1];
    for (int h = h0; h < min(h0 + WARP_ROWS_PER_CTA, p.Hout); ++h) {
      float v[NS + 1];
      if (WARP) {
        const int y0 = p.ytap.i0[h];
        const float fy = p.ytap.frac[h];
#pragma unroll
        for (int j = 0; j <= NS; ++j) {
          const float* base = (j == 0 ? p.mix : p.src[j - 1]) + in_off;
          const float eps = j == 0 ? 1e-10f : 0.f;             // mag_mix = mag_mix + 1e-10 before the warp (main.py:54)
          const float a = rc[j].get(base, y0, p.Hin, p.T, x0, fx, eps);
          const float c = rc[j].get(base, y0 + 1, p.Hin, p.T, x0, fx, eps);
          rc[j].put(y0, a, y0 + 1, c);
          v[j] = (1.f - fy) * a + fy * c;
        }
      } else {
#pragma unroll
        for (int j = 0; j <= NS; ++j) {
          const float* base = (j == 0 ? p.mix : p.src[j - 1]) + in_off;
          v[j] = __ldg(base + (size_t)h * p.T + w) + (j == 0 ? 1e-10f : 0.f);
        }
      }
      const size_t o = out_off + (size_t)h * p.T + w;
      const float mix = v[0];
      if (p.mix_w) p.mix_w[o] = mix;
      if (p.log_mix) p.log_mix[o] = logf(mix);                                   // main.py:85
      if (p.weights) p.weights[o] = p.weighted ? fminf(fmaxf(log1pf(mix), 1e-3f), 10.f) : 1.f;   // main.py:69-73
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        if (p.src_w[n]) p.src_w[n][o] = v[n + 1];
        if (p.gt[n]) {
          float g;
          if (p.binary) g = v[n + 1] > 0.5f * mix ? 1.f : 0.f;                   // main.py:79
          else g = fminf(fmaxf(v[n + 1] / mix, 0.f), 5.f);                       // main.py:81-82
          p.gt[n][o] = g;
        }
      }
    }
  }
}

template <int NS>
static int launch_prepare(const PrepareParams& p, bool warp, cudaStream_t stream) {
  dim3 grid(ceil_div(p.T, 256), ceil_div(p.Hout, WARP_ROWS_PER_CTA), p.B);
  if (warp) prepare_kernel<NS, true><<<grid, 256, 0, stream>>>(p);
  else prepare_kernel<NS, false><<<grid, 256, 0, stream>>>(p);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int prepare_impl(const float* mag_mix, const float* const* mags, int N, int B, int T, int log_freq, int weighted,
                 int binary, float* mix_w, float* const* mags_w, float* log_mix, float* weights, float* const* gt,
                 cudaStream_t stream) {
  PrepareParams p;
  memset(&p, 0, sizeof(p));
  p.mix = mag_mix;
  for (int n = 0; n < N; ++n) {
    p.src[n] = mags[n];
    p.src_w[n] = mags_w ? mags_w[n] : nullptr;
    p.gt[n] = gt ? gt[n] : nullptr;
  }
  p.mix_w = mix_w;
  p.log_mix = log_mix;
  p.weights = weights;
  p.B = B;
  p.Hin = AVSEP_N_BINS;
  p.Hout = log_freq ? 256 : AVSEP_N_BINS;
  p.T = T;
  p.weighted = weighted;
  p.binary = binary;
  if (log_freq) {
    int rc = get_taps(TAP_WARP_Y, AVSEP_N_BINS, 256, &p.ytap);
    if (rc) return rc;
    rc = get_taps(TAP_LINEAR, T, T, &p.xtap);
    if (rc) return rc;
  }
  switch (N) {
    case 0: return launch_prepare<0>(p, log_freq != 0, stream);
    case 1: return launch_prepare<1>(p, log_freq != 0, stream);
    case 2: return launch_prepare<2>(p, log_freq != 0, stream);
    case 3: return launch_prepare<3>(p, log_freq != 0, stream);
    default: return launch_prepare<4>(p, log_freq != 0, stream);
  }
}

}  // namespace avsep

using namespace avsep;

extern "C" {

int avsep_warp(const float* x_dev, int S, int Hin, int Hout, int T, int warp, float* y_dev, void* stream) {
  AVSEP_REQUIRE(x_dev && y_dev, "avsep_warp: null tensor");
  AVSEP_REQUIRE(S >= 0 && Hin >= 1 && Hout >= 1 && T >= 1, "avsep_warp: bad shape S=%d Hin=%d Hout=%d T=%d", S, Hin, Hout, T);
  AVSEP_REQUIRE(S <= 65535, "avsep_warp: S = %d exceeds 65535 per call", S);
  int rc = require_sm100();
  if (rc) return rc;
  if (S == 0) return AVSEP_OK;
  Taps yt, xt;
  rc = get_taps(warp ? TAP_WARP_Y : TAP_UNWARP_Y, Hin, Hout, &yt);
  if (rc) return rc;
  rc = get_taps(TAP_LINEAR, T, T, &xt);
  if (rc) return rc;
  dim3 grid(ceil_div(T, 256), ceil_div(Hout, WARP_ROWS_PER_CTA), S);
  warp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, Hin, Hout, T, yt, xt, y_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_prepare(const float* mag_mix_dev, const float* const* mags_dev, int N, int B, int T, int log_freq,
                  int weighted_loss, int binary_mask, float* mag_mix_w_dev, float* const* mags_w_dev,
                  float* log_mag_mix_dev, float* weights_dev, float* const* gt_masks_dev, void* stream) {
  AVSEP_REQUIRE(mag_mix_dev != nullptr, "avsep_prepare: mag_mix_dev is null");
  AVSEP_REQUIRE(N >= 0 && N <= AVSEP_MAX_SRC, "avsep_prepare: N = %d outside [0, %d]", N, AVSEP_MAX_SRC);
  AVSEP_REQUIRE(N == 0 || mags_dev, "avsep_prepare: mags_dev is null with N = %d", N);
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && T >= 1, "avsep_prepare: bad B = %d / T = %d", B, T);
  for (int n = 0; n < N; ++n) AVSEP_REQUIRE(mags_dev[n] != nullptr, "avsep_prepare: mags_dev[%d] is null", n);
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  return prepare_impl(mag_mix_dev, mags_dev, N, B, T, log_freq, weighted_loss, binary_mask, mag_mix_w_dev, mags_w_dev,
                      log_mag_mix_dev, weights_dev, gt_masks_dev, (cudaStream_t)stream);
}

}  // extern "C"
