// The criterion as a stand-alone op (models/criterion.py:6-49) for predictions that do not come
// from the synthesizer (root main.py feeds U-Net logits directly, main.py:131-136), and the
// S x S loss matrix of PitWrapper (models/criterion.py:159-176).  Pure streaming reductions.
#include <algorithm>

#include "common.cuh"

namespace avsep {

constexpr int LOSS_MAX_LIST = 8;
constexpr int LOSS_THREADS = 256;
constexpr int LOSS_ELEMS_PER_CTA = LOSS_THREADS * 4 * 4;

struct PtrList {
  const float* p[LOSS_MAX_LIST];
  const float* t[LOSS_MAX_LIST];
  float* d[LOSS_MAX_LIST];
};

__device__ __forceinline__ float l_elem(int kind, float p, float t) {
  if (kind == AVSEP_LOSS_BCE) {
    const float lp = fmaxf(logf(p), -100.f);
    const float l1p = fmaxf(log1pf(-p), -100.f);
    return (t - 1.f) * l1p - t * lp;
  }
  const float d = p - t;
  return kind == AVSEP_LOSS_L1 ? fabsf(d) : d * d;
}
__device__ __forceinline__ float l_grad(int kind, float p, float t) {
  if (kind == AVSEP_LOSS_BCE) return (p - t) / fmaxf((1.f - p) * p, 1e-12f);
  const float d = p - t;
  if (kind == AVSEP_LOSS_L1) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  return 2.f * d;
}

__device__ __forceinline__ double cta_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_fwd_kernel(int kind, PtrList L, const float* __restrict__ weight, int N,
                                                                 size_t numel, double* __restrict__ partial) {
  __shared__ double red[32];
  const size_t base = (size_t)blockIdx.x * LOSS_ELEMS_PER_CTA;
  float acc = 0.f;
  for (int it = 0; it < 4; ++it) {
    const size_t i = base + ((size_t)it * LOSS_THREADS + threadIdx.x) * 4;
    if (i + 4 <= numel && (numel & 3) == 0) {
      float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
      if (weight) w = *reinterpret_cast<const float4*>(weight + i);
      for (int n = 0; n < N; ++n) {
        const float4 p = ld_stream(reinterpret_cast<const float4*>(L.p[n] + i));
        const float4 t = ld_stream(reinterpret_cast<const float4*>(L.t[n] + i));
        acc += w.x * l_elem(kind, p.x, t.x) + w.y * l_elem(kind, p.y, t.y) + w.z * l_elem(kind, p.z, t.z) +
               w.w * l_elem(kind, p.w, t.w);
      }
    } else {
      for (size_t j = i; j < min(i + 4, numel); ++j) {
        const float w = weight ? weight[j] : 1.f;
        for (int n = 0; n < N; ++n) acc += w * l_elem(kind, L.p[n][j], L.t[n][j]);
      }
    }
  }
  const double tot = cta_sum((double)acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) loss_fin_kernel(const double* __restrict__ partial, int n, double inv_count,
                                                         float* __restrict__ loss) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  const double tot = cta_sum(s, red);
  if (threadIdx.x == 0) loss[0] = (float)(tot * inv_count);
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_kernel(int kind, PtrList L, const float* __restrict__ weight, int N,
                                                                 size_t numel, const float* __restrict__ gscale, float inv_count) {
  const float g = gscale[0] * inv_count;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (size_t)gridDim.x * blockDim.x) {
    const float w = weight ? weight[i] : 1.f;
    for (int n = 0; n < N; ++n)
      if (L.d[n]) L.d[n][i] = g * w * l_grad(kind, L.p[n][i], L.t[n][i]);
  }
}

// x *= g[0] unless g[0] == 1 (then the kernel touches no memory): lets a fused forward+backward hand out gradients
// computed for an upstream gradient of 1 and still honour an arbitrary upstream scalar without a host sync.
__global__ void __launch_bounds__(256) scale_if_needed_kernel(float* __restrict__ x, size_t n4, size_t n, const float* __restrict__ g) {
  const float s = g[0];
  if (s == 1.0f) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(x)[i] = v;
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) x[i] *= s;
}

// PIT: lossmat[b][i][j] = mean_p w[b,p,i] * l(pred[b,p,j], tgt[b,p,i]); sources innermost.
template <int S>
__global__ void __launch_bounds__(256) pit_lossmat_kernel(int kind, const float* __restrict__ preds, const float* __restrict__ tgts,
                                                           const float* __restrict__ wts, size_t P, double* __restrict__ partial) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  float acc[S][S];
#pragma unroll
  for (int i = 0; i < S; ++i)
#pragma unroll
    for (int j = 0; j < S; ++j) acc[i][j] = 0.f;
  const size_t off = (size_t)b * P * S;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float pr[S], tg[S], ww[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      pr[s] = preds[off + p * S + s];
      tg[s] = tgts[off + p * S + s];
      ww[s] = wts[off + p * S + s];
    }
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
      for (int j = 0; j < S; ++j) acc[i][j] += ww[i] * l_elem(kind, pr[j], tg[i]);
  }
#pragma unroll
  for (int i = 0; i < S; ++i)
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const double tot = cta_sum((double)acc[i][j], red);
      if (threadIdx.x == 0) partial[(((size_t)b * gridDim.x + blockIdx.x) * S + i) * S + j] = tot;
    }
}

// d loss[b] / d pred[b,p,j] for loss[b] = (1/S) sum_i lossmat[b][i][perm_b(i)]: prediction j is matched with
// target i = inv_perm[b][j]  ->  g[b] / (S P) * w[b,p,i] * l'(pred[b,p,j], tgt[b,p,i])
template <int S>
__global__ void __launch_bounds__(256) pit_bwd_kernel(int kind, const float* __restrict__ preds, const float* __restrict__ tgts,
                                                       const float* __restrict__ wts, const int* __restrict__ inv_perm,
                                                       const float* __restrict__ gloss, size_t P, float* __restrict__ dpreds) {
  const int b = blockIdx.y;
  int ip[S];
#pragma unroll
  for (int j = 0; j < S; ++j) ip[j] = inv_perm[b * S + j];
  const float g = gloss[b] / ((float)S * (float)P);
  const size_t off = (size_t)b * P * S;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float pr[S], tg[S], ww[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      pr[s] = preds[off + p * S + s];
      tg[s] = tgts[off + p * S + s];
      ww[s] = wts[off + p * S + s];
    }
#pragma unroll
    for (int j = 0; j < S; ++j) {
      float t = tg[0], w = ww[0];
#pragma unroll
      for (int i = 1; i < S; ++i)
        if (ip[j] == i) { t = tg[i]; w = ww[i]; }
      dpreds[off + p * S + j] = g * w * l_grad(kind, pr[j], t);
    }
  }
}

__global__ void pit_fin_kernel(const double* __restrict__ partial, int chunks, int SS, double inv_P, float* __restrict__ lossmat) {
  const int b = blockIdx.x;
  const int e = threadIdx.x;
  if (e >= SS) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[((size_t)b * chunks + c) * SS + e];
  lossmat[(size_t)b * SS + e] = (float)(s * inv_P);
}

// PitWrapper._fast_pit (criterion.py:113-136) on the device: one thread per sample walks the S! permutations in
// itertools.permutations order, c = mean_i lossmat[i][perm[i]] in fp32 (as torch's .mean() of S fp32 values), keeps the
// FIRST strict minimum (`loss > c_loss`), and writes loss, the permutation and its inverse.  No D2H of the matrix.
__global__ void pit_search_kernel(const float* __restrict__ lossmat, int B, int S, float* __restrict__ loss, int* __restrict__ perm,
                                  int* __restrict__ inv_perm) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* lm = lossmat + (size_t)b * S * S;
  int cur[4] = {0, 1, 2, 3}, best_p[4] = {0, 1, 2, 3};
  float best = 0.f;
  bool have = false;
  while (true) {
    float sum = 0.f;
    for (int i = 0; i < S; ++i) sum += lm[i * S + cur[i]];
    const float c = sum / (float)S;
    if (!have || best > c) {
      best = c;
      have = true;
      for (int i = 0; i < S; ++i) best_p[i] = cur[i];
    }
    // next permutation in lexicographic order (= itertools.permutations of range(S))
    int k = S - 2;
    while (k >= 0 && cur[k] > cur[k + 1]) --k;
    if (k < 0) break;
    int l = S - 1;
    while (cur[l] < cur[k]) --l;
    int t = cur[k]; cur[k] = cur[l]; cur[l] = t;
    for (int lo = k + 1, hi = S - 1; lo < hi; ++lo, --hi) { t = cur[lo]; cur[lo] = cur[hi]; cur[hi] = t; }
  }
  loss[b] = best;
  for (int i = 0; i < S; ++i) {
    perm[b * S + i] = best_p[i];
    inv_perm[b * S + best_p[i]] = i;
  }
}

// PitWrapper.reorder_tensor (criterion.py:180-200): y[b, ..., s] = x[b, ..., perm[b][s]], one streaming pass
template <int S>
__global__ void __launch_bounds__(256) pit_reorder_kernel(const float* __restrict__ x, const int* __restrict__ perm, size_t P,
                                                          float* __restrict__ y) {
  const int b = blockIdx.y;
  int pm[S];
#pragma unroll
  for (int s = 0; s < S; ++s) pm[s] = perm[b * S + s];
  const float* xb = x + (size_t)b * P * S;
  float* yb = y + (size_t)b * P * S;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
    float v[S];
#pragma unroll
    for (int s = 0; s < S; ++s) v[s] = xb[i * S + s];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float o = v[0];
#pragma unroll
      for (int q = 1; q < S; ++q) o = pm[s] == q ? v[q] : o;
      yb[i * S + s] = o;
    }
  }
}

}  // namespace avsep

using namespace avsep;

extern "C" {

size_t avsep_loss_scratch_bytes(int N, size_t numel) {
  (void)N;
  const size_t ctas = (numel + LOSS_ELEMS_PER_CTA - 1) / LOSS_ELEMS_PER_CTA;
  return sizeof(double) * (ctas + 1) + 256;
}

int avsep_loss_forward(int loss_kind, const float* const* preds_dev, const float* const* targets_dev, const float* weight_dev,
                       int N, size_t numel, float* loss_dev, void* scratch_dev, void* stream) {
  AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_L1 && loss_kind <= AVSEP_LOSS_BCE, "avsep_loss_forward: unknown loss kind %d", loss_kind);
  AVSEP_REQUIRE(N >= 1 && N <= LOSS_MAX_LIST, "avsep_loss_forward: N = %d outside [1, %d]", N, LOSS_MAX_LIST);
  AVSEP_REQUIRE(preds_dev && targets_dev && loss_dev && scratch_dev, "avsep_loss_forward: null argument");
  AVSEP_REQUIRE(numel >= 1, "avsep_loss_forward: empty input");
  int rc = require_sm100();
  if (rc) return rc;
  PtrList L;
  memset(&L, 0, sizeof(L));
  for (int n = 0; n < N; ++n) {
    AVSEP_REQUIRE(preds_dev[n] && targets_dev[n], "avsep_loss_forward: null tensor in list at %d", n);
    L.p[n] = preds_dev[n];
    L.t[n] = targets_dev[n];
  }
  const size_t ctas = (numel + LOSS_ELEMS_PER_CTA - 1) / LOSS_ELEMS_PER_CTA;
  AVSEP_REQUIRE(ctas <= 0x7fffffffULL, "avsep_loss_forward: input too large");
  double* partial = static_cast<double*>(scratch_dev);
  loss_fwd_kernel<<<(unsigned)ctas, LOSS_THREADS, 0, (cudaStream_t)stream>>>(loss_kind, L, weight_dev, N, numel, partial);
  AVSEP_LAUNCH_CHECK();
  loss_fin_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partial, (int)ctas, 1.0 / ((double)numel * (double)N), loss_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_loss_backward(int loss_kind, const float* const* preds_dev, const float* const* targets_dev, const float* weight_dev,
                        int N, size_t numel, const float* gscale_dev, float* const* dpreds_dev, void* stream) {
  AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_L1 && loss_kind <= AVSEP_LOSS_BCE, "avsep_loss_backward: unknown loss kind %d", loss_kind);
  AVSEP_REQUIRE(N >= 1 && N <= LOSS_MAX_LIST, "avsep_loss_backward: N = %d outside [1, %d]", N, LOSS_MAX_LIST);
  AVSEP_REQUIRE(preds_dev && targets_dev && gscale_dev && dpreds_dev, "avsep_loss_backward: null argument");
  int rc = require_sm100();
  if (rc) return rc;
  if (numel == 0) return AVSEP_OK;
  PtrList L;
  memset(&L, 0, sizeof(L));
  for (int n = 0; n < N; ++n) {
    L.p[n] = preds_dev[n];
    L.t[n] = targets_dev[n];
    L.d[n] = dpreds_dev[n];
  }
  const unsigned grid = (unsigned)std::min<size_t>((numel + LOSS_THREADS - 1) / LOSS_THREADS, (size_t)148 * 16);
  loss_bwd_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)stream>>>(loss_kind, L, weight_dev, N, numel, gscale_dev,
                                                                   (float)(1.0 / ((double)numel * (double)N)));
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_scale_if_needed(float* x_dev, size_t n, const float* gscale_dev, void* stream) {
  AVSEP_REQUIRE(x_dev && gscale_dev, "avsep_scale_if_needed: null argument");
  int rc = require_sm100();
  if (rc) return rc;
  if (n == 0) return AVSEP_OK;
  const size_t n4 = ((reinterpret_cast<uintptr_t>(x_dev) & 15) == 0) ? n / 4 : 0;
  const unsigned grid = (unsigned)std::min<size_t>((n4 + 255) / 256 + 1, (size_t)148 * 8);
  scale_if_needed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n4, n, gscale_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

// scratch is internal here: the loss matrix is tiny; partial sums live in a per-call allocation
// from the stream-ordered pool so the call stays asynchronous.
int avsep_pit_lossmat(int loss_kind, const float* preds_dev, const float* targets_dev, const float* weights_dev, int B, size_t P,
                      int S, float* lossmat_dev, void* stream_) {
  AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_L1 && loss_kind <= AVSEP_LOSS_BCE, "avsep_pit_lossmat: unknown loss kind %d", loss_kind);
  AVSEP_REQUIRE(preds_dev && targets_dev && weights_dev && lossmat_dev, "avsep_pit_lossmat: null tensor");
  AVSEP_REQUIRE(S >= 1 && S <= 4, "avsep_pit_lossmat: S = %d outside [1, 4]", S);
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && P >= 1, "avsep_pit_lossmat: bad B / P");
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int chunks = (int)std::min<size_t>((P + 255) / 256, 64);
  double* partial = nullptr;
  AVSEP_CUDA(cudaMallocAsync(&partial, sizeof(double) * (size_t)B * chunks * S * S, stream));
  dim3 grid(chunks, B);
  switch (S) {
    case 1: pit_lossmat_kernel<1><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, P, partial); break;
    case 2: pit_lossmat_kernel<2><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, P, partial); break;
    case 3: pit_lossmat_kernel<3><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, P, partial); break;
    default: pit_lossmat_kernel<4><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, P, partial); break;
  }
  AVSEP_LAUNCH_CHECK();
  pit_fin_kernel<<<B, 32, 0, stream>>>(partial, chunks, S * S, 1.0 / (double)P, lossmat_dev);
  AVSEP_LAUNCH_CHECK();
  AVSEP_CUDA(cudaFreeAsync(partial, stream));
  return AVSEP_OK;
}

int avsep_pit_backward(int loss_kind, const float* preds_dev, const float* targets_dev, const float* weights_dev,
                       const int* inv_perm_dev, const float* gloss_dev, int B, size_t P, int S, float* dpreds_dev, void* stream_) {
  AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_L1 && loss_kind <= AVSEP_LOSS_BCE, "avsep_pit_backward: unknown loss kind %d", loss_kind);
  AVSEP_REQUIRE(preds_dev && targets_dev && weights_dev && inv_perm_dev && gloss_dev && dpreds_dev, "avsep_pit_backward: null tensor");
  AVSEP_REQUIRE(S >= 1 && S <= 4, "avsep_pit_backward: S = %d outside [1, 4]", S);
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && P >= 1, "avsep_pit_backward: bad B / P");
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  dim3 grid((unsigned)std::min<size_t>((P + 255) / 256, 256), B);
  switch (S) {
    case 1: pit_bwd_kernel<1><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, inv_perm_dev, gloss_dev, P, dpreds_dev); break;
    case 2: pit_bwd_kernel<2><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, inv_perm_dev, gloss_dev, P, dpreds_dev); break;
    case 3: pit_bwd_kernel<3><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, inv_perm_dev, gloss_dev, P, dpreds_dev); break;
    default: pit_bwd_kernel<4><<<grid, 256, 0, stream>>>(loss_kind, preds_dev, targets_dev, weights_dev, inv_perm_dev, gloss_dev, P, dpreds_dev); break;
  }
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_pit_search(const float* lossmat_dev, int B, int S, float* loss_dev, int* perm_dev, int* inv_perm_dev, void* stream) {
  AVSEP_REQUIRE(lossmat_dev && loss_dev && perm_dev && inv_perm_dev, "avsep_pit_search: null tensor");
  AVSEP_REQUIRE(S >= 1 && S <= 4 && B >= 0, "avsep_pit_search: bad B = %d / S = %d (S in [1, 4])", B, S);
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  pit_search_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(lossmat_dev, B, S, loss_dev, perm_dev, inv_perm_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_pit_reorder(const float* x_dev, const int* perm_dev, int B, size_t P, int S, float* y_dev, void* stream_) {
  AVSEP_REQUIRE(x_dev && perm_dev && y_dev, "avsep_pit_reorder: null tensor");
  AVSEP_REQUIRE(S >= 1 && S <= 4 && B >= 0 && B <= 65535 && P >= 1, "avsep_pit_reorder: bad B / P / S");
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  dim3 grid((unsigned)std::min<size_t>((P + 255) / 256, 1024), B);
  switch (S) {
    case 1: pit_reorder_kernel<1><<<grid, 256, 0, stream>>>(x_dev, perm_dev, P, y_dev); break;
    case 2: pit_reorder_kernel<2><<<grid, 256, 0, stream>>>(x_dev, perm_dev, P, y_dev); break;
    case 3: pit_reorder_kernel<3><<<grid, 256, 0, stream>>>(x_dev, perm_dev, P, y_dev); break;
    default: pit_reorder_kernel<4><<<grid, 256, 0, stream>>>(x_dev, perm_dev, P, y_dev); break;
  }
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

}  // extern "C"
