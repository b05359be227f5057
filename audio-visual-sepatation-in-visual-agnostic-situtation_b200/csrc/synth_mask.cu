// K3: the synthesizer mask -- InnerProd / Bias (models/synthesizer_net.py:6-70) for all N sources of
// a mixture in one pass over feat_sound, fused with the caller's sigmoid (models/__init__.py:16-18)
// and the list-form weighted loss (models/criterion.py:10-49), plus the backward pass.
//
// HBM-bound: feat_sound [B,K,HW] (8.4 MB per mixture at K=32, 256x256) is the dominant stream.  The
// reference reads it once per source in the forward (cuBLAS M=1 batched GEMM) and materialises one
// [B,K,H,W] outer product per source in the backward; here it is streamed exactly once per pass
// with 128-bit loads, the K-vector of every source lives in shared memory, the per-channel
// reductions for d(vec)/d(scale)/d(bias) use warp shuffles, and the loss is reduced in the same pass.
#include <algorithm>

#include "common.cuh"

namespace avsep {

constexpr int SYN_FWD_THREADS = 256;
constexpr int SYN_BWD_THREADS = 128;
constexpr int SYN_BWD_ITERS = 4;     // float4 groups per thread
constexpr int SYN_KCHUNK = 32;       // channels per backward CTA (grid.z walks the chunks)
constexpr int SYN_BWD_LOADS = 8;     // 128-bit loads in flight per thread
constexpr int SYN_MAX_K = 256;

__device__ __forceinline__ float act_apply(int act, float z) {
  // torch.sigmoid in fp32: 1 / (1 + exp(-z)); saturates to exactly 0 / 1 like the reference
  return act == AVSEP_ACT_SIGMOID ? 1.0f / (1.0f + expf(-z)) : z;
}

__device__ __forceinline__ float loss_elem(int kind, float p, float t) {
  if (kind == AVSEP_LOSS_BCE) {
    // F.binary_cross_entropy: -(t * max(log p, -100) + (1 - t) * max(log(1 - p), -100))
    const float lp = fmaxf(logf(p), -100.f);
    const float l1p = fmaxf(log1pf(-p), -100.f);
    return (t - 1.f) * l1p - t * lp;
  }
  const float d = p - t;
  return kind == AVSEP_LOSS_L1 ? fabsf(d) : d * d;
}

// d loss_elem / d pred (torch autograd semantics), without the 1/(numel*N) mean factor
__device__ __forceinline__ float loss_grad(int kind, float p, float t) {
  if (kind == AVSEP_LOSS_BCE) return (p - t) / fmaxf((1.f - p) * p, 1e-12f);   // ATen EPSILON = 1e-12
  const float d = p - t;
  if (kind == AVSEP_LOSS_L1) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  return 2.f * d;
}

// kernel variants of the loss element / gradient (forward, fused training and PIT kernels): same formulas, MUFU-based logs
// (3 ulp; the correctly rounded logf / log1pf cost ~50 instructions per mask value and made the forward kernel issue-bound) and, for
// sigmoid + BCE, the closed form of torch's two-step gradient
//   (p - t) / max(p (1 - p), 1e-12) * p (1 - p)  ==  (p - t) * min(1, p (1 - p) * 1e12)
// which keeps the exact zero where the float32 sigmoid saturates.
__device__ __forceinline__ float loss_elem_fast(int kind, float p, float t) {
  if (kind == AVSEP_LOSS_BCE) {
    const float lp = fmaxf(fast_log(p), -100.f);
    const float l1p = fmaxf(fast_log1m(p), -100.f);
    return (t - 1.f) * l1p - t * lp;
  }
  const float d = p - t;
  return kind == AVSEP_LOSS_L1 ? fabsf(d) : d * d;
}
__device__ __forceinline__ float dz_elem_fast(int kind, int act, float p, float t) {
  if (kind == AVSEP_LOSS_BCE && act == AVSEP_ACT_SIGMOID) return (p - t) * fminf(1.f, (1.f - p) * p * 1e12f);
  float d = loss_grad(kind, p, t);
  if (act == AVSEP_ACT_SIGMOID) d *= (1.f - p) * p;
  return d;
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = 0;
  if (warp == 0) {
    r = lane < nwarps ? smem[lane] : (T)0;
    r = warp_sum(r);
  }
  return r;   // valid in warp 0
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int NS, int VEC>
__global__ void __launch_bounds__(SYN_FWD_THREADS)
synth_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ vecs, const float* __restrict__ scale,
                 const float* __restrict__ bias, int B, int K, int HW, int act, float* __restrict__ out, int loss_kind,
                 const float* __restrict__ gt, const float* __restrict__ weight, double* __restrict__ partial) {
  __shared__ float vs[NS][SYN_MAX_K];
  __shared__ double red[32];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < NS * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    vs[n][k] = vecs[((size_t)b * NS + n) * K + k] * (scale ? scale[k] : 1.f);
  }
  __syncthreads();
  const float bz = bias ? bias[0] : 0.f;
  const int px = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  double lsum = 0.0;
  if (px < HW) {
    float acc[NS][VEC];
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[n][e] = 0.f;
    const float* fp = feat + (size_t)b * K * HW + px;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      float f[VEC];
      if (VEC == 4) {
        const float4 t4 = ld_stream(reinterpret_cast<const float4*>(fp + (size_t)k * HW));
        f[0] = t4.x; f[1] = t4.y; f[2] = t4.z; f[3] = t4.w;
      } else {
        f[0] = __ldg(fp + (size_t)k * HW);
      }
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float c = vs[n][k];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[n][e] = fmaf(c, f[e], acc[n][e]);
      }
    }
    float w[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) w[e] = 1.f;
    if (loss_kind != AVSEP_LOSS_NONE && weight) {
      if (VEC == 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(weight + (size_t)b * HW + px);
        w[0] = t4.x; w[1] = t4.y; w[2] = t4.z; w[3] = t4.w;
      } else {
        w[0] = weight[(size_t)b * HW + px];
      }
    }
    float ls = 0.f;
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      float pr[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) pr[e] = act_apply(act, acc[n][e] + bz);
      const size_t o = ((size_t)n * B + b) * HW + px;
      if (VEC == 4) *reinterpret_cast<float4*>(out + o) = make_float4(pr[0], pr[1], pr[2], pr[3]);
      else out[o] = pr[0];
      if (loss_kind != AVSEP_LOSS_NONE) {
        float t[VEC];
        if (VEC == 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(gt + o);
          t[0] = t4.x; t[1] = t4.y; t[2] = t4.z; t[3] = t4.w;
        } else {
          t[0] = gt[o];
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) ls += w[e] * loss_elem_fast(loss_kind, pr[e], t[e]);
      }
    }
    lsum = (double)ls;
  }
  if (loss_kind != AVSEP_LOSS_NONE) {
    const double tot = block_sum<double>(lsum, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
  }
}

// ---------------------------------------------------------------------------------------------
// Forward (+ loss) with the feat stream staged through shared memory: the eval step and plain InnerProd.forward when
// K <= 32 and HW % 4 == 0.  synth_fwd_kernel keeps 8 register-staged 128-bit loads in flight per thread and stops at
// ~0.77-0.82 of the HBM roofline; here, as in synth_train_kernel, every thread sends all K 16-byte feat values of its
// 4 pixels to a thread-private shared-memory column with cp.async (K loads in flight per thread, no staging registers,
// no barrier), in two halves so that the second half is still in flight while the first is being accumulated.  Same
// accumulation order and the same sigmoid as synth_fwd_kernel: identical masks.
// ---------------------------------------------------------------------------------------------
constexpr int SYN_FS_THREADS = 128;
constexpr int SYN_FS_GROUPS = 4;                       // pixel groups of 128 x 4 px walked by one CTA
constexpr int SYN_FS_SMEM = SYN_KCHUNK * SYN_FS_THREADS * 16;  // 64 KB staging

// PIT = true: instead of the list-form loss, the per-sample NS x NS matrix  lm[i][j] = sum w * l(pred_j, target_i)  of the SoP++ audio-only
// step (PitWrapper._opt_perm_loss, criterion.py:138-176) leaves the CTA as NS^2 doubles (reduced by pit_fwd_fin_kernel).
template <int NS, bool FULLK, bool PIT>
__global__ void __launch_bounds__(SYN_FS_THREADS, (PIT && NS > 2) ? 2 : 3)
synth_fwd_staged_kernel(const float* __restrict__ feat, const float* __restrict__ vecs, const float* __restrict__ scale,
                        const float* __restrict__ bias, int B, int K, int HW, int act, float* __restrict__ out, int loss_kind,
                        const float* __restrict__ gt, const float* __restrict__ weight, double* __restrict__ partial) {
  extern __shared__ float4 stage[];                    // [32][SYN_FS_THREADS], column = thread (private)
  __shared__ float vs[NS][SYN_KCHUNK];
  __shared__ double red[32];
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < NS * SYN_KCHUNK; i += blockDim.x) {
    const int n = i / SYN_KCHUNK, k = i % SYN_KCHUNK;
    vs[n][k] = (k < K) ? vecs[((size_t)b * NS + n) * K + k] * (scale ? scale[k] : 1.f) : 0.f;
  }
  __syncthreads();
  const float bz = bias ? bias[0] : 0.f;
  const float* fbase = feat + (size_t)b * K * HW;
  constexpr int HALF = SYN_KCHUNK / 2;
  float lsum = 0.f;
  float lm[PIT ? NS : 1][PIT ? NS : 1];
#pragma unroll
  for (int i = 0; i < (PIT ? NS : 1); ++i)
#pragma unroll
    for (int j = 0; j < (PIT ? NS : 1); ++j) lm[i][j] = 0.f;
  auto issue = [&](int px, int k0) {                   // feat[k0 .. k0 + 16) of the thread's 4 pixels -> its column
#pragma unroll
    for (int k = k0; k < k0 + HALF; ++k)
      if (FULLK || k < K) cp_async_16(&stage[k * SYN_FS_THREADS + tid], fbase + (size_t)k * HW + px);
    cp_async_commit();
  };
  auto pixel = [&](int grp) { return ((blockIdx.x * SYN_FS_GROUPS + grp) * SYN_FS_THREADS + tid) * 4; };
  int px = pixel(0);
  if (px < HW) { issue(px, 0); issue(px, HALF); }
#pragma unroll 1
  for (int grp = 0; grp < SYN_FS_GROUPS; ++grp, px = pixel(grp)) {
    if (px >= HW) break;                               // pixel groups grow with grp: the later ones are outside too
    const int px_next = pixel(grp + 1);
    const bool more = grp + 1 < SYN_FS_GROUPS && px_next < HW;
    float w[4] = {1.f, 1.f, 1.f, 1.f};
    float4 tgt[NS];
    if (loss_kind != AVSEP_LOSS_NONE) {                // small streams overlap the feat loads
      if (weight) {
        const float4 t4 = *reinterpret_cast<const float4*>(weight + (size_t)b * HW + px);
        w[0] = t4.x; w[1] = t4.y; w[2] = t4.z; w[3] = t4.w;
      }
#pragma unroll
      for (int n = 0; n < NS; ++n) tgt[n] = *reinterpret_cast<const float4*>(gt + ((size_t)n * B + b) * HW + px);
    }
    float z[NS][4];
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) z[n][e] = 0.f;
    auto accumulate = [&](int k0) {
#pragma unroll
      for (int k = k0; k < k0 + HALF; ++k) {
        if (FULLK || k < K) {
          const float4 f = stage[k * SYN_FS_THREADS + tid];
#pragma unroll
          for (int n = 0; n < NS; ++n) {
            const float c = vs[n][k];
            z[n][0] = fmaf(c, f.x, z[n][0]);
            z[n][1] = fmaf(c, f.y, z[n][1]);
            z[n][2] = fmaf(c, f.z, z[n][2]);
            z[n][3] = fmaf(c, f.w, z[n][3]);
          }
        }
      }
    };
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // first half landed (the second may still be in flight)
    accumulate(0);
    if (more) issue(px_next, 0);                       // the column's first half is free again: next group's loads start now
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory"); else cp_async_wait_all();
    accumulate(HALF);
    if (more) issue(px_next, HALF);
    float prs[PIT ? NS : 1][4];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const size_t o = ((size_t)n * B + b) * HW + px;
      float pr[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float zz = z[n][e] + bz;
        pr[e] = act == AVSEP_ACT_SIGMOID ? __frcp_rn(1.0f + expf(-zz)) : zz;     // == 1 / (1 + exp(-z)), correctly rounded
        if (PIT) prs[n][e] = pr[e];
      }
      *reinterpret_cast<float4*>(out + o) = make_float4(pr[0], pr[1], pr[2], pr[3]);
      if (!PIT && loss_kind != AVSEP_LOSS_NONE) {
        const float tt[4] = {tgt[n].x, tgt[n].y, tgt[n].z, tgt[n].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) lsum += w[e] * loss_elem_fast(loss_kind, pr[e], tt[e]);
      }
    }
    if (PIT) {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const float tt[4] = {tgt[i].x, tgt[i].y, tgt[i].z, tgt[i].w};
#pragma unroll
        for (int j = 0; j < NS; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) lm[PIT ? i : 0][PIT ? j : 0] += w[e] * loss_elem_fast(loss_kind, prs[PIT ? j : 0][e], tt[e]);
      }
    }
  }
  if (PIT) {
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const double tot = block_sum<double>((double)lm[PIT ? i : 0][PIT ? j : 0], red);
        if (tid == 0) partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (NS * NS) + i * NS + j] = tot;
      }
  } else if (loss_kind != AVSEP_LOSS_NONE) {
    const double tot = block_sum<double>((double)lsum, red);
    if (tid == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
  }
}

// deterministic final reduction of the per-CTA partial sums: loss = sum / (numel * N)
__global__ void __launch_bounds__(1024) loss_finalize_kernel(const double* __restrict__ partial, int n, double inv_count,
                                                             float* __restrict__ loss) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  const double tot = block_sum<double>(s, red);
  if (threadIdx.x == 0) loss[0] = (float)(tot * inv_count);
}

// PIT forward (SoP++/main.py:227-236 + PitWrapper._opt_perm_loss, criterion.py:138-176): the N masks of every mixture AND the
// per-sample N x N loss matrix  lossmat[b][i][j] = mean_HW w * l(pred_j, target_i)  in ONE pass over feat_sound (the reference
// runs N synthesizer calls, stacks, and materialises [H, W, N, N] repeats per sample).  Same streaming structure as
// synth_fwd_kernel; the N^2 partial sums leave the CTA as doubles and are reduced in a fixed order by pit_fwd_fin_kernel.
template <int NS>
__global__ void __launch_bounds__(SYN_FWD_THREADS)
synth_fwd_pit_kernel(const float* __restrict__ feat, const float* __restrict__ vecs, const float* __restrict__ scale,
                     const float* __restrict__ bias, int B, int K, int HW, int act, float* __restrict__ out, int loss_kind,
                     const float* __restrict__ gt, const float* __restrict__ weight, double* __restrict__ partial) {
  __shared__ float vs[NS][SYN_MAX_K];
  __shared__ double red[32];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < NS * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    vs[n][k] = vecs[((size_t)b * NS + n) * K + k] * (scale ? scale[k] : 1.f);
  }
  __syncthreads();
  const float bz = bias ? bias[0] : 0.f;
  const int px = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  float lm[NS][NS];
#pragma unroll
  for (int i = 0; i < NS; ++i)
#pragma unroll
    for (int j = 0; j < NS; ++j) lm[i][j] = 0.f;
  if (px < HW) {
    float acc[NS][4];
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    const float* fp = feat + (size_t)b * K * HW + px;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float4 t4 = ld_stream(reinterpret_cast<const float4*>(fp + (size_t)k * HW));
      const float f[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float c = vs[n][k];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = fmaf(c, f[e], acc[n][e]);
      }
    }
    float w[4] = {1.f, 1.f, 1.f, 1.f};
    if (weight) {
      const float4 t4 = *reinterpret_cast<const float4*>(weight + (size_t)b * HW + px);
      w[0] = t4.x; w[1] = t4.y; w[2] = t4.z; w[3] = t4.w;
    }
    float pr[NS][4], tg[NS][4];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const size_t o = ((size_t)n * B + b) * HW + px;
#pragma unroll
      for (int e = 0; e < 4; ++e) pr[n][e] = act_apply(act, acc[n][e] + bz);
      *reinterpret_cast<float4*>(out + o) = make_float4(pr[n][0], pr[n][1], pr[n][2], pr[n][3]);
      const float4 t4 = *reinterpret_cast<const float4*>(gt + o);
      tg[n][0] = t4.x; tg[n][1] = t4.y; tg[n][2] = t4.z; tg[n][3] = t4.w;
    }
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) lm[i][j] += w[e] * loss_elem_fast(loss_kind, pr[j][e], tg[i][e]);
  }
#pragma unroll
  for (int i = 0; i < NS; ++i)
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const double tot = block_sum<double>((double)lm[i][j], red);
      if (threadIdx.x == 0) partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (NS * NS) + i * NS + j] = tot;
    }
}

// lossmat[b][q] = (sum over the CTAs of sample b) / HW, fixed order
__global__ void pit_fwd_fin_kernel(const double* __restrict__ partial, int blocks, int SS, double inv_hw, float* __restrict__ lossmat) {
  const int b = blockIdx.x, q = threadIdx.x;
  if (q >= SS) return;
  double s = 0.0;
  for (int c = 0; c < blocks; ++c) s += partial[((size_t)b * blocks + c) * SS + q];
  lossmat[(size_t)b * SS + q] = (float)(s * inv_hw);
}

// y[j][b][:] = x[sel[b][j]][b][:] (per-sample choice of source plane: targets re-ordered by a PIT permutation, masks by its
// inverse), optionally times rowscale[b] * factor (per-sample upstream gradients folded into the loss weights)
__global__ void __launch_bounds__(256) plane_gather_kernel(const float* __restrict__ x, const int* __restrict__ sel, const float* __restrict__ rowscale,
                                                            float factor, int N, int B, size_t P, float* __restrict__ y) {
  const int b = blockIdx.y, j = blockIdx.z;
  const int src = sel ? sel[b * N + j] : j;
  const float sc = rowscale ? rowscale[b] * factor : 1.f;
  const float* xp = x + ((size_t)src * B + b) * P;
  float* yp = y + ((size_t)j * B + b) * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) yp[i] = xp[i] * sc;
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// grid = (pixel chunks, B, K chunks of 32).  Each thread owns SYN_BWD_ITERS float4 pixel groups:
//   dz[n]      = gscale * w * dl/dp * act'(p) / (numel * N)        (fused)   or given
//   dfeat[k]   = sum_n dz[n] * vec[n][k] * scale[k]                streamed out
//   S[n][k]   += dz[n] . feat[k]                                   registers -> warp shuffle -> CTA partial
template <int NS, bool GIVEN>
__global__ void __launch_bounds__(SYN_BWD_THREADS, (NS <= 2 ? 4 : (NS == 3 ? 3 : 2)))
synth_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ vecs, const float* __restrict__ scale, int B,
                 int K, int HW, int act, int loss_kind, const float* __restrict__ pred, const float* __restrict__ gt,
                 const float* __restrict__ weight, const float* __restrict__ gscale, const float* __restrict__ dz_in,
                 float inv_count, float* __restrict__ dfeat, float* __restrict__ s_partial /* [B][chunks][NS][Kp] */,
                 float* __restrict__ dz_partial /* [B][chunks][NS] */, int Kp) {
  __shared__ float vs[NS][SYN_KCHUNK];
  __shared__ float red[SYN_BWD_THREADS / 32][NS * SYN_KCHUNK + NS];
  const int b = blockIdx.y;
  const int k0 = blockIdx.z * SYN_KCHUNK;
  const int kn = min(SYN_KCHUNK, K - k0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < NS * SYN_KCHUNK; i += blockDim.x) {
    const int n = i / SYN_KCHUNK, k = i % SYN_KCHUNK;
    vs[n][k] = (k < kn) ? vecs[((size_t)b * NS + n) * K + k0 + k] * (scale ? scale[k0 + k] : 1.f) : 0.f;
  }
  __syncthreads();
  const float g = GIVEN ? 1.f : gscale[0] * inv_count;
  const int base_px = blockIdx.x * (SYN_BWD_THREADS * 4 * SYN_BWD_ITERS) + threadIdx.x * 4;

  // 1. dz for every pixel group this thread owns (small streams: pred, gt, weight or the given dz)
  float dz[SYN_BWD_ITERS][NS][4];
  float dzs[NS];
#pragma unroll
  for (int n = 0; n < NS; ++n) dzs[n] = 0.f;
#pragma unroll
  for (int it = 0; it < SYN_BWD_ITERS; ++it) {
    const int px = base_px + it * SYN_BWD_THREADS * 4;
    const bool ok = px < HW;
    float w[4] = {1.f, 1.f, 1.f, 1.f};
    if (!GIVEN && weight && ok) {
      const float4 t4 = *reinterpret_cast<const float4*>(weight + (size_t)b * HW + px);
      w[0] = t4.x; w[1] = t4.y; w[2] = t4.z; w[3] = t4.w;
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) dz[it][n][e] = 0.f;
      if (ok) {
        const size_t o = ((size_t)n * B + b) * HW + px;
        if (GIVEN) {
          const float4 d4 = *reinterpret_cast<const float4*>(dz_in + o);
          dz[it][n][0] = d4.x; dz[it][n][1] = d4.y; dz[it][n][2] = d4.z; dz[it][n][3] = d4.w;
        } else {
          const float4 p4 = *reinterpret_cast<const float4*>(pred + o);
          const float4 t4 = *reinterpret_cast<const float4*>(gt + o);
          const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
          const float tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float d = g * w[e] * loss_grad(loss_kind, pp[e], tt[e]);
            if (act == AVSEP_ACT_SIGMOID) d *= (1.f - pp[e]) * pp[e];
            dz[it][n][e] = d;
          }
        }
      }
      dzs[n] += (dz[it][n][0] + dz[it][n][1]) + (dz[it][n][2] + dz[it][n][3]);
    }
  }

  // 2. stream feat / dfeat, SYN_BWD_LOADS channels at a time: all the 128-bit loads of a batch are issued
  //    before any is consumed (memory-level parallelism is what this kernel lives on)
  const float* fbase = feat + ((size_t)b * K + k0) * HW;
  float* dbase = dfeat ? dfeat + ((size_t)b * K + k0) * HW : nullptr;
#pragma unroll 1
  for (int kc = 0; kc < SYN_KCHUNK; kc += SYN_BWD_LOADS) {
    float S[NS][SYN_BWD_LOADS];
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
      for (int j = 0; j < SYN_BWD_LOADS; ++j) S[n][j] = 0.f;
    if (kc < kn) {
#pragma unroll
      for (int it = 0; it < SYN_BWD_ITERS; ++it) {
        const int px = base_px + it * SYN_BWD_THREADS * 4;
        if (px < HW) {
          float4 f[SYN_BWD_LOADS];
#pragma unroll
          for (int j = 0; j < SYN_BWD_LOADS; ++j)
            f[j] = (kc + j < kn) ? ld_stream(reinterpret_cast<const float4*>(fbase + (size_t)(kc + j) * HW + px))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < SYN_BWD_LOADS; ++j) {
            float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int n = 0; n < NS; ++n) {
              const float c = vs[n][kc + j];
              o4.x = fmaf(c, dz[it][n][0], o4.x);
              o4.y = fmaf(c, dz[it][n][1], o4.y);
              o4.z = fmaf(c, dz[it][n][2], o4.z);
              o4.w = fmaf(c, dz[it][n][3], o4.w);
              S[n][j] += (dz[it][n][0] * f[j].x + dz[it][n][1] * f[j].y) + (dz[it][n][2] * f[j].z + dz[it][n][3] * f[j].w);
            }
            if (dbase && kc + j < kn) st_stream(reinterpret_cast<float4*>(dbase + (size_t)(kc + j) * HW + px), o4);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
      for (int j = 0; j < SYN_BWD_LOADS; ++j) {
        const float r = warp_sum(S[n][j]);
        if (lane == 0) red[warp][n * SYN_KCHUNK + kc + j] = r;
      }
  }
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const float r = warp_sum(dzs[n]);
    if (lane == 0) red[warp][NS * SYN_KCHUNK + n] = r;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS * SYN_KCHUNK + NS; i += blockDim.x) {
    float r = 0.f;
#pragma unroll
    for (int wq = 0; wq < SYN_BWD_THREADS / 32; ++wq) r += red[wq][i];
    const size_t cta = (size_t)b * gridDim.x + blockIdx.x;
    if (i < NS * SYN_KCHUNK) {
      const int n = i / SYN_KCHUNK, k = i % SYN_KCHUNK;
      if (k < kn) s_partial[(cta * NS + n) * Kp + k0 + k] = r;
    } else if (blockIdx.z == 0) {
      dz_partial[cta * NS + (i - NS * SYN_KCHUNK)] = r;
    }
  }
}

// T[b][n][k] = sum_chunks S ; dvecs = scale * T ; tvec = T * vec (input of the dscale reduction)
__global__ void synth_bwd_reduce_kernel(const float* __restrict__ s_partial, const float* __restrict__ vecs,
                                        const float* __restrict__ scale, int B, int N, int K, int Kp, int chunks,
                                        float* __restrict__ dvecs, float* __restrict__ tvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N * K) return;
  const int k = i % K, n = (i / K) % N, b = i / (K * N);
  float t = 0.f;
  for (int c = 0; c < chunks; ++c) t += s_partial[(((size_t)b * chunks + c) * N + n) * Kp + k];
  if (dvecs) dvecs[i] = t * (scale ? scale[k] : 1.f);
  tvec[i] = t * vecs[i];
}

// block k < K: dscale[k] = sum_{b,n} tvec[b][n][k];  block K: dbias = sum dz_partial
__global__ void __launch_bounds__(256) synth_bwd_final_kernel(const float* __restrict__ tvec, const float* __restrict__ dz_partial,
                                                               int BN, int K, int n_dz, float* __restrict__ dscale,
                                                               float* __restrict__ dbias) {
  __shared__ double red[32];
  double s = 0.0;
  if ((int)blockIdx.x < K) {
    for (int i = threadIdx.x; i < BN; i += blockDim.x) s += (double)tvec[(size_t)i * K + blockIdx.x];
    const double tot = block_sum<double>(s, red);
    if (threadIdx.x == 0 && dscale) dscale[blockIdx.x] = (float)tot;
  } else {
    for (int i = threadIdx.x; i < n_dz; i += blockDim.x) s += (double)dz_partial[i];
    const double tot = block_sum<double>(s, red);
    if (threadIdx.x == 0 && dbias) dbias[0] = (float)tot;
  }
}

// ---------------------------------------------------------------------------------------------
// fused training step: forward + loss + backward with ONE read of feat_sound from HBM
// ---------------------------------------------------------------------------------------------
// The backward needs dz = dL/dz, which depends on z = sum over ALL K channels, so feat is needed twice.  Each
// thread parks its 32 x float4 feat values in a thread-private shared-memory column during pass 1 (z), then
// pass 2 reads them back for S[n][k] = sum dz * feat and streams dfeat = sum_n dz[n] * vec[n][k] * scale[k]
// out.  DRAM traffic: feat in + dfeat out + the small mask tensors -- the separate fwd + bwd kernels read feat
// twice.  The gradient is computed for upstream gradient gscale (1 when the loss is the root of backward).
constexpr int SYN_TR_THREADS = 128;
constexpr int SYN_TR_GROUPS = 4;                       // pixel groups of 128 x 4 px walked by one CTA
constexpr int SYN_TR_SMEM = 32 * SYN_TR_THREADS * 16;  // 64 KB staging

// BWD = true: the stand-alone backward (plain autograd: InnerProd.forward + a criterion, PitWrapper) on the same staged columns:
// no pass 1, dz from the SAVED masks with torch's two-step gradient (or given as dz_in), then pass 2 as in training.
template <int NS, bool FULLK, bool BWD>
__global__ void __launch_bounds__(SYN_TR_THREADS, (NS <= 2 ? 3 : 2))
synth_train_kernel(const float* __restrict__ feat, const float* __restrict__ vecs, const float* __restrict__ scale,
                   const float* __restrict__ bias, int B, int K, int HW, int act, int loss_kind, const float* __restrict__ gt,
                   const float* __restrict__ weight, const float* __restrict__ gscale, float inv_count, float* __restrict__ pred,
                   double* __restrict__ loss_partial, float* __restrict__ dfeat, float* __restrict__ s_partial,
                   float* __restrict__ dz_partial, int Kp, const float* __restrict__ dz_in) {
  extern __shared__ float4 stage[];                    // [32][SYN_TR_THREADS], column = thread (private)
  __shared__ float vs[NS][SYN_KCHUNK];
  __shared__ float red[SYN_TR_THREADS / 32][NS * SYN_KCHUNK + NS];
  __shared__ double lred[32];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NS * SYN_KCHUNK; i += blockDim.x) {
    const int n = i / SYN_KCHUNK, k = i % SYN_KCHUNK;
    vs[n][k] = (k < K) ? vecs[((size_t)b * NS + n) * K + k] * (scale ? scale[k] : 1.f) : 0.f;
  }
  __syncthreads();
  const float bz = bias ? bias[0] : 0.f;
  const float g = (BWD && dz_in) ? 1.f : (gscale ? gscale[0] : 1.f) * inv_count;
  const float* fbase = feat + (size_t)b * K * HW;
  float* dbase = dfeat + (size_t)b * K * HW;
  float S[NS][SYN_KCHUNK];
  float dzs[NS];
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    dzs[n] = 0.f;
#pragma unroll
    for (int k = 0; k < SYN_KCHUNK; ++k) S[n][k] = 0.f;
  }
  float lsum = 0.f;
#pragma unroll 1
  for (int grp = 0; grp < SYN_TR_GROUPS; ++grp) {
    const int px = ((blockIdx.x * SYN_TR_GROUPS + grp) * SYN_TR_THREADS + tid) * 4;
    if (px >= HW) continue;
    // ---- pass 0: all K 16-byte feat loads of this thread go straight to its shared-memory column with cp.async
    //      (no staging registers, 32 loads in flight per thread) ----------------------------------------------------
#pragma unroll
    for (int k = 0; k < SYN_KCHUNK; ++k)
      if (FULLK || k < K) cp_async_16(&stage[k * SYN_TR_THREADS + tid], fbase + (size_t)k * HW + px);
    cp_async_commit();
    // small streams overlap the feat loads
    float w[4] = {1.f, 1.f, 1.f, 1.f};
    if (weight && !(BWD && dz_in)) {
      const float4 t4 = *reinterpret_cast<const float4*>(weight + (size_t)b * HW + px);
      w[0] = t4.x; w[1] = t4.y; w[2] = t4.z; w[3] = t4.w;
    }
    float dz[NS][4];
    if (BWD) {
      // ---- dz from the saved masks: g * w * dl/dp * act'(p), torch's two-step form (synth_bwd_kernel computes the same), or given
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const size_t o = ((size_t)n * B + b) * HW + px;
        if (dz_in) {
          const float4 d4 = *reinterpret_cast<const float4*>(dz_in + o);
          dz[n][0] = d4.x; dz[n][1] = d4.y; dz[n][2] = d4.z; dz[n][3] = d4.w;
        } else {
          const float4 p4 = *reinterpret_cast<const float4*>(pred + o);
          const float4 t4 = *reinterpret_cast<const float4*>(gt + o);
          const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
          const float tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float d = g * w[e] * loss_grad(loss_kind, pp[e], tt[e]);
            if (act == AVSEP_ACT_SIGMOID) d *= (1.f - pp[e]) * pp[e];
            dz[n][e] = d;
          }
        }
        dzs[n] += (dz[n][0] + dz[n][1]) + (dz[n][2] + dz[n][3]);
      }
      cp_async_wait_all();
    } else {
    float4 tgt[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) tgt[n] = *reinterpret_cast<const float4*>(gt + ((size_t)n * B + b) * HW + px);
    cp_async_wait_all();
    // ---- pass 1: z[n] = sum_k vs[n][k] * feat[k] ---------------------------------------------------------------------
    float z[NS][4];
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) z[n][e] = 0.f;
#pragma unroll
    for (int k = 0; k < SYN_KCHUNK; ++k) {
      if (FULLK || k < K) {
        const float4 f = stage[k * SYN_TR_THREADS + tid];
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const float c = vs[n][k];
          z[n][0] = fmaf(c, f.x, z[n][0]);
          z[n][1] = fmaf(c, f.y, z[n][1]);
          z[n][2] = fmaf(c, f.z, z[n][2]);
          z[n][3] = fmaf(c, f.w, z[n][3]);
        }
      }
    }
    // ---- activation, loss, dz --------------------------------------------------------------------------------
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const size_t o = ((size_t)n * B + b) * HW + px;
      const float tt[4] = {tgt[n].x, tgt[n].y, tgt[n].z, tgt[n].w};
      float pr[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float zz = z[n][e] + bz;
        pr[e] = act == AVSEP_ACT_SIGMOID ? __frcp_rn(1.0f + expf(-zz)) : zz;     // == 1 / (1 + exp(-z)), correctly rounded
        lsum += w[e] * loss_elem_fast(loss_kind, pr[e], tt[e]);
        dz[n][e] = g * w[e] * dz_elem_fast(loss_kind, act, pr[e], tt[e]);
      }
      *reinterpret_cast<float4*>(pred + o) = make_float4(pr[0], pr[1], pr[2], pr[3]);
      dzs[n] += (dz[n][0] + dz[n][1]) + (dz[n][2] + dz[n][3]);
    }
    }
    // ---- pass 2: dfeat out, S accumulation from the parked feat ------------------------------------------------
#pragma unroll
    for (int k = 0; k < SYN_KCHUNK; ++k) {
      if (FULLK || k < K) {
        const float4 f = stage[k * SYN_TR_THREADS + tid];
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const float c = vs[n][k];
          o4.x = fmaf(c, dz[n][0], o4.x);
          o4.y = fmaf(c, dz[n][1], o4.y);
          o4.z = fmaf(c, dz[n][2], o4.z);
          o4.w = fmaf(c, dz[n][3], o4.w);
          float acc = S[n][k];
          acc = fmaf(dz[n][0], f.x, acc);
          acc = fmaf(dz[n][1], f.y, acc);
          acc = fmaf(dz[n][2], f.z, acc);
          acc = fmaf(dz[n][3], f.w, acc);
          S[n][k] = acc;
        }
        st_stream(reinterpret_cast<float4*>(dbase + (size_t)k * HW + px), o4);
      }
    }
  }
  // ---- CTA reductions --------------------------------------------------------------------------------------------
#pragma unroll
  for (int n = 0; n < NS; ++n) {
#pragma unroll
    for (int k = 0; k < SYN_KCHUNK; ++k) {
      const float r = warp_sum(S[n][k]);
      if (lane == 0) red[warp][n * SYN_KCHUNK + k] = r;
    }
    const float r = warp_sum(dzs[n]);
    if (lane == 0) red[warp][NS * SYN_KCHUNK + n] = r;
  }
  const double ltot = block_sum<double>((double)lsum, lred);      // contains the __syncthreads that publishes red[][]
  const size_t cta = (size_t)b * gridDim.x + blockIdx.x;
  if (tid == 0 && !BWD) loss_partial[cta] = ltot;
  for (int i = tid; i < NS * SYN_KCHUNK + NS; i += blockDim.x) {
    float r = 0.f;
#pragma unroll
    for (int wq = 0; wq < SYN_TR_THREADS / 32; ++wq) r += red[wq][i];
    if (i < NS * SYN_KCHUNK) {
      const int n = i / SYN_KCHUNK, k = i % SYN_KCHUNK;
      if (k < K) s_partial[(cta * NS + n) * Kp + k] = r;
    } else {
      dz_partial[cta * NS + (i - NS * SYN_KCHUNK)] = r;
    }
  }
}

// One finalize launch for everything small: block k < K: dvecs[:, :, k] and dscale[k]; block K: dbias; block K+1: loss.
__global__ void __launch_bounds__(256)
synth_finalize_kernel(const float* __restrict__ s_partial, const float* __restrict__ dz_partial, const double* __restrict__ loss_partial,
                      const float* __restrict__ vecs, const float* __restrict__ scale, int B, int N, int K, int Kp, int chunks,
                      int n_loss, double inv_count, float* __restrict__ dvecs, float* __restrict__ dscale, float* __restrict__ dbias,
                      float* __restrict__ loss) {
  __shared__ double red[32];
  const int blk = blockIdx.x;
  double acc = 0.0;
  if (blk < K) {
    for (int i = threadIdx.x; i < B * N; i += blockDim.x) {       // i = b * N + n
      const int bb = i / N, n = i % N;
      // the per-CTA partials of one (b, n, k) are N * Kp floats apart: four independent accumulators keep four loads
      // in flight (a single running sum made this kernel a chain of `chunks` L2 round trips: 12 us for 34 CTAs)
      const float* sp = s_partial + ((size_t)bb * chunks * N + n) * Kp + blk;
      const size_t cs = (size_t)N * Kp;
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
      int c = 0;
      for (; c + 4 <= chunks; c += 4) {
        t0 += sp[(size_t)c * cs];
        t1 += sp[(size_t)(c + 1) * cs];
        t2 += sp[(size_t)(c + 2) * cs];
        t3 += sp[(size_t)(c + 3) * cs];
      }
      for (; c < chunks; ++c) t0 += sp[(size_t)c * cs];
      const float t = (t0 + t1) + (t2 + t3);
      if (dvecs) dvecs[(size_t)i * K + blk] = t * (scale ? scale[blk] : 1.f);
      acc += (double)(t * vecs[(size_t)i * K + blk]);
    }
    const double tot = block_sum<double>(acc, red);
    if (threadIdx.x == 0 && dscale) dscale[blk] = (float)tot;
  } else if (blk == K) {
#pragma unroll 4
    for (int i = threadIdx.x; i < B * chunks * N; i += blockDim.x) acc += (double)dz_partial[i];
    const double tot = block_sum<double>(acc, red);
    if (threadIdx.x == 0 && dbias) dbias[0] = (float)tot;
  } else {
#pragma unroll 4
    for (int i = threadIdx.x; i < n_loss; i += blockDim.x) acc += loss_partial[i];
    const double tot = block_sum<double>(acc, red);
    if (threadIdx.x == 0 && loss) loss[0] = (float)(tot * inv_count);
  }
}

// ---------------------------------------------------------------------------------------------
// forward_nosum / forward_pixelwise (inference helpers, synthesizer_net.py:21-38)
// ---------------------------------------------------------------------------------------------
__global__ void synth_nosum_kernel(const float* __restrict__ feat, const float* __restrict__ vec, const float* __restrict__ scale,
                                   const float* __restrict__ bias, int K, size_t HW, float* __restrict__ out) {
  const int bk = blockIdx.y;   // b * K + k
  const int k = bk % K;
  const float c = vec[bk] * (scale ? scale[k] : 1.f);
  const float bz = bias ? bias[0] : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x)
    out[(size_t)bk * HW + i] = fmaf(c, feat[(size_t)bk * HW + i], bz);
}

__global__ void __launch_bounds__(256) synth_pixelwise_kernel(const float* __restrict__ feat, const float* __restrict__ fimg,
                                                               const float* __restrict__ scale, const float* __restrict__ bias,
                                                               int K, int P, int HW, float* __restrict__ out) {
  // out[b][p][hw] = sum_k fimg[b][k][p] * scale[k] * feat[b][k][hw] + bias ; one CTA = (hw block, p, b)
  extern __shared__ float vk[];
  const int b = blockIdx.z, pidx = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) vk[k] = fimg[((size_t)b * K + k) * P + pidx] * (scale ? scale[k] : 1.f);
  __syncthreads();
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= HW) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(vk[k], __ldg(feat + ((size_t)b * K + k) * HW + hw), acc);
  out[((size_t)b * P + pidx) * HW + hw] = acc + (bias ? bias[0] : 0.f);
}

struct SynthScratch {
  size_t loss_partial, s_partial, dz_partial, tvec, total;
  int fwd_blocks_x, bwd_chunks, Kp;
};

SynthScratch synth_scratch_layout(int B, int N, int K, int HW) {
  SynthScratch s;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  s.fwd_blocks_x = ceil_div(ceil_div(HW, 4), SYN_FWD_THREADS);
  s.bwd_chunks = ceil_div(HW, SYN_BWD_THREADS * 4 * SYN_BWD_ITERS);
  s.Kp = ceil_div(K, SYN_KCHUNK) * SYN_KCHUNK;
  size_t off = 0;
  s.loss_partial = off; off = up(off + sizeof(double) * (size_t)B * ceil_div(HW, SYN_FWD_THREADS) * (size_t)(N > 1 ? N * N : 1));   // VEC=1 worst case; N^2 sums per CTA for the PIT forward
  s.s_partial = off;    off = up(off + sizeof(float) * (size_t)B * s.bwd_chunks * N * s.Kp);
  s.dz_partial = off;   off = up(off + sizeof(float) * (size_t)B * s.bwd_chunks * N);
  s.tvec = off;         off = up(off + sizeof(float) * (size_t)B * N * K);
  s.total = off;
  return s;
}

template <int NS>
static int synth_forward_n(const float* feat, const float* vecs, const float* scale, const float* bias, int B, int K, int HW,
                           int act, float* out, int loss_kind, const float* gt, const float* weight, float* loss,
                           void* scratch, cudaStream_t stream) {
  const bool vec4 = (HW % 4 == 0);
  const SynthScratch lay = synth_scratch_layout(B, NS, K, HW);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(scratch) + lay.loss_partial);
  const bool staged = vec4 && K <= SYN_KCHUNK && (reinterpret_cast<uintptr_t>(feat) & 15u) == 0;
  const int bx = staged ? ceil_div(HW, SYN_FS_THREADS * 4 * SYN_FS_GROUPS) : (vec4 ? lay.fwd_blocks_x : ceil_div(HW, SYN_FWD_THREADS));
  dim3 grid(bx, B);
  if (staged) {
    auto kern = K == SYN_KCHUNK ? synth_fwd_staged_kernel<NS, true, false> : synth_fwd_staged_kernel<NS, false, false>;
    AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SYN_FS_SMEM));
    kern<<<grid, SYN_FS_THREADS, SYN_FS_SMEM, stream>>>(feat, vecs, scale, bias, B, K, HW, act, out, loss_kind, gt, weight, partial);
  } else if (vec4)
    synth_fwd_kernel<NS, 4><<<grid, SYN_FWD_THREADS, 0, stream>>>(feat, vecs, scale, bias, B, K, HW, act, out, loss_kind, gt,
                                                                  weight, partial);
  else
    synth_fwd_kernel<NS, 1><<<grid, SYN_FWD_THREADS, 0, stream>>>(feat, vecs, scale, bias, B, K, HW, act, out, loss_kind, gt,
                                                                  weight, partial);
  AVSEP_LAUNCH_CHECK();
  if (loss_kind != AVSEP_LOSS_NONE) {
    const double inv = 1.0 / ((double)B * (double)HW * (double)NS);
    loss_finalize_kernel<<<1, 1024, 0, stream>>>(partial, bx * B, inv, loss);
    AVSEP_LAUNCH_CHECK();
  }
  return AVSEP_OK;
}

template <int NS>
static int synth_backward_n(const float* feat, const float* vecs, const float* scale, int B, int K, int HW, int act,
                            int loss_kind, const float* pred, const float* gt, const float* weight, const float* gscale,
                            const float* dz, float* dfeat, float* dvecs, float* dscale, float* dbias, void* scratch,
                            cudaStream_t stream) {
  const SynthScratch lay = synth_scratch_layout(B, NS, K, HW);
  char* sc = static_cast<char*>(scratch);
  float* s_partial = reinterpret_cast<float*>(sc + lay.s_partial);
  float* dz_partial = reinterpret_cast<float*>(sc + lay.dz_partial);
  float* tvec = reinterpret_cast<float*>(sc + lay.tvec);
  dim3 grid(lay.bwd_chunks, B, lay.Kp / SYN_KCHUNK);
  const float inv_count = (float)(1.0 / ((double)B * (double)HW * (double)NS));
  // the staged kernel (thread-private cp.async columns of synth_train_kernel, pass 2 only) when it applies: K <= 32, dfeat
  // wanted, 16-byte rows; its partial sums have the layout the reduce / final kernels below read (same 2048 px per CTA)
  const bool staged = dfeat != nullptr && K <= SYN_KCHUNK && (HW % 4) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(dfeat) & 15u) == 0;
  if (staged) {
    auto kern = (K == SYN_KCHUNK) ? synth_train_kernel<NS, true, true> : synth_train_kernel<NS, false, true>;
    AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SYN_TR_SMEM));
    dim3 sgrid(lay.bwd_chunks, B);
    kern<<<sgrid, SYN_TR_THREADS, SYN_TR_SMEM, stream>>>(feat, vecs, scale, nullptr, B, K, HW, act, loss_kind, gt, weight, gscale, inv_count,
                                                         const_cast<float*>(pred), nullptr, dfeat, s_partial, dz_partial, lay.Kp, dz);
  } else if (dz)
    synth_bwd_kernel<NS, true><<<grid, SYN_BWD_THREADS, 0, stream>>>(feat, vecs, scale, B, K, HW, act, loss_kind, pred, gt,
                                                                     weight, gscale, dz, inv_count, dfeat, s_partial,
                                                                     dz_partial, lay.Kp);
  else
    synth_bwd_kernel<NS, false><<<grid, SYN_BWD_THREADS, 0, stream>>>(feat, vecs, scale, B, K, HW, act, loss_kind, pred, gt,
                                                                      weight, gscale, dz, inv_count, dfeat, s_partial,
                                                                      dz_partial, lay.Kp);
  AVSEP_LAUNCH_CHECK();
  if (dvecs || dscale) {
    const int n = B * NS * K;
    synth_bwd_reduce_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(s_partial, vecs, scale, B, NS, K, lay.Kp, lay.bwd_chunks,
                                                                  dvecs, tvec);
    AVSEP_LAUNCH_CHECK();
  }
  if (dscale || dbias) {
    synth_bwd_final_kernel<<<K + 1, 256, 0, stream>>>(tvec, dz_partial, B * NS, K, B * lay.bwd_chunks * NS, dscale, dbias);
    AVSEP_LAUNCH_CHECK();
  }
  return AVSEP_OK;
}

template <int NS>
static int synth_train_n(const float* feat, const float* vecs, const float* scale, const float* bias, int B, int K, int HW,
                         int act, int loss_kind, const float* gt, const float* weight, const float* gscale, float* pred,
                         float* loss, float* dfeat, float* dvecs, float* dscale, float* dbias, void* scratch, cudaStream_t stream) {
  const SynthScratch lay = synth_scratch_layout(B, NS, K, HW);
  char* sc = static_cast<char*>(scratch);
  double* loss_partial = reinterpret_cast<double*>(sc + lay.loss_partial);
  float* s_partial = reinterpret_cast<float*>(sc + lay.s_partial);
  float* dz_partial = reinterpret_cast<float*>(sc + lay.dz_partial);
  const int chunks = ceil_div(HW, SYN_TR_THREADS * 4 * SYN_TR_GROUPS);      // == lay.bwd_chunks (same 2048 px per CTA)
  static_assert(SYN_TR_THREADS * SYN_TR_GROUPS == SYN_BWD_THREADS * SYN_BWD_ITERS, "scratch layout shared with the backward kernel");
  auto kern = (K == SYN_KCHUNK) ? synth_train_kernel<NS, true, false> : synth_train_kernel<NS, false, false>;
  AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SYN_TR_SMEM));
  dim3 grid(chunks, B);
  const float inv_count = (float)(1.0 / ((double)B * (double)HW * (double)NS));
  kern<<<grid, SYN_TR_THREADS, SYN_TR_SMEM, stream>>>(feat, vecs, scale, bias, B, K, HW, act, loss_kind, gt, weight, gscale, inv_count,
                                                       pred, loss_partial, dfeat, s_partial, dz_partial, lay.Kp, nullptr);
  AVSEP_LAUNCH_CHECK();
  synth_finalize_kernel<<<K + 2, 256, 0, stream>>>(s_partial, dz_partial, loss_partial, vecs, scale, B, NS, K, lay.Kp, chunks,
                                                    chunks * B, 1.0 / ((double)B * (double)HW * (double)NS), dvecs, dscale, dbias, loss);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

// fused forward + loss + backward (K <= 32, HW % 4 == 0, dfeat required); callers fall back to fwd + bwd otherwise
bool synth_train_supported(int K, int HW, const float* dfeat) { return K <= SYN_KCHUNK && (HW % 4) == 0 && dfeat != nullptr; }

int synth_train_impl(const float* feat, const float* vecs, const float* scale, const float* bias, int B, int N, int K, int HW,
                     int act, int loss_kind, const float* gt, const float* weight, const float* gscale, float* pred, float* loss,
                     float* dfeat, float* dvecs, float* dscale, float* dbias, void* scratch, cudaStream_t stream) {
  switch (N) {
    case 1: return synth_train_n<1>(feat, vecs, scale, bias, B, K, HW, act, loss_kind, gt, weight, gscale, pred, loss, dfeat, dvecs, dscale, dbias, scratch, stream);
    case 2: return synth_train_n<2>(feat, vecs, scale, bias, B, K, HW, act, loss_kind, gt, weight, gscale, pred, loss, dfeat, dvecs, dscale, dbias, scratch, stream);
    case 3: return synth_train_n<3>(feat, vecs, scale, bias, B, K, HW, act, loss_kind, gt, weight, gscale, pred, loss, dfeat, dvecs, dscale, dbias, scratch, stream);
    default: return synth_train_n<4>(feat, vecs, scale, bias, B, K, HW, act, loss_kind, gt, weight, gscale, pred, loss, dfeat, dvecs, dscale, dbias, scratch, stream);
  }
}

int synth_forward_impl(const float* feat, const float* vecs, const float* scale, const float* bias, int B, int N, int K, int HW,
                       int act, float* out, int loss_kind, const float* gt, const float* weight, float* loss, void* scratch,
                       cudaStream_t stream) {
  switch (N) {
    case 1: return synth_forward_n<1>(feat, vecs, scale, bias, B, K, HW, act, out, loss_kind, gt, weight, loss, scratch, stream);
    case 2: return synth_forward_n<2>(feat, vecs, scale, bias, B, K, HW, act, out, loss_kind, gt, weight, loss, scratch, stream);
    case 3: return synth_forward_n<3>(feat, vecs, scale, bias, B, K, HW, act, out, loss_kind, gt, weight, loss, scratch, stream);
    default: return synth_forward_n<4>(feat, vecs, scale, bias, B, K, HW, act, out, loss_kind, gt, weight, loss, scratch, stream);
  }
}

int synth_backward_impl(const float* feat, const float* vecs, const float* scale, int B, int N, int K, int HW, int act,
                        int loss_kind, const float* pred, const float* gt, const float* weight, const float* gscale,
                        const float* dz, float* dfeat, float* dvecs, float* dscale, float* dbias, void* scratch,
                        cudaStream_t stream) {
  switch (N) {
    case 1: return synth_backward_n<1>(feat, vecs, scale, B, K, HW, act, loss_kind, pred, gt, weight, gscale, dz, dfeat, dvecs, dscale, dbias, scratch, stream);
    case 2: return synth_backward_n<2>(feat, vecs, scale, B, K, HW, act, loss_kind, pred, gt, weight, gscale, dz, dfeat, dvecs, dscale, dbias, scratch, stream);
    case 3: return synth_backward_n<3>(feat, vecs, scale, B, K, HW, act, loss_kind, pred, gt, weight, gscale, dz, dfeat, dvecs, dscale, dbias, scratch, stream);
    default: return synth_backward_n<4>(feat, vecs, scale, B, K, HW, act, loss_kind, pred, gt, weight, gscale, dz, dfeat, dvecs, dscale, dbias, scratch, stream);
  }
}

}  // namespace avsep

using namespace avsep;

extern "C" {

size_t avsep_synth_scratch_bytes(int B, int N, int K, int HW) {
  if (B < 1 || N < 1 || K < 1 || HW < 1) return 256;
  return synth_scratch_layout(B, N, K, HW).total;
}

int avsep_synth_forward(const float* feat_dev, const float* vecs_dev, const float* scale_dev, const float* bias_dev, int B,
                        int N, int K, int HW, int activation, float* out_dev, int loss_kind, const float* gt_dev,
                        const float* weight_dev, float* loss_dev, void* scratch_dev, void* stream) {
  AVSEP_REQUIRE(feat_dev && vecs_dev && out_dev, "avsep_synth_forward: feat/vecs/out must be non-null");
  AVSEP_REQUIRE(N >= 1 && N <= AVSEP_MAX_SRC, "avsep_synth_forward: N = %d outside [1, %d]", N, AVSEP_MAX_SRC);
  AVSEP_REQUIRE(K >= 1 && K <= SYN_MAX_K, "avsep_synth_forward: K = %d outside [1, %d]", K, SYN_MAX_K);
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && HW >= 1, "avsep_synth_forward: bad B = %d / HW = %d", B, HW);
  AVSEP_REQUIRE(activation == AVSEP_ACT_NONE || activation == AVSEP_ACT_SIGMOID, "avsep_synth_forward: unknown activation %d", activation);
  AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_NONE && loss_kind <= AVSEP_LOSS_BCE, "avsep_synth_forward: unknown loss kind %d", loss_kind);
  if (loss_kind != AVSEP_LOSS_NONE)
    AVSEP_REQUIRE(gt_dev && loss_dev && scratch_dev, "avsep_synth_forward: the fused loss needs gt_dev, loss_dev and scratch_dev");
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  return synth_forward_impl(feat_dev, vecs_dev, scale_dev, bias_dev, B, N, K, HW, activation, out_dev, loss_kind, gt_dev,
                            weight_dev, loss_dev, scratch_dev, (cudaStream_t)stream);
}

int avsep_synth_backward(const float* feat_dev, const float* vecs_dev, const float* scale_dev, int B, int N, int K, int HW,
                         int activation, int loss_kind, const float* pred_dev, const float* gt_dev, const float* weight_dev,
                         const float* gscale_dev, const float* dz_dev, float* dfeat_dev, float* dvecs_dev, float* dscale_dev,
                         float* dbias_dev, void* scratch_dev, void* stream) {
  AVSEP_REQUIRE(feat_dev && vecs_dev && scratch_dev, "avsep_synth_backward: feat/vecs/scratch must be non-null");
  AVSEP_REQUIRE(N >= 1 && N <= AVSEP_MAX_SRC, "avsep_synth_backward: N = %d outside [1, %d]", N, AVSEP_MAX_SRC);
  AVSEP_REQUIRE(K >= 1 && K <= SYN_MAX_K, "avsep_synth_backward: K = %d outside [1, %d]", K, SYN_MAX_K);
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && HW >= 4 && HW % 4 == 0, "avsep_synth_backward: bad B = %d / HW = %d (HW must be a multiple of 4)", B, HW);
  if (!dz_dev) {
    AVSEP_REQUIRE(pred_dev && gt_dev && gscale_dev, "avsep_synth_backward: fused mode needs pred_dev, gt_dev and gscale_dev");
    AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_L1 && loss_kind <= AVSEP_LOSS_BCE, "avsep_synth_backward: unknown loss kind %d", loss_kind);
  }
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  return synth_backward_impl(feat_dev, vecs_dev, scale_dev, B, N, K, HW, activation, loss_kind, pred_dev, gt_dev, weight_dev,
                             gscale_dev, dz_dev, dfeat_dev, dvecs_dev, dscale_dev, dbias_dev, scratch_dev, (cudaStream_t)stream);
}

int avsep_synth_train(const float* feat_dev, const float* vecs_dev, const float* scale_dev, const float* bias_dev, int B, int N,
                      int K, int HW, int activation, int loss_kind, const float* gt_dev, const float* weight_dev,
                      const float* gscale_dev, float* pred_dev, float* loss_dev, float* dfeat_dev, float* dvecs_dev,
                      float* dscale_dev, float* dbias_dev, void* scratch_dev, void* stream) {
  AVSEP_REQUIRE(feat_dev && vecs_dev && gt_dev && pred_dev && loss_dev && dfeat_dev && scratch_dev, "avsep_synth_train: null argument");
  AVSEP_REQUIRE(N >= 1 && N <= AVSEP_MAX_SRC, "avsep_synth_train: N = %d outside [1, %d]", N, AVSEP_MAX_SRC);
  AVSEP_REQUIRE(K >= 1 && K <= SYN_MAX_K, "avsep_synth_train: K = %d outside [1, %d]", K, SYN_MAX_K);
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && HW >= 4 && HW % 4 == 0, "avsep_synth_train: bad B = %d / HW = %d (HW must be a multiple of 4)", B, HW);
  AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_L1 && loss_kind <= AVSEP_LOSS_BCE, "avsep_synth_train: unknown loss kind %d", loss_kind);
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (synth_train_supported(K, HW, dfeat_dev))
    return synth_train_impl(feat_dev, vecs_dev, scale_dev, bias_dev, B, N, K, HW, activation, loss_kind, gt_dev, weight_dev, gscale_dev,
                            pred_dev, loss_dev, dfeat_dev, dvecs_dev, dscale_dev, dbias_dev, scratch_dev, st);
  // K > 32: two passes over feat (forward, then backward from the saved activations)
  rc = synth_forward_impl(feat_dev, vecs_dev, scale_dev, bias_dev, B, N, K, HW, activation, pred_dev, loss_kind, gt_dev, weight_dev,
                          loss_dev, scratch_dev, st);
  if (rc) return rc;
  AVSEP_REQUIRE(gscale_dev != nullptr, "avsep_synth_train: gscale_dev is required when K > 32");
  return synth_backward_impl(feat_dev, vecs_dev, scale_dev, B, N, K, HW, activation, loss_kind, pred_dev, gt_dev, weight_dev, gscale_dev,
                             nullptr, dfeat_dev, dvecs_dev, dscale_dev, dbias_dev, scratch_dev, st);
}

int avsep_synth_nosum(const float* feat_dev, const float* vec_dev, const float* scale_dev, const float* bias_dev, int B, int K,
                      int HW, float* out_dev, void* stream) {
  AVSEP_REQUIRE(feat_dev && vec_dev && out_dev, "avsep_synth_nosum: null tensor");
  AVSEP_REQUIRE(B >= 0 && K >= 1 && HW >= 1 && (long long)B * K <= 65535, "avsep_synth_nosum: bad shape");
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  dim3 grid(std::min(ceil_div(HW, 256), 1024), B * K);
  synth_nosum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat_dev, vec_dev, scale_dev, bias_dev, K, (size_t)HW, out_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_synth_pixelwise(const float* feat_dev, const float* feats_img_dev, const float* scale_dev, const float* bias_dev, int B,
                          int K, int P, int HW, float* out_dev, void* stream) {
  AVSEP_REQUIRE(feat_dev && feats_img_dev && out_dev, "avsep_synth_pixelwise: null tensor");
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && K >= 1 && K <= 4096 && P >= 1 && P <= 65535 && HW >= 1, "avsep_synth_pixelwise: bad shape");
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  dim3 grid(ceil_div(HW, 256), P, B);
  synth_pixelwise_kernel<<<grid, 256, K * sizeof(float), (cudaStream_t)stream>>>(feat_dev, feats_img_dev, scale_dev, bias_dev, K,
                                                                                 P, HW, out_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_synth_forward_pit(const float* feat_dev, const float* vecs_dev, const float* scale_dev, const float* bias_dev, int B, int N,
                            int K, int HW, int activation, float* out_dev, int loss_kind, const float* gt_dev,
                            const float* weight_dev, float* lossmat_dev, void* scratch_dev, void* stream_) {
  AVSEP_REQUIRE(feat_dev && vecs_dev && out_dev && gt_dev && lossmat_dev && scratch_dev, "avsep_synth_forward_pit: null tensor");
  AVSEP_REQUIRE(N >= 1 && N <= AVSEP_MAX_SRC && K >= 1 && K <= SYN_MAX_K && B >= 0 && B <= 65535, "avsep_synth_forward_pit: bad B / N / K");
  AVSEP_REQUIRE(HW >= 4 && (HW % 4) == 0, "avsep_synth_forward_pit: H * W = %d must be a multiple of 4", HW);
  AVSEP_REQUIRE(loss_kind >= AVSEP_LOSS_L1 && loss_kind <= AVSEP_LOSS_BCE, "avsep_synth_forward_pit: unknown loss kind %d", loss_kind);
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const SynthScratch lay = synth_scratch_layout(B, N, K, HW);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(scratch_dev) + lay.loss_partial);
  const bool staged = K <= SYN_KCHUNK && (reinterpret_cast<uintptr_t>(feat_dev) & 15u) == 0;
  const int bx = staged ? ceil_div(HW, SYN_FS_THREADS * 4 * SYN_FS_GROUPS) : lay.fwd_blocks_x;
  dim3 grid(bx, B);
  if (staged) {
    // staged columns (cp.async, half-K pipelined) like the plain forward: 0.93-0.96 of the HBM roofline instead of 0.8
    auto launch = [&](auto kern) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SYN_FS_SMEM);
      if (e != cudaSuccess) return e;
      kern<<<grid, SYN_FS_THREADS, SYN_FS_SMEM, stream>>>(feat_dev, vecs_dev, scale_dev, bias_dev, B, K, HW, activation, out_dev, loss_kind,
                                                          gt_dev, weight_dev, partial);
      return cudaSuccess;
    };
    const bool fullk = K == SYN_KCHUNK;
    cudaError_t e;
    switch (N) {
      case 1: e = fullk ? launch(synth_fwd_staged_kernel<1, true, true>) : launch(synth_fwd_staged_kernel<1, false, true>); break;
      case 2: e = fullk ? launch(synth_fwd_staged_kernel<2, true, true>) : launch(synth_fwd_staged_kernel<2, false, true>); break;
      case 3: e = fullk ? launch(synth_fwd_staged_kernel<3, true, true>) : launch(synth_fwd_staged_kernel<3, false, true>); break;
      default: e = fullk ? launch(synth_fwd_staged_kernel<4, true, true>) : launch(synth_fwd_staged_kernel<4, false, true>); break;
    }
    AVSEP_CUDA(e);
  } else {
    switch (N) {
      case 1: synth_fwd_pit_kernel<1><<<grid, SYN_FWD_THREADS, 0, stream>>>(feat_dev, vecs_dev, scale_dev, bias_dev, B, K, HW, activation, out_dev, loss_kind, gt_dev, weight_dev, partial); break;
      case 2: synth_fwd_pit_kernel<2><<<grid, SYN_FWD_THREADS, 0, stream>>>(feat_dev, vecs_dev, scale_dev, bias_dev, B, K, HW, activation, out_dev, loss_kind, gt_dev, weight_dev, partial); break;
      case 3: synth_fwd_pit_kernel<3><<<grid, SYN_FWD_THREADS, 0, stream>>>(feat_dev, vecs_dev, scale_dev, bias_dev, B, K, HW, activation, out_dev, loss_kind, gt_dev, weight_dev, partial); break;
      default: synth_fwd_pit_kernel<4><<<grid, SYN_FWD_THREADS, 0, stream>>>(feat_dev, vecs_dev, scale_dev, bias_dev, B, K, HW, activation, out_dev, loss_kind, gt_dev, weight_dev, partial); break;
    }
  }
  AVSEP_LAUNCH_CHECK();
  pit_fwd_fin_kernel<<<B, 32, 0, stream>>>(partial, bx, N * N, 1.0 / (double)HW, lossmat_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_plane_gather(const float* x_dev, const int* sel_dev, const float* rowscale_dev, float factor, int N, int B, size_t P,
                       float* y_dev, void* stream) {
  AVSEP_REQUIRE(x_dev && y_dev && N >= 1 && N <= 65535 && B >= 0 && B <= 65535 && P >= 1, "avsep_plane_gather: bad argument");
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  dim3 grid((unsigned)std::min<size_t>((P + 255) / 256, 256), B, N);
  plane_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, sel_dev, rowscale_dev, factor, N, B, P, y_dev);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

}  // extern "C"
