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

// Vectorised variant (T % 4 == 0): a thread owns 4 consecutive output columns, so the row samples are one
// aligned 128-bit load plus the two neighbours (x taps are w-1 / w by construction of the linear grid),
// every output stream is written with 128-bit stores and the index arithmetic is amortised over 4 pixels.
constexpr int PREP_ROWS_PER_THREAD = 4;
constexpr int PREP_TX = 64, PREP_TY = 4;

struct Row4 {
  float v[4];
};

__device__ __forceinline__ Row4 row_sample4(const float* __restrict__ base, int y, int H, int T, int w0, const int (&sel)[4],
                                            const float (&fx)[4], float eps) {
  Row4 r;
  if (y < 0 || y >= H) {
#pragma unroll
    for (int j = 0; j < 4; ++j) r.v[j] = 0.f;
    return r;
  }
  const float* p = base + (size_t)y * T + w0;
  const float4 c = __ldg(reinterpret_cast<const float4*>(p));
  float e[6];
  e[0] = w0 > 0 ? __ldg(p - 1) + eps : 0.f;         // zero padding outside the image: no epsilon there
  e[1] = c.x + eps; e[2] = c.y + eps; e[3] = c.z + eps; e[4] = c.w + eps;
  e[5] = w0 + 4 < T ? __ldg(p + 4) + eps : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = sel[j] ? e[j + 1] : e[j];        // tap x0     (x0 = w - 1 + sel)
    const float b = sel[j] ? e[j + 2] : e[j + 1];    // tap x0 + 1
    r.v[j] = (1.f - fx[j]) * a + fx[j] * b;
  }
  return r;
}

template <int NS>
__global__ void __launch_bounds__(PREP_TX * PREP_TY, NS >= 3 ? 2 : 4) prepare_vec_kernel(const PrepareParams p) {
  const int b = blockIdx.z;
  const int h0 = (blockIdx.y * PREP_TY + threadIdx.y) * PREP_ROWS_PER_THREAD;
  if (h0 >= p.Hout) return;
  const size_t in_off = (size_t)b * p.Hin * p.T;
  const size_t out_off = (size_t)b * p.Hout * p.T;
  for (int w0 = (blockIdx.x * PREP_TX + threadIdx.x) * 4; w0 < p.T; w0 += gridDim.x * PREP_TX * 4) {
    const int4 xi = __ldg(reinterpret_cast<const int4*>(p.xtap.i0 + w0));
    const float4 xf = __ldg(reinterpret_cast<const float4*>(p.xtap.frac + w0));
    const int sel[4] = {xi.x - (w0 - 1), xi.y - w0, xi.z - (w0 + 1), xi.w - (w0 + 2)};
    const float fx[4] = {xf.x, xf.y, xf.z, xf.w};
    int cy0[NS + 1], cy1[NS + 1];
    Row4 cv0[NS + 1], cv1[NS + 1];
#pragma unroll
    for (int j = 0; j <= NS; ++j) { cy0[j] = -2; cy1[j] = -2; }
#pragma unroll
    for (int r = 0; r < PREP_ROWS_PER_THREAD; ++r) {
      const int h = h0 + r;
      if (h >= p.Hout) break;
      const int y0 = __ldg(p.ytap.i0 + h);
      const float fy = __ldg(p.ytap.frac + h);
      Row4 v[NS + 1];
#pragma unroll
      for (int j = 0; j <= NS; ++j) {
        const float* base = (j == 0 ? p.mix : p.src[j - 1]) + in_off;
        const float eps = j == 0 ? 1e-10f : 0.f;
        const Row4 a = (y0 == cy0[j]) ? cv0[j] : ((y0 == cy1[j]) ? cv1[j] : row_sample4(base, y0, p.Hin, p.T, w0, sel, fx, eps));
        const Row4 c = (y0 + 1 == cy1[j]) ? cv1[j] : ((y0 + 1 == cy0[j]) ? cv0[j] : row_sample4(base, y0 + 1, p.Hin, p.T, w0, sel, fx, eps));
        cy0[j] = y0; cv0[j] = a; cy1[j] = y0 + 1; cv1[j] = c;
#pragma unroll
        for (int e = 0; e < 4; ++e) v[j].v[e] = (1.f - fy) * a.v[e] + fy * c.v[e];
      }
      const size_t o = out_off + (size_t)h * p.T + w0;
      const float* mx = v[0].v;
      if (p.mix_w) *reinterpret_cast<float4*>(p.mix_w + o) = make_float4(mx[0], mx[1], mx[2], mx[3]);
      if (p.log_mix) *reinterpret_cast<float4*>(p.log_mix + o) = make_float4(fast_log(mx[0]), fast_log(mx[1]), fast_log(mx[2]), fast_log(mx[3]));
      if (p.weights) {
        float wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) wv[e] = p.weighted ? fminf(fmaxf(fast_log1p_pos(mx[e]), 1e-3f), 10.f) : 1.f;
        *reinterpret_cast<float4*>(p.weights + o) = make_float4(wv[0], wv[1], wv[2], wv[3]);
      }
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float* sv = v[n + 1].v;
        if (p.src_w[n]) *reinterpret_cast<float4*>(p.src_w[n] + o) = make_float4(sv[0], sv[1], sv[2], sv[3]);
        if (p.gt[n]) {
          float gv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            gv[e] = p.binary ? (sv[e] > 0.5f * mx[e] ? 1.f : 0.f) : fminf(fmaxf(sv[e] / mx[e], 0.f), 5.f);
          *reinterpret_cast<float4*>(p.gt[n] + o) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged variant (the product path when T % 4 == 0 and the tile fits shared memory).
// prepare_vec_kernel is latency-bound: every thread chases its own row loads (ncu: 37 % issue-active, 31 % warps active).
// Here a CTA owns PREPB_ROWS output rows x all T frames of one mixture; the (at most 2 per output row) input rows it
// needs -- and only those: on the high-frequency half of the axis most input rows are skipped -- are pulled into shared
// memory by TMA bulk copies (cp.async.bulk, one 4 T-byte row each, issued by the lanes of one warp, completion on one
// mbarrier), for the mixture and all N sources at once.  With ~48 KB in flight per CTA and several CTAs per SM the HBM
// pipe stays full without any thread waiting on its own loads; the bilinear taps then read shared memory.
constexpr int PREPB_ROWS = 8;
constexpr int PREPB_THREADS = 256;
constexpr int PREPB_MAX_SMEM = 96 * 1024;

__device__ __forceinline__ void bulk_load_row(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(PREPB_THREADS) prepare_bulk_kernel(const PrepareParams p) {
  extern __shared__ __align__(128) float rows_sm[];            // [(NS + 1)][2 * PREPB_ROWS][T]
  __shared__ __align__(8) unsigned long long bar_storage;
  __shared__ int row_of_slot[2 * PREPB_ROWS];                  // input row held by each staged slot
  __shared__ int slot_a[PREPB_ROWS], slot_b[PREPB_ROWS];       // slots of taps y0 / y0 + 1 of each output row (-1: outside the image)
  const int b = blockIdx.y;
  const int h0 = blockIdx.x * PREPB_ROWS;
  const int T = p.T;
  const uint32_t bar = smem_u32(&bar_storage);
  const uint32_t row_bytes = (uint32_t)T * 4u;
  if (threadIdx.x < 32) {
    // Warp 0 sets the tile up and issues the copies.  Lane e < 16 owns tap (e & 1) of output row (e >> 1); equal input rows
    // share one staged slot: the lowest lane holding a row is its leader, slots are numbered by leader order.
    const int lane = threadIdx.x;
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    const int r = (lane & 15) >> 1, tap = lane & 1;
    const int h = min(h0 + r, p.Hout - 1);
    const int y = __ldg(p.ytap.i0 + h) + tap;
    const bool valid = lane < 2 * PREPB_ROWS && y >= 0 && y < p.Hin;
    const unsigned same = __match_any_sync(0xffffffffu, valid ? y : -1000 - lane);
    const int leader = __ffs(same) - 1;
    const unsigned leaders = __ballot_sync(0xffffffffu, valid && leader == lane);
    const int slot = valid ? __popc(leaders & ((1u << leader) - 1u)) : -1;
    if (lane < 2 * PREPB_ROWS) {
      if (tap == 0) slot_a[r] = slot; else slot_b[r] = slot;
      if (valid && leader == lane) row_of_slot[slot] = y;
    }
    const int ns = __popc(leaders);
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(ns * (NS + 1)) * row_bytes);
    __syncwarp();
    for (int i = lane; i < ns * (NS + 1); i += 32) {
      const int j = i / ns, k = i % ns;
      const float* base = (j == 0 ? p.mix : p.src[j - 1]) + ((size_t)b * p.Hin + row_of_slot[k]) * T;
      bulk_load_row(smem_u32(rows_sm + ((size_t)j * 2 * PREPB_ROWS + k) * T), base, row_bytes, bar);
    }
  }
  __syncthreads();                                                       // slot tables and the barrier are visible to every warp
  // thread = 4 consecutive columns x 2 output rows; the x taps do not depend on the row
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const size_t out_off = (size_t)b * p.Hout * T;
  // (columns beyond T: threads idle; T <= 256 per pass, wider clips loop)
  mbar_wait(bar, 0, nullptr, 61);
  for (int w0 = 4 * tx; w0 < T; w0 += 4 * 64) {
    const int4 xi = __ldg(reinterpret_cast<const int4*>(p.xtap.i0 + w0));
    const float4 xf = __ldg(reinterpret_cast<const float4*>(p.xtap.frac + w0));
    const int sel[4] = {xi.x - (w0 - 1), xi.y - w0, xi.z - (w0 + 1), xi.w - (w0 + 2)};
    const float fx[4] = {xf.x, xf.y, xf.z, xf.w};
    auto sample = [&](int j, int slot, float eps, float (&out)[4]) {     // x-interpolated staged row at the 4 columns
      if (slot < 0) { out[0] = out[1] = out[2] = out[3] = 0.f; return; }
      const float* r = rows_sm + ((size_t)j * 2 * PREPB_ROWS + slot) * T + w0;
      const float4 c = *reinterpret_cast<const float4*>(r);
      float e[6];
      e[0] = w0 > 0 ? r[-1] + eps : 0.f;                                 // zero padding outside the image: no epsilon there
      e[1] = c.x + eps; e[2] = c.y + eps; e[3] = c.z + eps; e[4] = c.w + eps;
      e[5] = w0 + 4 < T ? r[4] + eps : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = sel[q] ? e[q + 1] : e[q];
        const float c2 = sel[q] ? e[q + 2] : e[q + 1];
        out[q] = (1.f - fx[q]) * a + fx[q] * c2;
      }
    };
#pragma unroll
    for (int rr = 0; rr < PREPB_ROWS / 4; ++rr) {
      const int r = ty + 4 * rr;
      const int h = h0 + r;
      if (h >= p.Hout) continue;
      const float fy = __ldg(p.ytap.frac + h);
      const int sa = slot_a[r], sb = slot_b[r];
      float a[4], c[4], mx[4];
      sample(0, sa, 1e-10f, a);                                          // mag_mix + 1e-10 before the warp (main.py:54)
      sample(0, sb, 1e-10f, c);
#pragma unroll
      for (int q = 0; q < 4; ++q) mx[q] = (1.f - fy) * a[q] + fy * c[q];
      const size_t o = out_off + (size_t)h * T + w0;
      if (p.mix_w) *reinterpret_cast<float4*>(p.mix_w + o) = make_float4(mx[0], mx[1], mx[2], mx[3]);
      if (p.log_mix) *reinterpret_cast<float4*>(p.log_mix + o) = make_float4(fast_log(mx[0]), fast_log(mx[1]), fast_log(mx[2]), fast_log(mx[3]));
      if (p.weights) {
        float wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = p.weighted ? fminf(fmaxf(fast_log1p_pos(mx[q]), 1e-3f), 10.f) : 1.f;
        *reinterpret_cast<float4*>(p.weights + o) = make_float4(wv[0], wv[1], wv[2], wv[3]);
      }
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        float sv[4];
        sample(n + 1, sa, 0.f, a);
        sample(n + 1, sb, 0.f, c);
#pragma unroll
        for (int q = 0; q < 4; ++q) sv[q] = (1.f - fy) * a[q] + fy * c[q];
        if (p.src_w[n]) *reinterpret_cast<float4*>(p.src_w[n] + o) = make_float4(sv[0], sv[1], sv[2], sv[3]);
        if (p.gt[n]) {
          float gv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            gv[q] = p.binary ? (sv[q] > 0.5f * mx[q] ? 1.f : 0.f) : fminf(fmaxf(sv[q] / mx[q], 0.f), 5.f);
          *reinterpret_cast<float4*>(p.gt[n] + o) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
      }
    }
  }
}

static bool bulk_pointers_ok(const PrepareParams& p, int NS) {
  // cp.async.bulk needs 16-byte aligned global rows
  if ((reinterpret_cast<uintptr_t>(p.mix) & 15u) != 0) return false;
  for (int n = 0; n < NS; ++n)
    if ((reinterpret_cast<uintptr_t>(p.src[n]) & 15u) != 0) return false;
  return true;
}

template <int NS>
static int launch_prepare(const PrepareParams& p, bool warp, cudaStream_t stream) {
  const size_t bulk_smem = (size_t)(NS + 1) * 2 * PREPB_ROWS * p.T * sizeof(float);
  if (warp && (p.T & 3) == 0 && bulk_smem <= (size_t)PREPB_MAX_SMEM && bulk_pointers_ok(p, NS)) {
    auto kern = prepare_bulk_kernel<NS>;
    AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
    dim3 grid(ceil_div(p.Hout, PREPB_ROWS), p.B);
    kern<<<grid, PREPB_THREADS, bulk_smem, stream>>>(p);
    AVSEP_LAUNCH_CHECK();
    return AVSEP_OK;
  }
  if (warp && (p.T & 3) == 0) {
    dim3 block(PREP_TX, PREP_TY);
    dim3 grid(ceil_div(p.T, PREP_TX * 4), ceil_div(p.Hout, PREP_TY * PREP_ROWS_PER_THREAD), p.B);
    prepare_vec_kernel<NS><<<grid, block, 0, stream>>>(p);
    AVSEP_LAUNCH_CHECK();
    return AVSEP_OK;
  }
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
