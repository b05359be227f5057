// STFT / iSTFT plan: split-precision DFT bases, TMA descriptors, bf16 staging kernels and the
// C-ABI entry points avsep_stft / avsep_separate / avsep_istft.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dft_gemm.cuh"
#include "dft_fold.cuh"
#include "idft_fold.cuh"
#include "warp_taps.h"

namespace avsep {

// ---------------------------------------------------------------------------------------------
// staging kernels
// ---------------------------------------------------------------------------------------------
// Staging of the PLAIN forward GEMM (the in-library cross-check of the folded kernels):
// wave [S, L] fp32 -> reflect-padded (n_fft/2 each side, librosa center=True) bf16 hi/lo rows of
// length Lpad.  Frame t of signal s is then the 1024-sample window starting at 256 t, which the
// GEMM reads directly through an overlapping-row TMA descriptor (row stride 256 elements).
// mix_N > 0: `wave` holds only the N sources [N, mix_B, L]; signals s < mix_B are the mixtures, formed here as
// the sum of their sources (dataset/music.py:127), signals s >= mix_B are the sources themselves.
__global__ void __launch_bounds__(256) wave_split_kernel(const float* __restrict__ wave, int S, int L, int Lpad, size_t lo_stride,
                                                         __nv_bfloat16* __restrict__ out, int mix_B, int mix_N) {
  // 8 consecutive padded positions per thread: two 16-byte stores (hi, lo); Lpad is a multiple of 256
  const int s = blockIdx.y;
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i0 >= Lpad) return;
  const bool is_mix = mix_N > 0 && s < mix_B;
  const float* w = wave + (size_t)(mix_N > 0 ? (is_mix ? s : s - mix_B) : s) * L;
  const size_t src_stride = (size_t)mix_B * L;
  __align__(16) __nv_bfloat16 hi[8];
  __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int i = i0 + e;
    int j = i - AVSEP_N_FFT / 2;
    float x = 0.f;
    if (i < L + AVSEP_N_FFT) {                      // positions beyond the padded signal stay zero
      if (j < 0) j = -j;
      if (j >= L) j = 2 * (L - 1) - j;
      if (j >= 0 && j < L) {
        x = __ldg(w + j);
        if (is_mix)
          for (int n = 1; n < mix_N; ++n) x += __ldg(w + n * src_stride + j);
      }
    }
    split_bf16(x, hi[e], lo[e]);
  }
  const size_t o = (size_t)s * Lpad + i0;
  *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(out + lo_stride + o) = *reinterpret_cast<const uint4*>(lo);
}

// Mixture waveforms of a device-summed batch (dataset/music.py:127: the mixture IS the sum of its sources):
// sources [N, B, L] -> mix [B, L].  Feeds the folded forward kernel when the mixture spectrum cannot be taken as X0 + X1
// (N != 2, or phase / (re, im) outputs).
__global__ void __launch_bounds__(256) mix_sum_kernel(const float* __restrict__ src, int N, size_t BL, float* __restrict__ mix) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < BL; i += (size_t)gridDim.x * blockDim.x) {
    float x = __ldg(src + i);
    for (int n = 1; n < N; ++n) x += __ldg(src + (size_t)n * BL + i);
    mix[i] = x;
  }
}

// K4 front half (main.py:213-246): un-warp each predicted mask to the linear frequency axis, threshold, multiply
// the mixture spectrum and emit the inverse-DFT operand X[t, k] = (re | im), time-major, as bf16 hi/lo.
// HBM-bound streaming kernels, no shared-memory transpose.
// spec_prep_vec_kernel (T even): a thread owns a register tile of 2 CONSECUTIVE frames x 8 CONSECUTIVE bins.
//   * reads  [f][t]: one 64-bit load per bin row; a warp (16 frame pairs x 2 bin groups) reads 128-byte runs
//   * writes [t][f]: the 8 bins of one frame are one 128-bit store per (re | im, hi | lo) plane; the 2 bin groups of a
//     warp and the 4 warps of a CTA fill 128 contiguous bytes of every operand row
//   * the mixture (re, im) tile is loaded once and reused by all N + 1 signals
//   * consecutive bins share un-warp rows: the x-interpolated values (4 frames) of the two most recent mask rows stay
//     in registers (about 1.3 mask row fetches per bin instead of 4 taps per element);
//     small tile = few registers = many resident warps: the kernel lives on memory-level parallelism
// spec_prep_scalar_kernel: one frame x 16 bins per thread, any T (ragged clips).
struct SpecPrepParams {
  const float* masks;   // [N, B, Hm, T]
  const float* mag;     // [B, 512, T]
  const float* phase;   // [B, 512, T] or null
  const float* re;      // [B, 512, T] or null
  const float* im;
  int N, B, T, Hm;
  int include_mix;
  int binary;
  float thres;
  Taps ytap, xtap;      // un-warp taps (rows: 512 outputs over Hm inputs; cols: T over T)
  int log_freq;
  __nv_bfloat16* out;   // [2][S_cap][T][1024]
  size_t lo_stride;
};

constexpr int SPEC_BINS_PER_THREAD = 16;
constexpr int SPEC_FRAMES_PER_CTA = 16;
constexpr int SPEC_BINS_PER_CTA = 256;            // 8 warps x 2 bin groups x 16 bins

// 16 fp32 values -> bf16 hi / lo (x ~= hi + lo), two 128-bit stores each
__device__ __forceinline__ void spec_store16(const float (&v)[SPEC_BINS_PER_THREAD], __nv_bfloat16* __restrict__ hi_dst,
                                             __nv_bfloat16* __restrict__ lo_dst) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  uint4* hd = reinterpret_cast<uint4*>(hi_dst);
  uint4* ld = reinterpret_cast<uint4*>(lo_dst);
  hd[0] = make_uint4(h[0], h[1], h[2], h[3]);
  hd[1] = make_uint4(h[4], h[5], h[6], h[7]);
  ld[0] = make_uint4(l[0], l[1], l[2], l[3]);
  ld[1] = make_uint4(l[4], l[5], l[6], l[7]);
}

__global__ void __launch_bounds__(256) spec_prep_scalar_kernel(const SpecPrepParams p) {
  __shared__ int s_y0[SPEC_BINS_PER_CTA];
  __shared__ float s_fy[SPEC_BINS_PER_CTA];
  const int b = blockIdx.z;
  const int f_cta = blockIdx.y * SPEC_BINS_PER_CTA;
  const int t = blockIdx.x * SPEC_FRAMES_PER_CTA + (threadIdx.x & 15);
  const int grp = threadIdx.x >> 4;               // 0..15: bin group inside the CTA
  const int fb = f_cta + grp * SPEC_BINS_PER_THREAD;
  if (p.log_freq && p.N > 0) {
    s_y0[threadIdx.x] = __ldg(p.ytap.i0 + f_cta + threadIdx.x);
    s_fy[threadIdx.x] = __ldg(p.ytap.frac + f_cta + threadIdx.x);
    __syncthreads();
  }
  if (t >= p.T) return;
  // mixture spectrum of the thread's bins (shared by all signals)
  float bre[SPEC_BINS_PER_THREAD], bim[SPEC_BINS_PER_THREAD];
  {
    const size_t o = ((size_t)b * AVSEP_N_BINS + fb) * p.T + t;
    if (p.re) {
#pragma unroll
      for (int i = 0; i < SPEC_BINS_PER_THREAD; ++i) {
        bre[i] = __ldg(p.re + o + (size_t)i * p.T);
        bim[i] = __ldg(p.im + o + (size_t)i * p.T);
      }
    } else {
#pragma unroll
      for (int i = 0; i < SPEC_BINS_PER_THREAD; ++i) {
        const float mg = __ldg(p.mag + o + (size_t)i * p.T);
        float sn, cs;
        sincosf(__ldg(p.phase + o + (size_t)i * p.T), &sn, &cs);
        bre[i] = mg * cs;
        bim[i] = mg * sn;
      }
    }
  }
  int x0 = 0;
  float fx = 0.f;
  if (p.log_freq) {
    x0 = __ldg(p.xtap.i0 + t);
    fx = __ldg(p.xtap.frac + t);
  }
  const bool xa = x0 >= 0, xb = x0 + 1 < p.T;
  const float wa = xa ? 1.f - fx : 0.f, wb = xb ? fx : 0.f;
  const int xa_i = xa ? x0 : 0, xb_i = xb ? x0 + 1 : 0;
  const int nsig = p.N + p.include_mix;
  for (int jsig = 0; jsig < nsig; ++jsig) {
    const int n = jsig - p.include_mix;            // -1 = mixture (mask 1)
    float vre[SPEC_BINS_PER_THREAD], vim[SPEC_BINS_PER_THREAD];
    if (n < 0) {
#pragma unroll
      for (int i = 0; i < SPEC_BINS_PER_THREAD; ++i) { vre[i] = bre[i]; vim[i] = bim[i]; }
    } else {
      const float* mp = p.masks + ((size_t)n * p.B + b) * (size_t)p.Hm * p.T;
      if (p.log_freq) {
        int cy0 = -2, cy1 = -2;
        float cv0 = 0.f, cv1 = 0.f;
        auto row = [&](int y) -> float {           // x-interpolated mask row y at this frame (0 outside the image)
          if (y == cy0) return cv0;
          if (y == cy1) return cv1;
          if (y < 0 || y >= p.Hm) return 0.f;
          const float* r = mp + (size_t)y * p.T;
          return wa * __ldg(r + xa_i) + wb * __ldg(r + xb_i);
        };
#pragma unroll
        for (int i = 0; i < SPEC_BINS_PER_THREAD; ++i) {
          const int y0 = s_y0[grp * SPEC_BINS_PER_THREAD + i];
          const float fy = s_fy[grp * SPEC_BINS_PER_THREAD + i];
          const float r0 = row(y0), r1 = row(y0 + 1);
          cy0 = y0; cv0 = r0; cy1 = y0 + 1; cv1 = r1;
          float mk = (1.f - fy) * r0 + fy * r1;
          if (p.binary) mk = mk > p.thres ? 1.f : 0.f;
          vre[i] = bre[i] * mk;
          vim[i] = bim[i] * mk;
        }
      } else {
#pragma unroll
        for (int i = 0; i < SPEC_BINS_PER_THREAD; ++i) {
          float mk = __ldg(mp + (size_t)(fb + i) * p.T + t);
          if (p.binary) mk = mk > p.thres ? 1.f : 0.f;
          vre[i] = bre[i] * mk;
          vim[i] = bim[i] * mk;
        }
      }
    }
    const size_t sig = (size_t)b * nsig + jsig;
    __nv_bfloat16* orow = p.out + (sig * p.T + t) * DFT_K + fb;
    spec_store16(vre, orow, orow + p.lo_stride);
    spec_store16(vim, orow + AVSEP_N_BINS, orow + p.lo_stride + AVSEP_N_BINS);
  }
}

constexpr int SPECV_FRAMES = 2, SPECV_BINS = 8;
constexpr int SPECV_THREADS = 128;
constexpr int SPECV_FRAMES_PER_CTA = 32;          // 16 frame pairs per warp
constexpr int SPECV_BINS_PER_CTA = 64;            // 4 warps x 2 bin groups x 8 bins

// 8 fp32 values -> bf16 hi / lo (x ~= hi + lo), one 128-bit store each
__device__ __forceinline__ void spec_store8(const float (&v)[SPECV_BINS], __nv_bfloat16* __restrict__ hi_dst,
                                            __nv_bfloat16* __restrict__ lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(SPECV_THREADS, 6) spec_prep_vec_kernel(const SpecPrepParams p) {
  __shared__ int s_y0[SPECV_BINS_PER_CTA];
  __shared__ float s_fy[SPECV_BINS_PER_CTA];
  const int b = blockIdx.z;
  const int f_cta = blockIdx.y * SPECV_BINS_PER_CTA;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t0 = blockIdx.x * SPECV_FRAMES_PER_CTA + (lane & 15) * SPECV_FRAMES;      // first of this thread's 2 frames
  const int grp = warp * 2 + (lane >> 4);                                            // 0..7: bin group inside the CTA
  const int fb = f_cta + grp * SPECV_BINS;
  if (p.log_freq && p.N > 0) {
    if (threadIdx.x < SPECV_BINS_PER_CTA) {
      s_y0[threadIdx.x] = __ldg(p.ytap.i0 + f_cta + threadIdx.x);
      s_fy[threadIdx.x] = __ldg(p.ytap.frac + f_cta + threadIdx.x);
    }
    __syncthreads();
  }
  if (t0 >= p.T) return;                           // T % 2 == 0: a frame pair is inside the clip or outside it
  // mixture spectrum tile (shared by all signals): [bin][frame]
  float bre[SPECV_BINS][SPECV_FRAMES], bim[SPECV_BINS][SPECV_FRAMES];
  {
    const size_t o = ((size_t)b * AVSEP_N_BINS + fb) * p.T + t0;
#pragma unroll
    for (int i = 0; i < SPECV_BINS; ++i) {
      if (p.re) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(p.re + o + (size_t)i * p.T));
        const float2 m = __ldg(reinterpret_cast<const float2*>(p.im + o + (size_t)i * p.T));
        bre[i][0] = r.x; bre[i][1] = r.y;
        bim[i][0] = m.x; bim[i][1] = m.y;
      } else {
        const float2 mg = __ldg(reinterpret_cast<const float2*>(p.mag + o + (size_t)i * p.T));
        const float2 ph = __ldg(reinterpret_cast<const float2*>(p.phase + o + (size_t)i * p.T));
        float sn, cs;
        sincosf(ph.x, &sn, &cs);
        bre[i][0] = mg.x * cs; bim[i][0] = mg.x * sn;
        sincosf(ph.y, &sn, &cs);
        bre[i][1] = mg.y * cs; bim[i][1] = mg.y * sn;
      }
    }
  }
  // x taps of the 2 frames: tap 0 is column t - 1 or t by construction of the linear grid (sel = 0 / 1)
  int sel[SPECV_FRAMES] = {0, 0};
  float fx[SPECV_FRAMES] = {0.f, 0.f};
  if (p.log_freq) {
    const int2 xi = __ldg(reinterpret_cast<const int2*>(p.xtap.i0 + t0));
    const float2 xf = __ldg(reinterpret_cast<const float2*>(p.xtap.frac + t0));
    sel[0] = xi.x - (t0 - 1); sel[1] = xi.y - t0;
    fx[0] = xf.x; fx[1] = xf.y;
  }
  const int nsig = p.N + p.include_mix;
  for (int jsig = 0; jsig < nsig; ++jsig) {
    const int n = jsig - p.include_mix;            // -1 = mixture (mask 1)
    const size_t sig = (size_t)b * nsig + jsig;
    __nv_bfloat16* orow = p.out + (sig * p.T + t0) * DFT_K + fb;
    float mk[SPECV_BINS][SPECV_FRAMES];
    if (n < 0) {
#pragma unroll
      for (int i = 0; i < SPECV_BINS; ++i) { mk[i][0] = 1.f; mk[i][1] = 1.f; }
    } else {
      const float* mp = p.masks + ((size_t)n * p.B + b) * (size_t)p.Hm * p.T;
      if (p.log_freq) {
        int cy0 = -2, cy1 = -2;
        float cv0[SPECV_FRAMES] = {0.f, 0.f}, cv1[SPECV_FRAMES] = {0.f, 0.f};
        // x-interpolated mask row y at the 2 frames (0 outside the image): one 64-bit load + the two neighbours
        auto row = [&](int y, float (&r)[SPECV_FRAMES]) {
          if (y == cy0) { r[0] = cv0[0]; r[1] = cv0[1]; return; }
          if (y == cy1) { r[0] = cv1[0]; r[1] = cv1[1]; return; }
          if (y < 0 || y >= p.Hm) { r[0] = 0.f; r[1] = 0.f; return; }
          const float* q = mp + (size_t)y * p.T + t0;
          const float2 c = __ldg(reinterpret_cast<const float2*>(q));
          float v[4];
          v[0] = t0 > 0 ? __ldg(q - 1) : 0.f;
          v[1] = c.x; v[2] = c.y;
          v[3] = t0 + 2 < p.T ? __ldg(q + 2) : 0.f;
#pragma unroll
          for (int e = 0; e < SPECV_FRAMES; ++e) {
            const float a = sel[e] ? v[e + 1] : v[e];        // tap x0 (x0 = t - 1 + sel)
            const float c2 = sel[e] ? v[e + 2] : v[e + 1];   // tap x0 + 1
            r[e] = (1.f - fx[e]) * a + fx[e] * c2;           // out-of-image taps contribute 0
          }
        };
#pragma unroll
        for (int i = 0; i < SPECV_BINS; ++i) {
          const int y0 = s_y0[grp * SPECV_BINS + i];
          const float fy = s_fy[grp * SPECV_BINS + i];
          float r0[SPECV_FRAMES], r1[SPECV_FRAMES];
          row(y0, r0);
          row(y0 + 1, r1);
          cy0 = y0; cy1 = y0 + 1;
#pragma unroll
          for (int e = 0; e < SPECV_FRAMES; ++e) {
            cv0[e] = r0[e]; cv1[e] = r1[e];
            const float m = (1.f - fy) * r0[e] + fy * r1[e];
            mk[i][e] = p.binary ? (m > p.thres ? 1.f : 0.f) : m;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < SPECV_BINS; ++i) {
          const float2 c = __ldg(reinterpret_cast<const float2*>(mp + (size_t)(fb + i) * p.T + t0));
          mk[i][0] = p.binary ? (c.x > p.thres ? 1.f : 0.f) : c.x;
          mk[i][1] = p.binary ? (c.y > p.thres ? 1.f : 0.f) : c.y;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < SPECV_FRAMES; ++e) {
      float vre[SPECV_BINS], vim[SPECV_BINS];
#pragma unroll
      for (int i = 0; i < SPECV_BINS; ++i) {
        vre[i] = n < 0 ? bre[i][e] : bre[i][e] * mk[i][e];
        vim[i] = n < 0 ? bim[i][e] : bim[i][e] * mk[i][e];
      }
      spec_store8(vre, orow + (size_t)e * DFT_K, orow + (size_t)e * DFT_K + p.lo_stride);
      spec_store8(vim, orow + (size_t)e * DFT_K + AVSEP_N_BINS, orow + (size_t)e * DFT_K + p.lo_stride + AVSEP_N_BINS);
    }
  }
}

}  // namespace avsep
#include "spec_prep.cuh"
namespace avsep {

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace avsep

struct avsep_plan {
  int device = 0;
  int max_signals = 0;
  int wave_len = 0;
  int T = 0;
  int Lpad = 0;
  int out_len = 0;
  // product kernels: folded forward (dft_fold3 / dft_fold3m, operand generated in the kernel) and folded inverse
  // (idft_fold2); the plain 1024 x 1024 CTA-pair GEMM of dft_gemm.cuh is kept as their in-library cross-check
  bool fold = true;                     // forward transform through the folded kernels
  bool mix_linear = true;               // two-source device-summed mixtures: |X0 + X1| from the two source transforms (dft_fold3m)
  bool inverse_fold = true;             // inverse transform through the folded kernel
  bool use_spec_tma = true;             // inverse operand through the shared-memory staged kernel (spec_prep.cuh) when it applies
  int generation = 0;                   // bumped by every avsep_plan_set_*: captured CUDA graphs of a step are keyed on it
  int sm_budget = 0;                    // persistent GEMM grids use at most this many SMs (0 = the whole device): SM-partitioned pipelines
  __nv_bfloat16* basis_fold = nullptr;  // [4 = kind * 2 + split][512 bins][512]
  __nv_bfloat16* basis_ifold = nullptr; // [4 = kind * 2 + split][512 rows in CTA-group order][512 bins]
  __nv_bfloat16* basis_fwd = nullptr;   // plain: [2][1024][1024]  (hi rows 0..1023, lo rows 1024..2047)
  __nv_bfloat16* basis_inv = nullptr;
  __nv_bfloat16* wave_split = nullptr;  // plain forward operand [2][S_cap][Lpad], allocated on first use
  __nv_bfloat16* spec_split = nullptr;  // inverse operand [2][S_cap][T][1024], allocated on first inverse call
  float* wave_mix = nullptr;            // [mix_cap] mixture waveforms summed on the device (allocated on first use)
  size_t mix_cap = 0;
  float* inv_env = nullptr;             // [256 * (T + 3)]
  const float* last_wave = nullptr;     // arguments of the last forward transform, for avsep_debug_gemm_only
  int last_mix_B = 0, last_mix_N = 0;
  int* err_flag = nullptr;
  CUtensorMap tm_basis_fold32, tm_basis_ifold, tm_spec_ifold;      // folded kernels
  CUtensorMap tm_spec_store;                                       // spec_prep_tma_kernel: 64 bins x 32 frames store boxes
  int unwarp_rmax = -1;                 // most mask rows a 64-bin tile of the un-warp taps (spec_prep_tma_kernel), -1 = not computed
  CUtensorMap tm_basis_fwd, tm_basis_inv, tm_wave2, tm_spec2;      // plain CTA-pair GEMM (128-frame boxes)
  bool tm_wave_valid = false;
  size_t workspace = 0;
};

// SMs a persistent CTA-pair kernel of this plan may occupy (avsep_plan_set_sm_budget)
static int plan_pairs(const avsep_plan* pl) {
  const int sms = avsep::sm_count();
  return (pl->sm_budget > 0 ? std::min(pl->sm_budget, sms) : sms) / 2;
}


namespace avsep {

static int encode_map(CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, int kblock) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(AVSEP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kblock == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kblock == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(AVSEP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return AVSEP_OK;
}

constexpr int PLAIN_BK = 64;            // K block of the plain GEMM (SWIZZLE_128B)

static int build_maps(avsep_plan* pl) {
  {
    cuuint64_t dims[2] = {(cuuint64_t)DFT_K, (cuuint64_t)(2 * DFT_ROWS)};
    cuuint64_t strides[1] = {(cuuint64_t)DFT_K * 2};
    cuuint32_t box[2] = {(cuuint32_t)PLAIN_BK, (cuuint32_t)DFT_TILE_M};
    int rc = encode_map(&pl->tm_basis_fwd, pl->basis_fwd, 2, dims, strides, box, PLAIN_BK);
    if (rc) return rc;
    rc = encode_map(&pl->tm_basis_inv, pl->basis_inv, 2, dims, strides, box, PLAIN_BK);
    if (rc) return rc;
  }
  if (pl->wave_split) {
    // frame t of a signal = 1024 staged samples at offset 256 t: overlapping rows (row stride 512 B), no frame materialisation
    cuuint64_t dims[3] = {(cuuint64_t)DFT_K, (cuuint64_t)pl->T, (cuuint64_t)(2 * pl->max_signals)};
    cuuint64_t strides[2] = {(cuuint64_t)AVSEP_HOP * 2, (cuuint64_t)pl->Lpad * 2};
    cuuint32_t box2[3] = {(cuuint32_t)PLAIN_BK, (cuuint32_t)(DFT_TILE_N / 2), 1};
    int rc = encode_map(&pl->tm_wave2, pl->wave_split, 3, dims, strides, box2, PLAIN_BK);
    if (rc) return rc;
    pl->tm_wave_valid = true;
  }
  {
    cuuint64_t dimsA[2] = {(cuuint64_t)FOLD_K, (cuuint64_t)(4 * FOLD_BINS)};
    cuuint64_t stridesA[1] = {(cuuint64_t)FOLD_K * 2};
    cuuint32_t boxA32[2] = {(cuuint32_t)Fold3Cfg::BK, (cuuint32_t)FOLD_TILE_M};
    int rc = encode_map(&pl->tm_basis_fold32, pl->basis_fold, 2, dimsA, stridesA, boxA32, Fold3Cfg::BK);
    if (rc) return rc;
  }
  {
    cuuint64_t dimsA[2] = {(cuuint64_t)IFOLD_K, (cuuint64_t)(4 * IFOLD_ROWS)};
    cuuint64_t stridesA[1] = {(cuuint64_t)IFOLD_K * 2};
    cuuint32_t boxA[2] = {(cuuint32_t)IFoldCfg::BK, 128};
    int rc = encode_map(&pl->tm_basis_ifold, pl->basis_ifold, 2, dimsA, stridesA, boxA, IFoldCfg::BK);
    if (rc) return rc;
  }
  if (pl->spec_split) {
    cuuint64_t dims[3] = {(cuuint64_t)DFT_K, (cuuint64_t)pl->T, (cuuint64_t)(2 * pl->max_signals)};
    cuuint64_t strides[2] = {(cuuint64_t)DFT_K * 2, (cuuint64_t)pl->T * DFT_K * 2};
    cuuint32_t box2[3] = {(cuuint32_t)PLAIN_BK, (cuuint32_t)(DFT_TILE_N / 2), 1};
    int rc = encode_map(&pl->tm_spec2, pl->spec_split, 3, dims, strides, box2, PLAIN_BK);
    if (rc) return rc;
    cuuint32_t box3[3] = {(cuuint32_t)IFoldCfg::BK, (cuuint32_t)(IFOLD_TILE_N / 2), 1};
    rc = encode_map(&pl->tm_spec_ifold, pl->spec_split, 3, dims, strides, box3, IFoldCfg::BK);
    if (rc) return rc;
    cuuint32_t box_st[3] = {(cuuint32_t)SPT_BINS, (cuuint32_t)SPT_FRAMES, 1};
    rc = encode_map(&pl->tm_spec_store, pl->spec_split, 3, dims, strides, box_st, 64);
    if (rc) return rc;
  }
  return AVSEP_OK;
}

static void fill_basis(std::vector<uint16_t>& fwd, std::vector<uint16_t>& inv) {
  const int NF = AVSEP_N_FFT;
  std::vector<double> win(NF), cs(NF), sn(NF);
  for (int n = 0; n < NF; ++n) {
    win[n] = 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)NF);   // periodic Hann
    cs[n] = std::cos(2.0 * M_PI * (double)n / (double)NF);
    sn[n] = std::sin(2.0 * M_PI * (double)n / (double)NF);
  }
  fwd.assign((size_t)2 * DFT_ROWS * DFT_K, 0);
  inv.assign((size_t)2 * DFT_ROWS * DFT_K, 0);
  auto put = [](std::vector<uint16_t>& dst, int row, int k, double v) {
    const float f = (float)v;
    const uint16_t hi = host_f32_to_bf16(f);
    const uint16_t lo = host_f32_to_bf16(f - host_bf16_to_f32(hi));
    dst[(size_t)row * DFT_K + k] = hi;
    dst[(size_t)(DFT_ROWS + row) * DFT_K + k] = lo;
  };
  // forward: row R = 128 m + 32 q + 16 part + i  <->  bin f = 64 m + 16 q + i, part 0 = cos, 1 = -sin
  for (int R = 0; R < DFT_ROWS; ++R) {
    const int m = R >> 7, q = (R >> 5) & 3, part = (R >> 4) & 1, i = R & 15;
    const int f = 64 * m + 16 * q + i;
    for (int n = 0; n < NF; ++n) {
      const int a = (int)(((long long)f * n) % NF);
      put(fwd, R, n, part == 0 ? win[n] * cs[a] : -win[n] * sn[a]);
    }
  }
  // inverse: row R = 128 m + 32 q + jj  <->  sample n = 32 m + jj + 256 q of the frame;
  // column k < 512: coefficient of Re X[k]; k >= 512: coefficient of Im X[k - 512]   (irfft, n = 1022)
  for (int R = 0; R < DFT_ROWS; ++R) {
    const int m = R >> 7, q = (R >> 5) & 3, jj = R & 31;
    const int n = 32 * m + jj + 256 * q;
    if (n >= NF) continue;
    for (int f = 0; f < AVSEP_N_BINS; ++f) {
      const double c = (f == 0 || f == NF / 2) ? 1.0 : 2.0;
      const int a = (int)(((long long)f * n) % NF);
      put(inv, R, f, win[n] * c * cs[a] / (double)NF);
      put(inv, R, AVSEP_N_BINS + f, -win[n] * c * sn[a] / (double)NF);
    }
  }
}

static void fill_basis_fold(std::vector<uint16_t>& out) {
  const int NF = AVSEP_N_FFT;
  out.assign((size_t)4 * FOLD_BINS * FOLD_K, 0);
  for (int f = 0; f < FOLD_BINS; ++f)
    for (int n = 0; n < FOLD_K; ++n) {
      const double win = 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)NF);
      const int a = (int)(((long long)f * n) % NF);
      const double ang = 2.0 * M_PI * (double)a / (double)NF;
      const double vals[2] = {win * std::cos(ang), -win * std::sin(ang)};
      for (int kind = 0; kind < 2; ++kind) {
        const float v = (float)vals[kind];
        const uint16_t hi = host_f32_to_bf16(v);
        const uint16_t lo = host_f32_to_bf16(v - host_bf16_to_f32(hi));
        out[((size_t)(kind * 2 + 0) * FOLD_BINS + f) * FOLD_K + n] = hi;
        out[((size_t)(kind * 2 + 1) * FOLD_BINS + f) * FOLD_K + n] = lo;
      }
    }
}

// folded inverse basis (idft_fold.cuh): plane (kind * 2 + split), row R in CTA-group order, column = bin f
static void fill_basis_ifold(std::vector<uint16_t>& out) {
  const int NF = AVSEP_N_FFT;
  out.assign((size_t)4 * IFOLD_ROWS * IFOLD_K, 0);
  for (int R = 0; R < IFOLD_ROWS; ++R) {
    const int n = ifold_row(R);
    const double win = 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)NF);
    for (int f = 0; f < IFOLD_K; ++f) {
      const bool edge = (f == 0 || f == NF / 2);
      const double c = edge ? 1.0 : 2.0;
      const int a = (int)(((long long)f * n) % NF);
      const double ang = 2.0 * M_PI * (double)a / (double)NF;
      // irfft ignores the imaginary part of the DC and Nyquist bins
      const double vals[2] = {win * c * std::cos(ang) / (double)NF, edge ? 0.0 : -win * c * std::sin(ang) / (double)NF};
      for (int kind = 0; kind < 2; ++kind) {
        const float v = (float)vals[kind];
        const uint16_t hi = host_f32_to_bf16(v);
        const uint16_t lo = host_f32_to_bf16(v - host_bf16_to_f32(hi));
        out[((size_t)(kind * 2 + 0) * IFOLD_ROWS + R) * IFOLD_K + f] = hi;
        out[((size_t)(kind * 2 + 1) * IFOLD_ROWS + R) * IFOLD_K + f] = lo;
      }
    }
  }
}

static int launch_ifold(avsep_plan* pl, int S, float* wav_dev, cudaStream_t stream) {
  IFoldParams ip;
  memset(&ip, 0, sizeof(ip));
  ip.num_signals = S;
  ip.T = pl->T;
  ip.lo_offset = pl->max_signals;
  ip.wav = wav_dev;
  ip.out_len = pl->out_len;
  ip.inv_env = pl->inv_env;
  ip.err_flag = pl->err_flag;
  const int grid2 = 2 * std::min(2 * S, plan_pairs(pl));
  if (grid2 <= 0) return AVSEP_OK;
  auto kern = idft_fold2_kernel<IFoldCfg::BK>;
  AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, IFoldCfg::SMEM_BYTES));
  kern<<<grid2, IFOLD_THREADS, IFoldCfg::SMEM_BYTES, stream>>>(pl->tm_basis_ifold, pl->tm_spec_ifold, ip);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

// folded forward transform of S signals, E / O operand generated in the kernel from the fp32 waveforms
static int launch_fold3(avsep_plan* pl, const FoldParams& prm, cudaStream_t stream) {
  const int n_tiles = ceil_div(prm.T, FOLD2_TILE_N);
  const int pair_tiles = prm.num_signals * n_tiles * (FOLD_M_TILES / 2);
  const int grid2 = 2 * std::min(pair_tiles, plan_pairs(pl));
  if (grid2 <= 0) return AVSEP_OK;
  auto kern = dft_fold3_fwd_kernel<2>;
  AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Fold3Cfg::SMEM_BYTES));
  kern<<<grid2, FOLD3_THREADS, Fold3Cfg::SMEM_BYTES, stream>>>(pl->tm_basis_fold32, prm);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

// forward transform of a device-summed mixture batch from its NSRC sources alone (see dft_fold3m in dft_fold.cuh)
template <int NSRC>
static int launch_fold3m_n(avsep_plan* pl, const FoldParams& prm, cudaStream_t stream) {
  const int n_tiles = ceil_div(prm.T, Fold3mShape<NSRC>::TILE_F);
  const int pair_tiles = prm.num_signals * n_tiles * (FOLD_M_TILES / 2);
  const int grid2 = 2 * std::min(pair_tiles, plan_pairs(pl));
  if (grid2 <= 0) return AVSEP_OK;
  auto kern = dft_fold3m_fwd_kernel<2, NSRC>;
  AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Fold3Cfg::SMEM_BYTES));
  kern<<<grid2, FOLD3_THREADS, Fold3Cfg::SMEM_BYTES, stream>>>(pl->tm_basis_fold32, prm);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

static int launch_fold3m(avsep_plan* pl, const FoldParams& prm, int nsrc, cudaStream_t stream) {
  switch (nsrc) {
    case 2: return launch_fold3m_n<2>(pl, prm, stream);
    case 3: return launch_fold3m_n<3>(pl, prm, stream);
    case 4: return launch_fold3m_n<4>(pl, prm, stream);
    default: return set_error(AVSEP_ERR_INVALID, "launch_fold3m: %d sources per mixture (2..4 supported)", nsrc);
  }
}

// the plain 1024 x 1024 CTA-pair GEMM (cross-check of the folded kernels, raw-accumulator debug hook)
template <int EPI>
static int launch_gemm(avsep_plan* pl, const CUtensorMap& tmA, const CUtensorMap& tmB2, DftParams prm, cudaStream_t stream) {
  const int pair_tiles = prm.total_tiles / 2;
  const int grid2 = 2 * std::min(pair_tiles, plan_pairs(pl));
  if (grid2 <= 0) return AVSEP_OK;
  auto kern = dft_gemm2_kernel<PLAIN_BK, EPI>;
  AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Dft2Cfg<PLAIN_BK, EPI>::SMEM_BYTES));
  kern<<<grid2, DFT_THREADS, Dft2Cfg<PLAIN_BK, EPI>::SMEM_BYTES, stream>>>(tmA, tmB2, prm);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

static DftParams base_params(const avsep_plan* pl, int S) {
  DftParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.num_signals = S;
  prm.T = pl->T;
  prm.n_tiles = ceil_div(pl->T, DFT_TILE_N);
  prm.total_tiles = S * prm.n_tiles * DFT_M_TILES;
  prm.lo_offset = pl->max_signals;
  prm.out_len = pl->out_len;
  prm.inv_env = pl->inv_env;
  prm.err_flag = pl->err_flag;
  return prm;
}

// operand of the plain forward GEMM (allocated on first use: the product path never needs it)
static int stage_waves(avsep_plan* pl, const float* wave_dev, int S, cudaStream_t stream, int mix_B = 0, int mix_N = 0) {
  if (!pl->wave_split) {
    const size_t bytes = (size_t)2 * pl->max_signals * pl->Lpad * sizeof(__nv_bfloat16);
    AVSEP_CUDA(cudaMalloc(&pl->wave_split, bytes));
    pl->workspace += bytes;
    int rc = build_maps(pl);
    if (rc) return rc;
  }
  const size_t lo_stride = (size_t)pl->max_signals * pl->Lpad;
  dim3 grid(ceil_div(pl->Lpad, 256 * 8), S);
  wave_split_kernel<<<grid, 256, 0, stream>>>(wave_dev, S, pl->wave_len, pl->Lpad, lo_stride, pl->wave_split, mix_B, mix_N);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

static int ensure_spec(avsep_plan* pl) {
  if (pl->spec_split) return AVSEP_OK;
  const size_t bytes = (size_t)2 * pl->max_signals * pl->T * DFT_K * sizeof(__nv_bfloat16);
  AVSEP_CUDA(cudaMalloc(&pl->spec_split, bytes));
  pl->workspace += bytes;
  return build_maps(pl);
}

}  // namespace avsep

using namespace avsep;

extern "C" {

int avsep_plan_create(avsep_plan** plan, int max_signals, int wave_len) {
  AVSEP_REQUIRE(plan != nullptr, "avsep_plan_create: plan is null");
  *plan = nullptr;
  AVSEP_REQUIRE(max_signals >= 1, "avsep_plan_create: max_signals must be >= 1 (got %d)", max_signals);
  AVSEP_REQUIRE(wave_len > AVSEP_N_FFT / 2, "avsep_plan_create: wave_len %d too short for reflect padding of %d",
                wave_len, AVSEP_N_FFT / 2);
  int rc = require_sm100();
  if (rc) return rc;
  avsep_plan* pl = new avsep_plan();
  cudaGetDevice(&pl->device);
  pl->max_signals = max_signals;
  pl->wave_len = wave_len;
  pl->T = 1 + wave_len / AVSEP_HOP;
  pl->Lpad = AVSEP_HOP * (pl->T + 4);
  pl->out_len = AVSEP_HOP * (pl->T - 1);

  std::vector<uint16_t> fwd, inv;
  fill_basis(fwd, inv);
  const size_t basis_bytes = fwd.size() * sizeof(uint16_t);
  auto fail = [&](int code) { avsep_plan_destroy(pl); return code; };
#define PLAN_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(set_error(AVSEP_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)));   \
  } while (0)
  PLAN_CUDA(cudaMalloc(&pl->basis_fwd, basis_bytes));
  PLAN_CUDA(cudaMalloc(&pl->basis_inv, basis_bytes));
  PLAN_CUDA(cudaMemcpy(pl->basis_fwd, fwd.data(), basis_bytes, cudaMemcpyHostToDevice));
  PLAN_CUDA(cudaMemcpy(pl->basis_inv, inv.data(), basis_bytes, cudaMemcpyHostToDevice));
  pl->workspace += 2 * basis_bytes;

  // reciprocal window sum-square envelope (librosa.filters.window_sumsquare), padded coordinates
  {
    const int n_env = AVSEP_HOP * (pl->T + 3);
    std::vector<double> env(n_env, 0.0);
    for (int t = 0; t < pl->T; ++t)
      for (int n = 0; n < AVSEP_N_FFT; ++n) {
        const double w = 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)AVSEP_N_FFT);
        env[(size_t)t * AVSEP_HOP + n] += w * w;
      }
    std::vector<float> inv_env(n_env);
    for (int i = 0; i < n_env; ++i) inv_env[i] = env[i] > 2.2250738585072014e-308 ? (float)(1.0 / env[i]) : 1.0f;
    PLAN_CUDA(cudaMalloc(&pl->inv_env, sizeof(float) * n_env));
    PLAN_CUDA(cudaMemcpy(pl->inv_env, inv_env.data(), sizeof(float) * n_env, cudaMemcpyHostToDevice));
    pl->workspace += sizeof(float) * n_env;
  }
  PLAN_CUDA(cudaMalloc(&pl->err_flag, sizeof(int)));
  PLAN_CUDA(cudaMemset(pl->err_flag, 0, sizeof(int)));

  {
    std::vector<uint16_t> fb;
    fill_basis_fold(fb);
    PLAN_CUDA(cudaMalloc(&pl->basis_fold, fb.size() * 2));
    PLAN_CUDA(cudaMemcpy(pl->basis_fold, fb.data(), fb.size() * 2, cudaMemcpyHostToDevice));
    pl->workspace += fb.size() * 2;
  }
  {
    std::vector<uint16_t> ib;
    fill_basis_ifold(ib);
    PLAN_CUDA(cudaMalloc(&pl->basis_ifold, ib.size() * 2));
    PLAN_CUDA(cudaMemcpy(pl->basis_ifold, ib.data(), ib.size() * 2, cudaMemcpyHostToDevice));
    pl->workspace += ib.size() * 2;
  }
  rc = build_maps(pl);
  if (rc) return fail(rc);
#undef PLAN_CUDA
  *plan = pl;
  return AVSEP_OK;
}

int avsep_plan_destroy(avsep_plan* pl) {
  if (!pl) return AVSEP_OK;
  cudaFree(pl->basis_fwd);
  cudaFree(pl->basis_inv);
  cudaFree(pl->wave_split);
  cudaFree(pl->spec_split);
  cudaFree(pl->inv_env);
  cudaFree(pl->basis_fold);
  cudaFree(pl->basis_ifold);
  cudaFree(pl->wave_mix);
  cudaFree(pl->err_flag);
  delete pl;
  return AVSEP_OK;
}

int avsep_plan_frames(const avsep_plan* pl) { return pl ? pl->T : 0; }
int avsep_plan_istft_len(const avsep_plan* pl) { return pl ? pl->out_len : 0; }
size_t avsep_plan_workspace_bytes(const avsep_plan* pl) { return pl ? pl->workspace : 0; }

int avsep_plan_set_fold(avsep_plan* pl, int enable) {
  AVSEP_REQUIRE(pl != nullptr, "avsep_plan_set_fold: plan is null");
  ++pl->generation;
  pl->fold = enable != 0;
  return AVSEP_OK;
}

int avsep_plan_set_mix_linear(avsep_plan* pl, int enable) {
  AVSEP_REQUIRE(pl != nullptr, "avsep_plan_set_mix_linear: plan is null");
  ++pl->generation;
  pl->mix_linear = enable != 0;
  return AVSEP_OK;
}

int avsep_plan_set_sm_budget(avsep_plan* pl, int sms) {
  AVSEP_REQUIRE(pl != nullptr && sms >= 0 && (sms == 0 || sms >= 2), "avsep_plan_set_sm_budget: bad argument");
  ++pl->generation;
  pl->sm_budget = sms;
  return AVSEP_OK;
}

int avsep_plan_set_spec_staged(avsep_plan* pl, int enable) {
  AVSEP_REQUIRE(pl != nullptr, "avsep_plan_set_spec_staged: plan is null");
  ++pl->generation;
  pl->use_spec_tma = enable != 0;
  return AVSEP_OK;
}

int avsep_plan_set_inverse_fold(avsep_plan* pl, int enable) {
  AVSEP_REQUIRE(pl != nullptr, "avsep_plan_set_inverse_fold: plan is null");
  ++pl->generation;
  pl->inverse_fold = enable != 0;
  return AVSEP_OK;
}

}  // extern "C"

namespace avsep {
int plan_generation(const avsep_plan* pl) { return pl ? pl->generation : 0; }

static FoldParams fold_params(const avsep_plan* pl, int num_signals, float* mag_dev) {
  FoldParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.num_signals = num_signals;
  fp.T = pl->T;
  fp.mag = mag_dev;
  fp.prefetch = 1;
  fp.err_flag = pl->err_flag;
  fp.L = pl->wave_len;
  return fp;
}

int stft_impl(avsep_plan* pl, const float* wave_dev, int S, float* mag_dev, float* phase_dev, float* re_dev, float* im_dev,
              int reim_count, void* stream_, int mix_B, int mix_N) {
  AVSEP_REQUIRE(pl != nullptr, "avsep_stft: plan is null");
  AVSEP_REQUIRE(wave_dev && mag_dev, "avsep_stft: wave_dev and mag_dev are required");
  AVSEP_REQUIRE((re_dev == nullptr) == (im_dev == nullptr), "avsep_stft: re_dev and im_dev must be given together");
  AVSEP_REQUIRE(S >= 0 && S <= pl->max_signals, "avsep_stft: S = %d exceeds the plan capacity %d", S, pl->max_signals);
  if (S == 0) return AVSEP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  pl->last_wave = wave_dev;
  pl->last_mix_B = mix_B;
  pl->last_mix_N = mix_N;
  if (pl->fold && pl->mix_linear && mix_N >= 2 && mix_N <= 4 && S == (mix_N + 1) * mix_B && !phase_dev &&
      (!re_dev || reim_count == mix_B)) {
    // magnitudes of [mixture | source 0 | source 1 | ..] (and, for eval, (re, im) of the mixture) from the mix_N SOURCE
    // transforms alone: the mixture is the sum of its sources and the STFT is linear
    FoldParams fp = fold_params(pl, mix_B, mag_dev);
    fp.wave = wave_dev;
    fp.re = re_dev;
    fp.im = im_dev;
    fp.reim_count = re_dev ? mix_B : 0;
    return launch_fold3m(pl, fp, mix_N, stream);
  }
  if (pl->fold) {
    FoldParams fp = fold_params(pl, S, mag_dev);
    fp.wave = wave_dev;
    fp.phase = phase_dev;
    fp.re = re_dev;
    fp.im = im_dev;
    fp.reim_count = reim_count;
    if (mix_N > 0) {
      // `wave_dev` holds the N sources only: the mixture waveforms (signals 0..mix_B-1) are summed into a plan buffer
      const size_t BL = (size_t)mix_B * pl->wave_len;
      if (pl->mix_cap < BL) {
        AVSEP_CUDA(cudaStreamSynchronize(stream));
        cudaFree(pl->wave_mix);
        pl->wave_mix = nullptr;
        AVSEP_CUDA(cudaMalloc(&pl->wave_mix, BL * sizeof(float)));
        pl->workspace += (BL - pl->mix_cap) * sizeof(float);
        pl->mix_cap = BL;
      }
      mix_sum_kernel<<<(unsigned)std::min<size_t>((BL + 255) / 256, 4096), 256, 0, stream>>>(wave_dev, mix_N, BL, pl->wave_mix);
      AVSEP_LAUNCH_CHECK();
      fp.wave_mix = pl->wave_mix;
      fp.mix_B = mix_B;
    }
    return launch_fold3(pl, fp, stream);
  }
  int rc = stage_waves(pl, wave_dev, S, stream, mix_B, mix_N);
  if (rc) return rc;
  DftParams prm = base_params(pl, S);
  prm.mag = mag_dev;
  prm.phase = phase_dev;
  prm.re = re_dev;
  prm.im = im_dev;
  prm.reim_count = reim_count;
  return launch_gemm<EPI_STFT>(pl, pl->tm_basis_fwd, pl->tm_wave2, prm, stream);
}
}  // namespace avsep

extern "C" {

int avsep_stft(avsep_plan* pl, const float* wave_dev, int S, float* mag_dev, float* phase_dev, float* re_dev,
               float* im_dev, void* stream) {
  return stft_impl(pl, wave_dev, S, mag_dev, phase_dev, re_dev, im_dev, S, stream, 0, 0);
}

int avsep_stft_reim(avsep_plan* pl, const float* wave_dev, int S, float* mag_dev, float* re_dev, float* im_dev, int reim_count,
                    void* stream) {
  AVSEP_REQUIRE(reim_count >= 0 && reim_count <= S, "avsep_stft_reim: reim_count = %d outside [0, S = %d]", reim_count, S);
  return stft_impl(pl, wave_dev, S, mag_dev, nullptr, re_dev, im_dev, reim_count, stream, 0, 0);
}

int avsep_stft_mix(avsep_plan* pl, const float* sources_dev, int B, int N, float* mag_dev, float* phase_dev, void* stream) {
  AVSEP_REQUIRE(B >= 0 && N >= 1 && N <= AVSEP_MAX_SRC, "avsep_stft_mix: bad B = %d / N = %d", B, N);
  return stft_impl(pl, sources_dev, (N + 1) * B, mag_dev, phase_dev, nullptr, nullptr, 0, stream, B, N);
}

int avsep_stft_mix_reim(avsep_plan* pl, const float* sources_dev, int B, int N, float* mag_dev, float* re_dev, float* im_dev,
                        void* stream) {
  AVSEP_REQUIRE(B >= 0 && N >= 1 && N <= AVSEP_MAX_SRC, "avsep_stft_mix_reim: bad B = %d / N = %d", B, N);
  AVSEP_REQUIRE(re_dev && im_dev, "avsep_stft_mix_reim: re_dev and im_dev are required");
  return stft_impl(pl, sources_dev, (N + 1) * B, mag_dev, nullptr, re_dev, im_dev, B, stream, B, N);
}

int avsep_separate(avsep_plan* pl, const float* masks_dev, int N, int B, const float* mag_dev, const float* phase_dev,
                   const float* re_dev, const float* im_dev, int log_freq, int binary_mask, float mask_thres,
                   int include_mix, float* wav_dev, void* stream_) {
  AVSEP_REQUIRE(pl != nullptr, "avsep_separate: plan is null");
  AVSEP_REQUIRE(N >= 0 && N <= 64 && B >= 0, "avsep_separate: bad N = %d / B = %d", N, B);
  AVSEP_REQUIRE(N == 0 || masks_dev, "avsep_separate: masks_dev is null with N = %d", N);
  AVSEP_REQUIRE(wav_dev != nullptr, "avsep_separate: wav_dev is null");
  AVSEP_REQUIRE((re_dev && im_dev) || (mag_dev && phase_dev), "avsep_separate: give (mag, phase) or (re, im)");
  AVSEP_REQUIRE(pl->T <= DFT_TILE_N, "avsep_separate: the fused overlap-add supports up to %d frames per clip (plan has %d)",
                DFT_TILE_N, pl->T);
  include_mix = include_mix ? 1 : 0;
  const int nsig = N + include_mix;
  AVSEP_REQUIRE((long long)B * nsig <= pl->max_signals, "avsep_separate: B * (N + mix) = %lld exceeds the plan capacity %d",
                (long long)B * nsig, pl->max_signals);
  if (B == 0 || nsig == 0) return AVSEP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = ensure_spec(pl);
  if (rc) return rc;
  SpecPrepParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.masks = masks_dev;
  sp.mag = mag_dev;
  sp.phase = phase_dev;
  sp.re = (re_dev && im_dev) ? re_dev : nullptr;
  sp.im = (re_dev && im_dev) ? im_dev : nullptr;
  sp.N = N;
  sp.B = B;
  sp.T = pl->T;
  sp.Hm = log_freq ? 256 : AVSEP_N_BINS;
  sp.include_mix = include_mix;
  sp.binary = binary_mask;
  sp.thres = mask_thres;
  sp.log_freq = log_freq;
  if (log_freq && N > 0) {
    rc = get_taps(TAP_UNWARP_Y, 256, AVSEP_N_BINS, &sp.ytap);
    if (rc) return rc;
    rc = get_taps(TAP_LINEAR, pl->T, pl->T, &sp.xtap);
    if (rc) return rc;
  } else {
    sp.log_freq = 0;
    if (N > 0 && log_freq == 0) sp.Hm = AVSEP_N_BINS;
  }
  sp.out = pl->spec_split;
  sp.lo_stride = (size_t)pl->max_signals * pl->T * DFT_K;
  // staged kernel: (re, im) input, 16-byte aligned rows, and the widest 64-bin tile of the un-warp fits the staging buffer
  constexpr int kStageFloats = 4608;                     // 18 KB: one source of the widest tile (110 rows x 40), 3 sources elsewhere
  bool staged = sp.re != nullptr && (pl->T & 3) == 0 && pl->use_spec_tma &&
                ((reinterpret_cast<uintptr_t>(sp.re) | reinterpret_cast<uintptr_t>(sp.im) | reinterpret_cast<uintptr_t>(sp.masks)) & 15u) == 0;
  if (staged && sp.log_freq) {
    if (pl->unwarp_rmax < 0) {
      std::vector<int> i0(AVSEP_N_BINS);
      std::vector<float> fr(AVSEP_N_BINS);
      compute_taps(TAP_UNWARP_Y, 256, AVSEP_N_BINS, i0.data(), fr.data());
      int rmax = 0;
      for (int f = 0; f < AVSEP_N_BINS; f += SPT_BINS)
        rmax = std::max(rmax, std::min(i0[f + SPT_BINS - 1] + 1, 255) - std::max(i0[f], 0) + 1);
      pl->unwarp_rmax = rmax;
    }
    staged = pl->unwarp_rmax * SPT_COLS <= kStageFloats;
  }
  if (staged) {
    const int smem = 1024 + SPT_OUT_BYTES + 2 * SPT_TILE_FLOATS * 4 + kStageFloats * 4;   // 51 KB: 4 CTAs per SM
    AVSEP_CUDA(cudaFuncSetAttribute(spec_prep_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(ceil_div(pl->T, SPT_FRAMES), AVSEP_N_BINS / SPT_BINS, B);
    spec_prep_tma_kernel<<<grid, SPT_THREADS, smem, stream>>>(pl->tm_spec_store, sp, kStageFloats, pl->max_signals);
  } else if ((pl->T & 1) == 0) {
    dim3 grid(ceil_div(pl->T, SPECV_FRAMES_PER_CTA), AVSEP_N_BINS / SPECV_BINS_PER_CTA, B);
    spec_prep_vec_kernel<<<grid, SPECV_THREADS, 0, stream>>>(sp);
  } else {
    dim3 grid(ceil_div(pl->T, SPEC_FRAMES_PER_CTA), AVSEP_N_BINS / SPEC_BINS_PER_CTA, B);
    spec_prep_scalar_kernel<<<grid, 256, 0, stream>>>(sp);
  }
  AVSEP_LAUNCH_CHECK();
  if (pl->inverse_fold) return launch_ifold(pl, B * nsig, wav_dev, stream);
  DftParams prm = base_params(pl, B * nsig);
  prm.wav = wav_dev;
  return launch_gemm<EPI_ISTFT>(pl, pl->tm_basis_inv, pl->tm_spec2, prm, stream);
}

int avsep_istft(avsep_plan* pl, const float* mag_dev, const float* phase_dev, int S, float* wav_dev, void* stream) {
  return avsep_separate(pl, nullptr, 0, S, mag_dev, phase_dev, nullptr, nullptr, 0, 0, 0.5f, 1, wav_dev, stream);
}

// raw accumulator of the PLAIN forward / inverse GEMM on the operands staged in the plan (forward: by a prior avsep_stft
// with the fold switched off; inverse: by a prior avsep_separate / avsep_istft)
int avsep_debug_raw_gemm(avsep_plan* pl, int inverse, int S, float* out_dev, void* stream_) {
  AVSEP_REQUIRE(pl && out_dev, "avsep_debug_raw_gemm: null argument");
  AVSEP_REQUIRE(S >= 1 && S <= pl->max_signals, "avsep_debug_raw_gemm: bad S");
  if (inverse) {
    int rc = ensure_spec(pl);
    if (rc) return rc;
  } else {
    AVSEP_REQUIRE(pl->tm_wave_valid, "avsep_debug_raw_gemm: no staged forward operand (call avsep_stft with avsep_plan_set_fold(plan, 0) first)");
  }
  DftParams prm = base_params(pl, S);
  prm.raw = out_dev;
  return launch_gemm<EPI_RAW>(pl, inverse ? pl->tm_basis_inv : pl->tm_basis_fwd, inverse ? pl->tm_spec2 : pl->tm_wave2, prm,
                              (cudaStream_t)stream_);
}

// the transform kernel alone with its product epilogue (bench.py / tools time the kernels in isolation with this):
// inverse on the operand already staged in the plan; forward on the waveform pointer of the last avsep_stft call
int avsep_debug_gemm_only(avsep_plan* pl, int inverse, int S, float* out_dev, void* stream_) {
  AVSEP_REQUIRE(pl && out_dev, "avsep_debug_gemm_only: null argument");
  AVSEP_REQUIRE(S >= 1 && S <= pl->max_signals, "avsep_debug_gemm_only: bad S");
  DftParams prm = base_params(pl, S);
  if (inverse) {
    int rc = ensure_spec(pl);
    if (rc) return rc;
    AVSEP_REQUIRE(pl->T <= DFT_TILE_N, "avsep_debug_gemm_only: inverse needs T <= %d", DFT_TILE_N);
    if (pl->inverse_fold) return launch_ifold(pl, S, out_dev, (cudaStream_t)stream_);
    prm.wav = out_dev;
    return launch_gemm<EPI_ISTFT>(pl, pl->tm_basis_inv, pl->tm_spec2, prm, (cudaStream_t)stream_);
  }
  if (pl->fold) {
    AVSEP_REQUIRE(pl->last_wave != nullptr, "avsep_debug_gemm_only: call avsep_stft on this plan first");
    FoldParams fp = fold_params(pl, S, out_dev);
    fp.wave = pl->last_wave;
    if (pl->last_mix_N > 0) {                       // the mixture waveforms of the last call are still in the plan buffer
      fp.wave_mix = pl->wave_mix;
      fp.mix_B = pl->last_mix_B;
    }
    return launch_fold3(pl, fp, (cudaStream_t)stream_);
  }
  AVSEP_REQUIRE(pl->tm_wave_valid, "avsep_debug_gemm_only: no staged forward operand");
  prm.mag = out_dev;
  return launch_gemm<EPI_STFT>(pl, pl->tm_basis_fwd, pl->tm_wave2, prm, (cudaStream_t)stream_);
}

int avsep_debug_basis(avsep_plan* pl, int inverse, float* out_host) {
  AVSEP_REQUIRE(pl && out_host, "avsep_debug_basis: null argument");
  std::vector<uint16_t> h((size_t)2 * DFT_ROWS * DFT_K);
  AVSEP_CUDA(cudaMemcpy(h.data(), inverse ? pl->basis_inv : pl->basis_fwd, h.size() * 2, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < (size_t)DFT_ROWS * DFT_K; ++i)
    out_host[i] = host_bf16_to_f32(h[i]) + host_bf16_to_f32(h[(size_t)DFT_ROWS * DFT_K + i]);
  return AVSEP_OK;
}

// staged operand of the forward GEMM (hi + lo) for signal s, as fp32 [T, 1024]: lets the tests
// separate "staging is wrong" from "the tensor-core pipeline is wrong"
int avsep_debug_staged_frames(avsep_plan* pl, int s, float* out_host) {
  AVSEP_REQUIRE(pl && out_host && s >= 0 && s < pl->max_signals, "avsep_debug_staged_frames: bad argument");
  AVSEP_REQUIRE(pl->wave_split != nullptr, "avsep_debug_staged_frames: no staged forward operand");
  AVSEP_CUDA(cudaDeviceSynchronize());
  std::vector<uint16_t> hi(pl->Lpad), lo(pl->Lpad);
  const size_t lo_stride = (size_t)pl->max_signals * pl->Lpad;
  AVSEP_CUDA(cudaMemcpy(hi.data(), pl->wave_split + (size_t)s * pl->Lpad, pl->Lpad * 2, cudaMemcpyDeviceToHost));
  AVSEP_CUDA(cudaMemcpy(lo.data(), pl->wave_split + lo_stride + (size_t)s * pl->Lpad, pl->Lpad * 2, cudaMemcpyDeviceToHost));
  for (int t = 0; t < pl->T; ++t)
    for (int k = 0; k < DFT_K; ++k)
      out_host[(size_t)t * DFT_K + k] = host_bf16_to_f32(hi[t * AVSEP_HOP + k]) + host_bf16_to_f32(lo[t * AVSEP_HOP + k]);
  return AVSEP_OK;
}

}  // extern "C"
