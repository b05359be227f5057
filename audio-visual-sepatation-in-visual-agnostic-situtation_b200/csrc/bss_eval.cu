// SDR / SIR / SAR / SI-SDR of every clip of a batch on the device, hand-written.
//
// Replaces the per-clip host call of the reference, main.py:252-275:
//   get_metrics(mix, clean [N, L], estimate [N, L], sample_rate=11025, metrics_list=['sdr','sir','sar','si_sdr'])
// = asteroid -> mir_eval.separation.bss_eval_sources(compute_permutation=False) (BSS-eval v3: least-squares projection of the
// estimate onto the references delayed by 0..511 samples) and pb_bss_eval.evaluation.si_sdr.
//
// mir_eval builds the (N*512)^2 Gram matrix of the delayed references with 131072-point FFTs, solves it densely and
// measures the energies of the projected SIGNALS.  The same numbers come out of three much smaller pieces, all fp64:
//   1. bss_corr_kernel   the cross-correlations at lags 0..511 that the Gram matrix G and the right-hand sides D consist of,
//                        summed directly in the time domain (products of fp32 samples are exact in fp64)
//   2. bss_solve_kernel  G is block-Toeplitz (T[a][c] = R(c - a), R(k)[i][j] = sum_t s_i[t + k] s_j[t]): the block Levinson
//                        recursion solves G C = D in O(512^2 N^3) operations, one CTA per system, state in shared memory.
//                        The projection energies are quadratic forms, |P y|^2 = C . D, so no projected signal is formed:
//                          a_j = |projection of estimate j on its own reference|^2      (scalar Toeplitz system)
//                          b_j = |projection of estimate j on all references|^2         (block Toeplitz system)
//   3. bss_metric_kernel with E_j = |estimate j|^2:   SDR = a / (E - a),  SIR = a / (b - a),  SAR = b / (E - b)   (in dB),
//                        SI-SDR = (<s, y>^2 / |s|^2) / (E - <s, y>^2 / |s|^2)
// (orthogonality of least-squares residuals turns mir_eval's e_interf / e_artif energies into these differences).
// tests/test_bss_metrics.py pins all of it against the CPU restatement of mir_eval in oracle/bss_eval.py.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace avsep {

constexpr int BSS_F = 512;            // distortion filter length (mir_eval default)
constexpr int BSS_LB = 8;             // adjacent lags per thread
constexpr int BSS_GROUPS = 4;         // thread groups of 64 (x 8 lags = 512) splitting the samples of a staging block
constexpr int BSS_SUB = 2048;         // samples staged in shared memory at a time
constexpr int BSS_TCH = 8192;         // samples per CTA
constexpr int BSS_XS = BSS_SUB + BSS_F + 8;

__host__ __device__ inline int bss_chunks(int L) { return (L + BSS_TCH - 1) / BSS_TCH; }
__device__ __forceinline__ int bss_pad(int e) { return e + (e >> 3); }     // stride 9 between the lanes of a warp: no bank conflicts

struct BssParams {
  const void* ref;      // reference sources: element (b, i, t) at ref[b * ref_sb + i * ref_sn + t]
  const void* est;
  long long ref_sb, ref_sn, est_sb, est_sn;
  int B, N, L;
  double* partial;      // [B][2 N^2][chunks][512]  chunk sums of the lag products
  double* corr;         // [B][2 N^2][512]          pairs 0..N^2-1: R(k)[i][j]; N^2..2N^2-1: D[a][i][j] = <s_i delayed by a, y_j>
  double* proj;         // [B][2][N]                a_j, b_j
  double* out;          // [4][B][N]                sdr, sir, sar, si_sdr
};

// c(k) = sum_t x[t + k] * y[t], k = 0..511, for one (clip, ordered pair, chunk of samples)
template <typename T>
__global__ void __launch_bounds__(256) bss_corr_kernel(const BssParams p) {
  __shared__ double xs[BSS_XS + BSS_XS / 8 + 2];
  __shared__ double ys[BSS_SUB];
  const int chunk = blockIdx.x, pair = blockIdx.y, b = blockIdx.z;
  const int NN = p.N * p.N;
  const int i = (pair % NN) / p.N, j = (pair % NN) % p.N;
  const T* refs = static_cast<const T*>(p.ref) + (size_t)b * p.ref_sb;
  const T* ests = static_cast<const T*>(p.est) + (size_t)b * p.est_sb;
  // R(k)[i][j] = sum_t s_i[t + k] s_j[t]          -> x = s_i, y = s_j
  // D[a][i][j] = sum_t s_i[t] y_j[t + a]          -> x = y_j (estimate), y = s_i
  const T* x = pair < NN ? refs + (size_t)i * p.ref_sn : ests + (size_t)j * p.est_sn;
  const T* y = pair < NN ? refs + (size_t)j * p.ref_sn : refs + (size_t)i * p.ref_sn;
  const int grp = threadIdx.x >> 6, lt = threadIdx.x & 63;
  const int k0 = BSS_LB * lt;
  double acc[BSS_LB];
#pragma unroll
  for (int l = 0; l < BSS_LB; ++l) acc[l] = 0.0;
  for (int sub = 0; sub < BSS_TCH / BSS_SUB; ++sub) {
    const int t_base = chunk * BSS_TCH + sub * BSS_SUB;
    if (t_base >= p.L) break;                                  // block-uniform
    __syncthreads();
    for (int e = threadIdx.x; e < BSS_XS; e += 256) {
      const int t = t_base + e;
      xs[bss_pad(e)] = t < p.L ? (double)x[t] : 0.0;           // x beyond the clip is zero: the sum may run over every t < L
      if (e < BSS_SUB) ys[e] = t < p.L ? (double)y[t] : 0.0;
    }
    __syncthreads();
    const int u0 = (BSS_SUB / BSS_GROUPS) * grp;
    double w[BSS_LB];                                          // w[l] = x[u + k0 + l]
#pragma unroll
    for (int l = 0; l < BSS_LB - 1; ++l) w[l] = xs[bss_pad(u0 + k0 + l)];
    for (int u = u0; u < u0 + BSS_SUB / BSS_GROUPS; u += BSS_LB) {
#pragma unroll
      for (int r = 0; r < BSS_LB; ++r) {                       // rotating window: slot (r + l) % 8 holds x[u + r + k0 + l]
        w[(r + BSS_LB - 1) % BSS_LB] = xs[bss_pad(u + r + k0 + BSS_LB - 1)];
        const double yv = ys[u + r];
#pragma unroll
        for (int l = 0; l < BSS_LB; ++l) acc[l] = fma(w[(r + l) % BSS_LB], yv, acc[l]);
      }
    }
  }
  // the four groups of a CTA summed in a fixed order
  __syncthreads();
  double* red = xs;                                            // [4][512]
#pragma unroll
  for (int l = 0; l < BSS_LB; ++l) red[grp * BSS_F + k0 + l] = acc[l];
  __syncthreads();
  double* dst = p.partial + (((size_t)b * 2 * NN + pair) * gridDim.x + chunk) * BSS_F;
  for (int k = threadIdx.x; k < BSS_F; k += 256) dst[k] = ((red[k] + red[BSS_F + k]) + red[2 * BSS_F + k]) + red[3 * BSS_F + k];
}

// chunk sums -> correlations, fixed order
__global__ void __launch_bounds__(256) bss_reduce_kernel(const BssParams p, int chunks) {
  const size_t n = (size_t)p.B * 2 * p.N * p.N * BSS_F;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t bp = idx / BSS_F;
    const int k = (int)(idx % BSS_F);
    const double* src = p.partial + bp * chunks * BSS_F + k;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += src[(size_t)c * BSS_F];
    p.corr[idx] = s;
  }
}

// ---- small dense helpers on N x N blocks in registers ---------------------------------------------------------------
template <int N, int M>
__device__ __forceinline__ void small_solve(const double (&A)[N][N], double (&Bm)[N][M]) {
  // Bm <- A^-1 Bm by Gaussian elimination with partial pivoting (A is a symmetric positive definite error covariance; the
  // pivoting only guards degenerate inputs such as a silent reference)
  double a[N][N];
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) a[r][c] = A[r][c];
#pragma unroll
  for (int col = 0; col < N; ++col) {
    int piv = col;
    double best = fabs(a[col][col]);
#pragma unroll
    for (int r = col + 1; r < N; ++r)
      if (fabs(a[r][col]) > best) { best = fabs(a[r][col]); piv = r; }
#pragma unroll
    for (int r = col + 1; r < N; ++r) {
      if (r == piv) {
#pragma unroll
        for (int c = 0; c < N; ++c) { const double t = a[col][c]; a[col][c] = a[r][c]; a[r][c] = t; }
#pragma unroll
        for (int c = 0; c < M; ++c) { const double t = Bm[col][c]; Bm[col][c] = Bm[r][c]; Bm[r][c] = t; }
      }
    }
    const double inv = 1.0 / a[col][col];
#pragma unroll
    for (int r = col + 1; r < N; ++r) {
      const double f = a[r][col] * inv;
#pragma unroll
      for (int c = col; c < N; ++c) a[r][c] -= f * a[col][c];
#pragma unroll
      for (int c = 0; c < M; ++c) Bm[r][c] -= f * Bm[col][c];
    }
  }
#pragma unroll
  for (int r = N - 1; r >= 0; --r) {
#pragma unroll
    for (int c = 0; c < M; ++c) {
      double s = Bm[r][c];
#pragma unroll
      for (int q = r + 1; q < N; ++q) s -= a[r][q] * Bm[q][c];
      Bm[r][c] = s / a[r][r];
    }
  }
}

// sum of NV doubles per thread over the CTA (256 threads), result broadcast to all threads; fixed order.  The cross-warp
// stage is done by ONE thread per value (the fp64 pipe is the bottleneck of the recursion: no redundant additions).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch /* [9][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double s = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) scratch[warp * NV + q] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += scratch[w * NV + threadIdx.x];
    scratch[8 * NV + threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) v[q] = scratch[8 * NV + q];
}

// inverse of a small symmetric positive definite block: closed forms for N <= 3 (one division), elimination otherwise
template <int N>
__device__ __forceinline__ void small_inverse(const double (&A)[N][N], double (&inv)[N][N]) {
  if (N == 1) {
    inv[0][0] = 1.0 / A[0][0];
  } else if (N == 2) {
    const double r = 1.0 / (A[0][0] * A[1][1] - A[0][1] * A[1][0]);
    inv[0][0] = A[1][1] * r; inv[0][1] = -A[0][1] * r;
    inv[1][0] = -A[1][0] * r; inv[1][1] = A[0][0] * r;
  } else if (N == 3) {
    const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    const double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const double r = 1.0 / (A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02);
    inv[0][0] = c00 * r;
    inv[1][0] = c01 * r;
    inv[2][0] = c02 * r;
    inv[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * r;
    inv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * r;
    inv[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * r;
    inv[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * r;
    inv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * r;
    inv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * r;
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) inv[i][j] = i == j ? 1.0 : 0.0;
    small_solve<N, N>(A, inv);
  }
}
template <int N, int M>
__device__ __forceinline__ void small_mul(const double (&A)[N][N], const double (&Bm)[N][M], double (&C)[N][M]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < N; ++q) s += A[i][q] * Bm[q][j];
      C[i][j] = s;
    }
}

// Block Levinson recursion for the symmetric block-Toeplitz system  sum_c R(c - a) x[c] = d[a],  a = 0..F-1, with N x N blocks
// (R(-k) = R(k)^T) and NR right-hand-side columns.  Fv / Bv: forward / backward block vectors (T_m Fv = [Ef; 0..0], Fv[0] = I;
// T_m Bv = [0..0; Eb], Bv[m-1] = I), both updated with the reflection blocks Kf = Eb^-1 Df, Kb = Ef^-1 Db each order, and
// x_{m+1} = [x_m; 0] + Bv_{m+1} Eb_{m+1}^-1 (d[m] - eps).  Returns sum_a x[a] . d[a] per right-hand side (the projection energy).
//   Rg: [N][N][F] (pair-major as bss_corr writes it), dg: [N][NR-column pairs][F] through the `dcol` map
template <int N, int NR>
__device__ void block_levinson(const double* __restrict__ Rg, const double* __restrict__ dg, int d_pair_stride, const int (&dcol)[NR],
                               double* Fv, double* Bv, double* xv, double* scratch, double (&energy)[NR]) {
  constexpr int NN = N * N, NX = N * NR;
  const int tid = threadIdx.x;
  auto Rk = [&](int k, int i, int j) { return __ldg(Rg + ((size_t)(i * N + j)) * BSS_F + k); };                 // R(k)[i][j], k >= 0
  auto dv = [&](int a, int i, int r) { return __ldg(dg + (size_t)(i * d_pair_stride + dcol[r]) * BSS_F + a); };  // d[a][i][r]
  double Ef[N][N], Eb[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) { Ef[i][j] = Rk(0, i, j); Eb[i][j] = Ef[i][j]; }
  for (int c = tid; c < BSS_F; c += 256) {
#pragma unroll
    for (int q = 0; q < NN; ++q) { Fv[c * NN + q] = 0.0; Bv[c * NN + q] = 0.0; }
#pragma unroll
    for (int q = 0; q < NX; ++q) xv[c * NX + q] = 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    double x0[N][NR];
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = 0; j < N; ++j) { Fv[i * N + j] = i == j ? 1.0 : 0.0; Bv[i * N + j] = i == j ? 1.0 : 0.0; }
#pragma unroll
      for (int r = 0; r < NR; ++r) x0[i][r] = dv(0, i, r);
    }
    small_solve<N, NR>(Ef, x0);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int r = 0; r < NR; ++r) xv[i * NR + r] = x0[i][r];
  }
  __syncthreads();
  for (int m = 1; m < BSS_F; ++m) {
    // Df = sum_c R(c - m) Fv[c] = sum_c R(m - c)^T Fv[c];  Db = sum_c R(c + 1) Bv[c];  eps = sum_c R(m - c)^T x[c]     (c < m)
    double acc[2 * NN + NX];
#pragma unroll
    for (int q = 0; q < 2 * NN + NX; ++q) acc[q] = 0.0;
    for (int c = tid; c < m; c += 256) {
      double Rt[N][N], Rp[N][N];                               // Rt = R(m - c)^T, Rp = R(c + 1)
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) { Rt[i][j] = Rk(m - c, j, i); Rp[i][j] = Rk(c + 1, i, j); }
#pragma unroll
      for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double sf = 0.0, sb = 0.0;
#pragma unroll
          for (int q = 0; q < N; ++q) { sf += Rt[i][q] * Fv[c * NN + q * N + j]; sb += Rp[i][q] * Bv[c * NN + q * N + j]; }
          acc[i * N + j] += sf;
          acc[NN + i * N + j] += sb;
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          double se = 0.0;
#pragma unroll
          for (int q = 0; q < N; ++q) se += Rt[i][q] * xv[c * NX + q * NR + r];
          acc[2 * NN + i * NR + r] += se;
        }
      }
    }
    block_sum<2 * NN + NX>(acc, scratch);
    // every thread: reflection blocks, new error covariances, solution increment (N x N work, redundantly)
    double Kf[N][N], Kb[N][N], g[N][NR], Df[N][N], Db[N][N], inv[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) { Df[i][j] = acc[i * N + j]; Db[i][j] = acc[NN + i * N + j]; }
    small_inverse<N>(Eb, inv);
    small_mul<N, N>(inv, Df, Kf);                              // Kf = Eb^-1 Df
    small_inverse<N>(Ef, inv);
    small_mul<N, N>(inv, Db, Kb);                              // Kb = Ef^-1 Db
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double sf = 0.0, sb = 0.0;
#pragma unroll
        for (int q = 0; q < N; ++q) { sf += Db[i][q] * Kf[q][j]; sb += Df[i][q] * Kb[q][j]; }
        Ef[i][j] -= sf;                                        // Ef <- Ef - Db Kf
        Eb[i][j] -= sb;                                        // Eb <- Eb - Df Kb
      }
    double rhs[N][NR];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int r = 0; r < NR; ++r) rhs[i][r] = dv(m, i, r) - acc[2 * NN + i * NR + r];
    small_inverse<N>(Eb, inv);
    small_mul<N, NR>(inv, rhs, g);                             // g = Eb_new^-1 (d[m] - eps)
    // Fv'[c] = Fv[c] - Bv[c - 1] Kf,  Bv'[c] = Bv[c - 1] - Fv[c] Kb   (c = 0..m; Fv[m] = 0, Bv[-1] = 0): read, sync, write
    double nf[2][NN], nb[2][NN];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = tid + 256 * h;
      if (c <= m) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = 0; j < N; ++j) {
            double f = Fv[c * NN + i * N + j], bprev = c >= 1 ? Bv[(c - 1) * NN + i * N + j] : 0.0;
            double sf = 0.0, sb = 0.0;
#pragma unroll
            for (int q = 0; q < N; ++q) {
              sf += (c >= 1 ? Bv[(c - 1) * NN + i * N + q] : 0.0) * Kf[q][j];
              sb += Fv[c * NN + i * N + q] * Kb[q][j];
            }
            nf[h][i * N + j] = f - sf;
            nb[h][i * N + j] = bprev - sb;
          }
      }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = tid + 256 * h;
      if (c <= m) {
#pragma unroll
        for (int q = 0; q < NN; ++q) { Fv[c * NN + q] = nf[h][q]; Bv[c * NN + q] = nb[h][q]; }
        // x'[c] = x[c] + Bv'[c] g
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < N; ++q) s += nb[h][i * N + q] * g[q][r];
            xv[c * NX + i * NR + r] += s;
          }
      }
    }
    __syncthreads();
  }
  // projection energy per right-hand side: sum_a x[a] . d[a]
  double e[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) e[r] = 0.0;
  for (int c = tid; c < BSS_F; c += 256)
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int r = 0; r < NR; ++r) e[r] += xv[c * NX + i * NR + r] * dv(c, i, r);
  block_sum<NR>(e, scratch);
#pragma unroll
  for (int r = 0; r < NR; ++r) energy[r] = e[r];
}

// grid (1 + N, B): system 0 = all references (N x N blocks, N right-hand sides -> b_j), system 1 + j = reference j alone -> a_j
template <int N>
__global__ void __launch_bounds__(256) bss_solve_kernel(const BssParams p) {
  extern __shared__ double sm[];
  const int b = blockIdx.y, sys = blockIdx.x;
  const double* corr = p.corr + (size_t)b * 2 * N * N * BSS_F;
  double* scratch = sm;                                        // [9][2 N^2 + N^2]
  double* base = sm + 9 * 3 * N * N;
  if (sys == 0) {
    double* Fv = base;
    double* Bv = Fv + BSS_F * N * N;
    double* xv = Bv + BSS_F * N * N;
    int dcol[N];
#pragma unroll
    for (int r = 0; r < N; ++r) dcol[r] = r;
    double en[N];
    block_levinson<N, N>(corr, corr + (size_t)N * N * BSS_F, N, dcol, Fv, Bv, xv, scratch, en);
    if (threadIdx.x < N) p.proj[((size_t)b * 2 + 1) * N + threadIdx.x] = en[threadIdx.x];
  } else {
    const int j = sys - 1;
    double* Fv = base;
    double* Bv = Fv + BSS_F;
    double* xv = Bv + BSS_F;
    const int dcol[1] = {0};
    double en[1];
    // scalar system of source j: R_jj(k) is pair (j, j); d[a] = D[a][j][j] is pair N^2 + j * N + j
    block_levinson<1, 1>(corr + (size_t)(j * N + j) * BSS_F, corr + (size_t)(N * N + j * N + j) * BSS_F, 1, dcol, Fv, Bv, xv, scratch, en);
    if (threadIdx.x == 0) p.proj[((size_t)b * 2 + 0) * N + j] = en[0];
  }
}

// E_j = |estimate j|^2 and the four metrics; one warp per (clip, source)
template <typename T>
__global__ void __launch_bounds__(256) bss_metric_kernel(const BssParams p) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= p.B * p.N) return;
  const int b = gw / p.N, j = gw % p.N;
  const T* y = static_cast<const T*>(p.est) + (size_t)b * p.est_sb + (size_t)j * p.est_sn;
  double e = 0.0;
  for (int t = lane; t < p.L; t += 32) { const double v = (double)y[t]; e = fma(v, v, e); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if (lane != 0) return;
  const int NN = p.N * p.N;
  const double* corr = p.corr + (size_t)b * 2 * NN * BSS_F;
  const double a = p.proj[((size_t)b * 2 + 0) * p.N + j], bb = p.proj[((size_t)b * 2 + 1) * p.N + j];
  const double rjj = corr[(size_t)(j * p.N + j) * BSS_F], dot = corr[(size_t)(NN + j * p.N + j) * BSS_F];
  auto db = [](double num, double den) { return den > 0.0 ? 10.0 * log10(num / den) : INFINITY; };   // mir_eval's _safe_db: den = 0 -> inf
  const size_t o = (size_t)b * p.N + j, plane = (size_t)p.B * p.N;
  p.out[o] = db(a, e - a);
  p.out[plane + o] = db(a, bb - a);
  p.out[2 * plane + o] = db(bb, e - bb);
  const double tgt = dot * dot / rjj;
  p.out[3 * plane + o] = 10.0 * log10(tgt / (e - tgt));
}

template <int N>
static int launch_solve(const BssParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(double) * (9 * 3 * N * N + (size_t)BSS_F * (2 * N * N + N * N));
  auto kern = bss_solve_kernel<N>;
  AVSEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(1 + N, p.B), 256, smem, stream>>>(p);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

}  // namespace avsep

using namespace avsep;

extern "C" {

size_t avsep_bss_eval_workspace_bytes(int B, int N, int L) {
  if (B < 1 || N < 1 || L < 1) return 0;
  const size_t pairs = (size_t)B * 2 * N * N;
  return sizeof(double) * (pairs * bss_chunks(L) * BSS_F + pairs * BSS_F + (size_t)B * 2 * N) + 256;
}

int avsep_bss_eval(const void* ref_dev, long long ref_stride_b, long long ref_stride_n, const void* est_dev, long long est_stride_b,
                   long long est_stride_n, int is_double, int B, int N, int L, double* out_dev, void* workspace_dev, void* stream_) {
  AVSEP_REQUIRE(ref_dev && est_dev && out_dev && workspace_dev, "avsep_bss_eval: null argument");
  AVSEP_REQUIRE(B >= 0 && B <= 65535 && N >= 1 && N <= AVSEP_MAX_SRC && L >= 1, "avsep_bss_eval: bad B = %d / N = %d / L = %d", B, N, L);
  int rc = require_sm100();
  if (rc) return rc;
  if (B == 0) return AVSEP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  BssParams p;
  p.ref = ref_dev; p.est = est_dev;
  p.ref_sb = ref_stride_b; p.ref_sn = ref_stride_n; p.est_sb = est_stride_b; p.est_sn = est_stride_n;
  p.B = B; p.N = N; p.L = L;
  const int chunks = bss_chunks(L);
  const size_t pairs = (size_t)B * 2 * N * N;
  uintptr_t w = ((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255;
  p.partial = reinterpret_cast<double*>(w);
  p.corr = p.partial + pairs * chunks * BSS_F;
  p.proj = p.corr + pairs * BSS_F;
  p.out = out_dev;
  dim3 grid(chunks, 2 * N * N, B);
  if (is_double) bss_corr_kernel<double><<<grid, 256, 0, stream>>>(p);
  else bss_corr_kernel<float><<<grid, 256, 0, stream>>>(p);
  AVSEP_LAUNCH_CHECK();
  bss_reduce_kernel<<<std::min<size_t>((pairs * BSS_F + 255) / 256, 2048), 256, 0, stream>>>(p, chunks);
  AVSEP_LAUNCH_CHECK();
  switch (N) {
    case 1: rc = launch_solve<1>(p, stream); break;
    case 2: rc = launch_solve<2>(p, stream); break;
    case 3: rc = launch_solve<3>(p, stream); break;
    default: rc = launch_solve<4>(p, stream); break;
  }
  if (rc) return rc;
  const int warps = B * N;
  if (is_double) bss_metric_kernel<double><<<ceil_div(warps * 32, 256), 256, 0, stream>>>(p);
  else bss_metric_kernel<float><<<ceil_div(warps * 32, 256), 256, 0, stream>>>(p);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

}  // extern "C"
