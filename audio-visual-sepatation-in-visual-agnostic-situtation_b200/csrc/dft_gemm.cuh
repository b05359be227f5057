// The windowed DFT / inverse DFT as a split-precision (bf16 x 3) tcgen05 GEMM for sm_100a.
//
//   D[row, t] = sum_k  Basis[row, k] * X[t, k]            row < 1024, t < T, k < 1024
//
// forward (STFT, dataset/base.py:142-147):  Basis = window * (cos | -sin) rows, X[t, k] = padded wave[256 t + k]
// inverse (iSTFT, utils.py:101-104)      :  Basis = window * irfft rows,        X[t, k] = masked (re | im) spectrum
//
// Both operands are pre-split x = hi + lo (bf16 each, 16 significant bits); the kernel accumulates
// hi*hi + hi*lo + lo*hi in fp32 TMEM accumulators, which holds the 1e-4 tolerance of the north star
// for n_fft = 1022 (K padded to 1024 with two zero basis columns).
//
// This PLAIN (unfolded) GEMM is the in-library cross-check of the folded product kernels (dft_fold.cuh, idft_fold.cuh) and
// the raw-accumulator debug hook; round 1 shipped it as the inverse transform.  CTA pairs (cta_group::2), persistent over
// (signal, frame-tile, row-tile) work items, 256 x 256 output tile per pair (UMMA M=256, N=256, K=16), K pipelined in
// 64-wide stages filled by TMA (128B swizzle), two 256-column TMEM accumulator buffers so the epilogue of tile i overlaps
// the MMAs of tile i+1.
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane of the leader CTA)
//   warps 2..5  : epilogue (each owns the 32 TMEM lanes selected by warp_id % 4)
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace avsep {

constexpr int DFT_K = 1024;        // n_fft padded
constexpr int DFT_ROWS = 1024;     // basis rows (forward: 512 re + 512 im; inverse: 1022 samples + 2 zero rows)
constexpr int DFT_TILE_M = 128;
constexpr int DFT_TILE_N = 256;
constexpr int DFT_M_TILES = DFT_ROWS / DFT_TILE_M;   // 8
constexpr int DFT_THREADS = 192;
constexpr int DFT_SCRATCH_BYTES = 2 * 4 * 32 * 32 * 4;   // inverse epilogue overlap-add exchange

enum DftEpilogue : int { EPI_STFT = 0, EPI_ISTFT = 1, EPI_RAW = 2 };

constexpr int DFT_SMEM_LIMIT = 232448;   // 227 KB opt-in shared memory per CTA on sm_100

struct DftParams {
  int num_signals;   // S
  int T;             // frames per signal
  int n_tiles;       // ceil(T / 256) frame tiles per signal
  int total_tiles;   // S * n_tiles * 8
  int lo_offset;     // index of the lo split along the outermost tensor-map dimension of X
  // EPI_STFT outputs [S, 512, T] (any may be null except mag)
  float* mag;
  float* phase;
  float* re;
  float* im;
  int reim_count;   // re / im are written for signals s < reim_count only
  // EPI_ISTFT output [S, out_len]
  float* wav;
  int out_len;            // 256 * (T - 1)
  const float* inv_env;   // [256 * (T + 3)] reciprocal window sum-square envelope
  // EPI_RAW output [S, 1024, T]
  float* raw;
  int* err_flag;
};

// One tile of epilogue work for the warp owning TMEM lane group q (shared by the 1-CTA and the 2-CTA kernels).
// `release_tmem()` is called as soon as the accumulator has been read out (the MMA issuer may then reuse it).
template <int EPI, class Release>
__device__ __forceinline__ void dft_epilogue_tile(const DftParams& p, int s, int nt, int m, uint32_t taddr, float* scratch,
                                                  int q, int lane, Release release_tmem) {
  const int t_tile = nt * DFT_TILE_N;

  if (EPI == EPI_STFT) {
    // rows of this lane group: lanes 0..15 = re, lanes 16..31 = im of bins f = 64 m + 16 q + (lane & 15)
    const int part = lane >> 4;
    const int f = 64 * m + 16 * q + (lane & 15);
    const size_t row_off = ((size_t)s * AVSEP_N_BINS + f) * (size_t)p.T;
    const bool vec_ok = (p.T & 3) == 0;
    const bool want_reim = p.re != nullptr && s < p.reim_count;
    for (int c0 = 0; c0 < DFT_TILE_N; c0 += 32) {
      if (t_tile + c0 >= p.T) break;
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
      float re[16], im[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t send = part ? v[j] : v[16 + j];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 16);
        re[j] = __uint_as_float(part ? recv : v[j]);
        im[j] = __uint_as_float(part ? v[16 + j] : recv);
      }
      const int t0 = t_tile + c0 + part * 16;
      float mg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) mg[j] = sqrtf(fmaf(re[j], re[j], im[j] * im[j]));
      if (vec_ok && t0 + 16 <= p.T) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(p.mag + row_off + t0 + j) = make_float4(mg[j], mg[j + 1], mg[j + 2], mg[j + 3]);
        if (p.phase) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(p.phase + row_off + t0 + j) =
                make_float4(atan2f(im[j], re[j]), atan2f(im[j + 1], re[j + 1]), atan2f(im[j + 2], re[j + 2]), atan2f(im[j + 3], re[j + 3]));
        }
        if (want_reim) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            *reinterpret_cast<float4*>(p.re + row_off + t0 + j) = make_float4(re[j], re[j + 1], re[j + 2], re[j + 3]);
            *reinterpret_cast<float4*>(p.im + row_off + t0 + j) = make_float4(im[j], im[j + 1], im[j + 2], im[j + 3]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (t0 + j < p.T) {
            p.mag[row_off + t0 + j] = mg[j];
            if (p.phase) p.phase[row_off + t0 + j] = atan2f(im[j], re[j]);
            if (want_reim) { p.re[row_off + t0 + j] = re[j]; p.im[row_off + t0 + j] = im[j]; }
          }
        }
      }
    }
    tcgen05_fence_before();
    release_tmem();
  } else if (EPI == EPI_RAW) {
    const int row = DFT_TILE_M * m + 32 * q + lane;
    for (int c0 = 0; c0 < DFT_TILE_N; c0 += 32) {
      if (t_tile + c0 >= p.T) break;
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int t = t_tile + c0 + j;
        if (t < p.T) p.raw[((size_t)s * DFT_ROWS + row) * (size_t)p.T + t] = __uint_as_float(v[j]);
      }
    }
    tcgen05_fence_before();
    release_tmem();
  } else {
    // EPI_ISTFT: row (q, lane) is sample n = 32 m + lane + 256 q of every frame; column = frame t.
    // Padded output position 256 (t + q) + 32 m + lane: the four lane groups hold the four
    // overlapping frames of an output block P = t + q.  Exchange through shared memory in a fixed
    // order (deterministic), scale by the reciprocal envelope, trim n_fft/2, clip, store.
    float* wav = p.wav + (size_t)s * p.out_len;
    const int j_in_blk = 32 * m + lane;
    auto emit = [&](int P, float sum) {
      const int pos = 256 * P + j_in_blk;
      const int idx = pos - (AVSEP_N_FFT / 2);
      if (idx >= 0 && idx < p.out_len) {
        float y = sum * __ldg(p.inv_env + pos);
        y = fminf(fmaxf(y, -1.0f), 1.0f);
        wav[idx] = y;
      }
    };
    for (int i = 0; i < DFT_TILE_N / 32; ++i) {
      const int c0 = 32 * i;
      const int buf = i & 1;
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
      float* mine = scratch + ((buf * 4 + q) * 32) * 32 + lane;
#pragma unroll
      for (int c = 0; c < 32; ++c) mine[c * 32] = __uint_as_float(v[c]);
      if (i == DFT_TILE_N / 32 - 1) {
        tcgen05_fence_before();
        release_tmem();
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int P = c0 + 8 * q + e;
        float sum = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int t = P - qq;
          if (t >= c0) sum += scratch[((buf * 4 + qq) * 32 + (t - c0)) * 32 + lane];
          else if (t >= 0) sum += scratch[(((buf ^ 1) * 4 + qq) * 32 + (t - c0 + 32)) * 32 + lane];
        }
        emit(P, sum);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    if (q < 3) {
      // tail blocks P = 256 + q need frames P - qq <= 255, all in the last exchange buffer (buf 1, c0 = 224)
      const int P = DFT_TILE_N + q;
      float sum = 0.f;
#pragma unroll
      for (int qq = 1; qq < 4; ++qq) {
        const int t = P - qq;
        if (t < DFT_TILE_N) sum += scratch[((1 * 4 + qq) * 32 + (t - 224)) * 32 + lane];
      }
      emit(P, sum);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): one 256 x 256 output tile per pair of SMs.
//
// With cta_group::1 (round 1's first kernel, since removed) every MMA re-reads its A (128 x 16) and B (256 x 16) tiles from the SM's own shared memory
// and TMA writes 2A + 2B per K step into the same shared memory: ~158 B/clk against the 128 B/clk the shared
// memory can move, which caps the tensor pipe at ~81 % (ncu measured 78-83 %).  A CTA pair splits the B rows
// (frames) between the two SMs, so each SM stages and reads only half of B: ~104 B/clk, below the limit.
//   * every CTA TMA-loads its own 128 basis rows and its half (128 frames) of B; all transaction bytes are
//     credited to the LEADER's full barrier
//   * the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256: 128 accumulator rows per CTA) and
//     commits with a cluster multicast, which frees the stage in both CTAs / wakes both epilogues
//   * each CTA runs the unchanged epilogue on its own 128 rows; accumulator release goes to the leader
template <int BK, int EPI>
struct Dft2Cfg {
  static_assert(BK == 32 || BK == 64, "K block is 32 (SWIZZLE_64B) or 64 (SWIZZLE_128B)");
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = DFT_TILE_M * ROW_BYTES;         // this CTA's 128 basis rows, one split
  static constexpr int B_BYTES = (DFT_TILE_N / 2) * ROW_BYTES;   // this CTA's half of the frames, one split
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // per CTA
  static constexpr int SCRATCH_BYTES = (EPI == 1) ? DFT_SCRATCH_BYTES : 0;
  static constexpr int STAGES = (DFT_SMEM_LIMIT - 1024 - 256 - SCRATCH_BYTES) / STAGE_BYTES;
  static constexpr int K_BLOCKS = DFT_K / BK;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + SCRATCH_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;
  static_assert(STAGES >= 2 && (2 * STAGES + 5) * 8 <= 256, "barrier block overflow");
};

template <int BK, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DFT_THREADS, 1)
dft_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DftParams p) {
  using Cfg = Dft2Cfg<BK, EPI>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* scratch = reinterpret_cast<float*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES);
  const uint32_t bar_base = smem_base + Cfg::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::BAR_OFFSET + 8 * (2 * Cfg::STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int pair_tiles = p.num_signals * p.n_tiles * (DFT_M_TILES / 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);              // 128 epilogue threads of each CTA of the pair
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  cluster_sync_all();                             // barriers of BOTH CTAs are initialised before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < pair_tiles; tile += npairs) {
        const int mp = tile % (DFT_M_TILES / 2);
        const int rest = tile / (DFT_M_TILES / 2);
        const int nt = rest % p.n_tiles;
        const int s = rest / p.n_tiles;
        const int m = 2 * mp + (int)rank;
        const int t0 = nt * DFT_TILE_N + (int)rank * (DFT_TILE_N / 2);
        for (int kb = 0; kb < Cfg::K_BLOCKS; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 21);
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          tma2_load_2d(&tmA, lead_full, st, kb * BK, m * DFT_TILE_M);
          tma2_load_2d(&tmA, lead_full, st + Cfg::A_BYTES, kb * BK, DFT_ROWS + m * DFT_TILE_M);
          tma2_load_3d(&tmB, lead_full, st + 2 * Cfg::A_BYTES, kb * BK, t0, s);
          tma2_load_3d(&tmB, lead_full, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, kb * BK, t0, p.lo_offset + s);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =======================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * DFT_TILE_M, DFT_TILE_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < pair_tiles; tile += npairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err_flag, 22);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * DFT_TILE_N);
        for (int kb = 0; kb < Cfg::K_BLOCKS; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 23);
          tcgen05_fence_after();
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_hi = umma_smem_desc(st, Cfg::ROW_BYTES);
          const uint64_t a_lo = umma_smem_desc(st + Cfg::A_BYTES, Cfg::ROW_BYTES);
          const uint64_t b_hi = umma_smem_desc(st + 2 * Cfg::A_BYTES, Cfg::ROW_BYTES);
          const uint64_t b_lo = umma_smem_desc(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, Cfg::ROW_BYTES);
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) {
            const uint64_t adv = (uint64_t)(k16 * 2);
            umma_f16_2sm(d_tmem, a_hi + adv, b_lo + adv, idesc, (kb | k16) != 0 ? 1u : 0u);
            umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
            umma_f16_2sm(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          umma_commit_2sm(empty_bar(stage), 3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(tfull_bar(acc), 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue (both CTAs, own 128 rows) =================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < pair_tiles; tile += npairs) {
      const int mp = tile % (DFT_M_TILES / 2);
      const int rest = tile / (DFT_M_TILES / 2);
      const int nt = rest % p.n_tiles;
      const int s = rest / p.n_tiles;
      const int m = 2 * mp + (int)rank;
      mbar_wait(tfull_bar(acc), acc_phase, p.err_flag, 24);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * DFT_TILE_N);
      const uint32_t lead_tempty = mapa_shared(tempty_bar(acc), 0);
      dft_epilogue_tile<EPI>(p, s, nt, m, taddr, scratch, q, lane, [&]() { mbar_arrive_cluster_relaxed(lead_tempty); });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                             // no CTA frees TMEM / exits while its peer may still signal it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace avsep
