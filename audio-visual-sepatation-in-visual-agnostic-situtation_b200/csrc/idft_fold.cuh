// Inverse STFT (utils.py:101-104 -> librosa.istft) as a FOLDED split-precision tcgen05 GEMM with the overlap-add
// in the epilogue: half the tensor work of the plain 1024 x 1024 inverse contraction.
//
// irfft of a 1022-point frame, times the symmetric periodic Hann window w (w[n] = w[1022 - n], w[0] = 0):
//   y[n]        = Pw[n] + Qw[n]          n = 0 .. 511
//   y[1022 - n] = Pw[n] - Qw[n]          n = 1 .. 510
//   Pw[n] =  w[n] / 1022 * sum_f c_f Re X[f] cos(2 pi f n / 1022)      (even in n)
//   Qw[n] = -w[n] / 1022 * sum_f c_f Im X[f] sin(2 pi f n / 1022)      (odd in n),  c_0 = c_511 = 1, else 2
// so each signal needs two 512 x 512 x T contractions (cos rows x Re, sin rows x Im) instead of one 1024 x 1024 x T.
//
// CTA pair tile (cta_group::2, M = 256, N = 256): 256 folded rows x all (<= 256) frames of one signal; Pw accumulates
// in TMEM columns [0, 256), Qw in [256, 512) of the same lanes, so an epilogue thread owns (Pw, Qw) of its (n, t).
//
// Overlap-add.  Padded output position 256 t + n; output block P = samples [256 P, 256 P + 256).  Row n of frame t
// feeds  a = Pw + Qw  into block t + (n >= 256) at sample j = n mod 256, and  b = Pw - Qw  into block t + 2 or t + 3
// at sample (254 - j) mod 256.  A CTA owns a set of 64 samples j that is closed under j -> 254 - j and the 128 rows
// {j, j + 256}, so every output sample is completed inside one CTA, in a fixed summation order (bit-reproducible, no
// atomics):  out[P][j] = a(j, P) + a(j + 256, P - 1) + b(254 - j + 256, P - 2) + b(254 - j, P - 3).
// The four kinds of contribution go through a 36-frame shared-memory ring indexed by DESTINATION sample slot, so both
// the scatter and the gather are conflict-free 32-lane rows; then x reciprocal window-sum-square envelope, trim
// n_fft / 2, clip to [-1, 1], coalesced store.
//   warp 0: TMA producer | warp 1: TMEM allocator + MMA issuer (leader CTA) | warps 2..9: epilogue
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace avsep {

constexpr int IFOLD_THREADS = 320;
constexpr int IFOLD_ROWS = 512;                    // folded rows n = 0 .. 511
constexpr int IFOLD_K = 512;                       // bins per kind (Re, Im)
constexpr int IFOLD_TILE_N = 256;                  // frames per pair tile (= the most the fused overlap-add supports)
constexpr int IFOLD_SMEM_LIMIT = 232448;

struct IFoldCfg {
  static constexpr int BK = 32;                    // SWIZZLE_64B
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = 128 * ROW_BYTES;  // one split of this CTA's 128 basis rows
  static constexpr int B_BYTES = 128 * ROW_BYTES;  // one split of this CTA's 128 frames
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;     // 32 KB
  static constexpr int STAGES = 5;
  static constexpr int KB_HALF = IFOLD_K / BK;     // K blocks per accumulator (Pw, then Qw)
  static constexpr int RING_FRAMES = 36;           // 32 frames of a step + the 3-frame look-back (+1)
  static constexpr int RING_BYTES = 4 * RING_FRAMES * 64 * 4;       // kinds a0, a1, b2, b3
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + RING_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;
  static_assert(SMEM_BYTES <= IFOLD_SMEM_LIMIT, "shared memory budget");
  static_assert((2 * STAGES + 3) * 8 <= 256, "barrier block overflow");
};

struct IFoldParams {
  int num_signals;
  int T;                  // frames per signal (<= 256)
  int lo_offset;          // plane index of the lo split of the operand
  float* wav;             // [S, out_len]
  int out_len;            // 256 (T - 1)
  const float* inv_env;   // [256 (T + 3)]
  int* err_flag;
};

// sample j (0..255) of output blocks owned by CTA group c4 (0..3) at ring slot s (0..63):
// slots 0..31 ascend from 32 c4, slots 32..63 are their mirrors 254 - j (ascending too); the self-mirrored
// samples 127 and 255 share group 3 (slot 31 = 127, slot 32 = 255)
__host__ __device__ inline int ifold_sample(int c4, int slot) {
  if (slot < 32) return 32 * c4 + slot;
  if (c4 == 3 && slot == 32) return 255;
  return 191 - 32 * c4 + slot;
}
// folded row n of basis row R = 128 c4 + l  (TMEM lane l of CTA group c4)
__host__ __device__ inline int ifold_row(int R) {
  const int c4 = R >> 7, l = R & 127;
  return ifold_sample(c4, l & 63) + 256 * (l >> 6);
}

// basis tensor map: rows = (kind * 2 + split) * 512 + R, cols = f (512)          kind 0 = cos (Pw), 1 = -sin (Qw)
// operand tensor map: dims (k = 1024 = [Re | Im] x 512 bins, t = T, plane = split * lo_offset + s)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(IFOLD_THREADS, 1)
idft_fold2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const IFoldParams p) {
  using Cfg = IFoldCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* ring = reinterpret_cast<float*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES);
  const uint32_t bar_base = smem_base + Cfg::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tempty_bar = bar_base + 8u * (2 * Cfg::STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 2);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::BAR_OFFSET + 8 * (2 * Cfg::STAGES + 2));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int pair_tiles = p.num_signals * 2;       // (signal, row half)

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 512);                   // 256 epilogue threads of each CTA of the pair
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < pair_tiles; tile += npairs) {
        const int mp = tile & 1;
        const int s = tile >> 1;
        const int row0 = (2 * mp + (int)rank) * 128;
        const int t0 = (int)rank * (IFOLD_TILE_N / 2);
        for (int kb = 0; kb < 2 * Cfg::KB_HALF; ++kb) {
          const int kind = kb / Cfg::KB_HALF;
          const int k0 = (kb % Cfg::KB_HALF) * Cfg::BK;
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 51);
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          tma2_load_2d(&tmA, lead_full, st, k0, (kind * 2 + 0) * IFOLD_ROWS + row0);
          tma2_load_2d(&tmA, lead_full, st + Cfg::A_BYTES, k0, (kind * 2 + 1) * IFOLD_ROWS + row0);
          tma2_load_3d(&tmB, lead_full, st + 2 * Cfg::A_BYTES, kind * IFOLD_K + k0, t0, s);
          tma2_load_3d(&tmB, lead_full, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, kind * IFOLD_K + k0, t0, p.lo_offset + s);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =======================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, IFOLD_TILE_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_phase = 0;
      for (int tile = pair; tile < pair_tiles; tile += npairs) {
        mbar_wait(tempty_bar, tile_phase ^ 1u, p.err_flag, 52);
        tcgen05_fence_after();
        for (int kb = 0; kb < 2 * Cfg::KB_HALF; ++kb) {
          const int kind = kb / Cfg::KB_HALF;
          const uint32_t d_tmem = tmem_base + (uint32_t)(kind * IFOLD_TILE_N);
          const bool first = (kb % Cfg::KB_HALF) == 0;
          mbar_wait(full_bar(stage), phase, p.err_flag, 53);
          tcgen05_fence_after();
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_hi = umma_smem_desc(st, Cfg::ROW_BYTES);
          const uint64_t a_lo = umma_smem_desc(st + Cfg::A_BYTES, Cfg::ROW_BYTES);
          const uint64_t b_hi = umma_smem_desc(st + 2 * Cfg::A_BYTES, Cfg::ROW_BYTES);
          const uint64_t b_lo = umma_smem_desc(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, Cfg::ROW_BYTES);
#pragma unroll
          for (int k16 = 0; k16 < Cfg::BK / 16; ++k16) {
            const uint64_t adv = (uint64_t)(k16 * 2);
            umma_f16_2sm(d_tmem, a_hi + adv, b_lo + adv, idesc, (first && k16 == 0) ? 0u : 1u);
            umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
            umma_f16_2sm(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          umma_commit_2sm(empty_bar(stage), 3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(tfull_bar, 3);
        tile_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue (8 warps per CTA): overlap-add ============
    const int q = warp & 3;                        // TMEM lane group this warp may touch
    const int half = (warp - 2) >> 2;              // which 16 of a step's 32 frame columns
    const int etid = (warp - 2) * 32 + lane;       // 0..255
    const int l = 32 * q + lane;                   // TMEM lane = basis row within the CTA's 128
    const int h = l >> 6, slot = l & 63;
    const uint32_t lead_tempty = mapa_shared(tempty_bar, 0);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nsteps = (p.T + 31) >> 5;
    constexpr int RF = Cfg::RING_FRAMES;
    uint32_t tile_phase = 0;
    for (int tile = pair; tile < pair_tiles; tile += npairs) {
      const int c4 = 2 * (tile & 1) + (int)rank;
      const int s = tile >> 1;
      float* wav = p.wav + (size_t)s * p.out_len;
      // destination of this row's mirrored contribution b: ring kind 2 (block t + 2) for rows n >= 255, kind 3
      // (block t + 3) for rows 1..254, at the slot of sample (254 - j) mod 256
      int kind_b = h ? 2 : 3;
      int slot_b = 63 - slot;
      bool zero_b = false;
      if (c4 == 3) {
        if (slot == 31) slot_b = 31;                               // sample 127 mirrors onto itself
        if (slot == 32) {                                          // sample 255 mirrors onto itself
          slot_b = 32;
          if (h == 0) kind_b = 2;                                  // row 255 -> y[767]: block t + 2
          else { kind_b = 3; zero_b = true; }                      // row 511 -> y[511] is the a term itself
        }
      }
      float* ring_a = ring + (h * RF) * 64 + slot;
      float* ring_b = ring + (kind_b * RF) * 64 + slot_b;
      const int gslot = etid & 63;
      const int gj = ifold_sample(c4, gslot);
      // The window sum-square envelope of a block that all four overlapping frames P-3..P reach (3 <= P <= T-1) depends on
      // the sample j only (same four addends, same order on the host): one load per tile instead of one per output.
      const float env_mid = __ldg(p.inv_env + 256 * min(3, p.T - 1) + gj);
      auto emit = [&](int P, float sum) {
        const int pos = 256 * P + gj;
        const int idx = pos - (AVSEP_N_FFT / 2);
        if (idx >= 0 && idx < p.out_len) {
          const float env = (P >= 3 && P < p.T) ? env_mid : __ldg(p.inv_env + pos);
          float y = sum * env;
          y = fminf(fmaxf(y, -1.0f), 1.0f);
          wav[idx] = y;
        }
      };
      mbar_wait(tfull_bar, tile_phase, p.err_flag, 54);
      tcgen05_fence_after();
      for (int st = 0; st < nsteps; ++st) {
        const int c0 = 32 * st + 16 * half;
        uint32_t vp[16], vq[16];
        tmem_ld_32x32b_x16(taddr + c0, vp);
        tmem_ld_32x32b_x16(taddr + IFOLD_TILE_N + c0, vq);
        tmem_ld_wait();
        int fs = c0 % RF;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pw = __uint_as_float(vp[i]), qw = __uint_as_float(vq[i]);
          ring_a[fs * 64] = pw + qw;
          ring_b[fs * 64] = zero_b ? 0.f : pw - qw;
          if (++fs == RF) fs = 0;
        }
        if (st == nsteps - 1) {                    // accumulators read out: the next tile's MMAs may start
          tcgen05_fence_before();
          mbar_arrive_cluster_relaxed(lead_tempty);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int P = 32 * st + (etid >> 6) + 4 * i;
          float sum = ring[(0 * RF + P % RF) * 64 + gslot];
          if (P >= 1) sum += ring[(1 * RF + (P - 1) % RF) * 64 + gslot];
          if (P >= 2) sum += ring[(2 * RF + (P - 2) % RF) * 64 + gslot];
          if (P >= 3) sum += ring[(3 * RF + (P - 3) % RF) * 64 + gslot];
          emit(P, sum);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      // tail blocks P = 32 nsteps + e, e = 0..2: only the look-back terms exist
      if (etid < 192) {
        const int e = etid >> 6;
        const int P = 32 * nsteps + e;
        float sum = 0.f;
        if (e == 0 && P >= 1) sum += ring[(1 * RF + (P - 1) % RF) * 64 + gslot];
        if (e <= 1 && P >= 2) sum += ring[(2 * RF + (P - 2) % RF) * 64 + gslot];
        if (P >= 3) sum += ring[(3 * RF + (P - 3) % RF) * 64 + gslot];
        emit(P, sum);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tile_phase ^= 1u;
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace avsep
