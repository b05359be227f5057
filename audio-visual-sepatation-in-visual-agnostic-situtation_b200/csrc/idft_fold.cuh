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
// Overlap-add in registers.  Padded output position 256 t + n; output block P = samples [256 P, 256 P + 256).  Row n of
// frame t feeds  a = Pw + Qw  into block t + (n >= 256) at sample j = n mod 256, and  b = Pw - Qw  into block t + 2
// (n >= 255) or t + 3 at sample (254 - j) mod 256:
//   out[P][j] = a(j, P) + a(j + 256, P - 1) + b(254 - j + 256, P - 2) + b(254 - j, P - 3).
// The basis rows are ORDERED so that these four rows are TMEM lanes of the same warp: lane = 16 h + 8 m + i holds row
// n = j + 256 h with j = 32 c + 8 q + i (m = 0) or its mirror 254 - j (m = 1), for CTA group c (0..3) and lane group q.
// A thread holds its row's a / b for 32 consecutive frames in registers, so the frame offsets are register-index
// offsets and the row exchange is three warp shuffles per frame (xor 16: other h, xor 8 / 24: mirror) -- no shared
// memory, no barriers, fixed summation order (bit-reproducible).  The self-mirrored samples 127 and 255 share the
// last lane octet of group (c = 3, q = 3), which re-routes its mirrored terms with two extra shuffles.  Then x
// reciprocal window-sum-square envelope, trim n_fft / 2, clip to [-1, 1], store.
//
// Epilogue / MMA overlap with all 512 TMEM columns in use: the overlap-add is linear, out = OLA(Pw) + OLA(+-Qw), so the
// two accumulators are drained SEPARATELY.  While the Qw MMAs of tile i run, the epilogue reads Pw of tile i and keeps
// its partial block sums in registers (64 per thread); while the Pw MMAs of tile i + 1 run, it reads Qw of tile i, adds
// the partials and stores.  Each accumulator has its own full / empty barrier pair, so the tensor pipe never waits for
// the epilogue as long as half an epilogue fits in half a main loop.
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

template <int BK_>
struct IFoldCfgT {
  static_assert(BK_ == 32 || BK_ == 64, "K block is 32 (SWIZZLE_64B) or 64 (SWIZZLE_128B)");
  static constexpr int BK = BK_;
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = 128 * ROW_BYTES;  // one split of this CTA's 128 basis rows
  static constexpr int B_BYTES = 128 * ROW_BYTES;  // one split of this CTA's 128 frames
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;     // 32 KB
  static constexpr int STAGES = (IFOLD_SMEM_LIMIT - 1024 - 256) / STAGE_BYTES;    // 7 (BK = 32) or 3 (BK = 64)
  static constexpr int KB_HALF = IFOLD_K / BK;     // K blocks per accumulator (Pw, then Qw)
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;
  static_assert(SMEM_BYTES <= IFOLD_SMEM_LIMIT, "shared memory budget");
  static_assert((2 * STAGES + 5) * 8 <= 256, "barrier block overflow");
};
using IFoldCfg = IFoldCfgT<64>;

struct IFoldParams {
  int num_signals;
  int T;                  // frames per signal (<= 256)
  int lo_offset;          // plane index of the lo split of the operand
  float* wav;             // [S, out_len]
  int out_len;            // 256 (T - 1)
  const float* inv_env;   // [256 (T + 3)]
  int* err_flag;
};

// sample j (0..255) of the output blocks that TMEM lane l = 32 q + 16 h + 8 m + i of CTA group c4 works on:
// m = 0: j = 32 c4 + 8 q + i, m = 1: its mirror 254 - j; sample 127 mirrors onto itself, so its m = 1 lane takes
// the other self-mirrored sample, 255
__host__ __device__ inline int ifold_sample(int c4, int l) {
  const int q = l >> 5, m = (l >> 3) & 1, i = l & 7;
  const int jf = 32 * c4 + 8 * q + i;
  if (!m) return jf;
  return jf == 127 ? 255 : 254 - jf;
}
// folded row n of basis row R = 128 c4 + l
__host__ __device__ inline int ifold_row(int R) {
  const int c4 = R >> 7, l = R & 127;
  return ifold_sample(c4, l) + 256 * ((l >> 4) & 1);
}

// Overlap-add of one accumulator (Pw: QPHASE = false, Qw: QPHASE = true) of one tile for one epilogue warp: frames
// [128 HALF, 128 HALF + 128) of its 32 rows, straight-line code (a tcgen05.ld writes its registers asynchronously until
// tcgen05.wait::ld, so no loop-carried copy of them may exist; and without branches in the per-frame body the compiler
// interleaves the shuffles of many frames).
//   block P, sample j:  x(j, P) + x(j + 256, P - 1) +- [ x(mirror + 256, P - 2) + x(mirror, P - 3) ]      (+ for Pw, - for Qw)
// The two lanes of a sample (h = 0: row j, h = 1: row j + 256) own alternate blocks: even blocks are summed by the
// h = 0 lane, odd blocks by the h = 1 lane, so each thread carries 64 partial sums from the Pw pass to the Qw pass.
// SPECIAL: the warp whose last lane octet holds the self-mirrored samples 127 (m = 0) and 255 (m = 1): the mirrored
// term of row 127 -> sample 127 (block t + 3), of row 383 -> sample 127 (t + 2), of row 255 -> sample 255 (t + 2), and
// that of row 511 is the direct term itself (dropped); two extra shuffles re-route them.
template <bool SPECIAL, int HALF, bool QPHASE>
__device__ __forceinline__ void ifold_overlap_add(uint32_t taddr, int lane, int j, int T, int out_len, float* __restrict__ wav,
                                                  const float* __restrict__ inv_env, uint32_t lead_tempty, float (&part)[65]) {
  constexpr int F0 = 128 * HALF;
  const bool hb = (lane & 16) != 0;
  const bool sp = SPECIAL && (lane & 7) == 7;
  const bool sp_a0 = sp && (lane & 24) == 0, sp_a1 = sp && (lane & 24) == 16;
  // The window sum-square envelope of a block that all four overlapping frames P-3..P reach (3 <= P <= T-1) depends on
  // the sample j only (same four addends, same order on the host); block T is the only later one that reaches the
  // output (256 (T - 1) samples after the trim): two loads per tile instead of one per output.
  float env_mid = 0.f, env_T = 0.f;
  if (QPHASE) {
    env_mid = __ldg(inv_env + 256 * min(3, T - 1) + j);
    env_T = __ldg(inv_env + 256 * T + j);
  }
  float* const out = wav + (j - AVSEP_N_FFT / 2);  // out[256 P] = sample j of block P (dereferenced only inside the clip)
  // look-back terms of the previous frames: own value, partner row (lane ^ 16), mirror rows (lanes ^ 8, ^ 24)
  float X1 = 0.f, XA1 = 0.f, M8_1 = 0.f, M8_2 = 0.f, M8_3 = 0.f, M24_1 = 0.f, M24_2 = 0.f, M24_3 = 0.f;
  auto exchange = [&](float x, float& xa, float& m8, float& m24) {
    float xs = x;
    if (SPECIAL) {
      const float s8 = __shfl_xor_sync(0xffffffffu, x, 8), s24 = __shfl_xor_sync(0xffffffffu, x, 24);
      xs = sp ? (sp_a0 ? 0.f : (sp_a1 ? s24 : s8)) : x;
    }
    xa = __shfl_xor_sync(0xffffffffu, x, 16);
    m8 = __shfl_xor_sync(0xffffffffu, xs, 8);
    m24 = __shfl_xor_sync(0xffffffffu, xs, 24);
  };
  auto push = [&](float x, float xa, float m8, float m24) {
    X1 = x; XA1 = xa;
    M8_3 = M8_2; M8_2 = M8_1; M8_1 = m8;
    M24_3 = M24_2; M24_2 = M24_1; M24_1 = m24;
  };
  // block sum owned by this lane for the frame with compile-time parity `odd`
  auto block_sum = [&](bool odd, float x, float xa) {
    // even (h = 0 lane = row j):       x(P) + partner(P-1) +- (mirror+256 = lane^24 (P-2) + mirror = lane^8 (P-3))
    // odd  (h = 1 lane = row j + 256): partner(P) + x(P-1) +- (lane^8 (P-2) + lane^24 (P-3))
    const float direct = odd ? (xa + X1) : (x + XA1);
    const float mirror = odd ? (M8_2 + M24_3) : (M24_2 + M8_3);
    return QPHASE ? direct - mirror : direct + mirror;
  };
  constexpr uint32_t COL0 = QPHASE ? IFOLD_TILE_N : 0;
  uint32_t v[32];
  if (HALF) {                                      // look-back frames 125..127 come from the other half's columns
    uint32_t cv[4];
    tmem_ld_32x32b_x4(taddr + COL0 + F0 - 4, cv);
    tmem_ld_32x32b_x32(taddr + COL0 + F0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 1; i < 4; ++i) {
      const float x = __uint_as_float(cv[i]);
      float xa, m8, m24;
      exchange(x, xa, m8, m24);
      push(x, xa, m8, m24);
    }
  } else {
    tmem_ld_32x32b_x32(taddr + COL0 + F0, v);
    tmem_ld_wait();
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int c0 = F0 + 32 * c;
    float fx[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) fx[i] = __uint_as_float(v[i]);
    if (c < 3) {                                   // the next chunk's TMEM load flies while this one is exchanged
      tmem_ld_32x32b_x32(taddr + COL0 + c0 + 32, v);
    } else {
      tcgen05_fence_before();                      // accumulator read out: the MMAs of the next tile may overwrite it
      mbar_arrive_cluster_relaxed(lead_tempty);
    }
    // interior chunk (warp-uniform): every block has the periodic envelope and lies inside the clip
    const bool interior = (c0 >= 3) && (c0 + 32 <= T);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int k = (32 * c + i) >> 1;             // this thread's partial-sum slot of the block pair (i, i + 1)
      float xa0, m8_0, m24_0, xa1, m8_1, m24_1;
      exchange(fx[i], xa0, m8_0, m24_0);
      const float g_even = block_sum(false, fx[i], xa0);
      push(fx[i], xa0, m8_0, m24_0);
      exchange(fx[i + 1], xa1, m8_1, m24_1);
      const float g_odd = block_sum(true, fx[i + 1], xa1);
      push(fx[i + 1], xa1, m8_1, m24_1);
      const float g = hb ? g_odd : g_even;         // the block this lane owns
      if (!QPHASE) {
        part[k] = g;
      } else {
        const int P = c0 + i + (hb ? 1 : 0);
        const float sum = part[k] + g;
        if (interior) {
          out[256 * P] = fminf(fmaxf(sum * env_mid, -1.0f), 1.0f);
        } else {
          const float env = P < 3 ? __ldg(inv_env + 256 * P + j) : (P >= T ? env_T : env_mid);
          const int idx = 256 * P + j - AVSEP_N_FFT / 2;
          if ((unsigned)idx < (unsigned)out_len) out[256 * P] = fminf(fmaxf(sum * env, -1.0f), 1.0f);
        }
      }
    }
    if (c < 3) tmem_ld_wait();
  }
  if (HALF) {                                      // block 256 (reaches the output when T = 256): look-back terms only, even -> h = 0 lane
    const float g = QPHASE ? XA1 - (M24_2 + M8_3) : XA1 + (M24_2 + M8_3);
    if (!QPHASE) {
      part[64] = g;
    } else {
      const int idx = 256 * 256 + j - AVSEP_N_FFT / 2;
      if (!hb && (unsigned)idx < (unsigned)out_len) out[256 * 256] = fminf(fmaxf((part[64] + g) * env_T, -1.0f), 1.0f);
    }
  }
}

// basis tensor map: rows = (kind * 2 + split) * 512 + R, cols = f (512)          kind 0 = cos (Pw), 1 = -sin (Qw)
// operand tensor map: dims (k = 1024 = [Re | Im] x 512 bins, t = T, plane = split * lo_offset + s)
template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(IFOLD_THREADS, 1)
idft_fold2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const IFoldParams p) {
  using Cfg = IFoldCfgT<BK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  // one full / empty pair per accumulator: kind 0 = Pw (TMEM columns [0, 256)), 1 = Qw ([256, 512))
  auto tfull_bar = [&](int kind) { return bar_base + 8u * (2 * Cfg::STAGES + kind); };
  auto tempty_bar = [&](int kind) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + kind); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::BAR_OFFSET + 8 * (2 * Cfg::STAGES + 4));

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
    for (int kind = 0; kind < 2; ++kind) {
      mbar_init(tfull_bar(kind), 1);
      mbar_init(tempty_bar(kind), 512);           // 256 epilogue threads of each CTA of the pair
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
        for (int kb = 0; kb < 2 * Cfg::KB_HALF; ++kb) {
          const int kind = kb / Cfg::KB_HALF;
          const uint32_t d_tmem = tmem_base + (uint32_t)(kind * IFOLD_TILE_N);
          const bool first = (kb % Cfg::KB_HALF) == 0;
          if (first) {                            // the epilogue has read this accumulator of the previous tile
            mbar_wait(tempty_bar(kind), tile_phase ^ 1u, p.err_flag, 52);
            tcgen05_fence_after();
          }
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
            // hi * lo and hi * hi share A = hi: read from shared memory once (operand collector), then lo * hi
            umma_f16_2sm_keep_a(d_tmem, a_hi + adv, b_lo + adv, idesc, (first && k16 == 0) ? 0u : 1u);
            umma_f16_2sm_reuse_a(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
          }
          umma_commit_2sm(empty_bar(stage), 3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          if ((kb % Cfg::KB_HALF) == Cfg::KB_HALF - 1) umma_commit_2sm(tfull_bar(kind), 3);   // this accumulator is complete
        }
        tile_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue (8 warps per CTA): overlap-add in registers
    const int q = warp & 3;                        // TMEM lane group this warp may touch
    const int half = (warp - 2) >> 2;              // frames [128 half, 128 half + 128)
    const uint32_t lead_tempty_p = mapa_shared(tempty_bar(0), 0), lead_tempty_q = mapa_shared(tempty_bar(1), 0);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t tile_phase = 0;
    for (int tile = pair; tile < pair_tiles; tile += npairs) {
      const int c4 = 2 * (tile & 1) + (int)rank;
      const int s = tile >> 1;
      float* wav = p.wav + (size_t)s * p.out_len;
      const int j = ifold_sample(c4, 32 * q + lane);
      const bool special_warp = (c4 == 3 && q == 3);
      float part[65];
      // Columns >= T are zero (TMA zero-fills out-of-range frames) and outputs beyond the clip are dropped, so ragged
      // clips need no special case.
      mbar_wait(tfull_bar(0), tile_phase, p.err_flag, 54);
      tcgen05_fence_after();
      if (half == 0) {
        if (special_warp) ifold_overlap_add<true, 0, false>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_p, part);
        else ifold_overlap_add<false, 0, false>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_p, part);
      } else {
        if (special_warp) ifold_overlap_add<true, 1, false>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_p, part);
        else ifold_overlap_add<false, 1, false>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_p, part);
      }
      mbar_wait(tfull_bar(1), tile_phase, p.err_flag, 55);
      tcgen05_fence_after();
      if (half == 0) {
        if (special_warp) ifold_overlap_add<true, 0, true>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_q, part);
        else ifold_overlap_add<false, 0, true>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_q, part);
      } else {
        if (special_warp) ifold_overlap_add<true, 1, true>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_q, part);
        else ifold_overlap_add<false, 1, true>(taddr, lane, j, p.T, p.out_len, wav, p.inv_env, lead_tempty_q, part);
      }
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
