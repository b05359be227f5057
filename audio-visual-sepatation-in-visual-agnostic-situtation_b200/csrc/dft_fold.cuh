// Forward STFT as a FOLDED split-precision tcgen05 GEMM: half the tensor work of the plain GEMM in dft_gemm.cuh.
//
// For a real frame x[0..1021] and the symmetric periodic Hann window w (w[n] = w[1022 - n]):
//   Re X[f] =  sum_{n=0}^{511} w[n] cos(2 pi f n / 1022) * E[n],   E[n] = x[n] + x[1022 - n]  (E[0] = x[0], E[511] = x[511])
//   Im X[f] = -sum_{n=1}^{510} w[n] sin(2 pi f n / 1022) * O[n],   O[n] = x[n] - x[1022 - n]  (O[0] = O[511] = 0)
// so the 1024 x 1024 x T contraction per signal becomes two 512 x 512 x T ones (cos rows x E, sin rows x O).
//
// Two kernels, both CTA pairs (cta_group::2) that GENERATE the E / O operand themselves from the fp32 waveform:
//   dft_fold3_fwd_kernel   any batch of signals (eval step, stft / stft_mix with phase or (re, im) outputs)
//   dft_fold3m_fwd_kernel  mixtures formed on the device from their 2..4 sources: the training / eval step (mixture spectrum = sum Xn)
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace avsep {

constexpr int FOLD_K = 512;            // folded samples per frame
constexpr int FOLD_BINS = 512;
constexpr int FOLD_TILE_M = 128;
constexpr int FOLD_TILE_N = 128;
constexpr int FOLD_M_TILES = FOLD_BINS / FOLD_TILE_M;   // 4
constexpr int FOLD_THREADS = 320;
constexpr int FOLD_SMEM_LIMIT = 232448;

struct FoldParams {
  int num_signals;
  int T;
  int n_tiles;       // ceil(T / 128)
  int total_tiles;   // S * n_tiles * 4
  int s_cap;         // signals per (kind, split) plane of the staged operand
  int sig_base;      // global index of the first signal of this launch (chunked launches reuse the staging slots)
  float* mag;        // [S, 512, T]
  float* phase;      // optional
  float* re;         // optional (with im), written for signals s < reim_count
  float* im;
  int reim_count;
  int prefetch;      // L2-prefetch the next tile's operand while loading the current one
  int* err_flag;
  const float* wave; // [S, L] fp32 waveforms (the E / O operand is generated inside the kernel)
  int L;
  // device-summed mixtures (dft_fold3): signals s < mix_B are read from wave_mix [mix_B, L], signal s >= mix_B is
  // wave[s - mix_B] (the sources); mix_B = 0: every signal comes from `wave`
  const float* wave_mix;
  int mix_B;
};

// waveform of signal s of a dft_fold3 launch
__device__ __forceinline__ const float* fold3_signal(const FoldParams& p, int s) {
  return s < p.mix_B ? p.wave_mix + (size_t)s * p.L : p.wave + (size_t)(s - p.mix_B) * p.L;
}

constexpr int FOLD2_TILE_N = 256;      // frames per pair tile of dft_fold3

// ======================================================================================================
// dft_fold3: the CTA-pair folded GEMM with the E / O operand GENERATED IN THE KERNEL from the raw waveform.
//
// Staging E[n] = x[n] + x[1022 - n], O[n] = x[n] - x[1022 - n] (bf16 hi / lo) in HBM would cost 1 MB written and read back
// per signal, 4x the waveform itself, because folded frames no longer overlap (round 1 measured that variant: 0.25 ms for
// 240 signals against 0.20 ms here).  Instead NPW producer warps per CTA read the fp32 waveform (63 MB for 240 signals: it
// stays in L2), fold, split and store straight into the swizzled shared-memory tile the MMA reads.
//
// Stage (64 KB, K block 32, SWIZZLE_64B, both kinds so one waveform load feeds E and O):
//   [cos hi | cos lo | -sin hi | -sin lo]  4 x 8 KB, this CTA's 128 bins, by TMA from the basis
//   [E hi   | E lo   |  O hi   |  O lo  ]  4 x 8 KB, this CTA's 128 frames, by the producer warps
// full barrier (leader CTA): 1 expect_tx arrival (TMA bytes of both CTAs) + one arrival per producer warp of both
// CTAs, each after fence.proxy.async so the generic-proxy stores are visible to the tensor core's async proxy.
//   warp 0: TMA (basis) | warp 1: MMA issuer | warps 2..9: epilogue | warps 10..10+NPW-1: operand producers
// ======================================================================================================
constexpr int FOLD_XT_FLOATS = 32 * 17;   // per-warp transpose tile of the epilogue: 32 bins x (16 frames + 1 pad)

struct Fold3Cfg {
  static constexpr int BK = 32;
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int SUB_BYTES = FOLD_TILE_M * ROW_BYTES;          // 8 KB: 128 rows x 32 bf16
  static constexpr int STAGE_BYTES = 8 * SUB_BYTES;
  static constexpr int XPOSE_BYTES = 8 * FOLD_XT_FLOATS * 4;          // one [32][17] transpose tile per epilogue warp
  static constexpr int STAGES = (FOLD_SMEM_LIMIT - 1024 - 256 - XPOSE_BYTES) / STAGE_BYTES;
  static constexpr int KSTEPS = FOLD_K / BK;                         // 16
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + XPOSE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;
  static_assert(STAGES >= 2 && (2 * STAGES + 3) * 8 <= 256, "barrier block overflow");
};

// A waveform load the compiler may neither hoist, sink nor speculate: the producers rely on the loads of one K step
// being issued back to back (no live index registers, no spills), and on nothing consuming them before the convert.
__device__ __forceinline__ float fold3_ldg(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// centre reflect padding of sample index j: -j for j < 0, 2 (L - 1) - j for j >= L (one reflection suffices: L >= 512)
__device__ __forceinline__ int fold3_reflect(int j, int L) {
  j = abs(j);
  return max(min(j, 2 * (L - 1) - j), 0);
}

// |re + i im| with MUFU.SQRT (sqrt.approx: max relative error 2^-23, i.e. within 1 ulp): the correctly rounded sqrtf costs ~18
// instructions, and the epilogue -- which cannot overlap the MMAs because the accumulators fill all 512 TMEM columns -- was
// instruction-bound on it (clock stamps: 24.5 K cycles of epilogue after 25.5 K cycles of MMAs per tile)
__device__ __forceinline__ float fold_mag(float re, float im) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(re, re, im * im)));
  return r;
}

// A lane owns one bin (TMEM lane) and holds 16 consecutive frames of it; written as they are, the 32 lanes of a store instruction
// would touch 32 different cache lines (the L1 tag stage pays per line: measured 12 K cycles per tile).  Each warp therefore
// transposes its 32 bins x 16 frames block through a private padded shared-memory tile, so that one store instruction writes two
// 64-byte row segments.  xt: this warp's [32][17] floats.
// TC: frames per clip when known at compile time (256 for the 65 535-sample clips of every BASELINE config), 0 = use T.  With
// a runtime T every store needs its own 64-bit row * T multiply-add -- ncu's source view of round 2 showed 5.3 IMAD + 2 LEA
// per STG, 43 % of the epilogue's instructions; with TC the 16 row addresses are one base pointer plus immediates.
template <int TC>
__device__ __forceinline__ void fold_store_block(float* __restrict__ xt, int lane, const float (&v)[16], float* __restrict__ plane_row0,
                                                 int T, int t0) {
  const int Ts = TC ? TC : T;
#pragma unroll
  for (int j = 0; j < 16; ++j) xt[lane * 17 + j] = v[j];
  __syncwarp();
  const int col = lane & 15, sub = lane >> 4;
  if (t0 + col < Ts) {
    float* q = plane_row0 + (size_t)sub * Ts + t0 + col;
    const float* x = xt + sub * 17 + col;
#pragma unroll
    for (int r = 0; r < 32; r += 2) __stcs(q + (size_t)r * Ts, x[r * 17]);
  }
  __syncwarp();
}

// Waveform loads of one K step of one producer warp.  Lane (sub, c) owns samples n0 = 32 ks + 2 c and n0 + 1 of the
// frames t_even + 2 i + sub: forward terms x[n0], x[n0 + 1] and mirrored terms x[1022 - n0], x[1021 - n0].
// Interior frames (bit i of lin_mask) use two base pointers plus compile-time offsets; the first / last two frames
// of a clip take the reflecting index path.
template <int ITER>
__device__ __forceinline__ void fold3_load(const float* __restrict__ w, int L, int T, int tbase, int n0, uint32_t lin_mask,
                                           float (&a0)[ITER], float (&a1)[ITER], float (&b0)[ITER], float (&b1)[ITER]) {
  const int jf0 = tbase * AVSEP_HOP - AVSEP_N_FFT / 2 + n0;     // x[n0] of frame tbase in signal coordinates
  const int jm0 = tbase * AVSEP_HOP + AVSEP_N_FFT / 2 - n0;     // x[1022 - n0]
  const float* pf = w + jf0;                                    // only dereferenced where in range
  const float* pm = w + jm0;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    if (lin_mask & (1u << i)) {                                 // warp-uniform, computed once per tile
      const int d = 2 * i * AVSEP_HOP;
      a0[i] = fold3_ldg(pf + d);
      a1[i] = fold3_ldg(pf + d + 1);
      b0[i] = fold3_ldg(pm + d);
      b1[i] = fold3_ldg(pm + d - 1);
    } else {
      const int d = (min(tbase + 2 * i, T - 1) - tbase) * AVSEP_HOP;     // frames past T are masked by the convert
      a0[i] = fold3_ldg(w + fold3_reflect(jf0 + d, L));
      a1[i] = fold3_ldg(w + fold3_reflect(jf0 + d + 1, L));
      b0[i] = fold3_ldg(w + fold3_reflect(jm0 + d, L));
      b1[i] = fold3_ldg(w + fold3_reflect(jm0 + d - 1, L));
    }
  }
}

constexpr int FOLD3_THREADS = 640;      // 5 warpgroups: {TMA, MMA, 2 idle} | epilogue | epilogue | producers | producers
constexpr int FOLD3_PW = 8;             // producer warps per CTA

static_assert(4 * 48 + 8 * 112 + 8 * 104 <= (FOLD3_THREADS / 32) * 96, "setmaxnreg targets exceed the CTA's register pool");
template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(REGS)); }

// NG = producer groups.  NG = 1: all 8 producer warps work on every K step (16 frames each).  NG = 2: warps 0..3
// convert the even K steps and warps 4..7 the odd ones (32 frames each), so a group has two MMA step times for
// "loads in flight -> convert -> store".  (More loads in flight per THREAD do not help: the fence.proxy.async that
// publishes a stage also waits for the thread's outstanding global loads -- measured.)
// A 640-thread CTA owns 640 x 96 registers and setmaxnreg only moves registers inside that pool (an .inc the pool
// cannot satisfy blocks forever): warpgroup 0 shrinks to 48 per thread, the two epilogue warpgroups grow to 112 and
// the two producer warpgroups to 104  (4 x 48 + 8 x 112 + 8 x 104 = 1920 = 20 warps x 96).
template <int NG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FOLD3_THREADS, 1)
dft_fold3_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const FoldParams p) {
  using Cfg = Fold3Cfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* xpose_all = reinterpret_cast<float*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES);
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
  const int n_tiles = (p.T + FOLD2_TILE_N - 1) / FOLD2_TILE_N;
  const int pair_tiles = p.num_signals * n_tiles * (FOLD_M_TILES / 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1 + 2 * (FOLD3_PW / NG));   // TMA expect_tx arrival + the producer warps of the step's group, both CTAs
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 512);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp < 4) {
  // warpgroup 0 (TMA warp, MMA warp, two idle warps) hands registers to the producer warpgroup
  setmaxnreg_dec<48>();
  if (warp == 0) {
    // ===================================== TMA producer: basis rows (both CTAs) ================
    // All 32 lanes first pull the waveform segment of a tile's 128 frames (135 KB, contiguous) into L2, one tile
    // ahead of the producer warps, so their loads never wait on DRAM.
    auto prefetch_tile = [&](int tile) {
      if (!p.prefetch || tile >= pair_tiles) return;
      const int rest = tile / (FOLD_M_TILES / 2);
      const int nt = rest % n_tiles;
      const int s = rest / n_tiles + p.sig_base;
      const float* w = fold3_signal(p, s);
      const int t_first = nt * FOLD2_TILE_N + (int)rank * (FOLD2_TILE_N / 2);
      const int lo = max(t_first * AVSEP_HOP - AVSEP_N_FFT / 2, 0);
      const int hi = min((t_first + FOLD2_TILE_N / 2 - 1) * AVSEP_HOP + AVSEP_N_FFT / 2, p.L - 1);
      const char* b = reinterpret_cast<const char*>(w + lo);
      const char* e = reinterpret_cast<const char*>(w + hi + 1);
      b -= reinterpret_cast<uintptr_t>(b) & 127u;
      for (const char* q = b + lane * 128; q < e; q += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
    };
    prefetch_tile(pair);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < pair_tiles; tile += npairs) {
      prefetch_tile(tile + npairs);
      if (lane == 0) {
        const int mp = tile % (FOLD_M_TILES / 2);
        const int row0 = (2 * mp + (int)rank) * FOLD_TILE_M;
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 41);
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * 4 * Cfg::SUB_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma2_load_2d(&tmA, lead_full, st + j * Cfg::SUB_BYTES, ks * Cfg::BK, j * FOLD_BINS + row0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =======================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * FOLD_TILE_M, FOLD2_TILE_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_phase = 0;
      for (int tile = pair; tile < pair_tiles; tile += npairs) {
        mbar_wait(tempty_bar, tile_phase ^ 1u, p.err_flag, 42);
        tcgen05_fence_after();
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 43);
          tcgen05_fence_after();
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
#pragma unroll
          for (int kind = 0; kind < 2; ++kind) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(kind * FOLD2_TILE_N);
            const uint64_t a_hi = umma_smem_desc(st + (2 * kind) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
            const uint64_t a_lo = umma_smem_desc(st + (2 * kind + 1) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
            const uint64_t b_hi = umma_smem_desc(st + (4 + 2 * kind) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
            const uint64_t b_lo = umma_smem_desc(st + (5 + 2 * kind) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
#pragma unroll
            for (int k16 = 0; k16 < Cfg::BK / 16; ++k16) {
              const uint64_t adv = (uint64_t)(k16 * 2);
              // hi * lo and hi * hi share A = hi: read from shared memory once (operand collector), then lo * hi
              umma_f16_2sm_keep_a(d_tmem, a_hi + adv, b_lo + adv, idesc, (ks == 0 && k16 == 0) ? 0u : 1u);
              umma_f16_2sm_reuse_a(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
              umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
            }
          }
          umma_commit_2sm(empty_bar(stage), 3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(tfull_bar, 3);
        tile_phase ^= 1u;
      }
    }
    __syncwarp();
  }
  } else if (warp >= 12) {
    // ===================================== operand producers: waveform -> E / O hi / lo tiles ==
    constexpr int WPG = FOLD3_PW / NG;                 // warps per group
    constexpr int RPW = (FOLD2_TILE_N / 2) / WPG;      // frames (rows) per producer warp: 16 or 32
    constexpr int ITER = RPW / 2;                      // a warp covers 2 rows x 32 samples per iteration
    setmaxnreg_inc<104>();
    const int grp = (warp - 12) / WPG;
    const int pw = (warp - 12) % WPG;
    const int sub = lane >> 4, c = lane & 15;
    const int r0 = pw * RPW;
    const int my_tiles = pair < pair_tiles ? (pair_tiles - pair + npairs - 1) / npairs : 0;
    const int steps = my_tiles * Cfg::KSTEPS;
    float a0[ITER], a1[ITER], b0[ITER], b1[ITER];
    const float* w = p.wave;
    int tbase = 0, tile_i = -1;
    uint32_t lin_mask = 0;
    auto issue_loads = [&](int g, float (&x0)[ITER], float (&x1)[ITER], float (&y0)[ITER], float (&y1)[ITER], int& tb) {
      const int ks = g % Cfg::KSTEPS;                  // g = flat (tile, k step) index of this pair, increasing
      if (g / Cfg::KSTEPS != tile_i) {
        tile_i = g / Cfg::KSTEPS;
        const int tile = pair + tile_i * npairs;
        const int rest = tile / (FOLD_M_TILES / 2);
        const int nt = rest % n_tiles;
        const int s = rest / n_tiles + p.sig_base;
        w = fold3_signal(p, s);
        tbase = nt * FOLD2_TILE_N + (int)rank * (FOLD2_TILE_N / 2) + r0 + sub;
        // bit i: both frames of iteration i lie in [2, (L - 512) / 256], where no sample index leaves the signal
        lin_mask = 0;
        const int t_even = tbase - sub;
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
          const int t = t_even + 2 * i;
          if (t >= 2 && (t + 1) * AVSEP_HOP + AVSEP_N_FFT / 2 <= p.L - 1 && t + 1 < p.T) lin_mask |= 1u << i;
        }
      }
      tb = tbase;
      fold3_load<ITER>(w, p.L, p.T, tbase, ks * Cfg::BK + 2 * c, lin_mask, x0, x1, y0, y1);
    };
    // fold + split the loaded samples of step g and write the bf16 pairs into the step's stage
    auto convert_store = [&](int g, const float (&x0)[ITER], const float (&x1)[ITER], const float (&y0)[ITER],
                             const float (&y1)[ITER], int tb) {
      const int stage = g % Cfg::STAGES;
      const uint32_t phase = (uint32_t)(g / Cfg::STAGES) & 1u;
      const int n0 = (g % Cfg::KSTEPS) * Cfg::BK + 2 * c;
      const bool pair0 = n0 >= 1, pair1 = n0 + 1 <= AVSEP_N_FFT / 2 - 1;
      mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 44);   // the MMAs that read the stage's previous contents are done
      uint8_t* bt = smem_gen + stage * Cfg::STAGE_BYTES + 4 * Cfg::SUB_BYTES;
      // Fast path (14 of 16 K steps of a full tile): every frame exists and every sample has its mirror partner, so
      // no masking.  The generic path handles K step 0 (n = 0 has no partner), K step 15 (n = 511) and ragged tiles.
      const bool plain = (n0 - 2 * c) != 0 && (n0 - 2 * c) != (Cfg::KSTEPS - 1) * Cfg::BK && tb - sub + RPW <= p.T;
      auto emit = [&](int i, float E0, float E1, float O0, float O1) {
        const int r = r0 + 2 * i + sub;
        // SWIZZLE_64B: the 16-byte chunk index (address bits 4..5) is XORed with address bits 7..8 = (r >> 1) & 3
        const int off = r * Cfg::ROW_BYTES + ((4 * c) ^ ((((r0 >> 1) + i) & 3) << 4));
        const __nv_bfloat162 eh = __floats2bfloat162_rn(E0, E1);
        const float2 ehf = __bfloat1622float2(eh);
        const __nv_bfloat162 el = __floats2bfloat162_rn(E0 - ehf.x, E1 - ehf.y);
        const __nv_bfloat162 oh = __floats2bfloat162_rn(O0, O1);
        const float2 ohf = __bfloat1622float2(oh);
        const __nv_bfloat162 ol = __floats2bfloat162_rn(O0 - ohf.x, O1 - ohf.y);
        *reinterpret_cast<__nv_bfloat162*>(bt + off) = eh;
        *reinterpret_cast<__nv_bfloat162*>(bt + Cfg::SUB_BYTES + off) = el;
        *reinterpret_cast<__nv_bfloat162*>(bt + 2 * Cfg::SUB_BYTES + off) = oh;
        *reinterpret_cast<__nv_bfloat162*>(bt + 3 * Cfg::SUB_BYTES + off) = ol;
      };
      if (plain) {
#pragma unroll
        for (int i = 0; i < ITER; ++i) emit(i, x0[i] + y0[i], x1[i] + y1[i], x0[i] - y0[i], x1[i] - y1[i]);
      } else {
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
          const bool ok = tb + 2 * i < p.T;            // frames past the end of a ragged clip are zero rows
          const float u0 = ok ? x0[i] : 0.f, u1 = ok ? x1[i] : 0.f;
          const float v0 = (ok && pair0) ? y0[i] : 0.f, v1 = (ok && pair1) ? y1[i] : 0.f;
          emit(i, u0 + v0, u1 + v1, pair0 ? u0 - v0 : 0.f, pair1 ? u1 - v1 : 0.f);
        }
      }
      fence_proxy_async_smem();                        // generic-proxy stores -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(full_bar(stage), 0));
    };
    int tb = 0;
    if (grp < steps) issue_loads(grp, a0, a1, b0, b1, tb);
    for (int g = grp; g < steps; g += NG) {
      convert_store(g, a0, a1, b0, b1, tb);
      // refill AFTER the store: its fence.proxy.async would otherwise wait for these loads
      if (g + NG < steps) issue_loads(g + NG, a0, a1, b0, b1, tb);
    }
  } else {
    // ===================================== epilogue (8 warps per CTA, own 128 bins) ===========
    setmaxnreg_inc<112>();
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t lead_tempty = mapa_shared(tempty_bar, 0);
    // the whole tile loop once per compile-time frame count: TC = 256 (every BASELINE clip) turns the store addresses into
    // base + immediate; TC = 0 is the general path
    auto run = [&](auto tc) {
    constexpr int TC = decltype(tc)::value;
    const int Tt = TC ? TC : p.T;
    uint32_t tile_phase = 0;
    for (int tile = pair; tile < pair_tiles; tile += npairs) {
      const int mp = tile % (FOLD_M_TILES / 2);
      const int rest = tile / (FOLD_M_TILES / 2);
      const int nt = rest % n_tiles;
      const int s = rest / n_tiles;
      mbar_wait(tfull_bar, tile_phase, p.err_flag, 45);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const int f0 = (2 * mp + (int)rank) * FOLD_TILE_M + 32 * q;             // first bin of this warp's 32 TMEM lanes
      const bool want_reim = p.re != nullptr && (s + p.sig_base) < p.reim_count;
      const size_t row_off = ((size_t)(s + p.sig_base) * AVSEP_N_BINS + f0) * (size_t)Tt;
      float* xt = xpose_all + (warp - 4) * FOLD_XT_FLOATS;
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 16) {
        const int t0 = nt * FOLD2_TILE_N + c0;
        if (t0 >= Tt) break;
        uint32_t vr[16], vi[16];
        tmem_ld_32x32b_x16(taddr + c0, vr);
        tmem_ld_32x32b_x16(taddr + FOLD2_TILE_N + c0, vi);
        tmem_ld_wait();
        float val[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) val[j] = fold_mag(__uint_as_float(vr[j]), __uint_as_float(vi[j]));
        fold_store_block<TC>(xt, lane, val, p.mag + row_off, Tt, t0);
        if (p.phase) {
#pragma unroll
          for (int j = 0; j < 16; ++j) val[j] = atan2f(__uint_as_float(vi[j]), __uint_as_float(vr[j]));
          fold_store_block<TC>(xt, lane, val, p.phase + row_off, Tt, t0);
        }
        if (want_reim) {
#pragma unroll
          for (int j = 0; j < 16; ++j) val[j] = __uint_as_float(vr[j]);
          fold_store_block<TC>(xt, lane, val, p.re + row_off, Tt, t0);
#pragma unroll
          for (int j = 0; j < 16; ++j) val[j] = __uint_as_float(vi[j]);
          fold_store_block<TC>(xt, lane, val, p.im + row_off, Tt, t0);
        }
      }
      tcgen05_fence_before();
      mbar_arrive_cluster_relaxed(lead_tempty);
      tile_phase ^= 1u;
    }
    };
    if (p.T == 256) run(std::integral_constant<int, 256>{}); else run(std::integral_constant<int, 0>{});
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}


// ======================================================================================================
// dft_fold3m: dft_fold3 for a MIXTURE THAT IS THE SUM OF ITS NSRC SOURCES (dataset/music.py:127 builds every training and
// evaluation mixture that way).  The STFT is linear, so the mixture spectrum is X0 + .. + X(NSRC-1) and only the NSRC source
// transforms are computed: 1 / (NSRC + 1) less tensor work, waveform traffic and operand conversion.
// Pair tile = 256 bins x TILE_F frames of ALL sources of one mixture.  The B tiles of a stage hold the CTA's FPC frames of
// source 0 in rows [0, FPC), of source 1 in [FPC, 2 FPC), ... (same 64 KB stage, same producer code as dft_fold3), one
// N = 2 NSRC FPC MMA per product covers all sources, so the Re (TMEM columns [0, N)) and Im ([256, 256 + N)) accumulators
// hold [CTA0: source 0 | source 1 | .. | CTA1: source 0 | ..] x FPC frames, and the epilogue writes the magnitude planes
// |sum Xn|, |X0|, |X1|, .. -- the [mixture | sources] layout the rest of the step reads -- plus, on request (eval), (re, im)
// of the mixture spectrum for the inverse transform.
//   NSRC = 2: FPC = 64 frames per CTA and source, N = 256 (the training step);  NSRC = 3: FPC = 32, N = 192;  NSRC = 4: FPC = 32, N = 256
// ======================================================================================================
template <int NSRC>
struct Fold3mShape {
  static_assert(NSRC >= 2 && NSRC <= 4, "two to four sources per mixture");
  static constexpr int FPC = NSRC == 2 ? 64 : 32;     // frames per CTA and source
  static constexpr int ROWS = NSRC * FPC;             // operand rows per CTA: 128 / 96 / 128
  static constexpr int TILE_F = 2 * FPC;              // frames per pair tile
  static constexpr int UMMA_N = 2 * ROWS;             // 256 / 192 / 256
  static constexpr int WPG = ROWS / 32;               // producer warps per group: each converts 32 rows = 32 frames of ONE source
  static_assert(UMMA_N % 16 == 0 && UMMA_N <= 256 && 2 * WPG <= FOLD3_PW, "tile shape");
};
constexpr int FOLD3M_IM_COL = 256;                    // TMEM column of the Im accumulator

template <int NG, int NSRC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FOLD3_THREADS, 1)
dft_fold3m_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const FoldParams p) {
  static_assert(NG == 2, "two producer groups alternate K steps");
  using Cfg = Fold3Cfg;
  using Sh = Fold3mShape<NSRC>;
  constexpr int FPC = Sh::FPC, ROWS = Sh::ROWS, TILE_F = Sh::TILE_F, WPG = Sh::WPG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* xpose_all = reinterpret_cast<float*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES);
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
  const int n_tiles = (p.T + TILE_F - 1) / TILE_F;
  const int nmix = p.num_signals;                  // mixtures; signal planes of p.mag: [mixture | source 0 | source 1 | ..] x nmix
  const int pair_tiles = p.num_signals * n_tiles * (FOLD_M_TILES / 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1 + 2 * WPG);         // TMA expect_tx arrival + the producer warps of the step's group, both CTAs
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 512);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp < 4) {
  // warpgroup 0 (TMA warp, MMA warp, two idle warps) hands registers to the producer warpgroup
  setmaxnreg_dec<48>();
  if (warp == 0) {
    // ===================================== TMA producer: basis rows (both CTAs) ================
    // All 32 lanes first pull the waveform segments of a tile's frames (contiguous per source) into L2, one tile
    // ahead of the producer warps, so their loads never wait on DRAM.
    auto prefetch_tile = [&](int tile) {
      if (!p.prefetch || tile >= pair_tiles) return;
      const int rest = tile / (FOLD_M_TILES / 2);
      const int nt = rest % n_tiles;
      const int b = rest / n_tiles;
      const int t_first = nt * TILE_F + (int)rank * FPC;
      const int lo = max(t_first * AVSEP_HOP - AVSEP_N_FFT / 2, 0);
      const int hi = min((t_first + FPC - 1) * AVSEP_HOP + AVSEP_N_FFT / 2, p.L - 1);
      for (int src = 0; src < NSRC; ++src) {
        const float* w = p.wave + ((size_t)src * nmix + b) * p.L;
        const char* bb = reinterpret_cast<const char*>(w + lo);
        const char* e = reinterpret_cast<const char*>(w + hi + 1);
        bb -= reinterpret_cast<uintptr_t>(bb) & 127u;
        for (const char* q = bb + lane * 128; q < e; q += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
      }
    };
    prefetch_tile(pair);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < pair_tiles; tile += npairs) {
      prefetch_tile(tile + npairs);
      if (lane == 0) {
        const int mp = tile % (FOLD_M_TILES / 2);
        const int row0 = (2 * mp + (int)rank) * FOLD_TILE_M;
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 41);
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * 4 * Cfg::SUB_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma2_load_2d(&tmA, lead_full, st + j * Cfg::SUB_BYTES, ks * Cfg::BK, j * FOLD_BINS + row0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =======================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * FOLD_TILE_M, Sh::UMMA_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_phase = 0;
      for (int tile = pair; tile < pair_tiles; tile += npairs) {
        mbar_wait(tempty_bar, tile_phase ^ 1u, p.err_flag, 42);
        tcgen05_fence_after();
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 43);
          tcgen05_fence_after();
          const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
          // One MMA per product covers ALL sources: the B tile's rows per CTA are [source 0: FPC frames | source 1 | ..], so
          // accumulator columns come out as [CTA0: src0 | src1 | .. | CTA1: src0 | src1 | ..] x FPC frames.
          // (One MMA per source would read the basis tile from shared memory NSRC times.)
#pragma unroll
          for (int kind = 0; kind < 2; ++kind) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(kind * FOLD3M_IM_COL);              // Re: columns [0, N), Im: [256, 256 + N)
            const uint64_t a_hi = umma_smem_desc(st + (2 * kind) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
            const uint64_t a_lo = umma_smem_desc(st + (2 * kind + 1) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
            const uint64_t b_hi = umma_smem_desc(st + (4 + 2 * kind) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
            const uint64_t b_lo = umma_smem_desc(st + (5 + 2 * kind) * Cfg::SUB_BYTES, Cfg::ROW_BYTES);
#pragma unroll
            for (int k16 = 0; k16 < Cfg::BK / 16; ++k16) {
              const uint64_t adv = (uint64_t)(k16 * 2);
              // hi * lo and hi * hi share A = hi: read from shared memory once (operand collector), then lo * hi
              umma_f16_2sm_keep_a(d_tmem, a_hi + adv, b_lo + adv, idesc, (ks == 0 && k16 == 0) ? 0u : 1u);
              umma_f16_2sm_reuse_a(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
              umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
            }
          }
          umma_commit_2sm(empty_bar(stage), 3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(tfull_bar, 3);
        tile_phase ^= 1u;
      }
    }
    __syncwarp();
  }
  } else if (warp >= 12) {
    // ===================================== operand producers: waveform -> E / O hi / lo tiles ==
    // warp (grp, pw): K steps grp, grp + 2, .., rows [32 pw, 32 pw + 32) = 32 frames of source (32 pw) / FPC
    constexpr int RPW = 32;
    constexpr int ITER = RPW / 2;                      // a warp covers 2 rows x 32 samples per iteration
    setmaxnreg_inc<104>();
    const int grp = (warp - 12) / WPG;
    const int pw = (warp - 12) % WPG;
    const int sub = lane >> 4, c = lane & 15;
    const int r0 = pw * RPW;
    const int my_tiles = (pair < pair_tiles && grp < NG) ? (pair_tiles - pair + npairs - 1) / npairs : 0;   // warps beyond 2 WPG: idle
    const int steps = my_tiles * Cfg::KSTEPS;
    float a0[ITER], a1[ITER], b0[ITER], b1[ITER];
    const float* w = p.wave;
    int tbase = 0, tile_i = -1;
    uint32_t lin_mask = 0;
    auto issue_loads = [&](int g, float (&x0)[ITER], float (&x1)[ITER], float (&y0)[ITER], float (&y1)[ITER], int& tb) {
      const int ks = g % Cfg::KSTEPS;                  // g = flat (tile, k step) index of this pair, increasing
      if (g / Cfg::KSTEPS != tile_i) {
        tile_i = g / Cfg::KSTEPS;
        const int tile = pair + tile_i * npairs;
        const int rest = tile / (FOLD_M_TILES / 2);
        const int nt = rest % n_tiles;
        const int b = rest / n_tiles;
        w = p.wave + ((size_t)(r0 / FPC) * nmix + b) * p.L;                                   // rows [n FPC, (n + 1) FPC): source n
        tbase = nt * TILE_F + (int)rank * FPC + r0 % FPC + sub;
        // bit i: both frames of iteration i lie in [2, (L - 512) / 256], where no sample index leaves the signal
        lin_mask = 0;
        const int t_even = tbase - sub;
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
          const int t = t_even + 2 * i;
          if (t >= 2 && (t + 1) * AVSEP_HOP + AVSEP_N_FFT / 2 <= p.L - 1 && t + 1 < p.T) lin_mask |= 1u << i;
        }
      }
      tb = tbase;
      fold3_load<ITER>(w, p.L, p.T, tbase, ks * Cfg::BK + 2 * c, lin_mask, x0, x1, y0, y1);
    };
    // fold + split the loaded samples of step g and write the bf16 pairs into the step's stage
    auto convert_store = [&](int g, const float (&x0)[ITER], const float (&x1)[ITER], const float (&y0)[ITER],
                             const float (&y1)[ITER], int tb) {
      const int stage = g % Cfg::STAGES;
      const uint32_t phase = (uint32_t)(g / Cfg::STAGES) & 1u;
      const int n0 = (g % Cfg::KSTEPS) * Cfg::BK + 2 * c;
      const bool pair0 = n0 >= 1, pair1 = n0 + 1 <= AVSEP_N_FFT / 2 - 1;
      mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 44);   // the MMAs that read the stage's previous contents are done
      uint8_t* bt = smem_gen + stage * Cfg::STAGE_BYTES + 4 * Cfg::SUB_BYTES;
      // Fast path (14 of 16 K steps of a full tile): every frame exists and every sample has its mirror partner, so
      // no masking.  The generic path handles K step 0 (n = 0 has no partner), K step 15 (n = 511) and ragged tiles.
      const bool plain = (n0 - 2 * c) != 0 && (n0 - 2 * c) != (Cfg::KSTEPS - 1) * Cfg::BK && tb - sub + RPW <= p.T;
      auto emit = [&](int i, float E0, float E1, float O0, float O1) {
        const int r = r0 + 2 * i + sub;
        // SWIZZLE_64B: the 16-byte chunk index (address bits 4..5) is XORed with address bits 7..8 = (r >> 1) & 3
        const int off = r * Cfg::ROW_BYTES + ((4 * c) ^ ((((r0 >> 1) + i) & 3) << 4));
        const __nv_bfloat162 eh = __floats2bfloat162_rn(E0, E1);
        const float2 ehf = __bfloat1622float2(eh);
        const __nv_bfloat162 el = __floats2bfloat162_rn(E0 - ehf.x, E1 - ehf.y);
        const __nv_bfloat162 oh = __floats2bfloat162_rn(O0, O1);
        const float2 ohf = __bfloat1622float2(oh);
        const __nv_bfloat162 ol = __floats2bfloat162_rn(O0 - ohf.x, O1 - ohf.y);
        *reinterpret_cast<__nv_bfloat162*>(bt + off) = eh;
        *reinterpret_cast<__nv_bfloat162*>(bt + Cfg::SUB_BYTES + off) = el;
        *reinterpret_cast<__nv_bfloat162*>(bt + 2 * Cfg::SUB_BYTES + off) = oh;
        *reinterpret_cast<__nv_bfloat162*>(bt + 3 * Cfg::SUB_BYTES + off) = ol;
      };
      if (plain) {
#pragma unroll
        for (int i = 0; i < ITER; ++i) emit(i, x0[i] + y0[i], x1[i] + y1[i], x0[i] - y0[i], x1[i] - y1[i]);
      } else {
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
          const bool ok = tb + 2 * i < p.T;            // frames past the end of a ragged clip are zero rows
          const float u0 = ok ? x0[i] : 0.f, u1 = ok ? x1[i] : 0.f;
          const float v0 = (ok && pair0) ? y0[i] : 0.f, v1 = (ok && pair1) ? y1[i] : 0.f;
          emit(i, u0 + v0, u1 + v1, pair0 ? u0 - v0 : 0.f, pair1 ? u1 - v1 : 0.f);
        }
      }
      fence_proxy_async_smem();                        // generic-proxy stores -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(full_bar(stage), 0));
    };
    int tb = 0;
    if (grp < steps) issue_loads(grp, a0, a1, b0, b1, tb);
    for (int g = grp; g < steps; g += NG) {
      convert_store(g, a0, a1, b0, b1, tb);
      // refill AFTER the store: its fence.proxy.async would otherwise wait for these loads
      if (g + NG < steps) issue_loads(g + NG, a0, a1, b0, b1, tb);
    }
  } else {
    // ===================================== epilogue (8 warps per CTA, own 128 bins) ===========
    setmaxnreg_inc<112>();
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t lead_tempty = mapa_shared(tempty_bar, 0);
    // the whole tile loop once per compile-time frame count (see fold_store_block)
    auto run = [&](auto tc) {
    constexpr int TC = decltype(tc)::value;
    const int Tt = TC ? TC : p.T;
    uint32_t tile_phase = 0;
    for (int tile = pair; tile < pair_tiles; tile += npairs) {
      const int mp = tile % (FOLD_M_TILES / 2);
      const int rest = tile / (FOLD_M_TILES / 2);
      const int nt = rest % n_tiles;
      const int b = rest / n_tiles;
      mbar_wait(tfull_bar, tile_phase, p.err_flag, 45);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const int f0 = (2 * mp + (int)rank) * FOLD_TILE_M + 32 * q;             // first bin of this warp's 32 TMEM lanes
      float* xt = xpose_all + (warp - 4) * FOLD_XT_FLOATS;
      // planes of p.mag: [mixture | source 0 | source 1 | ..] x nmix; mixture = |X0 + X1 + ..| (the STFT is linear and the mixture IS
      // the sum of its sources).  Every accumulator value is read from TMEM once per 16-frame step and feeds its source's and
      // the mixture's magnitude (and, for eval, the mixture's (re, im)).
      auto plane = [&](int j) { return p.mag + ((size_t)(j * nmix + b) * AVSEP_N_BINS + f0) * (size_t)Tt; };
      const bool want_reim = p.re != nullptr;
      const size_t reim_off = ((size_t)b * AVSEP_N_BINS + f0) * (size_t)Tt;
      // frames [half FPC, half FPC + FPC) of the pair tile belong to CTA `half`: source n sits in columns half ROWS + n FPC + frame,
      // Im 256 columns after Re
      for (int c0 = 0; c0 < FPC; c0 += 16) {
        const int t0 = nt * TILE_F + half * FPC + c0;
        if (t0 >= Tt) break;
        const uint32_t col = taddr + (uint32_t)(half * ROWS + c0);
        if constexpr (NSRC == 2) {
          uint32_t r0v[16], i0v[16], r1v[16], i1v[16];
          tmem_ld_32x32b_x16(col, r0v);
          tmem_ld_32x32b_x16(col + FOLD3M_IM_COL, i0v);
          tmem_ld_32x32b_x16(col + FPC, r1v);
          tmem_ld_32x32b_x16(col + FPC + FOLD3M_IM_COL, i1v);
          tmem_ld_wait();
          float vm[16], v0[16], v1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a0 = __uint_as_float(r0v[j]), b0 = __uint_as_float(i0v[j]);
            const float a1 = __uint_as_float(r1v[j]), b1 = __uint_as_float(i1v[j]);
            vm[j] = fold_mag(a0 + a1, b0 + b1);
            v0[j] = fold_mag(a0, b0);
            v1[j] = fold_mag(a1, b1);
          }
          fold_store_block<TC>(xt, lane, vm, plane(0), Tt, t0);
          fold_store_block<TC>(xt, lane, v0, plane(1), Tt, t0);
          fold_store_block<TC>(xt, lane, v1, plane(2), Tt, t0);
          if (want_reim) {
#pragma unroll
            for (int j = 0; j < 16; ++j) vm[j] = __uint_as_float(r0v[j]) + __uint_as_float(r1v[j]);
            fold_store_block<TC>(xt, lane, vm, p.re + reim_off, Tt, t0);
#pragma unroll
            for (int j = 0; j < 16; ++j) vm[j] = __uint_as_float(i0v[j]) + __uint_as_float(i1v[j]);
            fold_store_block<TC>(xt, lane, vm, p.im + reim_off, Tt, t0);
          }
        } else {
          float sr[16], si[16];
#pragma unroll
          for (int n = 0; n < NSRC; ++n) {
            uint32_t rv[16], iv[16];
            tmem_ld_32x32b_x16(col + n * FPC, rv);
            tmem_ld_32x32b_x16(col + n * FPC + FOLD3M_IM_COL, iv);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = __uint_as_float(rv[j]), bi = __uint_as_float(iv[j]);
              v[j] = fold_mag(a, bi);
              sr[j] = n == 0 ? a : sr[j] + a;
              si[j] = n == 0 ? bi : si[j] + bi;
            }
            fold_store_block<TC>(xt, lane, v, plane(1 + n), Tt, t0);
          }
          float vm[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) vm[j] = fold_mag(sr[j], si[j]);
          fold_store_block<TC>(xt, lane, vm, plane(0), Tt, t0);
          if (want_reim) {
            fold_store_block<TC>(xt, lane, sr, p.re + reim_off, Tt, t0);
            fold_store_block<TC>(xt, lane, si, p.im + reim_off, Tt, t0);
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive_cluster_relaxed(lead_tempty);
      tile_phase ^= 1u;
    }
    };
    if (p.T == 256) run(std::integral_constant<int, 256>{}); else run(std::integral_constant<int, 0>{});
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace avsep
