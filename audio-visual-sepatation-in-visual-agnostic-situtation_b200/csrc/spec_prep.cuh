// K4 front half, shared-memory staged (the product path of avsep_separate when (re, im) of the mixture are given and
// T % 4 == 0): un-warp the predicted masks to the linear frequency axis, threshold, multiply the mixture spectrum and
// emit the time-major bf16 hi / lo operand of the inverse DFT (main.py:213-246).
//
// spec_prep_vec_kernel (dft_plan.cu) was latency-bound: every thread chased its own mask-row loads (ncu: 12 of 16.8 stall
// cycles per instruction on the long scoreboard, 31 % issue-active) and each 128-bit operand store touched 16 different
// cache lines (16 frames x 32 B), so the LSU data pipe ran 11 wavefronts per request.  Here a CTA owns a
// 64 bins x 32 frames tile of one mixture:
//   * loads: the (re, im) tile and the contiguous range of mask rows the tile's bins tap (for all sources at once when they
//     fit the staging buffer, else a few sources per pass) are copied to shared memory with 16-byte cp.async -- everything
//     in flight at once, no thread waits on its own chain; the mixture signal (mask 1) is converted while the masks fly
//   * compute: taps, threshold, multiply and the bf16 hi / lo split read shared memory only
//   * stores: the four [32 frames][64 bins] bf16 planes of a signal are assembled in shared memory in the SWIZZLE_128B
//     layout and written by TMA tensor stores (cp.async.bulk.tensor .. bulk_group), double-buffered across signals:
//     whole 128-byte rows, no LSU wavefronts, no store instructions in the threads
// Bit-identical to spec_prep_vec_kernel (same expressions in the same order).
#pragma once
#include "common.cuh"

namespace avsep {

constexpr int SPT_BINS = 64, SPT_FRAMES = 32, SPT_THREADS = 128;
constexpr int SPT_COLS = 40;                                   // staged mask columns: frames [t0 - 4, t0 + 36), 16-byte chunks
constexpr int SPT_PART_BYTES = SPT_FRAMES * SPT_BINS * 2;      // 4 KB: one (re | im, hi | lo) plane of a signal's tile
constexpr int SPT_OUT_BYTES = 4 * SPT_PART_BYTES;              // 16 KB: the four planes of one signal's tile
constexpr int SPT_TILE_FLOATS = SPT_BINS * SPT_FRAMES;         // (re) or (im) tile, [bin][frame]
static_assert(2 * SPT_TILE_FLOATS * 4 == SPT_OUT_BYTES, "output buffer 1 aliases the (re, im) tile exactly");

__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :: "l"(tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// 8 fp32 values -> bf16 hi / lo (x ~= hi + lo) as two 16-byte chunks
__device__ __forceinline__ void spt_split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// stage_floats: capacity of the mask staging buffer; planes: signals per (hi | lo) plane group of the operand (= plan capacity)
__global__ void __launch_bounds__(SPT_THREADS) spec_prep_tma_kernel(const __grid_constant__ CUtensorMap tm_out, const SpecPrepParams p,
                                                                    int stage_floats, int planes) {
  extern __shared__ uint8_t spt_smem_raw[];
  __shared__ int s_y0[SPT_BINS];
  __shared__ float s_fy[SPT_BINS];
  const uint32_t base_u32 = (smem_u32(spt_smem_raw) + 1023u) & ~1023u;
  uint8_t* base = spt_smem_raw + (base_u32 - smem_u32(spt_smem_raw));
  // output buffers [4 parts][32 frames][128 B], swizzled: buffer 0 is its own, buffer 1 ALIASES the (re, im) tile, which is
  // dead once every thread holds its cells in registers (load_tile + the barrier of the first finished signal)
  uint8_t* out_sm = base;
  float* re_sm = reinterpret_cast<float*>(base + SPT_OUT_BYTES);           // [64 bins][32 frames]
  float* im_sm = re_sm + SPT_TILE_FLOATS;
  float* mk_sm = im_sm + SPT_TILE_FLOATS;                                  // staged masks, stage_floats

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int f_cta = blockIdx.y * SPT_BINS;
  const int t0 = blockIdx.x * SPT_FRAMES;
  const int T = p.T;
  const int nsig = p.N + p.include_mix;
  const bool warped = p.log_freq && p.N > 0;

  // rows of the (warped) masks this tile taps: y0 is non-decreasing in the bin
  int ylo = 0, R = SPT_BINS;                                               // linear masks: the tile's own 64 rows
  if (warped) {
    if (tid < SPT_BINS) {
      s_y0[tid] = __ldg(p.ytap.i0 + f_cta + tid);
      s_fy[tid] = __ldg(p.ytap.frac + f_cta + tid);
    }
    ylo = max(__ldg(p.ytap.i0 + f_cta), 0);
    const int yhi = min(__ldg(p.ytap.i0 + f_cta + SPT_BINS - 1) + 1, p.Hm - 1);
    R = max(yhi - ylo + 1, 0);
  }
  const int cols = warped ? SPT_COLS : SPT_FRAMES;
  const int per_src = R * cols;
  const int spp = per_src > 0 ? max(min(p.N, stage_floats / per_src), 1) : max(p.N, 1);     // sources per staging pass

  // ---- (re, im) tile: 64 rows x 8 chunks each ---------------------------------------------------------------------------
  {
    const size_t o = ((size_t)b * AVSEP_N_BINS + f_cta) * T + t0;
    for (int i = tid; i < SPT_BINS * (SPT_FRAMES / 4); i += SPT_THREADS) {
      const int r = i >> 3, c = (i & 7) * 4;
      if (t0 + c < T) {
        cp_async_16(re_sm + r * SPT_FRAMES + c, p.re + o + (size_t)r * T + c);
        cp_async_16(im_sm + r * SPT_FRAMES + c, p.im + o + (size_t)r * T + c);
      } else {
        *reinterpret_cast<float4*>(re_sm + r * SPT_FRAMES + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(im_sm + r * SPT_FRAMES + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  }
  // division-free: a warp covers 3 rows x 10 chunks (warped masks) or 4 rows x 8 chunks (linear masks) per iteration
  auto stage_masks = [&](int n0, int cnt) {
    const int cpr = cols / 4;                                              // chunks per row: 10 or 8
    const int rows_per_warp = 32 / cpr;                                    // 3 or 4
    const int lr = lane / cpr, ch = lane - lr * cpr;
    const int col = (warped ? t0 - 4 : t0) + 4 * ch;
    const bool col_ok = col >= 0 && col < T;
    const int row0 = warped ? ylo : f_cta;
    if (lr < rows_per_warp) {
      for (int k = 0; k < cnt; ++k) {
        const float* src = p.masks + (((size_t)(n0 + k) * p.B + b) * p.Hm + row0) * (size_t)T + col;
        float* dst = mk_sm + (size_t)k * per_src + 4 * ch;
        for (int rr = warp * rows_per_warp + lr; rr < R; rr += (SPT_THREADS / 32) * rows_per_warp) {
          if (col_ok) cp_async_16(dst + rr * cols, src + (size_t)rr * T);
          else *reinterpret_cast<float4*>(dst + rr * cols) = make_float4(0.f, 0.f, 0.f, 0.f);   // zero padding outside the image
        }
      }
    }
    cp_async_commit();
  };
  if (p.N > 0) stage_masks(0, min(spp, p.N));

  // thread = frames (fp, fp + 16) x 8 consecutive bins
  const int fp = lane & 15, grp = warp * 2 + (lane >> 4);
  const int fr[2] = {fp, fp + 16};
  int sel[2] = {0, 0};
  float fx[2] = {0.f, 0.f};
  if (warped) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int t = min(t0 + fr[e], T - 1);
      sel[e] = __ldg(p.xtap.i0 + t) - (t - 1);           // tap 0 is column t - 1 (sel = 0) or t (sel = 1)
      fx[e] = __ldg(p.xtap.frac + t);
    }
  }
  float bre[8][2], bim[8][2];
  bool have_tile = false;
  auto load_tile = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        bre[i][e] = re_sm[(grp * 8 + i) * SPT_FRAMES + fr[e]];
        bim[i][e] = im_sm[(grp * 8 + i) * SPT_FRAMES + fr[e]];
      }
    have_tile = true;
  };
  // write the thread's 2 frames x 8 bins of one signal into output buffer `buf` (SWIZZLE_128B: chunk ^= frame & 7)
  auto emit = [&](int buf, const float (&mk)[8][2], bool unit) {
    uint8_t* ob = out_sm + buf * SPT_OUT_BYTES;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float vre[8], vim[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        vre[i] = unit ? bre[i][e] : bre[i][e] * mk[i][e];
        vim[i] = unit ? bim[i][e] : bim[i][e] * mk[i][e];
      }
      uint4 h, l;
      const int off = fr[e] * 128 + ((grp ^ (fr[e] & 7)) << 4);
      spt_split8(vre, h, l);
      *reinterpret_cast<uint4*>(ob + 0 * SPT_PART_BYTES + off) = h;
      *reinterpret_cast<uint4*>(ob + 1 * SPT_PART_BYTES + off) = l;
      spt_split8(vim, h, l);
      *reinterpret_cast<uint4*>(ob + 2 * SPT_PART_BYTES + off) = h;
      *reinterpret_cast<uint4*>(ob + 3 * SPT_PART_BYTES + off) = l;
    }
  };
  int issued = 0;                                        // signals handed to TMA so far (buffer = issued & 1)
  auto begin_signal = [&]() {                            // the buffer's previous store (two signals ago) has read shared memory
    if (issued >= 2) {
      if (tid == 0) bulk_wait_read<1>();
      __syncthreads();
    }
  };
  auto finish_signal = [&](int jsig) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      const uint32_t src = base_u32 + (issued & 1) * SPT_OUT_BYTES;
      const int sig = b * nsig + jsig;
      tma_store_3d(&tm_out, src + 0 * SPT_PART_BYTES, f_cta, t0, sig);
      tma_store_3d(&tm_out, src + 1 * SPT_PART_BYTES, f_cta, t0, planes + sig);
      tma_store_3d(&tm_out, src + 2 * SPT_PART_BYTES, AVSEP_N_BINS + f_cta, t0, sig);
      tma_store_3d(&tm_out, src + 3 * SPT_PART_BYTES, AVSEP_N_BINS + f_cta, t0, planes + sig);
      bulk_commit();
    }
    ++issued;
  };

  // ---- mixture signal (mask 1): needs only the (re, im) tile, converted while the masks are still in flight --------------
  if (p.N > 0) cp_async_wait<1>(); else cp_async_wait<0>();
  __syncthreads();
  load_tile();
  if (p.include_mix) {
    float mk[8][2];
    begin_signal();
    emit(issued & 1, mk, true);
    finish_signal(0);
  }
  // ---- sources, spp per staging pass ---------------------------------------------------------------------------------------
  for (int n0 = 0; n0 < p.N; n0 += spp) {
    const int cnt = min(spp, p.N - n0);
    if (n0 > 0) {
      __syncthreads();                                   // every thread is done reading the previous pass
      stage_masks(n0, cnt);
    }
    cp_async_wait<0>();
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      const float* ms = mk_sm + (size_t)k * per_src;
      float mk[8][2];
      if (warped) {
        int cy0 = -2, cy1 = -2;
        float cv0[2] = {0.f, 0.f}, cv1[2] = {0.f, 0.f};
        auto row = [&](int y, float (&r)[2]) {           // x-interpolated mask row y at the 2 frames (0 outside the image)
          if (y == cy0) { r[0] = cv0[0]; r[1] = cv0[1]; return; }
          if (y == cy1) { r[0] = cv1[0]; r[1] = cv1[1]; return; }
          if (y < 0 || y >= p.Hm) { r[0] = 0.f; r[1] = 0.f; return; }
          const float* q = ms + (y - ylo) * SPT_COLS + 3;                // q[j] = column t0 - 1 + j
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float a = q[fr[e] + sel[e]];
            const float c2 = q[fr[e] + sel[e] + 1];
            r[e] = (1.f - fx[e]) * a + fx[e] * c2;
          }
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int y0 = s_y0[grp * 8 + i];
          const float fy = s_fy[grp * 8 + i];
          float r0[2], r1[2];
          row(y0, r0);
          row(y0 + 1, r1);
          cy0 = y0; cy1 = y0 + 1;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            cv0[e] = r0[e]; cv1[e] = r1[e];
            const float m = (1.f - fy) * r0[e] + fy * r1[e];
            mk[i][e] = p.binary ? (m > p.thres ? 1.f : 0.f) : m;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float c = ms[(grp * 8 + i) * SPT_FRAMES + fr[e]];
            mk[i][e] = p.binary ? (c > p.thres ? 1.f : 0.f) : c;
          }
      }
      begin_signal();
      emit(issued & 1, mk, false);
      finish_signal(p.include_mix + n0 + k);
    }
  }
  (void)have_tile;
  if (tid == 0) bulk_wait_read<0>();                     // shared memory stays valid until the last store has read it
}

}  // namespace avsep
