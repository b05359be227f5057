// Separable tap tables of `F.grid_sample(x, warpgrid(...), align_corners=False)`.
//
// utils.warpgrid (utils.py:12-26) builds grid_x from the column index only and grid_y from the row
// index only, so the bilinear, zero-padded resample factorises: out = Wy * in * Wx^T with two taps
// per output row / column.  The tables are computed on the host exactly as the reference does
// (float64 linspace + pow / log, cast to float32, then ATen's float32 un-normalisation
// i = ((g + 1) * size - 1) / 2) and cached on the device per (device, n_in, n_out, kind).
#pragma once
#include <stdint.h>

namespace avsep {

enum TapKind : int {
  TAP_WARP_Y = 0,    // warpgrid(warp=True)  rows:  gy = (21^((y+1)/2) - 11) / 10
  TAP_UNWARP_Y = 1,  // warpgrid(warp=False) rows:  gy = log(10 y + 11) / log(21) * 2 - 1
  TAP_LINEAR = 2,    // grid_x (and grid_y when log_freq is off): g = linspace(-1, 1, n_out)
};

struct Taps {
  const int* i0;      // device: lower tap index per output index (may be -1 or n_in - 1 => a tap is out of range)
  const float* frac;  // device: weight of tap i0 + 1; tap i0 has weight 1 - frac
};

// host computation (also exported for CPU tests)
void compute_taps(int kind, int n_in, int n_out, int* i0, float* frac);
// device-cached tables; returns 0 on success
int get_taps(int kind, int n_in, int n_out, Taps* out);

}  // namespace avsep
