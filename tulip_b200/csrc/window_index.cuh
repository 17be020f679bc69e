// Index arithmetic of the (shifted) window partition -- the only place it is defined.
// Used inside the fused attention kernels and by the stand-alone bit-exact index ops.
#pragma once
#include "common.cuh"

struct WinGeom {
  int H, W;        // token grid
  int Mh, Mw;      // window used for partitioning
  int sh, sw;      // cyclic shift, torch.roll(x, (-sh,-sw)) before partition (tulip.py:290)
};

// token i (row-major in the window) of window (b, wh, ww) on the rolled grid -> flat NHWC token index
// rolled[h] = x[(h + sh) % H]                                   (tulip.py:290, 295, 248-252)
__host__ __device__ __forceinline__ int win_token_index(const WinGeom& g, int b, int wh, int ww, int i) {
  const int r = i / g.Mw, c = i % g.Mw;
  int hs = wh * g.Mh + r + g.sh;
  int ws = ww * g.Mw + c + g.sw;
  if (hs >= g.H) hs -= g.H;
  if (ws >= g.W) ws -= g.W;
  return (b * g.H + hs) * g.W + ws;
}

// region id assigned by create_mask on the rolled grid (tulip.py:261-271); a zero shift component
// makes python's slice(-0, None) cover everything, i.e. that component is constant.
__host__ __device__ __forceinline__ int win_region_id(const WinGeom& g, int wh, int ww, int i) {
  const int hr = wh * g.Mh + i / g.Mw, wr = ww * g.Mw + i % g.Mw;
  const int hreg = g.sh > 0 ? (hr >= g.H - g.Mh) + (hr >= g.H - g.sh) : 0;
  const int wreg = g.sw > 0 ? (wr >= g.W - g.Mw) + (wr >= g.W - g.sw) : 0;
  return 3 * hreg + wreg;
}

// relative_position_index[i][j] for the window the bias table was built for (tulip.py:228-240)
__host__ __device__ __forceinline__ int rel_bias_index(int bMh, int bMw, int i, int j) {
  const int ri = i / bMw, ci = i % bMw, rj = j / bMw, cj = j % bMw;
  return (ri - rj + bMh - 1) * (2 * bMw - 1) + (ci - cj + bMw - 1);
}
