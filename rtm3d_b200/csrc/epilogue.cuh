// Epilogues shared by the generic and the streaming decode kernels.  Out-of-line (one copy per kernel): they run once per
// image and must not bloat the instruction footprint of the scan loops.
#pragma once
#include "common.cuh"
#include "params.h"
#include "tier_math.cuh"

namespace rtm3d {

// Tier A rows of image b (models/model.py:47-50 gather + sub-pixel, :63-73 regress / scale / 2D box), run by a whole block.
// One lane per (detection, vertex): the 2V+2 scattered 4-byte gathers of a detection are issued by neighbouring lanes at
// the same time (one DRAM round trip instead of a serial chain), the 2D box is a shuffle reduction over the vertex lanes.
//   sorted : `cnt` keys in descending order.  Rows >= cnt are zero-filled (cls = flat = -1).
template <typename T>
static __device__ __noinline__ void block_emit_main(const DecodeParams& p, int b, const uint64_t* sorted, int cnt) {
  const int V = p.n_vert, K = p.K, HW = p.H * p.W;
  if (p.off == nullptr) {
    // selection only (rtm3d_select_main): score, flat index and count; the caller's own epilogue follows
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
      const size_t row = static_cast<size_t>(b) * K + j;
      const bool valid = j < cnt;
      p.score[row] = valid ? key_score(sorted[j]) : 0.f;
      p.flat[row] = valid ? static_cast<int32_t>(key_flat(sorted[j])) : -1;
    }
    if (threadIdx.x == 0) p.counts[b] = cnt;
    return;
  }
  int vp = 1;
  while (vp < V) vp <<= 1;                       // lanes per detection (power of two <= 16)
  const int items = K * vp;
  const int items_pad = (items + 31) & ~31;
  const T* off2 = reinterpret_cast<const T*>(p.off2) + static_cast<size_t>(b) * 2 * HW;
  const T* off = reinterpret_cast<const T*>(p.off) + static_cast<size_t>(b) * 2 * V * HW;
  for (int idx = threadIdx.x; idx < items_pad; idx += blockDim.x) {
    const int j = idx / vp, v = idx - j * vp;
    const bool row_ok = j < K;
    const bool valid = row_ok && j < cnt;
    const bool vert = v < V;
    const size_t row = static_cast<size_t>(b) * K + (row_ok ? j : 0);
    float vx = 0.f, vy = 0.f, mx = 0.f, my = 0.f;
    uint32_t flat = 0;
    int c = 0;
    uint64_t key = 0;
    if (valid) {
      key = sorted[j];
      flat = key_flat(key);
      c = flat / HW;
      const int rem = flat - c * HW;
      const int yi = rem / p.W;
      const int xi = rem - yi * p.W;
      // issue every gather before the first use
      const float r0 = to_f32(off2[rem]);
      const float r1 = to_f32(off2[HW + rem]);
      float ox = 0.f, oy = 0.f;
      if (vert) {
        ox = to_f32(off[static_cast<size_t>(2 * v) * HW + rem]);
        oy = to_f32(off[static_cast<size_t>(2 * v + 1) * HW + rem]);
      }
      mx = subpixel(xi, r0);
      my = subpixel(yi, r1);
      vx = regress_coord(p.down, ox, mx);
      vy = regress_coord(p.down, oy, my);
    }
    float lo_x = (valid && vert) ? vx : INFINITY, hi_x = (valid && vert) ? vx : -INFINITY;
    float lo_y = (valid && vert) ? vy : INFINITY, hi_y = (valid && vert) ? vy : -INFINITY;
    for (int d = 1; d < vp; d <<= 1) {
      lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, d));
      hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, d));
      lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, d));
      hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, d));
    }
    if (!row_ok) continue;
    if (vert) {
      float* vout = p.verts + (row * V + v) * 2;
      vout[0] = valid ? vx : 0.f;
      vout[1] = valid ? vy : 0.f;
    }
    if (v == 0) {
      p.cls[row] = valid ? c : -1;
      p.score[row] = valid ? key_score(key) : 0.f;
      p.proj[row * 2 + 0] = valid ? scale_coord(p.down, mx) : 0.f;
      p.proj[row * 2 + 1] = valid ? scale_coord(p.down, my) : 0.f;
      p.bbox[row * 4 + 0] = valid ? lo_x : 0.f;
      p.bbox[row * 4 + 1] = valid ? lo_y : 0.f;
      p.bbox[row * 4 + 2] = valid ? hi_x : 0.f;
      p.bbox[row * 4 + 3] = valid ? hi_y : 0.f;
      if (p.flat) p.flat[row] = valid ? static_cast<int32_t>(flat) : -1;
    }
  }
  if (threadIdx.x == 0) p.counts[b] = cnt;
}

// Tier B candidate j of plane (b,c): index split + sub-pixel add (models/model.py:113-114 and the commented :55-57).
template <typename T>
__device__ __forceinline__ void emit_kpt_row(const DecodeParams& p, int b, int c, int j, float score, uint32_t flat) {
  const int HW = p.H * p.W;
  const int yi = flat / p.W;
  const int xi = flat - yi * p.W;
  const T* off2 = reinterpret_cast<const T*>(p.off2) + static_cast<size_t>(b) * 2 * HW;
  const float r0 = to_f32(off2[flat]);
  const float r1 = to_f32(off2[HW + flat]);
  const size_t row = (static_cast<size_t>(b) * p.C + c) * p.K + j;
  p.kscore[row] = score;
  p.kxy[row * 2 + 0] = subpixel(xi, r0);
  p.kxy[row * 2 + 1] = subpixel(yi, r1);
  p.kflat[row] = static_cast<int32_t>(flat);
}

// Tier B rows of plane (b,c), run by a whole block.  Rows cnt..K-1 are 0.0-score fillers = the lowest flat indices that
// are not positive-score peaks (what a top-K over the zero-filled peak map returns).
//   scratch: shared u32 scratch with >= 3*K + 8 words
template <typename T>
static __device__ __noinline__ void block_emit_kpt(const DecodeParams& p, int b, int c, const uint64_t* sorted, int cnt,
                                                   uint32_t* scratch) {
  const int K = p.K;
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) emit_kpt_row<T>(p, b, c, j, key_score(sorted[j]), key_flat(sorted[j]));
  if (cnt < K) {
    const int HW = p.H * p.W;
    const int span = min(K + cnt, HW);          // the first K-cnt non-candidate indices lie in [0, K+cnt)
    uint32_t* taken = scratch;                  // [span]
    uint32_t* fill = scratch + 2 * K;           // [K]
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
      uint32_t t = 0;
      for (int q = 0; q < cnt; ++q) t |= (key_flat(sorted[q]) == static_cast<uint32_t>(i));
      taken[i] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int r = cnt;
      for (int i = 0; i < span && r < K; ++i)
        if (!taken[i]) fill[r++] = i;
    }
    __syncthreads();
    for (int j = cnt + threadIdx.x; j < K; j += blockDim.x) emit_kpt_row<T>(p, b, c, j, 0.0f, fill[j]);
  }
}

template <typename T, int MODE>
__device__ __forceinline__ void block_emit(const DecodeParams& p, int b, int c, const uint64_t* sorted, int cnt,
                                           uint32_t* scratch) {
  if (MODE == kModeMain) block_emit_main<T>(p, b, sorted, cnt);
  else block_emit_kpt<T>(p, b, c, sorted, cnt, scratch);
}

}  // namespace rtm3d
