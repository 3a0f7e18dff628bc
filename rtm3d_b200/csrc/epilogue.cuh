// Per-detection epilogues shared by the generic and the streaming decode kernels.
#pragma once
#include "common.cuh"
#include "params.h"

namespace rtm3d {

// Tier A row j of image b (models/model.py:47-50 gather + sub-pixel, :63-73 regress / scale / 2D box).
// `valid` false -> the row is zero-filled (cls = flat = -1).
template <typename T>
__device__ __forceinline__ void emit_main_row(const DecodeParams& p, int b, int j, uint64_t key, bool valid) {
  const int V = p.n_vert;
  const size_t row = static_cast<size_t>(b) * p.K + j;
  float* vout = p.verts + row * V * 2;
  if (!valid) {
    p.cls[row] = -1;
    p.score[row] = 0.f;
    p.proj[row * 2 + 0] = 0.f; p.proj[row * 2 + 1] = 0.f;
    for (int v = 0; v < 2 * V; ++v) vout[v] = 0.f;
    p.bbox[row * 4 + 0] = 0.f; p.bbox[row * 4 + 1] = 0.f; p.bbox[row * 4 + 2] = 0.f; p.bbox[row * 4 + 3] = 0.f;
    if (p.flat) p.flat[row] = -1;
    return;
  }
  const int HW = p.H * p.W;
  const uint32_t flat = key_flat(key);
  const int c = flat / HW;
  const int rem = flat - c * HW;
  const int yi = rem / p.W;
  const int xi = rem - yi * p.W;
  const size_t pix = static_cast<size_t>(yi) * p.W + xi;
  const T* off2 = reinterpret_cast<const T*>(p.off2) + static_cast<size_t>(b) * 2 * HW;
  const T* off = reinterpret_cast<const T*>(p.off) + static_cast<size_t>(b) * 2 * V * HW;
  // issue every gather before the first use: they are the only uncoalesced loads of the decode
  float raw[2 * kMaxVerts];
  const float r0 = to_f32(off2[pix]);
  const float r1 = to_f32(off2[HW + pix]);
#pragma unroll 4
  for (int v = 0; v < 2 * V; ++v) raw[v] = to_f32(off[static_cast<size_t>(v) * HW + pix]);
  const float mx = __fadd_rn(static_cast<float>(xi), sigmoid_ref(r0));
  const float my = __fadd_rn(static_cast<float>(yi), sigmoid_ref(r1));
  float lo_x = INFINITY, lo_y = INFINITY, hi_x = -INFINITY, hi_y = -INFINITY;
  for (int v = 0; v < V; ++v) {
    const float vx = __fmul_rn(p.down, __fadd_rn(raw[2 * v], mx));
    const float vy = __fmul_rn(p.down, __fadd_rn(raw[2 * v + 1], my));
    vout[2 * v] = vx;
    vout[2 * v + 1] = vy;
    lo_x = fminf(lo_x, vx); hi_x = fmaxf(hi_x, vx);
    lo_y = fminf(lo_y, vy); hi_y = fmaxf(hi_y, vy);
  }
  p.cls[row] = c;
  p.score[row] = key_score(key);
  p.proj[row * 2 + 0] = __fmul_rn(p.down, mx);
  p.proj[row * 2 + 1] = __fmul_rn(p.down, my);
  p.bbox[row * 4 + 0] = lo_x; p.bbox[row * 4 + 1] = lo_y; p.bbox[row * 4 + 2] = hi_x; p.bbox[row * 4 + 3] = hi_y;
  if (p.flat) p.flat[row] = static_cast<int32_t>(flat);
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
  p.kxy[row * 2 + 0] = __fadd_rn(static_cast<float>(xi), sigmoid_ref(r0));
  p.kxy[row * 2 + 1] = __fadd_rn(static_cast<float>(yi), sigmoid_ref(r1));
  p.kflat[row] = static_cast<int32_t>(flat);
}

// Final stage for one selection problem, run by a whole block.
//   sorted : shared array holding `cnt` keys sorted descending (capacity >= next_pow2(K))
//   scratch: shared u32 scratch with >= 3*K + 8 words (only used for kModeKpt fillers)
// kModeMain: rows 0..cnt-1 valid, rest zero-filled, counts[b] = cnt.
// kModeKpt : rows cnt..K-1 are 0.0-score fillers = the lowest flat indices that are not positive-score peaks.
template <typename T, int MODE>
__device__ __forceinline__ void block_emit(const DecodeParams& p, int b, int c, const uint64_t* sorted, int cnt,
                                           uint32_t* scratch) {
  const int K = p.K;
  if (MODE == kModeMain) {
    for (int j = threadIdx.x; j < K; j += blockDim.x) emit_main_row<T>(p, b, j, j < cnt ? sorted[j] : 0ull, j < cnt);
    if (threadIdx.x == 0) p.counts[b] = cnt;
    return;
  }
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

}  // namespace rtm3d
