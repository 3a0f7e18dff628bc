// The arithmetic of the Tier A / Tier B epilogues, stated ONCE: every kernel that emits these values (the epilogue and
// grouping kernels, the fused post kernels, the strip kernels' in-kernel epilogues) calls these functions, so the
// association order -- and with it the bit pattern of every output float -- cannot drift between them.
// Explicit __fadd_rn / __fmul_rn / __fsub_rn: no contraction, IEEE round-to-nearest, as ATen's separate ops on CUDA.
#pragma once
#include "common.cuh"

namespace rtm3d {

// position = integer coordinate + sigmoid(offset logit): the centre's sub-pixel add (models/model.py:48-50) and the keypoint
// candidates' (:113-114 with the commented :55-57)
__device__ __forceinline__ float subpixel(int i, float offset_logit) { return __fadd_rn(static_cast<float>(i), sigmoid_ref(offset_logit)); }

// vertex coordinate = down * (regressed offset + centre)   (models/model.py:63-69; also the grouped `verts` of :160)
__device__ __forceinline__ float regress_coord(float down, float off, float m) { return __fmul_rn(down, __fadd_rn(off, m)); }

// output scaling of a centre / candidate coordinate (models/model.py:69-70)
__device__ __forceinline__ float scale_coord(float down, float v) { return __fmul_rn(down, v); }

// two fp32 lanes in one 64-bit register for the packed fp32x2 pipe (sub.rn / mul.rn per lane: the scalar results)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return static_cast<unsigned long long>(__float_as_uint(lo)) | (static_cast<unsigned long long>(__float_as_uint(hi)) << 32);
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// _group_vertexs_kf (models/model.py:144-161): j* = argmin_j ||(v_j - m) - off||^2 over K candidates (x, y pairs, 8-byte
// aligned), first minimal index on ties (torch.argmin).  d = dx*dx + dy*dy with (v - m) - off in that order (:147,149).
// Four independent (best, index) chains over j = 4i + u keep the compare/select dependency off the critical path.
__device__ __forceinline__ int nearest_candidate(const float* cand, int K, float mx, float my, float ox, float oy) {
  const unsigned long long* cand2 = reinterpret_cast<const unsigned long long*>(cand);
  const unsigned long long m2 = pack2(mx, my), o2 = pack2(ox, oy);
  auto dist = [&](int j) {
    const unsigned long long df = sub2(sub2(cand2[j], m2), o2);
    const unsigned long long sq = mul2(df, df);
    return __fadd_rn(__uint_as_float(static_cast<uint32_t>(sq)), __uint_as_float(static_cast<uint32_t>(sq >> 32)));
  };
  float bd[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
  int bi[4] = {0, 0, 0, 0};
  int j = 0;
#pragma unroll 2
  for (; j + 4 <= K; j += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d = dist(j + u);
      if (d < bd[u]) { bd[u] = d; bi[u] = j + u; }
    }
  }
  for (; j < K; ++j) {
    const float d = dist(j);
    if (d < bd[0]) { bd[0] = d; bi[0] = j; }      // (j is past every index chain 0 has seen)
  }
  float best = bd[0];
  int bj = bi[0];
#pragma unroll
  for (int u = 1; u < 4; ++u)
    if (bd[u] < best || (bd[u] == best && bi[u] < bj)) { best = bd[u]; bj = bi[u]; }
  return bj;
}

}  // namespace rtm3d
