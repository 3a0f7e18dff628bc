// Device helpers shared by the streaming kernels (scan_planes.cu and the legacy decode_planes.cu): PTX wrappers for the
// bulk-copy ring (cp.async.bulk + mbarrier), the 16-byte pixel group and the rows of its 3x3 window.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtm3d {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
namespace pl {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// The heat-maps are read once: their lines are the first to leave L2 (`policy` = evict_first), which keeps the
// selection outputs the post kernel reads next (score / flat / kflat) resident.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_normal() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void fence_sc_cta() { asm volatile("fence.sc.cta;" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a pipeline bug must surface as a launch failure, never as a hung GPU.  `backoff_ns`: sleep between
// polls (waiting warps share issue slots with the scanners).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* status, uint32_t code, unsigned backoff_ns) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (true) {
    __nanosleep(backoff_ns);
    if (mbar_try_wait(bar, parity)) return;
    if (clock64() - t0 > 4000000000LL) {
      if (status) atomicExch(status, code);
      __threadfence_system();
      __trap();
    }
  }
}
}  // namespace pl

template <typename T> struct Grp;  // one 16-byte group of a row
template <> struct Grp<float> {
  static constexpr int E = 4;
  __device__ static __forceinline__ void load(const unsigned char* p, float (&v)[4]) {
    const float4 f = *reinterpret_cast<const float4*>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
  __device__ static __forceinline__ float elem(const unsigned char* p, int i) { return reinterpret_cast<const float*>(p)[i]; }
};
template <> struct Grp<__nv_bfloat16> {
  static constexpr int E = 8;
  __device__ static __forceinline__ void load(const unsigned char* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ static __forceinline__ float elem(const unsigned char* p, int i) {
    return __uint_as_float(static_cast<uint32_t>(reinterpret_cast<const unsigned short*>(p)[i]) << 16);
  }
};

// One row of the 3x3 window around a group: r[0] = left neighbour of the group's first pixel, r[1..E] = the pixels above /
// below the group, r[E+1] = right neighbour of its last pixel; -inf where the image ends (max_pool2d's implicit padding).
template <typename T>
__device__ __forceinline__ void load_window_row(const unsigned char* p, bool has_row, bool has_l, bool has_r,
                                                float (&r)[Grp<T>::E + 2]) {
  constexpr int E = Grp<T>::E;
  if (has_row) {
    float v[E];
    Grp<T>::load(p, v);
#pragma unroll
    for (int i = 0; i < E; ++i) r[i + 1] = v[i];
    r[0] = has_l ? Grp<T>::elem(p, -1) : -INFINITY;
    r[E + 1] = has_r ? Grp<T>::elem(p, E) : -INFINITY;
  } else {
#pragma unroll
    for (int i = 0; i < E + 2; ++i) r[i] = -INFINITY;
  }
}

}  // namespace rtm3d
