// Shared device helpers of the RTM3D decode kernels (sm_100a).
//
// Numerics contract (DESIGN.md "Exactness"):
//  * s(x) = 1.0f / (1.0f + expf(-x))  -- libdevice expf, IEEE add and divide, no fast-math.  Measured bit-identical
//    to torch's CUDA sigmoid on 13 M values (profiles/r01_probe_sigmoid_topk.json), i.e. to models/model.py:85.
//  * ordering key = (score bits << 32) | (0xFFFFFFFF - flat index): descending u64 order == (score desc, index asc),
//    the order torch.topk yields on CUDA (same probe) -- models/model.py:90.
//  * every arithmetic step that reaches an output is written with explicit round-to-nearest intrinsics so that no
//    FMA contraction can change a bit relative to the reference's separate ATen ops.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtm3d {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs a driver call: made once per (kernel, device), and again only when
// a launch asks for more than the kernel was granted before.
template <auto Kern>
inline cudaError_t ensure_dynamic_smem(size_t bytes) {
  static int granted[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool tracked = dev >= 0 && dev < 64;
  if (tracked && granted[dev] >= static_cast<int>(bytes)) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e == cudaSuccess && tracked) granted[dev] = static_cast<int>(bytes);
  return e;
}


constexpr int kMaxTopK = 1024;
constexpr int kMaxVerts = 16;

__device__ __forceinline__ float sigmoid_ref(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
// Same function, one out-of-line copy: for rare paths, so that they do not bloat the instruction footprint of the kernel
// (a 137 KB kernel was measured to run its cold paths at instruction-fetch speed).
static __device__ __noinline__ float sigmoid_cold(float x) { return sigmoid_ref(x); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ uint64_t make_key(float score, uint32_t flat) {
  return (static_cast<uint64_t>(__float_as_uint(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - flat);
}
__device__ __forceinline__ float key_score(uint64_t k) { return __uint_as_float(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ uint32_t key_flat(uint64_t k) { return 0xFFFFFFFFu - static_cast<uint32_t>(k); }

// order-preserving map float -> u32 and back (NaN patterns land beyond +-inf and never compare as candidates)
__host__ __device__ __forceinline__ uint32_t f32_ord_bits(uint32_t b) { return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ uint32_t f32_ord(float f) { return f32_ord_bits(__float_as_uint(f)); }
__device__ __forceinline__ float f32_unord(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u); }

// ---- conservative logit-domain filters --------------------------------------------------------------------------
// The scan compares LOGITS and evaluates the exact sigmoid only for the few pixels that survive.  All margins below
// are chosen so that, with the computed sigmoid within a few ulp of the real one, a pixel rejected in the logit
// domain could not have changed the result in the sigmoid domain (DESIGN.md "Exactness" has the derivation).
constexpr float kTieTol = 6.103515625e-05f;  // 2^-14: logits closer than this may collapse to one fp32 sigmoid (x <= 2)
constexpr float kSatKnee = 2.0f;             // above this the collapse distance grows like e^x: handled by exact compares
constexpr float kDenormKnee = -80.0f;        // below this sigmoid is (nearly) denormal: always exact compares

// A neighbour logit xn can tie with / exceed the centre xc IN THE SIGMOID DOMAIN only if this holds; every other
// neighbour is strictly smaller after the sigmoid as well.
__device__ __forceinline__ bool neighbour_needs_exact(float xn, float xc) {
  return (xn > fminf(xc, kSatKnee) - kTieTol) || (xc < kDenormKnee);
}

// Largest logit bound T with: x < T  =>  s(x) < s(xk) strictly (xk = logit of the current K-th best candidate).
__device__ __forceinline__ float filter_from_kth_logit(float xk) {
  if (xk < kDenormKnee) return -INFINITY;
  float m = (xk <= kSatKnee) ? kTieTol : 2.44140625e-04f * __expf(xk);  // 2^-12 * e^x
  return xk - m;                                                        // inf - inf cannot occur: __expf(>88) = inf -> -inf
}

// ---- block-level primitives (all threads of the block must call) ------------------------------------------------

// In-place descending bitonic sort of a[0..npad), npad a power of two.
static __device__ __noinline__ void block_bitonic_sort_desc(uint64_t* a, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        int p = i ^ j;
        if (p > i) {
          uint64_t x = a[i], y = a[p];
          bool first_block = ((i & k) == 0);
          if (first_block ? (x < y) : (x > y)) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// Exact selection of the K largest of n DISTINCT u64 keys (MSB-first radix select, 8-bit digits).
//   keys : n keys in shared memory (not modified)        out : receives min(n,K) keys, unordered
//   hist : 256 + 4 words of shared scratch
// Returns min(n, K).
static __device__ __noinline__ int block_select_topk(const uint64_t* keys, int n, int K, uint64_t* out, uint32_t* hist) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (n <= K) {
    for (int i = tid; i < n; i += nt) out[i] = keys[i];
    __syncthreads();
    return n;
  }
  uint64_t prefix = 0, mask = 0;
  uint32_t need = static_cast<uint32_t>(K);
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = tid; i < 256; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      uint64_t k = keys[i];
      if ((k & mask) == prefix) atomicAdd(&hist[static_cast<uint32_t>(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns digits 255-8l .. 248-8l (descending); find the digit where the running count reaches `need`
      uint32_t c[8], s = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { c[q] = hist[255 - 8 * tid - q]; s += c[q]; }
      uint32_t incl = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (tid >= d) incl += v;
      }
      uint32_t excl = incl - s;
      if (excl < need && incl >= need) {
        uint32_t run = excl;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (run < need && run + c[q] >= need) { hist[256] = 255 - 8 * tid - q; hist[257] = need - run; }
          run += c[q];
        }
      }
    }
    __syncthreads();
    const uint32_t digit = hist[256];
    need = hist[257];
    prefix |= static_cast<uint64_t>(digit) << shift;
    mask |= 0xFFull << shift;
    __syncthreads();
  }
  // keys are distinct, so exactly K of them are >= prefix (the K-th largest key)
  if (tid == 0) hist[258] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    uint64_t k = keys[i];
    if (k >= prefix) out[atomicAdd(&hist[258], 1u)] = k;
  }
  __syncthreads();
  return K;
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace rtm3d
