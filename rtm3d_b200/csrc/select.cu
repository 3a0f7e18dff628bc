// Selection after the scan kernel (general path, any K <= 1024 and any number of strips per plane): one CTA per selection
// problem -- the C planes of an image (flat top-K over C*H*W, models/model.py:87-98) or one keypoint plane (per-channel
// top-K, models/model.py:109-114).  The CTA merges the candidate lists of the problem's strips (each provably contains its
// strip's K best peaks, scan_planes.cu), selects the exact K best keys, sorts them by (score desc, index asc) and writes
// score / flat / counts (Tier A) or kscore / kflat with the 0.0-score fillers (Tier B).
#include "common.cuh"
#include "params.h"
#include "select_common.cuh"

namespace rtm3d {

constexpr int kSelThreads = 128;

// Tier B rows cnt..K-1: the lowest flat indices that are not among the plane's positive-score peaks (what a top-K over the
// zero-filled peak map returns).  One warp; scratch >= 3K+8 words.
static __device__ __noinline__ void warp_fill_kpt(const SelectParams& p, size_t row0, const uint64_t* sorted, int cnt,
                                                  uint32_t* scratch, int lane) {
  const int K = p.K, HW = p.H * p.W;
  uint32_t* fill = scratch + 2 * K;             // [K] filler indices (rows cnt..K-1)
  if (cnt < K) {
    const int span = min(K + cnt, HW);          // the first K-cnt non-candidate indices lie in [0, K+cnt)
    uint32_t* taken = scratch;                  // [span]
#pragma unroll 1
    for (int i = lane; i < span; i += 32) {
      uint32_t t = 0;
#pragma unroll 1
      for (int q = 0; q < cnt; ++q) t |= (key_flat(sorted[q]) == static_cast<uint32_t>(i));
      taken[i] = t;
    }
    __syncwarp();
    if (lane == 0) {
      int r = cnt;
#pragma unroll 1
      for (int i = 0; i < span && r < K; ++i)
        if (!taken[i]) fill[r++] = i;
    }
    __syncwarp();
  }
#pragma unroll 1
  for (int j = lane; j < K; j += 32) {
    p.kscore[row0 + j] = j < cnt ? key_score(sorted[j]) : 0.0f;
    p.kflat[row0 + j] = static_cast<int32_t>(j < cnt ? key_flat(sorted[j]) : fill[j]);
  }
}

__global__ void __launch_bounds__(kSelThreads) select_kernel(const SelectParams p, int ns) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_n;
  __shared__ int s_cnt[kFastLists];
  const int K = p.K, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                  // [ns] sort buffer | fast path: [n_lists + 1][kFastPad]
  const int n_main = p.C > 0 ? p.B : 0;
  const int q = blockIdx.x;
  const bool is_main = q < n_main;
  const int first = is_main ? q * p.C * p.Sp : (p.B * p.C + (q - n_main)) * p.Sp;
  const int n_lists = is_main ? p.C * p.Sp : p.Sp;
  // strict `score > thresh` of models/model.py:91; 0.0 for the keypoint planes: zero-score pixels are fillers, not peaks
  const float lim = is_main ? p.thresh : 0.0f;
  const uint64_t* top = keys;
  int have;
  bool fast = K <= kFastPad && n_lists <= kFastLists && (n_lists + 1) * kFastPad <= ns;
  {
    int too_long = 0;
    if (fast && tid < n_lists) too_long = (p.cand_count[first + tid] & ~kCandScoreKeys) > static_cast<uint32_t>(4 * kFastKeys);
    fast = fast && !__syncthreads_or(too_long);
  }
  if (fast) {
    // one warp per list (register sort), then the rank-merge of the problem's lists
    for (int t = warp; t < n_lists; t += kSelThreads / 32) {
      const int cnt = warp_sort_list(p.cand, p.cand_count, p.list_cap, first + t, lim, p.K, keys + t * kFastPad, lane);
      if (lane == 0) s_cnt[t] = cnt;
    }
    __syncthreads();
    int total = 0;
    for (int o = 0; o < n_lists; ++o) total += s_cnt[o];
    have = min(K, total);
    if (n_lists > 1) {
      block_merge_lists<kSelThreads>(keys, s_cnt, n_lists, K, keys + n_lists * kFastPad);
      __syncthreads();
      top = keys + n_lists * kFastPad;
    }
  } else {
    have = block_select_sorted<kSelThreads>(p.cand, p.cand_count, p.list_cap, first, n_lists, K, lim, keys, ns, &s_n);
  }
  if (is_main) {
    // every key of a main list has score > thresh: counts = number of keys
    const int b = q;
    for (int j = tid; j < K; j += kSelThreads) {
      const size_t row = static_cast<size_t>(b) * K + j;
      const bool valid = j < have;
      p.score[row] = valid ? key_score(top[j]) : 0.f;
      p.flat[row] = valid ? static_cast<int32_t>(key_flat(top[j])) : -1;
    }
    if (tid == 0) p.counts[b] = have;
  } else if (tid < 32) {
    warp_fill_kpt(p, static_cast<size_t>(q - n_main) * K, top, have, reinterpret_cast<uint32_t*>(keys + ns), tid);
  }
}

static int select_ns(int K) {
  int ns = 4 * next_pow2(K);
  return ns < 1024 ? 1024 : ns;
}

size_t select_smem_bytes(int C, int Cv, int Sp, int list_cap, int K) {
  (void)C; (void)Cv; (void)Sp; (void)list_cap;       // lists that do not fit the sort buffer are merged in rounds
  return static_cast<size_t>(select_ns(K)) * 8 + (3 * static_cast<size_t>(K) + 8) * 4;
}

int launch_select(const SelectParams& p, cudaStream_t s) {
  const size_t smem = select_smem_bytes(p.C, p.Cv, p.Sp, p.list_cap, p.K);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const unsigned grid = static_cast<unsigned>((p.C > 0 ? p.B : 0) + p.B * p.Cv);
  select_kernel<<<grid, kSelThreads, smem, s>>>(p, select_ns(p.K));
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rtm3d
