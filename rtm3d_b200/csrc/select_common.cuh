// Selection of a problem's K best keys from the candidate lists the scan kernel wrote (scan_planes.cu), shared by the
// select kernel (select.cu) and the fused select + post kernel (postproc.cu).
//
// A selection problem = the C planes of an image (flat top-K over C*H*W, models/model.py:87-98) or one keypoint plane
// (per-channel top-K, models/model.py:109-114); its candidates are the lists of its strips, each of which provably contains
// the strip's K best peaks.  Keys arrive as (ordered logit, index) pairs: the sigmoid of models/model.py:85,107 is evaluated
// HERE, for the few hundred listed pixels only, and pixels at or below the score floor drop out (strict `score > thresh`,
// models/model.py:91; 0.0 for keypoint planes: zero-score pixels are fillers, not peaks).  The order is the canonical
// (score desc, index asc) = descending u64 key order (common.cuh), what torch.topk yields on CUDA.
#pragma once
#include "common.cuh"
#include "params.h"

namespace rtm3d {

// In-place descending bitonic sort of a[0..npad) in shared memory, npad a power of two, by all NT threads of the block.
// One thread per PAIR; stages with partner distance <= 32 touch 64-element blocks that one warp owns, so a warp barrier is
// enough between them (a 256-key sort has 5 block barriers instead of 36).
template <int NT>
__device__ __forceinline__ void block_sort_desc(uint64_t* a, int npad) {
  const int tid = threadIdx.x;
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (npad >> 1); t += NT) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int q = i | j;
        const uint64_t x = a[i], y = a[q];
        const bool desc = (i & k) == 0;
        if (desc ? (x < y) : (x > y)) { a[i] = y; a[q] = x; }
      }
      if (j > 32) __syncthreads(); else __syncwarp();
    }
    // the next merge starts with a partner distance of k: block-wide unless the pairs stay inside one warp's elements
    if (k >= 64) __syncthreads();
  }
  __syncthreads();
}

// The K best keys of a selection problem, sorted, in buf[0..have): returns have = min(K, number of valid candidates).
//   first / n_lists : the problem's candidate lists (consecutive strips)
//   buf             : shared memory, `ns` keys (a power of two >= 2 * next_pow2(K))
//   s_n             : one shared word of scratch
// Lists that do not fit the buffer are merged in rounds (sort, keep the K best, refill).
template <int NT>
__device__ __forceinline__ int block_select_sorted(const unsigned long long* cand, const uint32_t* cand_count, int list_cap, int first,
                                                   int n_lists, int K, float lim, uint64_t* buf, int ns, uint32_t* s_n) {
  const int tid = threadIdx.x;
  int have = 0;
  int l = 0, done_in_list = 0;
  bool sorted_once = false;
  while (true) {
    if (tid == 0) *s_n = static_cast<uint32_t>(have);
    __syncthreads();
    // take raw keys while they are guaranteed to fit
    int room = ns - have;
    while (l < n_lists && room > 0) {
      const uint32_t raw = cand_count[first + l];
      const int cnt = static_cast<int>(raw & ~kCandScoreKeys);
      const bool score_keys = (raw & kCandScoreKeys) != 0u;
      const int take = min(cnt - done_in_list, room);
      const unsigned long long* src = cand + static_cast<size_t>(first + l) * list_cap + done_in_list;
      for (int i = tid; i < take; i += NT) {
        unsigned long long k = src[i];
        bool valid = true;
        if (!score_keys) {
          const float sc = sigmoid_ref(f32_unord(static_cast<uint32_t>(k >> 32)));
          valid = sc > lim;
          k = make_key(sc, key_flat(k));
        }
        if (valid) buf[atomicAdd(s_n, 1u)] = k;
      }
      room -= take;
      done_in_list += take;
      if (done_in_list == cnt) { ++l; done_in_list = 0; }
    }
    __syncthreads();
    const int n = static_cast<int>(*s_n);
    const bool more = l < n_lists;
    if (n == have && sorted_once && !more) break;               // nothing new since the last sort
    int npad = 64;
    while (npad < n) npad <<= 1;
    for (int i = n + tid; i < npad; i += NT) buf[i] = 0ull;
    __syncthreads();
    block_sort_desc<NT>(buf, npad);
    sorted_once = true;
    have = min(n, K);
    if (!more) break;
  }
  return have;
}

// ---------------------------------------------------------------------------------------------------------------
// Fast path: one warp sorts one list in registers, the lists of a problem are rank-merged.
constexpr int kFastKeys = 256;               // keys of one list a round of the register sort takes (8 per lane)
constexpr int kFastPad = 128;                // sorted keys kept per list (>= K on the fast path)
constexpr int kFastLists = 16;               // lists per CTA on the fast path
constexpr int kFastProblems = 4;             // selection problems per CTA on the fast path

// Descending bitonic sort of 256 keys held 8 per lane (network element e = lane * 8 + k) in registers: partner distances below
// 8 are exchanges inside a lane, the others one 64-bit shuffle per key.
__device__ __forceinline__ void warp_sort256_desc(unsigned long long (&x)[8], int lane) {
#pragma unroll
  for (int kk = 2; kk <= 256; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 8) {
        const int lj = j >> 3;                                   // partner lane distance
        const bool lower = (lane & lj) == 0;
        const bool desc = (lane & (kk >> 3)) == 0;               // (e & kk) == 0  (kk = 256: always)
        const bool take_max = lower == desc;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, x[k], lj);
          const bool gt = x[k] > o;
          x[k] = (gt == take_max) ? x[k] : o;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if ((k & j) == 0) {
            const int q = k | j;
            const bool desc = (kk < 8) ? ((k & kk) == 0) : ((lane & (kk >> 3)) == 0);
            const unsigned long long a = x[k], c = x[q];
            const bool keep = (a > c) == desc;                    // a stays in front
            x[k] = keep ? a : c;
            x[q] = keep ? c : a;
          }
        }
      }
    }
  }
}

// number of keys of the descending list r[0..len) that are larger than `key`
__device__ __forceinline__ int count_greater_u64(const uint64_t* r, int len, uint64_t key) {
  int lo = 0, hi = len;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (r[mid] > key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One warp: the kFastPad best (score, index) keys of candidate list `strip`, sorted, into dst[0..kFastPad) (zeros behind the
// valid ones); returns min(valid, kFastPad).  (logit, index) keys become (score, index) keys here -- the sigmoid of
// models/model.py:85,107 for the listed pixels -- and pixels at or below the score floor `lim` drop out.  The first round
// takes 256 keys; a longer list goes on in rounds of 128 new keys merged with the 128 best so far.
__device__ __forceinline__ int warp_sort_list(const unsigned long long* cand, const uint32_t* cand_count, int list_cap, int strip, float lim,
                                              uint64_t* dst, int lane) {
  const uint32_t raw = cand_count[strip];
  const int cnt = static_cast<int>(raw & ~kCandScoreKeys);
  const bool score_keys = (raw & kCandScoreKeys) != 0u;
  const unsigned long long* src = cand + static_cast<size_t>(strip) * list_cap;
  int valid = 0;
  for (int base = 0; base == 0 || base < cnt; base += (base == 0 ? kFastKeys : kFastPad)) {
    unsigned long long x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned long long key = 0ull;
      if (base != 0 && k < 4) {
        key = dst[k * 32 + lane];                                // the best so far (any order)
      } else {
        const int i = base + (base == 0 ? k : k - 4) * 32 + lane;
        if (i < cnt) {
          key = src[i];
          if (!score_keys) {
            const float sc = sigmoid_ref(f32_unord(static_cast<uint32_t>(key >> 32)));
            key = sc > lim ? make_key(sc, key_flat(key)) : 0ull;
          }
        }
        valid += key != 0ull ? 1 : 0;
      }
      x[k] = key;
    }
    warp_sort256_desc(x, lane);
    __syncwarp();
    if (lane < kFastPad / 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[lane * 8 + k] = x[k];
    }
    __syncwarp();
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  return min(valid, kFastPad);
}

// Rank-merge of `nl` sorted lists (lists[l * kFastPad ..], cnt[l] valid keys each) into top[0..K): a key's rank = its
// position in its own list + the number of larger keys in the others (keys are distinct).  All NT threads; the caller
// synchronises afterwards.
template <int NT>
__device__ __forceinline__ void block_merge_lists(const uint64_t* lists, const int* cnt, int nl, int K, uint64_t* top) {
  for (int idx = threadIdx.x; idx < nl * kFastPad; idx += NT) {
    const int li = idx / kFastPad, pos = idx - li * kFastPad;
    if (pos >= cnt[li]) continue;
    const uint64_t key = lists[li * kFastPad + pos];
    int rk = pos;
    for (int o = 0; o < nl; ++o)
      if (o != li) rk += count_greater_u64(lists + o * kFastPad, cnt[o], key);
    if (rk < K) top[rk] = key;
  }
}

}  // namespace rtm3d
