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

// One warp: the best (score, index) keys of candidate list `strip`, sorted, into dst[0..kFastPad) (zeros behind the valid
// ones); returns the number of valid keys stored (<= kFastPad).  Only the list's K best matter to the caller (K <= kFastPad):
// dst[0..min(K, valid)) are exactly the list's best, what follows is sorted and smaller but need not be the next best.
// (logit, index) keys become (score, index) keys here -- the sigmoid of models/model.py:85,107 for the listed pixels -- and
// pixels at or below the score floor `lim` drop out.  The first 256 keys go through the register sort; a longer list goes on
// in rounds of 128 new keys, of which only those above the K-th best so far are inserted (one shuffle + four
// compare-selects per lane and insertion: a list of 300 keys inserts ~17, a second full sort would cost as much as the first).
__device__ __forceinline__ int warp_sort_list(const unsigned long long* cand, const uint32_t* cand_count, int list_cap, int strip, float lim,
                                              int K, uint64_t* dst, int lane) {
  const uint32_t raw = cand_count[strip];
  const int cnt = static_cast<int>(raw & ~kCandScoreKeys);
  const bool score_keys = (raw & kCandScoreKeys) != 0u;
  const unsigned long long* src = cand + static_cast<size_t>(strip) * list_cap;
  auto load_key = [&](int i) -> unsigned long long {
    unsigned long long key = 0ull;
    if (i < cnt) {
      key = src[i];
      if (!score_keys) {
        const float sc = sigmoid_ref(f32_unord(static_cast<uint32_t>(key >> 32)));
        key = sc > lim ? make_key(sc, key_flat(key)) : 0ull;
      }
    }
    return key;
  };
  int stored;
  {
    unsigned long long x[8];
    int valid = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[k] = load_key(k * 32 + lane);
      valid += x[k] != 0ull ? 1 : 0;
    }
    warp_sort256_desc(x, lane);
    if (lane < kFastPad / 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[lane * 8 + k] = x[k];
    }
    __syncwarp();
    stored = min(__reduce_add_sync(0xffffffffu, valid), kFastPad);
  }
  if (cnt > kFastKeys) {
    // r[i] = dst[4 * lane + i]: the sorted list, four consecutive keys per lane
    unsigned long long r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = dst[4 * lane + i];
    for (int base = kFastKeys; base < cnt; base += kFastPad) {
      const unsigned long long pivot = stored >= K ? dst[K - 1] : 0ull;      // (dst is current: written back after every round)
      unsigned long long x[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = load_key(base + k * 32 + lane);
      int inserted = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned mask = __ballot_sync(0xffffffffu, x[k] > pivot);             // (valid keys are > 0)
        inserted += __popc(mask);
        while (mask) {
          const int from = __ffs(mask) - 1;
          mask &= mask - 1u;
          const unsigned long long c = __shfl_sync(0xffffffffu, x[k], from);
          // every key below c moves one place down (the last one falls off), c takes the place that opens
          unsigned long long left = __shfl_up_sync(0xffffffffu, r[3], 1);
          if (lane == 0) left = ~0ull;
#pragma unroll
          for (int i = 3; i >= 0; --i) {
            const unsigned long long l = i == 0 ? left : r[i - 1];
            r[i] = r[i] > c ? r[i] : (l > c ? c : l);
          }
        }
      }
      if (inserted) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[4 * lane + i] = r[i];
        __syncwarp();
        stored = min(stored + inserted, kFastPad);
      }
    }
  }
  return stored;
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
