// Persistent plane-streaming decode kernel for sm_100a.
//
// Work item = one strip (H/split rows) of
//   the C main planes of image b  -> selection problem "image b"      : flat top-K over C*H*W      (models/model.py:87-98)
//   keypoint plane (b, c < Cv)    -> selection problem "plane (b,c)"  : per-channel top-K over H*W (models/model.py:109-114)
// One CTA per SM walks the items blockIdx.x, blockIdx.x + gridDim.x, ...  Both heat-maps of a batch can be decoded by
// ONE launch (hm_main and hm_kpt both set); the two public entry points use the same kernel with one of them absent.
//
// Warp roles (one CTA = kAWarps + kBWarps + kFinWarps + 1 warps):
//   producer (1 lane)  streams every item as chunks of `chunk_rows` rows plus the row above and below (rows of an NCHW
//                      plane are contiguous: a chunk is ONE bulk async copy, SASS UBLKCP) into a shared-memory ring, running
//                      ahead across item boundaries; flow control by full/empty mbarriers.
//   A-warps            phase A: per 16-byte group one LDS.128, a max and one compare against the running LOGIT threshold
//                      of the item (four groups per lane in flight); groups holding a survivor are recorded in the
//                      stage's worklist (one shared-memory atomic per 128 groups).
//   B-warps            phase B: as soon as a chunk's worklist is complete they pull dense batches of 32 recorded groups
//                      (one group per lane): exact 3x3 peak test (utils/model_utils.py:17-26 applied to the sigmoid of
//                      models/model.py:85, decided in the logit domain where that is provably the same), sigmoid, score
//                      threshold; the candidate key (score, index) is appended to the item's list and counted in a score
//                      histogram (128 bins per octave).  Every 48 appended keys the logit threshold is re-derived from the
//                      histogram: the lower edge of the highest bin with >= K candidates at or above it bounds the K-th
//                      best from below, so exactness never depends on timing.  The stage returns to the producer when
//                      every B-warp has left it.
//                      A plane STARTS at the threshold remembered from the previous plane of the same index (a few bins
//                      lower) instead of -inf.  The finisher checks that at least K candidates were found above that
//                      start; if not, the plane is redone without speculation in a second pass, so the result is exact
//                      either way.
//   finishers          take over an item once every scanner has left it (hist/list are double-buffered, the scanners go
//                      straight on to the next item): cut the list at the final threshold, sort the survivors, then either
//                      emit the rows (problem with one part) or publish the part's top-K and let the last part to arrive
//                      merge and write the selection (score, flat index).  The gathers of the regression maps, sub-pixel
//                      add, vertex regress and 2D box (models/model.py:47-50,63-73,117-132) follow in two small, wide
//                      epilogue kernels (postproc.cu).
//
// Adversarial inputs (plateaus, saturated or sorted maps) can produce more candidates than the list holds.  Then one
// scanner warp takes the lock, waits until every handed-out slot is written, selects the exact K-th key (radix select),
// compacts the list and publishes that key as a second, exact filter (`kstar`).  Slow, but bounded and still exact.
#include <cuda_runtime.h>

#include "common.cuh"
#include "params.h"
#include "scan_common.cuh"

namespace rtm3d {

constexpr int kAWarps = 8;               // threshold-filter warps (phase A), two per SM sub-partition
constexpr int kAGroups = 1;              // A-warp groups: group g scans the chunks with (chunk number % kAGroups) == g, so that
                                         // kAGroups chunks are being filtered at any time (one group of all eight: half the scan latency per chunk)
constexpr int kAPerGroup = kAWarps / kAGroups;
constexpr int kBWarps = 7;               // peak-test / candidate warps (phase B)
constexpr int kFinWarps = 4;              // each finishes whole items on its own (ticket order)
// Warp ids: finishers 0..3, B-warps 4..10, producer 11, A-warps 12..19 (two A-warps per SM sub-partition).  The order of
// the roles had no measurable effect on B200.
constexpr int kFinWarp0 = 0;
constexpr int kBWarp0 = kFinWarps;
constexpr int kProdWarp = kBWarp0 + kBWarps;
constexpr int kAWarp0 = kProdWarp + 1;
constexpr int kPlaneThreads = (kAWarp0 + kAWarps) * 32;
constexpr int kAUnroll = 4;              // 16-byte groups per lane per phase-A iteration
constexpr int kNBuf = 4;                 // selection buffers (hist + list): items in flight between the A-warps and the finishers
constexpr int kBufShift = 2;
constexpr int kHistBins = 1664;          // score histogram: 64 bins per octave over [2^-24, 1] (1537 used; 13 * 128)
constexpr int kScoreShift = 17;
constexpr uint32_t kScoreBase = 0x33800000u >> kScoreShift; // bits of 2^-24
constexpr int kMaxStages = 4;
constexpr int kDefaultStages = 2;        // few, large stages: a chunk costs every role a fixed ~1 us of hand-offs whatever its size
                                         // (measured: 2 x 64 KB beats 3 x 43 KB and 4 x 33 KB by 2-4 % on cfg2/4/5)
constexpr int kMaxPlanes = 64;           // plane indices with a remembered threshold (speculation)
constexpr int kSpecMargin = 1;           // bins below the remembered boundary the speculative threshold starts at
constexpr int kUpdateEvery = 48;         // appended keys between two threshold updates

struct PlaneGeom {
  int stages;         // ring depth (power of two)
  int speculate;      // 1: start items at the threshold remembered from the previous item of the same plane index
  int chunk_rows;     // centre rows per chunk
  int copy_rows;      // rows per bulk copy (a chunk = ceil((chunk_rows + 2) / copy_rows) copies on one barrier)
  int stage_bytes;    // (chunk_rows + 2) * row_bytes
  int row_bytes;
  int gpr;            // 16-byte groups per row
  unsigned gpr_magic; // ceil(2^32 / gpr)
  int split;          // strips per plane (power of two)
  int split_shift;    // log2(split)
  int rows_lo, nch_lo, nch_hi;  // rows of the shorter strips and chunks per strip (shorter / longer strips)
  int list_cap;       // keys per candidate list
  int fin_cap;        // keys per finisher buffer (two per finisher warp)
  int wl_cap;         // worklist entries per stage (kAPerGroup segments)
  int wl_seg;         // worklist entries per A-warp segment (= the groups one A-warp scans in a chunk, rounded up)
  int n_items;
  int max_ctas;       // 0 = one CTA per SM
  int debug;          // timing experiments only (results are then wrong): 1 no B batches, 2 finisher stops after the release,
                      // 3 = 1+2, 4 A records nothing, 5 finisher stops after the sort, 6 ... after publish/merge, 7 = 1+2+4,
                      // 8 producer + A only, 9 / 10 no Tier A / Tier B emit, 12 = 8+4
  unsigned smem;
};


// Developer counters (decode_planes_kernel<T, true> only; tools/plane_stats.py): summed over all CTAs of a launch.
enum StatSlot { kStItems = 0, kStRetried, kStWlEntries, kStBatches, kStPushed, kStUpdates, kStCompactions, kStWaitBufFree,
                kStWaitFull, kStWaitScanned, kStFinBusy, kStFinWait, kStATotal, kStProdWait, kStBBusy, kStBTotal,
                kStFinBoundary, kStFinCompact, kStFinRelease, kStFinSort, kStFinPublish, kStFinEmit, kStALoop, kStASetup, kStSlots };
#define RTM3D_ACC(slot, val) do { if constexpr (STATS) acc_[slot] += static_cast<long long>(val); } while (0)
#define RTM3D_CLK() (STATS ? clock64() : 0ll)
#define RTM3D_FLUSH(slot) do { if constexpr (STATS) { if (p.stats && acc_[slot] != 0) { atomicAdd(&p.stats[slot], static_cast<unsigned long long>(acc_[slot])); acc_[slot] = 0; } } } while (0)
// Per-chunk timestamps of CTA 0 (tools/trace_cta.py): compiled in only with -DRTM3D_CHUNK_TRACE -- even a predicated-off
// mark in the per-chunk loops of every role costs ~10% of the kernel.
#ifdef RTM3D_CHUNK_TRACE
#define RTM3D_MARK(kind, idx) do { if (!STATS && p.stats && blockIdx.x == 0 && (idx) < 96u) p.stats[(kind) * 96 + (idx)] = static_cast<unsigned long long>(clock64()); } while (0)
#else
#define RTM3D_MARK(kind, idx) do { } while (0)
#endif
#define RTM3D_TRACE(ev) do { if constexpr (STATS) { if (p.stats && blockIdx.x == 0 && lane == 0 && trace_n < 960) { p.stats[64 + trace_n] = (static_cast<unsigned long long>(ev) << 56) | (static_cast<unsigned long long>(clock64()) & 0x00FFFFFFFFFFFFFFull); ++trace_n; } } } while (0)
#define RTM3D_FIN_LAP(slot) do { if constexpr (STATS) { if (lane == 0) { const long long now_ = clock64(); acc_[slot] += now_ - lap_; lap_ = now_; } } } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Per-item selection state (double-buffered) and CTA control block.
struct __align__(16) Sel {
  volatile uint32_t reserve;    // list slots handed out (runs past list_cap when the list is full: those are void)
  uint32_t lock;                // threshold update / compaction mutex
  volatile uint32_t last_upd;   // value of `reserve` at the last threshold update
  volatile float t_filter;      // running logit threshold
  volatile unsigned long long kstar;  // exact key bound after a compaction (0 = none)
  volatile int spec_bin;        // speculative start threshold of the item (histogram bin, -1 = none); set by the finisher
  volatile float spec_t;        // its logit bound filter_from_bin(spec_bin) (-inf when none), computed once per item
  volatile int last_bin;        // histogram boundary at the last threshold update (-1: fewer than K keys so far)
  uint32_t b_done;              // B-warps that have finished the item's last chunk
  uint32_t pad[2];
};

struct __align__(16) PlaneCtl {
  unsigned long long full[kMaxStages];      // producer -> A-warps: chunk landed
  unsigned long long scanned[kMaxStages];   // A-warps -> B-warps: phase A of the chunk done, worklist complete
  unsigned long long empty[kMaxStages];     // B-warps -> producer: phase B done, stage free
  unsigned long long item_done[kNBuf];
  unsigned long long buf_free[kNBuf];
  Sel sel[kNBuf];
  volatile uint32_t wl_count[kMaxStages][kAPerGroup];   // worklist entries per stage and A-warp (written by that A-warp after its scan)
  uint32_t wl_next[kMaxStages];             // next batch of the stage's worklist to hand out (reset with wl_count)
  uint32_t rsel[kNBuf + kFinWarps][264];   // radix-select scratch: [buf] B-warp compaction of that buffer, [kNBuf + w] finisher warp w
  uint32_t fin_next;            // next item ordinal to hand to a finisher warp
  volatile uint32_t fin_released[kNBuf];   // how many times the finishers have handed selection buffer [buf] back
  uint32_t n_retry;             // items of this CTA whose speculative start threshold failed (redone in pass 1)
  volatile int guess_bin[kMaxPlanes];       // final boundary bin of the last finished item of each plane index (-1 = unknown)
};

struct ItemInfo {
  int b, slot, strip;           // slot 0 = the image's main planes (when C > 0), the following slots = keypoint planes
  int plane;                    // index into the per-plane threshold memory: 0 for the main item, C + kc for keypoint plane kc
  int kc;                       // keypoint plane (slot items), -1 for the main item
  bool is_main;
  int nplanes;                  // planes streamed by this item: C for the main item (one flat top-K over C*H*W), else 1
  int ys, ye;                   // rows of the strip
  int cpp;                      // chunks per plane
  int nchunks;                  // nplanes * cpp
  const unsigned char* base;    // base address of the item's first plane
  size_t plane_bytes;
  uint32_t flat_base;           // added to y*W+x to form the key's flat index (first plane)
  uint32_t hw;                  // H*W: flat_base advances by it from plane to plane
};

// Items are numbered heavy-first: the main items (one per image and strip: C planes in one flat top-K) and then the
// keypoint items (one plane each), the strips of one problem next to each other:
//   main item  i           = b * split + strip                      (i < n_main)
//   keypoint   n_main + j,   j = (b * Cv + kc) * split + strip
// A CTA takes main items blockIdx.x, blockIdx.x + gridDim.x, ... and then a contiguous run of keypoint items sized so
// that every CTA streams about the same number of planes (a plain stride over image-major items gave some CTAs every
// main item they could get and others none: 26 vs 17 planes per CTA at B=256, C=3, Cv=9 on 148 SMs).
struct CtaItems {
  int n_items, n_main, cta, grid;
  int mains;            // main items of this CTA
  int kpt_first;        // its first keypoint item (index j) ...
  int count;            // ... and its total number of items
  __device__ __forceinline__ void init(const PlaneParams& p, const PlaneGeom& g, int cta_, int grid_) {
    n_items = g.n_items; cta = cta_; grid = grid_;
    n_main = p.C > 0 ? p.B * g.split : 0;
    const int n_kpt = n_items - n_main;
    const int m_lo = n_main / grid, n_hi = n_main - m_lo * grid;
    const long long units = static_cast<long long>(n_main) * p.C + n_kpt;
    const int target = static_cast<int>((units + grid - 1) / grid);               // planes per CTA, rounded up
    const int cap_hi = max(0, target - p.C * (m_lo + 1)), cap_lo = max(0, target - p.C * m_lo);
    mains = m_lo + (cta < n_hi ? 1 : 0);
    const long long first = cta < n_hi ? static_cast<long long>(cta) * cap_hi
                                       : static_cast<long long>(n_hi) * cap_hi + static_cast<long long>(cta - n_hi) * cap_lo;
    const long long left = static_cast<long long>(n_kpt) - first;
    const int cap = cta < n_hi ? cap_hi : cap_lo;
    kpt_first = static_cast<int>(first < n_kpt ? first : n_kpt);
    count = mains + static_cast<int>(left < 0 ? 0 : (left < cap ? left : cap));
  }
  // the CTA's kl-th item (n_items past the end)
  __device__ __forceinline__ int at(int kl) const {
    if (kl < mains) return cta + kl * grid;
    return kl < count ? n_main + kpt_first + (kl - mains) : n_items;
  }
};

struct ItemIter {
  const CtaItems* ci;
  int kl, item;
  __device__ __forceinline__ void init(const CtaItems* c) { ci = c; kl = 0; item = c->at(0); }
  __device__ __forceinline__ void next() { ++kl; item = ci->at(kl); }
};

// index into the per-plane threshold memory of an item: 0 for a main item, C + kc for keypoint plane kc
__device__ __forceinline__ int item_plane(const PlaneParams& p, const PlaneGeom& g, int n_main, int item) {
  if (item < n_main) return 0;
  return p.C + ((item - n_main) >> g.split_shift) % p.Cv;
}

__device__ __forceinline__ ItemInfo decode_item(const PlaneParams& p, const PlaneGeom& g, const ItemIter& ii, int elem_bytes) {
  ItemInfo it;
  const int n_main = ii.ci->n_main;
  it.is_main = ii.item < n_main;
  if (it.is_main) {
    it.b = ii.item >> g.split_shift;
    it.strip = ii.item & (g.split - 1);
    it.slot = 0;
    it.kc = -1;
  } else {
    const int j = ii.item - n_main, u = j >> g.split_shift;
    it.strip = j & (g.split - 1);
    it.b = u / p.Cv;
    it.kc = u - it.b * p.Cv;
    it.slot = it.kc + (p.C > 0 ? 1 : 0);
  }
  it.plane = it.is_main ? 0 : p.C + it.kc;
  it.nplanes = it.is_main ? p.C : 1;
  it.ys = (it.strip * p.H) >> g.split_shift;
  it.ye = ((it.strip + 1) * p.H) >> g.split_shift;
  it.cpp = (it.ye - it.ys == g.rows_lo) ? g.nch_lo : g.nch_hi;
  it.nchunks = it.nplanes * it.cpp;
  it.hw = static_cast<uint32_t>(p.H * p.W);
  it.plane_bytes = static_cast<size_t>(p.H) * p.W * elem_bytes;
  it.flat_base = 0u;
  if (it.is_main) it.base = reinterpret_cast<const unsigned char*>(p.hm_main) + static_cast<size_t>(it.b) * p.C * it.plane_bytes;
  else it.base = reinterpret_cast<const unsigned char*>(p.hm_kpt) + (static_cast<size_t>(it.b) * p.Cv + it.kc) * it.plane_bytes;
  return it;
}

// ---------------------------------------------------------------------------------------------------------------
// Score histogram: bin = top 15 bits of the (positive) fp32 score, offset so that bin 0 collects everything below
// 2^-24.  Bin edges are exact floats, monotone in the score and therefore in the sort key.
__device__ __forceinline__ uint32_t score_bin(float sc) {
  const uint32_t h = __float_as_uint(sc) >> kScoreShift;
  return h > kScoreBase ? h - kScoreBase : 0u;
}
__device__ __forceinline__ uint32_t bin_edge_bits(int bin) { return (static_cast<uint32_t>(bin) + kScoreBase) << kScoreShift; }

// Largest usable logit bound T for a histogram boundary bin (bin >= 1):  x < T  =>  sigmoid_ref(x) < edge(bin) strictly.
// T = logit(edge) - delta, delta >= 2^-16/(1-edge): 64x the worst few-ulp error of the computed sigmoid
// (tests/test_sigmoid_gpu.py checks the property for every bin through rtm3d_threshold_table).
static __device__ __noinline__ float filter_from_bin(int bin) {
  if (bin < 1) return -INFINITY;
  const float se = __uint_as_float(bin_edge_bits(bin));
  if (se >= 1.0f) return 15.0f;                        // sigmoid_ref(x) < 1.0f for every x < 15
  const double sd = static_cast<double>(se);
  const double L = log(sd / (1.0 - sd));
  const double e = 1.52587890625e-05;                  // 2^-16
  const double d = e / (1.0 - sd) + e * fabs(L) + e;
  return __double2float_rd(L - d);
}

// Highest bin `bs` with  sum(hist[bs..]) >= K  (one warp; returns -1 when fewer than K keys are counted; bin 0 has no
// lower edge and means "no threshold").
// Counts only grow while the item is scanned, so a concurrent scan still yields a valid lower bound.
__device__ __forceinline__ int warp_hist_boundary(const uint32_t* hist, int K, int lane) {
  uint32_t cum = 0;
  for (int blk = kHistBins / 128 - 1; blk >= 0; --blk) {
    const uint4 h = reinterpret_cast<const uint4*>(hist)[blk * 32 + lane];   // bins blk*128 + 4*lane + {0,1,2,3}
    const uint32_t s = h.x + h.y + h.z + h.w;
    uint32_t suf = s;                                                         // sum over lanes >= lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_down_sync(0xffffffffu, suf, d);
      if (lane + d < 32) suf += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, suf, 0);
    if (cum + total >= static_cast<uint32_t>(K)) {
      const uint32_t ok = __ballot_sync(0xffffffffu, cum + suf >= static_cast<uint32_t>(K));
      const int owner = 31 - __clz(ok);                                       // highest lane whose suffix reaches K
      int bin = -1;
      if (lane == owner) {
        uint32_t above = cum + suf - s;                                       // keys in bins above this lane's four
        const uint32_t c[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int q = 3; q >= 0; --q) {
          above += c[q];
          if (bin < 0 && above >= static_cast<uint32_t>(K)) bin = blk * 128 + 4 * lane + q;
        }
      }
      return __shfl_sync(0xffffffffu, bin, owner);
    }
    cum += total;
  }
  return -1;
}

// K-th largest of the n keys in list[] (zeros = empty slots) by MSB-first radix select; ONE warp.  Returns 0 when the
// list holds fewer than K non-zero keys.  Keys are distinct.
static __device__ __noinline__ unsigned long long warp_radix_kth(const unsigned long long* list, int n, int K,
                                                                  uint32_t* hist, int lane) {
  unsigned long long prefix = 0, mask = 0;
  uint32_t need = static_cast<uint32_t>(K);
#pragma unroll 1
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
#pragma unroll 1
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
      const unsigned long long k = list[i];
      if (k != 0ull && (k & mask) == prefix) atomicAdd(&hist[static_cast<uint32_t>(k >> shift) & 255u], 1u);
    }
    __syncwarp();
    // lane l owns digits 255-8l .. 248-8l (descending)
    uint32_t c[8], s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { c[q] = hist[255 - 8 * lane - q]; s += c[q]; }
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total < need) return 0ull;          // fewer than K keys under this prefix: only possible in pass 0
    const uint32_t excl = incl - s;
    const bool mine = excl < need && incl >= need;
    uint32_t digit = 0, rest = 0;
    if (mine) {
      uint32_t run = excl;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (run < need && run + c[q] >= need) { digit = 255 - 8 * lane - q; rest = need - run; }
        run += c[q];
      }
    }
    const uint32_t owner = __ballot_sync(0xffffffffu, mine);
    const int src = __ffs(owner) - 1;
    digit = __shfl_sync(0xffffffffu, digit, src);
    need = __shfl_sync(0xffffffffu, rest, src);
    prefix |= static_cast<unsigned long long>(digit) << shift;
    mask |= 0xFFull << shift;
    __syncwarp();
  }
  return prefix;
}

// Descending bitonic sort of one key per lane (registers + shuffles).
__device__ __forceinline__ unsigned long long warp_sort_desc_u64(unsigned long long key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, j);
      const bool lower_lane = (lane & j) == 0;
      const bool desc_block = (lane & k) == 0;
      const bool mine_ge = key >= o;
      const bool keep_mine = (lower_lane == desc_block) ? mine_ge : !mine_ge;
      if (!keep_mine) key = o;
    }
  }
  return key;
}

// number of keys in the descending run r[0..len) that are larger than `key`
__device__ __forceinline__ int count_greater(const unsigned long long* r, int len, unsigned long long key) {
  int lo = 0, hi = len;
#pragma unroll 1
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (r[mid] > key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------------------------
// Scanner side.

// Compaction of a full list (caller holds L.lock, warp-converged, L.reserve >= cap so no new slot is handed out).
static __device__ __noinline__ void compact_list(Sel& L, unsigned long long* list, int cap, int K, uint32_t* rsel, int lane) {
  // every slot below cap has an owner: wait until all of them are written (keys are non-zero)
  {
    const long long t0 = clock64();
    while (true) {
      bool all = true;
#pragma unroll 1
      for (int i = lane; i < cap; i += 32) all &= (reinterpret_cast<volatile unsigned long long*>(list)[i] != 0ull);
      if (__all_sync(0xffffffffu, all)) break;
      __nanosleep(64);
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
  __threadfence_block();
  const unsigned long long kth = warp_radix_kth(list, cap, K, rsel, lane);
  // stable in-place compaction (writes trail reads): keep keys >= kth
  int w = 0;
#pragma unroll 1
  for (int base = 0; base < cap; base += 32) {
    const int i = base + lane;
    unsigned long long k = 0ull;
    if (i < cap) k = list[i];
    const bool keep = (k != 0ull) && (k >= kth);
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) list[w + __popc(bal & ((1u << lane) - 1u))] = k;
    w += __popc(bal);
    __syncwarp();
  }
#pragma unroll 1
  for (int i = w + lane; i < cap; i += 32) list[i] = 0ull;
  __syncwarp();
  __threadfence_block();
  if (lane == 0) {
    if (kth > L.kstar) L.kstar = kth;
    L.last_upd = static_cast<uint32_t>(w);
    __threadfence_block();
    L.reserve = static_cast<uint32_t>(w);      // last: from here on slots are handed out again
  }
  __syncwarp();
}

// filter_from_bin through the workspace's table when rtm3d_workspace_init has filled it (ft != nullptr)
__device__ __forceinline__ float filter_lookup(const float* ft, int bin) {
  if (bin < 1) return -INFINITY;
  return ft ? __ldg(ft + bin) : filter_from_bin(bin);
}

// Re-derive the item's logit threshold from its histogram (caller holds L.lock, warp-converged).
static __device__ __noinline__ void update_threshold(Sel& L, const uint32_t* hist, int K, int lane, const float* ft) {
  const uint32_t r = L.reserve;
  const int bin = warp_hist_boundary(hist, K, lane);
  if (lane == 0) {
    if (bin >= 0) {
      const float t = filter_lookup(ft, bin);
      if (t > L.t_filter) L.t_filter = t;
    }
    L.last_bin = bin;
    L.last_upd = r;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Finisher side: ONE warp finishes an item on its own (no block barriers).

// Sort m keys of a[] (descending, zero-padded to a multiple of 32) into out[]: runs of 32 sorted in registers, then every
// key's rank = its position in its run + the number of larger keys in every other run (keys are distinct).
__device__ __forceinline__ void warp_fin_sort(unsigned long long* a, int m, unsigned long long* out, int lane) {
  const int nruns = (m + 31) >> 5;
#pragma unroll 1
  for (int r = 0; r < nruns; ++r) {
    const int i = r * 32 + lane;
    unsigned long long k = (i < m) ? a[i] : 0ull;
    k = warp_sort_desc_u64(k, lane);
    a[i] = k;
  }
  __syncwarp();
#pragma unroll 1
  for (int i = lane; i < nruns * 32; i += 32) {
    const unsigned long long k = a[i];
    if (k == 0ull) continue;
    const int own = i >> 5;
    int rank = i & 31;
#pragma unroll 1
    for (int r = 0; r < nruns; ++r) {
      if (r == own) continue;
      rank += count_greater(a + r * 32, 32, k);
    }
    out[rank] = k;
  }
  __syncwarp();
}

// Tier A selection of image b: score, flat index (and the count) of the `cnt` sorted keys; rows >= cnt get score 0 and
// flat -1.  The gathers / regress / 2D box of models/model.py:47-50,63-73 run afterwards in epilogue_main_kernel, wide
// over the batch, so that their scattered reads do not stall a streaming CTA.
static __device__ __noinline__ void warp_write_main(const PlaneParams& p, int b, const unsigned long long* sorted, int cnt, int lane) {
  const int K = p.K;
#pragma unroll 1
  for (int j = lane; j < K; j += 32) {
    const size_t row = static_cast<size_t>(b) * K + j;
    const bool valid = j < cnt;
    const unsigned long long key = valid ? sorted[j] : 0ull;
    p.score[row] = valid ? key_score(key) : 0.f;
    p.flat[row] = valid ? static_cast<int32_t>(key_flat(key)) : -1;
  }
  if (lane == 0) p.counts[b] = cnt;
}

// Tier B selection of plane (b,c): score and flat index of the K candidates.  Rows cnt..K-1 are 0.0-score fillers = the
// lowest flat indices that are not positive-score peaks (what a top-K over the zero-filled peak map returns, SURVEY
// App. A).  The sub-pixel add (models/model.py:113-114, :55-57) runs afterwards in epilogue_kpt_kernel.
//   scratch: >= 3K+8 words of shared memory of this warp.
static __device__ __noinline__ void warp_write_kpt(const PlaneParams& p, int b, int c, const unsigned long long* sorted, int cnt,
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
    const size_t row = (static_cast<size_t>(b) * p.Cv + c) * K + j;
    p.kscore[row] = j < cnt ? key_score(sorted[j]) : 0.0f;
    p.kflat[row] = static_cast<int32_t>(j < cnt ? key_flat(sorted[j]) : fill[j]);
  }
  __syncwarp();
}

// DBG: the timing experiments of PlaneGeom::debug are compiled in (tools/debug_time.py); the production instantiation
// carries none of their tests -- every instruction on a role's per-chunk path counts.
// SPLIT: planes are cut into strips (the publish + merge path of the finishers is compiled in).
template <typename T, bool STATS, bool DBG, bool SPLIT>
__global__ void __launch_bounds__(kPlaneThreads, 1) decode_planes_kernel(const PlaneParams p, const PlaneGeom g) {
  const int dbg = DBG ? g.debug : 0;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ PlaneCtl ctl;
  __shared__ PlaneParams sp;     // copy for the out-of-line finisher code (keeps the kernel parameters out of local memory)

  constexpr int E = Grp<T>::E;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!STATS && p.stats && tid == 0) p.stats[3840 + 4 * blockIdx.x + 0] = pl::globaltimer_ns();
  const int W = p.W, H = p.H, K = p.K;
  const int S = g.stages;                        // power of two

  // shared carve-up: ring | worklists | hist[kNBuf] | list[kNBuf] | per finisher warp: finA (also the emit scratch), finB
  unsigned char* ring = smem;
  size_t o = static_cast<size_t>(S) * g.stage_bytes;
  unsigned short* wl_all = reinterpret_cast<unsigned short*>(smem + o);    o += static_cast<size_t>(S) * g.wl_cap * 2;
  uint32_t* hist_all = reinterpret_cast<uint32_t*>(smem + o);               o += static_cast<size_t>(kNBuf) * kHistBins * 4;
  unsigned long long* list_all = reinterpret_cast<unsigned long long*>(smem + o);  o += static_cast<size_t>(kNBuf) * g.list_cap * 8;
  unsigned long long* fin_all = reinterpret_cast<unsigned long long*>(smem + o);   // per finisher warp: finA, finB [fin_cap]

  if (tid == 0) {
    sp = p;
    for (int s = 0; s < S; ++s) {
      pl::mbar_init(pl::smem_u32(&ctl.full[s]), 1);
      pl::mbar_init(pl::smem_u32(&ctl.scanned[s]), kAPerGroup);
      pl::mbar_init(pl::smem_u32(&ctl.empty[s]), (dbg == 8 || dbg == 12) ? kAPerGroup : kBWarps);
      ctl.wl_next[s] = 0u;
    }
    for (int q = 0; q < kNBuf; ++q) {
      pl::mbar_init(pl::smem_u32(&ctl.item_done[q]), kBWarps);
      pl::mbar_init(pl::smem_u32(&ctl.buf_free[q]), 1);
      Sel& L = ctl.sel[q];
      L.reserve = 0; L.lock = 0; L.last_upd = 0; L.t_filter = -INFINITY; L.kstar = 0ull; L.spec_bin = -1; L.spec_t = -INFINITY;
      L.last_bin = -1; L.b_done = 0;
      ctl.fin_released[q] = 0u;
    }
    ctl.fin_next = 0u;
    ctl.n_retry = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid >= 64 && tid < 64 + kMaxPlanes) {
    // remembered from the previous launch; anything that is not a bin number (a workspace reused with another shape) = none
    const uint32_t gw = __ldcg(&p.guess[tid - 64]);
    ctl.guess_bin[tid - 64] = gw <= static_cast<uint32_t>(kHistBins) ? static_cast<int>(gw) - 1 : -1;
  }
#pragma unroll 1
  for (int i = tid; i < kNBuf * kHistBins; i += kPlaneThreads) hist_all[i] = 0u;
#pragma unroll 1
  for (int i = tid; i < kNBuf * g.list_cap; i += kPlaneThreads) list_all[i] = 0ull;
  __syncthreads();

  // the workspace's threshold table, if it is there (a workspace that was only zeroed falls back to computing the bounds)
  const float* ftable = (p.ftable && __float_as_uint(__ldg(p.ftable + kFilterTableWords - 1)) == kFilterTableMagic) ? p.ftable : nullptr;
  const int planes_per_img = p.C + p.Cv;
  CtaItems cta_items;
  cta_items.init(p, g, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x));
  // speculative start threshold for a plane index: a few bins below the boundary remembered for it, or for its segment
  auto spec_for_plane = [&](int pl) -> int {
    if (pl >= kMaxPlanes) return -1;
    int gb = ctl.guess_bin[pl];
    if (gb < 0) {
      const int lo = pl < p.C ? 0 : p.C, hi = min(pl < p.C ? p.C : planes_per_img, kMaxPlanes);
      int mn = 0x7fffffff;
#pragma unroll 1
      for (int q = lo; q < hi; ++q) { const int v = ctl.guess_bin[q]; if (v >= 0 && v < mn) mn = v; }
      if (mn != 0x7fffffff) gb = mn;
    }
    return gb > kSpecMargin ? gb - kSpecMargin : -1;
  };
  if (tid < kNBuf && g.speculate) {
    // the first items of this CTA start from what the previous launch remembered
    const int q = tid;
    const int item = cta_items.at(q);
    if (item < g.n_items) {
      const int sb = spec_for_plane(item_plane(p, g, cta_items.n_main, item));
      ctl.sel[q].spec_bin = sb;
      ctl.sel[q].spec_t = filter_lookup(ftable, sb);
    }
  }
  __syncthreads();
  long long acc_[STATS ? kStSlots : 1];
  if constexpr (STATS) {
#pragma unroll
    for (int i = 0; i < kStSlots; ++i) acc_[i] = 0;
  }
  // Counters that run through both passes (pass 0: every item of this CTA; pass 1: the items whose speculative start
  // threshold turned out too high, redone without speculation).
  uint32_t gq = 0;     // chunks so far (producer / A / B)
  uint32_t ring_s = 0, ring_ph = 0;   // ... and the ring stage / barrier phase parity of chunk gq (any number of stages)
  auto ring_advance = [&]() { if (++ring_s == static_cast<uint32_t>(S)) { ring_s = 0u; ring_ph ^= 1u; } };
  uint32_t k = 0;      // items so far (A / B / finishers)
  uint32_t my_ticket = 0;   // finisher warps: ordinal of the item this warp finishes next
  int trace_n = 0;
  (void)trace_n;

  if (!STATS && p.stats && tid == 0) p.stats[3840 + 4 * blockIdx.x + 1] = pl::globaltimer_ns();
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      __syncthreads();               // every role is done with pass 0; the retry flags (this CTA's own writes) are visible
      if (!STATS && p.stats && tid == 0) p.stats[3840 + 4 * blockIdx.x + 2] = pl::globaltimer_ns();
      if (ctl.n_retry == 0u) break;  // (nearly always)
    }
    ItemIter ii;
    ii.init(&cta_items);

    if (warp == kProdWarp) {
      // ================================ producer ================================
      if (lane == 0) {
        const unsigned long long stream_policy = dbg == 13 ? pl::l2_policy_evict_normal() : pl::l2_policy_evict_first();
        for (; ii.item < g.n_items; ii.next()) {
          if (pass == 1 && __ldcg(&p.retry[ii.item]) == 0u) continue;
          const ItemInfo it = decode_item(p, g, ii, static_cast<int>(sizeof(T)));
          int qq = 0, pli = 0;                          // chunk within the plane, plane within the item
          for (int q = 0; q < it.nchunks; ++q, ++gq, ring_advance()) {
            const uint32_t s = ring_s;
            if (gq >= static_cast<uint32_t>(S)) {
              const long long w0 = RTM3D_CLK();
              pl::mbar_wait(pl::smem_u32(&ctl.empty[s]), ring_ph ^ 1u, p.status, 0xE1000001u, 64);
              RTM3D_ACC(kStProdWait, RTM3D_CLK() - w0);
            }
            const int c_lo = it.ys + qq * g.chunk_rows, c_hi = min(c_lo + g.chunk_rows, it.ye);
            const int top = max(c_lo - 1, 0), bot = min(c_hi + 1, H);      // rows [top, bot) incl. the halo rows
            const uint32_t bytes = static_cast<uint32_t>(bot - top) * g.row_bytes;
            // stage row r holds image row (c_lo - 1 + r): a missing top halo leaves stage row 0 unused
            const uint32_t dst = pl::smem_u32(ring + static_cast<size_t>(s) * g.stage_bytes) +
                                 static_cast<uint32_t>(top - (c_lo - 1)) * g.row_bytes;
            const uint32_t bar = pl::smem_u32(&ctl.full[s]);
            ctl.wl_next[s] = 0u;                        // every B-warp has left the stage (empty) / nobody has entered it yet
            RTM3D_TRACE(8);
            pl::mbar_arrive_expect_tx(bar, bytes);
            {
              // the chunk goes out as a few bulk copies on one barrier: several copies in flight per stage keep the
              // copy engine's request stream dense
              const unsigned char* src = it.base + static_cast<size_t>(pli) * it.plane_bytes + static_cast<size_t>(top) * g.row_bytes;
              const uint32_t piece = static_cast<uint32_t>(g.copy_rows) * g.row_bytes;
              for (uint32_t o2 = 0; o2 < bytes; o2 += piece)
                pl::bulk_g2s(dst + o2, src + o2, min(piece, bytes - o2), bar, stream_policy);
              RTM3D_MARK(0, gq);
            }
            if (++qq == it.cpp) { qq = 0; ++pli; }
          }
        }
        RTM3D_FLUSH(kStProdWait);
      }
    } else if (warp >= kAWarp0) {
      // ================================ A-warps: threshold filter ================================
      const int gpr = g.gpr;
      const long long a_t0 = RTM3D_CLK();
      for (; ii.item < g.n_items; ii.next()) {
        if (pass == 1 && __ldcg(&p.retry[ii.item]) == 0u) continue;
        const ItemInfo it = decode_item(p, g, ii, static_cast<int>(sizeof(T)));
        const int buf = static_cast<int>(k & (kNBuf - 1));
        if (k >= static_cast<uint32_t>(kNBuf) && !(dbg == 8 || dbg == 12)) {
          const long long w0 = RTM3D_CLK();
          pl::mbar_wait(pl::smem_u32(&ctl.buf_free[buf]), ((k >> kBufShift) - 1u) & 1u, p.status, 0xE1000003u, 64);
          if (lane == 0) RTM3D_ACC(kStWaitBufFree, RTM3D_CLK() - w0);
        }
        ++k;
        const long long as0 = RTM3D_CLK();
        const Sel& L = ctl.sel[buf];
        // start threshold: the score threshold's logit (main planes) and, speculatively, a few bins below where the
        // previous item of the same plane index ended (verified by the finisher; a miss is redone in pass 1)
        float t_floor = it.is_main ? p.t0 : -INFINITY;
        t_floor = fmaxf(t_floor, L.spec_t);
        if (lane == 0) RTM3D_ACC(kStASetup, RTM3D_CLK() - as0);
        int qq = 0;                                     // chunk within the plane
        for (int q = 0; q < it.nchunks; ++q, ++gq, ring_advance()) {
          const uint32_t s = ring_s;
          if ((gq & (kAGroups - 1)) != static_cast<uint32_t>((warp - kAWarp0) / kAPerGroup)) {   // the other group's chunk
            if (++qq == it.cpp) qq = 0;
            continue;
          }
          {
            const long long w0 = RTM3D_CLK();
            pl::mbar_wait(pl::smem_u32(&ctl.full[s]), ring_ph, p.status, 0xE1000002u, 100);
            if (lane == 0) RTM3D_MARK(1 + (warp - kAWarp0) % kAPerGroup, gq);
            if (lane == 0) RTM3D_ACC(kStWaitFull, RTM3D_CLK() - w0);
          }
          const long long al0 = RTM3D_CLK();
          if (warp == kAWarp0) RTM3D_TRACE(1);
          const int c_lo = it.ys + qq * g.chunk_rows, c_hi = min(c_lo + g.chunk_rows, it.ye);
          if (++qq == it.cpp) qq = 0;
          const unsigned char* centre = ring + static_cast<size_t>(s) * g.stage_bytes + g.row_bytes;  // image row c_lo
          const int aw = (warp - kAWarp0) % kAPerGroup;          // index within the group that scans this chunk
          // this A-warp's private segment of the stage's worklist: no atomics, the fill count lives in a register
          unsigned short* wl = wl_all + static_cast<size_t>(s) * g.wl_cap + static_cast<size_t>(aw) * g.wl_seg;
          int wn = 0;
          const int ng = (c_hi - c_lo) * gpr;
          const uint32_t lt = (1u << lane) - 1u;
          // tasks of 32 groups: aw, aw + kAPerGroup, ...; kAUnroll of them per round, out-of-range groups read as -inf
          const unsigned char* lp = centre + (static_cast<size_t>(aw) * 32 + lane) * 16;
#pragma unroll 1
          for (int g0 = aw * 32 + lane; g0 - lane < ng; g0 += kAUnroll * kAPerGroup * 32, lp += kAUnroll * kAPerGroup * 512) {
            const float tf = fmaxf(L.t_filter, t_floor);
            float m[kAUnroll];
#pragma unroll
            for (int u = 0; u < kAUnroll; ++u) {
              m[u] = -INFINITY;
              if (g0 + u * kAPerGroup * 32 < ng) {
                float v[E];
                Grp<T>::load(lp + u * kAPerGroup * 512, v);
                m[u] = v[0];
#pragma unroll
                for (int i = 1; i < E; ++i) m[u] = fmaxf(m[u], v[i]);
              }
            }
            bool any = false;
#pragma unroll
            for (int u = 0; u < kAUnroll; ++u) any |= (m[u] >= tf) && (g0 + u * kAPerGroup * 32 < ng);
            if (!(dbg == 4 || dbg == 7 || dbg == 12) && __any_sync(0xffffffffu, any)) {
#pragma unroll
              for (int u = 0; u < kAUnroll; ++u) {
                const bool hit = (m[u] >= tf) && (g0 + u * kAPerGroup * 32 < ng);
                const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                if (hit) wl[wn + __popc(bal & lt)] = static_cast<unsigned short>(g0 + u * kAPerGroup * 32);
                wn += __popc(bal);
              }
            }
          }
          if (lane == 0) ctl.wl_count[s][aw] = static_cast<uint32_t>(wn);
          if (warp == kAWarp0) RTM3D_TRACE(2);
          __syncwarp();
          if (lane == 0) pl::mbar_arrive(pl::smem_u32((dbg == 8 || dbg == 12) ? &ctl.empty[s] : &ctl.scanned[s]));
          if (lane == 0) RTM3D_MARK(5 + (warp - kAWarp0) % kAPerGroup, gq);
          if (warp == kAWarp0) RTM3D_TRACE(4);
          if (lane == 0) RTM3D_ACC(kStALoop, RTM3D_CLK() - al0);
        }
      }
      if (lane == 0) RTM3D_ACC(kStATotal, RTM3D_CLK() - a_t0);
      RTM3D_FLUSH(kStATotal); RTM3D_FLUSH(kStWaitBufFree); RTM3D_FLUSH(kStWaitFull); RTM3D_FLUSH(kStALoop); RTM3D_FLUSH(kStASetup);
    } else if (dbg == 8 || dbg == 12) {
      // (timing experiment: producer + A-warps only)
    } else if (warp >= kBWarp0) {
      // ================================ B-warps: peak test, candidates ================================
      const int gpr = g.gpr;
      const long long b_t0 = RTM3D_CLK();
      for (; ii.item < g.n_items; ii.next()) {
        if (pass == 1 && __ldcg(&p.retry[ii.item]) == 0u) continue;
        const ItemInfo it = decode_item(p, g, ii, static_cast<int>(sizeof(T)));
        const int buf = static_cast<int>(k & (kNBuf - 1));
        if (k >= static_cast<uint32_t>(kNBuf)) pl::mbar_wait(pl::smem_u32(&ctl.buf_free[buf]), ((k >> kBufShift) - 1u) & 1u, p.status, 0xE1000006u, 100);
        ++k;
        Sel& L = ctl.sel[buf];
        uint32_t* hist = hist_all + buf * kHistBins;
        unsigned long long* list = list_all + static_cast<size_t>(buf) * g.list_cap;
        const float lim = it.is_main ? p.thresh : 0.0f;      // strict: score > lim (models/model.py:91; 0.0 = filler)
        float t_floor = it.is_main ? p.t0 : -INFINITY;
        t_floor = fmaxf(t_floor, L.spec_t);
        int qq = 0;                                     // chunk within the plane
        uint32_t flat_base = it.flat_base;              // of the plane the chunk belongs to
        for (int q = 0; q < it.nchunks; ++q, ++gq, ring_advance()) {
          const uint32_t s = ring_s;
          {
            const long long w0 = RTM3D_CLK();
            pl::mbar_wait(pl::smem_u32(&ctl.scanned[s]), ring_ph, p.status, 0xE1000005u, 100);
            if (lane == 0) RTM3D_MARK(10 + (warp - kBWarp0), gq);
            if (lane == 0) RTM3D_ACC(kStWaitScanned, RTM3D_CLK() - w0);
          }
          const long long bb0 = RTM3D_CLK();
          const int c_lo = it.ys + qq * g.chunk_rows;
          const uint32_t chunk_flat_base = flat_base;
          if (++qq == it.cpp) { qq = 0; flat_base += it.hw; }
          const unsigned short* wl = wl_all + static_cast<size_t>(s) * g.wl_cap;
          const unsigned char* centre = ring + static_cast<size_t>(s) * g.stage_bytes + g.row_bytes;
          // the stage's worklist = the concatenation of the A-warps' segments
          int seg_end[kAPerGroup];
          int n = 0;
#pragma unroll
          for (int a = 0; a < kAPerGroup; ++a) { n += static_cast<int>(ctl.wl_count[s][a]); seg_end[a] = n; }
          const int nb_batches = (dbg == 1 || dbg == 3 || dbg == 7) ? 0 : ((n + 31) >> 5);
          if (warp == kBWarp0 && lane == 0) RTM3D_ACC(kStWlEntries, n);
          // Batches are handed out by a shared counter.  The stage is only needed for a batch's loads: the next batch is
          // claimed right after them, and when there is none the stage goes back to the producer BEFORE this warp works
          // through its candidates (sigmoid, list append, threshold update).
          auto grab = [&]() {
            int b2 = 0;
            if (lane == 0) b2 = static_cast<int>(atomicAdd(&ctl.wl_next[s], 1u));
            return __shfl_sync(0xffffffffu, b2, 0);
          };
          bool released = false;
          int bt = grab();
          while (bt < nb_batches) {
            const int wi = bt * 32 + lane;
            uint32_t pm = 0;                 // pixels of this lane's group that may be candidates
            float v[E], nb[E];
            int y = 0, c4 = 0;
            const unsigned char* gp = centre;
            bool alive = false;
            const float tf = fmaxf(L.t_filter, t_floor);      // the threshold has usually risen since phase A
            if (wi < n) {
              int seg = 0, seg_begin = 0;
#pragma unroll
              for (int a = 0; a < kAPerGroup - 1; ++a) { if (wi >= seg_end[a]) { seg = a + 1; seg_begin = seg_end[a]; } }
              const int gi = wl[seg * g.wl_seg + (wi - seg_begin)];
              gp = centre + static_cast<size_t>(gi) * 16;
              Grp<T>::load(gp, v);
              float m = v[0];
#pragma unroll
              for (int i = 1; i < E; ++i) m = fmaxf(m, v[i]);
              alive = m >= tf;
              const int rl = static_cast<int>(__umulhi(static_cast<uint32_t>(gi), g.gpr_magic));   // row within the chunk
              c4 = gi - rl * gpr;                                                                   // group within the row
              y = c_lo + rl;
            }
            if (!__any_sync(0xffffffffu, alive)) { bt = grab(); continue; }
            if (lane == 0) RTM3D_ACC(kStBatches, 1);
            if (alive) {
              const bool has_up = y > 0, has_dn = y + 1 < H, has_l = c4 > 0, has_r = c4 + 1 < gpr;
              float up[E + 2], dn[E + 2];
              load_window_row<T>(gp - g.row_bytes, has_up, has_l, has_r, up);
              load_window_row<T>(gp + g.row_bytes, has_dn, has_l, has_r, dn);
              const float ml = has_l ? Grp<T>::elem(gp, -1) : -INFINITY;
              const float mr = has_r ? Grp<T>::elem(gp, E) : -INFINITY;
#pragma unroll
              for (int i = 0; i < E; ++i) {
                const float a = fmaxf(fmaxf(up[i], up[i + 1]), up[i + 2]);
                const float c = fmaxf(fmaxf(dn[i], dn[i + 1]), dn[i + 2]);
                const float left = (i == 0) ? ml : v[i - 1];
                const float right = (i == E - 1) ? mr : v[i + 1];
                nb[i] = fmaxf(fmaxf(a, c), fmaxf(left, right));
                const float xc = v[i];
                // a neighbour this much larger is larger after the sigmoid too (logits <= 2 do not collapse that far)
                const bool dead = (xc <= kSatKnee && xc >= kDenormKnee && nb[i] > xc + kTieTol);
                if (xc >= tf && !dead) pm |= 1u << i;
              }
            }
            bt = grab();
            if (bt >= nb_batches && !released) {
              __syncwarp();
              if (lane == 0) pl::mbar_arrive(pl::smem_u32(&ctl.empty[s]));
              if (lane == 0) RTM3D_MARK(17 + (warp - kBWarp0), gq);
              released = true;
            }
            // per-lane candidate loop; a lane that finds the list full parks (`stuck`) until the warp has made room
            while (__any_sync(0xffffffffu, pm != 0u)) {
              bool stuck = false;
              const unsigned long long kstar = L.kstar;
              while (pm != 0u && !stuck) {
                const int i = __ffs(pm) - 1;
                float xc = v[0], xn = nb[0];
#pragma unroll
                for (int j = 1; j < E; ++j) { if (j == i) { xc = v[j]; xn = nb[j]; } }
                const float sc = sigmoid_cold(xc);
                bool cand = sc > lim;
                // sigmoid_ref is monotone (tests/test_sigmoid_gpu.py sweeps every float): the largest neighbour decides
                if (cand && neighbour_needs_exact(xn, xc) && sigmoid_cold(xn) > sc) cand = false;
                if (cand) {
                  const uint32_t flat = chunk_flat_base + static_cast<uint32_t>(y) * W + static_cast<uint32_t>(c4 * E + i);
                  const unsigned long long key = make_key(sc, flat);
                  if (key >= kstar) {
                    const uint32_t slot = atomicAdd(const_cast<uint32_t*>(&L.reserve), 1u);
                    if (slot < static_cast<uint32_t>(g.list_cap)) {
                      list[slot] = key;
                      atomicAdd(&hist[score_bin(sc)], 1u);
                    } else {
                      stuck = true;
                    }
                  }
                }
                if (!stuck) pm &= pm - 1;
              }
              if (__any_sync(0xffffffffu, stuck)) {
                // list full: become the compactor, or wait for whoever is
                int got = 0;
                if (lane == 0) got = (atomicCAS(&L.lock, 0u, 1u) == 0u);
                got = __shfl_sync(0xffffffffu, got, 0);
                if (got) {
                  if (L.reserve >= static_cast<uint32_t>(g.list_cap)) {
                    compact_list(L, list, g.list_cap, K, ctl.rsel[buf], lane);
                    if (lane == 0) RTM3D_ACC(kStCompactions, 1);
                  }
                  __syncwarp();
                  if (lane == 0) { __threadfence_block(); atomicExch(&L.lock, 0u); }
                } else {
                  __nanosleep(200);
                }
              }
            }
            // threshold update every kUpdateEvery appended keys
            int upd = 0;
            if (lane == 0) {
              const uint32_t now = L.reserve;
              if (now >= static_cast<uint32_t>(K) && now - L.last_upd >= static_cast<uint32_t>(kUpdateEvery) &&
                  now <= static_cast<uint32_t>(g.list_cap))
                upd = (atomicCAS(&L.lock, 0u, 1u) == 0u);
            }
            upd = __shfl_sync(0xffffffffu, upd, 0);
            if (upd) {
              if (lane == 0) RTM3D_ACC(kStUpdates, 1);
              update_threshold(L, hist, K, lane, ftable);
              __syncwarp();
              if (lane == 0) { __threadfence_block(); atomicExch(&L.lock, 0u); }
            }
          }
          __syncwarp();
          if (!released && lane == 0) pl::mbar_arrive(pl::smem_u32(&ctl.empty[s]));
          if (!released && lane == 0) RTM3D_MARK(17 + (warp - kBWarp0), gq);
          if (lane == 0) RTM3D_MARK(24 + (warp - kBWarp0), gq);
          if (q == it.nchunks - 1) {
            // (the finisher derives the final histogram boundary itself: a B-warp that did it here would be late for the
            // next chunks, and every stage waits for every B-warp)
            if (lane == 0) { __threadfence_block(); pl::mbar_arrive(pl::smem_u32(&ctl.item_done[buf])); }
          }
          if (lane == 0) RTM3D_ACC(kStBBusy, RTM3D_CLK() - bb0);
        }
      }
      if (lane == 0) RTM3D_ACC(kStBTotal, RTM3D_CLK() - b_t0);
      RTM3D_FLUSH(kStBTotal); RTM3D_FLUSH(kStBBusy); RTM3D_FLUSH(kStWaitScanned); RTM3D_FLUSH(kStWlEntries); RTM3D_FLUSH(kStBatches);
      RTM3D_FLUSH(kStUpdates); RTM3D_FLUSH(kStCompactions);
    } else {
      // ================================ finishers ================================
      // Each finisher warp takes whole items in ticket order and finishes them on its own.
      const int fw = warp - kFinWarp0;
      unsigned long long* finA = fin_all + static_cast<size_t>(2 * fw) * g.fin_cap;
      unsigned long long* finB = finA + g.fin_cap;
      uint32_t* fin_scratch = reinterpret_cast<uint32_t*>(finA);   // emit scratch: finA is dead once finB holds the sorted keys
      uint32_t* rsel = ctl.rsel[kNBuf + fw];
      if (pass == 0) {
        if (lane == 0) my_ticket = atomicAdd(&ctl.fin_next, 1u);
        my_ticket = __shfl_sync(0xffffffffu, my_ticket, 0);
      }
      for (; ii.item < g.n_items; ii.next()) {
        const int item = ii.item;
        if (pass == 1 && __ldcg(&p.retry[item]) == 0u) continue;
        const uint32_t ord = k++;
        if (ord != my_ticket) continue;
        const ItemInfo it = decode_item(p, g, ii, static_cast<int>(sizeof(T)));
        const int buf = static_cast<int>(ord & (kNBuf - 1));
        const long long f0 = RTM3D_CLK();
        // the previous use of this buffer must have been taken over by its finisher before this use's barrier phase can
        // be waited for by parity
        while (ctl.fin_released[buf] != (ord >> kBufShift)) __nanosleep(200);
        pl::mbar_wait(pl::smem_u32(&ctl.item_done[buf]), (ord >> kBufShift) & 1u, p.status, 0xE1000004u, 200);
        const long long f1 = RTM3D_CLK();
        long long lap_ = f1;
        (void)lap_;
        Sel& L = ctl.sel[buf];
        uint32_t* hist = hist_all + buf * kHistBins;
        unsigned long long* list = list_all + static_cast<size_t>(buf) * g.list_cap;
        const int n = min(static_cast<int>(L.reserve), g.list_cap);
        const float lim = it.is_main ? p.thresh : 0.0f;

        // ---- final boundary of the histogram; was the speculative start threshold justified?
        const int bin = warp_hist_boundary(hist, K, lane);
        const int sb = L.spec_bin;
        // speculation skipped pixels only when its edge lies above the score floor; it was right iff at least K
        // candidates were found at or above that edge
        const bool spec_active = sb >= 1 && __uint_as_float(bin_edge_bits(sb)) > lim;
        const bool failed = spec_active && bin < sb && dbg == 0;   // (the timing experiments starve the histogram)
        const uint32_t cut = (bin >= 1) ? bin_edge_bits(bin) : 0u;
        const unsigned long long kstar = L.kstar;
        if (lane == 0 && it.plane < kMaxPlanes) ctl.guess_bin[it.plane] = failed ? -1 : bin;
        RTM3D_FIN_LAP(kStFinBoundary);
        // ---- cut the list into finA (order irrelevant: keys are sorted next), then clear it
        int m = 0;
        if (!failed) {
#pragma unroll 2
          for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            unsigned long long key = 0ull;
            if (i < n) key = list[i];
            const bool keep = key != 0ull && static_cast<uint32_t>(key >> 32) >= cut && key >= kstar;
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              const int pos = m + __popc(bal & ((1u << lane) - 1u));
              if (pos < g.fin_cap) finA[pos] = key;
            }
            m += __popc(bal);
          }
          if (m > g.fin_cap) {
            // more keys at the cut than the sort buffer holds (adversarial ties): exact K-th key first, then cut again
            const unsigned long long kth = warp_radix_kth(list, n, K, rsel, lane);
            m = 0;
#pragma unroll 1
            for (int base = 0; base < n; base += 32) {
              const int i = base + lane;
              unsigned long long key = 0ull;
              if (i < n) key = list[i];
              const bool keep = key != 0ull && key >= kth;
              const uint32_t bal = __ballot_sync(0xffffffffu, keep);
              if (keep) finA[m + __popc(bal & ((1u << lane) - 1u))] = key;
              m += __popc(bal);                       // <= K <= fin_cap
            }
          }
        }
#pragma unroll 1
        for (int i = lane; i < n; i += 32) list[i] = 0ull;
        RTM3D_FIN_LAP(kStFinCompact);
#pragma unroll 1
        for (int i = lane; i < kHistBins / 4; i += 32) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        if (lane == 0) {
          L.reserve = 0; L.lock = 0; L.last_upd = 0; L.t_filter = -INFINITY; L.kstar = 0ull; L.last_bin = -1; L.b_done = 0;
          // speculative start threshold of the item that gets this buffer next (kNBuf items ahead, pass 0 only)
          int nsb = -1;
          const int nxt = cta_items.at(ii.kl + kNBuf);           // (pass 0 walks every item: kl == ord)
          if (g.speculate && pass == 0 && nxt < g.n_items) nsb = spec_for_plane(item_plane(p, g, cta_items.n_main, nxt));
          L.spec_bin = nsb;
          L.spec_t = filter_lookup(ftable, nsb);    // (-inf for nsb < 1)
          __threadfence_block();
          ctl.fin_released[buf] = (ord >> kBufShift) + 1u;
          pl::mbar_arrive(pl::smem_u32(&ctl.buf_free[buf]));
          my_ticket = atomicAdd(&ctl.fin_next, 1u);
        }
        my_ticket = __shfl_sync(0xffffffffu, my_ticket, 0);
        RTM3D_FIN_LAP(kStFinRelease);
        if (dbg == 2 || dbg == 3 || dbg == 7) continue;
        if (failed) {
          if (lane == 0) { p.retry[item] = 1u; atomicAdd(&ctl.n_retry, 1u); RTM3D_ACC(kStRetried, 1); }   // redone in pass 1 (no speculation there)
          continue;
        }
        // ---- sort the survivors
        const int mpad = ((m + 31) >> 5) << 5;
#pragma unroll 1
        for (int i = m + lane; i < mpad; i += 32) finA[i] = 0ull;
        __syncwarp();
        warp_fin_sort(finA, m, finB, lane);
        int have = min(m, K);
        RTM3D_FIN_LAP(kStFinSort);
        if (dbg == 5) continue;

        // ---- emit, or publish + merge by the last part of the problem
        const int parts = SPLIT ? g.split : 1;         // strips of the problem (image, or keypoint plane)
        const int kc = it.kc;
        bool do_emit = true;
        if (parts > 1) {
          const size_t unit = static_cast<size_t>(item);
#pragma unroll 1
          for (int i = lane; i < have; i += 32) p.keys[unit * K + i] = finB[i];
          if (lane == 0) p.key_counts[unit] = static_cast<uint32_t>(have);
          __threadfence();
          __syncwarp();
          const int ticket_id = it.is_main ? it.b : p.B + it.b * p.Cv + kc;
          int last = 0;
          if (lane == 0) {
            const uint32_t t = atomicAdd(&p.tickets[ticket_id], 1u);
            last = (t == static_cast<uint32_t>(parts - 1));
            if (last) p.tickets[ticket_id] = 0u;               // leave the workspace clean for the next call
          }
          last = __shfl_sync(0xffffffffu, last, 0);
          do_emit = last != 0;
          if (do_emit) {
            __threadfence();
            // units of this problem are contiguous items: the strips of the image's main item, or of one keypoint plane
            const size_t unit0 = static_cast<size_t>(item - it.strip);
            // gather the parts' sorted lists into finA at offsets u*K, then rank-merge into finB
            int total = 0;
#pragma unroll 1
            for (int u = 0; u < parts; ++u) {
              const int cu = static_cast<int>(__ldcg(&p.key_counts[unit0 + u]));
              total += cu;
#pragma unroll 2
              for (int i = lane; i < K; i += 32)
                finA[static_cast<size_t>(u) * K + i] = (i < cu) ? __ldcg(&p.keys[(unit0 + u) * K + i]) : 0ull;
            }
            __syncwarp();
            // rank-merge: a key's rank = its position in its own (sorted) list + the larger keys of every other list;
            // four keys per lane search in lockstep so that the dependent shared-memory reads overlap
#pragma unroll 1
            for (int i0 = lane; i0 < parts * K; i0 += 32 * 4) {
              unsigned long long key[4];
              int own[4], rank[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = i0 + 32 * e;
                key[e] = (i < parts * K) ? finA[i] : 0ull;
                own[e] = i / K;
                rank[e] = i - own[e] * K;
              }
#pragma unroll 1
              for (int u = 0; u < parts; ++u) {
                const unsigned long long* lst = finA + static_cast<size_t>(u) * K;
                int lo[4], hi[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) { lo[e] = 0; hi[e] = (u == own[e] || key[e] == 0ull) ? 0 : K; }
                for (int span = K; span > 0; span >>= 1) {          // enough halvings for K entries (lo == hi ends early)
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    if (lo[e] < hi[e]) {
                      const int mid = (lo[e] + hi[e]) >> 1;
                      if (lst[mid] > key[e]) lo[e] = mid + 1; else hi[e] = mid;   // zero padding never counts as larger
                    }
                  }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) rank[e] += lo[e];
              }
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (key[e] != 0ull && rank[e] < K) finB[rank[e]] = key[e];
            }
            __syncwarp();
            have = min(total, K);
          }
        }
        RTM3D_FIN_LAP(kStFinPublish);
        if (dbg == 6 || (dbg == 9 && it.is_main) || (dbg == 10 && !it.is_main)) continue;
        if (do_emit) {
          if (it.is_main) warp_write_main(sp, it.b, finB, have, lane);
          else warp_write_kpt(sp, it.b, kc, finB, have, fin_scratch, lane);
        }
        __syncwarp();
        RTM3D_FIN_LAP(kStFinEmit);
        if constexpr (STATS) {
          if (lane == 0) { acc_[kStItems] += 1; acc_[kStPushed] += n; acc_[kStFinWait] += f1 - f0; acc_[kStFinBusy] += clock64() - f1; }
        }
      }
      RTM3D_FLUSH(kStItems); RTM3D_FLUSH(kStPushed); RTM3D_FLUSH(kStFinWait); RTM3D_FLUSH(kStFinBusy); RTM3D_FLUSH(kStRetried);
      RTM3D_FLUSH(kStFinBoundary); RTM3D_FLUSH(kStFinCompact); RTM3D_FLUSH(kStFinRelease); RTM3D_FLUSH(kStFinSort);
      RTM3D_FLUSH(kStFinPublish); RTM3D_FLUSH(kStFinEmit);
    }
  }
  // leave the workspace clean: clear this CTA's retry flags once every role has finished reading them
  __syncthreads();
  if (!STATS && p.stats && tid == 0) p.stats[3840 + 4 * blockIdx.x + 3] = pl::globaltimer_ns();
  // remember the boundaries for the next launch.  Every CTA writes the planes it has seen (a CTA of a small batch sees
  // only a few plane indices); CTAs that disagree race benignly -- any of their values is a valid starting guess.
  if (tid < kMaxPlanes && ctl.guess_bin[tid] >= 0) p.guess[tid] = static_cast<uint32_t>(ctl.guess_bin[tid] + 1);
#pragma unroll 1
  if (ctl.n_retry != 0u) {
    for (int kl = tid; kl < cta_items.count; kl += kPlaneThreads) {
      const int item = cta_items.at(kl);
      if (__ldcg(&p.retry[item]) != 0u) p.retry[item] = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
#ifdef RTM3D_DEV
// developer instrumentation: process-global knobs, compiled only into `make DEV=1` builds
static unsigned long long* g_stats = nullptr;
static int g_stages_override = 0;
static unsigned long long* g_trace = nullptr;
void debug_set_trace(unsigned long long* t) { g_trace = t; }
static int g_copy_rows = 0;   // developer knob (debug_set_copy_rows): rows per bulk copy, 0 = whole chunk
void debug_set_copy_rows(int r) { if (r >= 1000) { g_stages_override = r - 1000; g_copy_rows = 0; } else g_copy_rows = r; }
void debug_set_stats(unsigned long long* dev_u64_16) { g_stats = dev_u64_16; }
unsigned long long* debug_get_stats() { return g_stats; }
#else
static unsigned long long* const g_stats = nullptr;
static unsigned long long* const g_trace = nullptr;
constexpr int g_stages_override = 0;
constexpr int g_copy_rows = 0;
#endif
// multiprocessor count of the current device (a process may drive several GPUs)
static int plane_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

static bool make_plane_geom(const PlaneParams& p, int dtype, int split_override, int speculate, PlaneGeom& g) {
  const int es = dtype == 0 ? 4 : 2;
  const int E = 16 / es;
  if (p.W % E != 0) return false;
  if ((p.hm_main && reinterpret_cast<uintptr_t>(p.hm_main) % 16 != 0) || (p.hm_kpt && reinterpret_cast<uintptr_t>(p.hm_kpt) % 16 != 0)) return false;
  const int row_bytes = p.W * es;
  if (row_bytes > 8192) return false;
  if (row_bytes < 32) return false;                        // a one-group row: ceil(2^32 / gpr) does not fit the 32-bit magic (generic kernels)
  const int g_sm_count = plane_sm_count();
  if (g_sm_count <= 0) return false;
  const long long planes = static_cast<long long>(p.B) * ((p.C > 0 ? 1 : 0) + p.Cv);   // items per strip index (the C main planes are one item)
  // Strips per plane.  Splitting a selection problem over several CTAs costs a publish + merge through global memory on
  // the finisher warps, which are the busiest part of a CTA: measured, split = 1 is never slower for batches of 1..256
  // images of 96x320 or 192x640 planes (tools/split_time.py), and 2-15x faster once there are more items than SMs.
  // Only a handful of very large planes (>= 1 MB per strip, fewer items than a quarter of the SMs) are cut into strips.
  int split = 1;
  if (split_override > 0) {
    split = split_override;
  } else {
    while (split < 8 && p.H / (2 * split) >= 8 && planes * split * 4 <= g_sm_count &&
           static_cast<long long>(p.H) * row_bytes / split >= (1LL << 20))
      split *= 2;
  }
  while (split > 1 && p.H < split) split /= 2;
  const int max_parts = split;
  g.split = split;
  g.split_shift = split == 8 ? 3 : split == 4 ? 2 : split == 2 ? 1 : 0;
  g.row_bytes = row_bytes;
  g.gpr = row_bytes / 16;
  g.gpr_magic = static_cast<unsigned>((0x100000000ULL + g.gpr - 1) / g.gpr);
  g.list_cap = 2 * p.K + 384;
  g.list_cap = (g.list_cap + 31) & ~31;
  int fin = max_parts * p.K;
  if (fin < p.K + 96) fin = p.K + 96;
  g.fin_cap = (fin + 31) & ~31;
  // finA doubles as the filler scratch of the keypoint planes: 3K+8 words
  {
    const int words = 3 * p.K + 8;
    if (g.fin_cap * 2 < words) g.fin_cap = ((words + 1) / 2 + 31) & ~31;
  }
  const size_t fixed = static_cast<size_t>(kNBuf) * kHistBins * 4 + static_cast<size_t>(kNBuf) * g.list_cap * 8 +
                       2ull * kFinWarps * g.fin_cap * 8;
  const size_t budget = 220 * 1024;
  if (fixed + 2ull * 3 * row_bytes > budget) return false;
  const int strip = (p.H + split - 1) / split;
  // ring: kDefaultStages stages sharing ~144 KB (+ a 2-byte worklist slot per 16-byte centre group); smaller stages
  // when the selection buffers are large
  int stages = g_stages_override > 0 ? g_stages_override : kDefaultStages;
  const size_t ring = budget - fixed;
  auto rows_for = [&](int st) {
    const long long per_stage = static_cast<long long>(ring / st) - 2LL * row_bytes - 64 - 2LL * kAPerGroup * 64;
    return static_cast<int>(per_stage * 8 / (9LL * row_bytes));       // cr*row_bytes + cr*row_bytes/8 <= per_stage
  };
  int cr = rows_for(stages);
  const int cr_target = (36864 * 4 / stages) / row_bytes - 2;
  if (cr > cr_target && cr_target >= 1) cr = cr_target;
  if (cr < 1) { stages = 2; cr = rows_for(stages); }
  if (cr < 1) return false;
  if (cr > strip) cr = strip;
  while (static_cast<long long>(cr) * g.gpr > 65535) --cr;           // worklist entries are 16-bit group indices
  const int nch = (strip + cr - 1) / cr;
  cr = (strip + nch - 1) / nch;                       // even out the chunks of a strip
  g.rows_lo = p.H >> g.split_shift;
  g.nch_lo = (g.rows_lo + cr - 1) / cr;
  g.nch_hi = (g.rows_lo + 1 + cr - 1) / cr;
  g.stages = stages;
  g.speculate = speculate;
  g.chunk_rows = cr;
  g.copy_rows = g_copy_rows > 0 ? g_copy_rows : cr + 2;
  g.stage_bytes = (cr + 2) * row_bytes;
  {
    const int tasks = (cr * g.gpr + 31) / 32;                             // tasks of 32 groups in a chunk
    const int per_warp = (tasks + kAPerGroup - 1) / kAPerGroup;                   // most tasks one A-warp gets
    g.wl_seg = per_warp * 32;
    g.wl_cap = kAPerGroup * g.wl_seg;
  }
  g.n_items = static_cast<int>(static_cast<long long>(p.B) * ((p.C > 0 ? 1 : 0) + p.Cv) * split);
  g.smem = static_cast<unsigned>(static_cast<size_t>(stages) * (g.stage_bytes + 2ull * g.wl_cap) + fixed);
  // dynamic + static shared memory (the control block and the parameter copy) must fit the 227 KB a block may use
  return static_cast<size_t>(g.smem) + sizeof(PlaneCtl) + sizeof(PlaneParams) + 128 <= 232448;
}

bool planes_eligible(const PlaneParams& p, int dtype) {
  PlaneGeom g{};
  return make_plane_geom(p, dtype, 0, 1, g);
}

template <typename T, bool STATS, bool DBG, bool SPLIT>
static int launch_planes_t(const PlaneParams& p, const PlaneGeom& g, cudaStream_t s) {
  auto kern = decode_planes_kernel<T, STATS, DBG, SPLIT>;
  cudaError_t e = ensure_dynamic_smem<decode_planes_kernel<T, STATS, DBG, SPLIT>>(g.smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int g_sm_count = plane_sm_count();
  int grid = g.n_items < g_sm_count ? g.n_items : g_sm_count;
  if (g.max_ctas > 0 && grid > g.max_ctas) grid = g.max_ctas;
  kern<<<grid, kPlaneThreads, g.smem, s>>>(p, g);
  return static_cast<int>(cudaGetLastError());
}

int launch_planes(const PlaneParams& p, int dtype, int split_override, int speculate, int max_ctas, int debug, cudaStream_t s) {
  PlaneGeom g{};
  if (!make_plane_geom(p, dtype, split_override, speculate, g)) return -1000;
  g.max_ctas = max_ctas;
  g.debug = debug;
  PlaneParams q = p;
  q.stats = g_stats ? g_stats : g_trace;
  if (g_stats) return dtype == 0 ? launch_planes_t<float, true, false, true>(q, g, s) : launch_planes_t<__nv_bfloat16, true, false, true>(q, g, s);
  if (debug != 0) return dtype == 0 ? launch_planes_t<float, false, true, true>(q, g, s) : launch_planes_t<__nv_bfloat16, false, true, true>(q, g, s);
  if (g.split > 1) return dtype == 0 ? launch_planes_t<float, false, false, true>(q, g, s) : launch_planes_t<__nv_bfloat16, false, false, true>(q, g, s);
  return dtype == 0 ? launch_planes_t<float, false, false, false>(q, g, s) : launch_planes_t<__nv_bfloat16, false, false, false>(q, g, s);
}

// Verification aid: the logit threshold and the score edge of every histogram bin (rtm3d_threshold_table).
__global__ void threshold_table_kernel(float* t, uint32_t* edge, int n) {
  const int bin = blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= n) return;
  t[bin] = filter_from_bin(bin);
  edge[bin] = bin >= 1 ? bin_edge_bits(bin) : 0u;
}
__global__ void filter_table_kernel(float* t) {
  const int bin = blockIdx.x * blockDim.x + threadIdx.x;
  if (bin < kHistBins) t[bin] = bin >= 1 ? filter_from_bin(bin) : -INFINITY;
  else if (bin < kFilterTableWords - 1) t[bin] = 15.0f;
  else if (bin == kFilterTableWords - 1) t[bin] = __uint_as_float(kFilterTableMagic);
}
int launch_filter_table(float* table, cudaStream_t s) {
  static_assert(kHistBins < kFilterTableWords, "table too small");
  filter_table_kernel<<<kFilterTableWords / 128, 128, 0, s>>>(table);
  return static_cast<int>(cudaGetLastError());
}
int threshold_table_bins() { return static_cast<int>((0x3F800000u >> kScoreShift) - kScoreBase) + 1; }
int launch_threshold_table(float* t, uint32_t* edge, cudaStream_t s) {
  const int n = threshold_table_bins();
  threshold_table_kernel<<<(n + 127) / 128, 128, 0, s>>>(t, edge, n);
  return static_cast<int>(cudaGetLastError());
}

// items (= workspace units) the kernel may use for this shape, for the workspace size
long long planes_max_units(int B, int planes_per_image) { return static_cast<long long>(B) * planes_per_image * 8; }

}  // namespace rtm3d
