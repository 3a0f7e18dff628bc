// Streaming decode kernel for sm_100a: the heat-map is streamed from HBM by bulk async copies (TMA engine,
// cp.async.bulk -> SASS UBLKCP) into a shared-memory ring and reduced to the exact top-K on the fly.
//
// One CTA (or a thread-block cluster of S CTAs, each owning H/S rows) per selection problem:
//   kModeMain : problem = image          (flat top-K over C*H*W, models/model.py:87-98)
//   kModeKpt  : problem = (image, plane) (per-channel top-K,     models/model.py:109-114)
//
// A chunk = `chunk_rows` centre rows of one plane plus the row above and below (rows of an NCHW plane are contiguous,
// so a chunk is ONE bulk copy).  Chunks are self-contained: a stage is handed back to the producer as soon as it has
// been scanned, so kStages-1 chunks are always in flight.  The two halo rows are read again by the neighbouring chunk
// a few microseconds later and are served by L2, not HBM.
//
// Warp roles (320 threads):
//   warp 9 lane 0 : producer.  Waits for a free stage (mbarrier `empty`), arms `full` with the byte count, issues the copy.
//   warps 0..7    : scanners.  Pass 1: per 16-byte group one LDS.128, a max and one compare against the running LOGIT
//                   threshold; groups holding a survivor go to a per-warp worklist (one vote per 32 groups).  Pass 2: the
//                   compacted worklist, one hit pixel per lane through one code path: exact sigmoid-domain 3x3 peak test
//                   (utils/model_utils.py:17-26 on models/model.py:85), then a push of (key, logit) into a shared queue.
//   warp 8        : selector.  Drains the queue in batches of 32 into a candidate list and about every K arrivals picks
//                   a pivot by sampling, drops what is below it and raises the threshold.  Any pivot with >= K candidates
//                   at or above it bounds the K-th best from below, so exactness never depends on the pivot quality.
//                   After the scan it sorts the (<= 2K) survivors itself, without block-wide barriers.
//   all warps     : cluster merge through distributed shared memory when S > 1, then the gather / regress / 2D-box
//                   epilogue (epilogue.cuh).
#include <cuda_runtime.h>

#include <type_traits>

#include "common.cuh"
#include "epilogue.cuh"
#include "params.h"

namespace rtm3d {

constexpr int kScanWarps = 8;
constexpr int kSelWarp = 8;
constexpr int kProdWarp = 9;
constexpr int kStreamThreads = 320;
constexpr int kStages = 3;           // ring depth: one stage being scanned, two in flight
constexpr int kSmemBudget = 110 * 1024;  // per CTA, so that two CTAs fit one SM (227 KB)
constexpr int kQCap = 512;           // candidate queue slots (power of two)
constexpr int kMaxCluster = 8;
constexpr int kUnroll = 5;           // 16-byte groups per lane per scan iteration
constexpr int kWorkList = 384;       // per-warp worklist entries (u16 group indices)
constexpr int kPeakList = 256;       // per-warp peak list entries (flat index, logit); flushed so that 32*4 more always fit
constexpr int kWarpSortMax = 256;    // the selector sorts up to this many survivors on its own

struct StreamGeom {
  int chunk_rows;     // centre rows per chunk
  int stage_bytes;    // (chunk_rows + 2) * row_bytes (multiple of 16)
  int row_bytes;
  int list_cap;       // candidate list capacity (>= 2K + 256)
  int cluster;        // CTAs per problem
  int fin_cap;        // capacity (u64) of the final sort/merge buffer that aliases the ring
  unsigned gpr_magic; // ceil(2^32 / groups_per_row): gi / gpr == __umulhi(gi, gpr_magic) for gi < 65536
  unsigned smem;      // dynamic shared bytes
  unsigned long long* timeline;  // developer instrumentation (16 u64 per CTA) or nullptr
  int debug;          // developer switches: 1 = no warm-up rounds, 2 = block bitonic instead of the rank merge
};

static unsigned long long* g_timeline = nullptr;  // set through rtm3d_debug_set_timeline (tools/ only)

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint64_t ld_dsmem_u64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_dsmem_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RTM3D_TL(slot, val) do { if (g.timeline) g.timeline[static_cast<size_t>(blockIdx.x) * 16 + (slot)] = (val); } while (0)

// Bounded wait: a pipeline bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* status, uint32_t code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (status) atomicExch(status, code);
      __threadfence_system();
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct __align__(16) StreamCtl {
  unsigned long long full[kStages];
  unsigned long long empty[kStages];
  volatile uint32_t q_tail;      // slots reserved by the scanners
  volatile uint32_t q_head;      // slots consumed by the selector
  volatile uint32_t done;        // scanner warps finished
  volatile float t_filter;       // running logit-domain threshold
  uint32_t list_count;           // candidates left in the list when the scan ends
  uint32_t best_count;           // keys in best[] after the local exact select
  uint32_t misc[2];              // [0]: survivors already sorted by the selector
};

// element access into a ring row (raw element type T in shared memory)
template <typename T> __device__ __forceinline__ float ring_elem(const unsigned char* row, int col) {
  return to_f32(reinterpret_cast<const T*>(row)[col]);
}

template <typename T> struct Group;  // one 16-byte group of a row
template <> struct Group<float> {
  static constexpr int kElems = 4;
  __device__ static __forceinline__ void load(const unsigned char* p, float (&v)[4]) {
    const float4 f = *reinterpret_cast<const float4*>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
};
template <> struct Group<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void load(const unsigned char* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Selector-side helpers (one warp).

struct SampleKey { uint32_t hi, lo; float x; };

__device__ __forceinline__ bool key_ge(uint32_t ahi, uint32_t alo, uint32_t bhi, uint32_t blo) {
  return (ahi > bhi) || (ahi == bhi && alo >= blo);
}

// Descending bitonic sort of one (hi,lo,x) triple per lane.
__device__ __forceinline__ void warp_sort_desc(SampleKey& s, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint32_t ohi = __shfl_xor_sync(0xffffffffu, s.hi, j);
      const uint32_t olo = __shfl_xor_sync(0xffffffffu, s.lo, j);
      const float ox = __shfl_xor_sync(0xffffffffu, s.x, j);
      const bool lower_lane = (lane & j) == 0;
      const bool desc_block = (lane & k) == 0;
      const bool mine_ge = key_ge(s.hi, s.lo, ohi, olo);
      // in a descending block the lower lane keeps the larger key
      const bool keep_mine = (lower_lane == desc_block) ? mine_ge : !mine_ge;
      if (!keep_mine) { s.hi = ohi; s.lo = olo; s.x = ox; }
    }
  }
}

__device__ __forceinline__ int warp_count_ge(const uint32_t* lhi, const uint32_t* llo, int count, uint32_t phi, uint32_t plo,
                                             int lane) {
  int c = 0;
  for (int i = lane; i < count; i += 32) c += key_ge(lhi[i], llo[i], phi, plo) ? 1 : 0;
  return __reduce_add_sync(0xffffffffu, c);
}

// Exact K-th largest 64-bit key of the list by bitwise descent (fallback when sampling cannot find a usable pivot).
__device__ __forceinline__ void warp_exact_kth(const uint32_t* lhi, const uint32_t* llo, int count, int K, int lane,
                                               uint32_t& phi, uint32_t& plo) {
  uint64_t prefix = 0, mask = 0;
  int need = K;
  for (int bit = 63; bit >= 0; --bit) {
    const uint64_t m1 = mask | (1ull << bit), want = prefix | (1ull << bit);
    int c = 0;
    for (int i = lane; i < count; i += 32) {
      const uint64_t k = (static_cast<uint64_t>(lhi[i]) << 32) | llo[i];
      c += ((k & m1) == want) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= need) prefix = want; else need -= c;
    mask = m1;
  }
  phi = static_cast<uint32_t>(prefix >> 32);
  plo = static_cast<uint32_t>(prefix);
}

// Drop every candidate below a pivot that still has >= K candidates at or above it; returns the new count and the logit
// of the pivot (the new running threshold is derived from it).
__device__ __forceinline__ int warp_prune(uint32_t* lhi, uint32_t* llo, float* lx, int count, int K, int max_keep, bool compact, int lane,
                                          float& x_pivot) {
  SampleKey s;
  {
    const int i = static_cast<int>((static_cast<long long>(lane) * count) >> 5);
    s.hi = lhi[i]; s.lo = llo[i]; s.x = lx[i];
  }
  warp_sort_desc(s, lane);
  // expected rank of the K-th best among 32 evenly spread samples, with head-room
  int r = (40 * K) / count + 2;
  if (r > 31) r = 31;
  uint32_t phi = 0, plo = 0;
  float px = 0.f;
  bool found = false;
  for (int attempt = 0; attempt < 6; ++attempt) {
    phi = __shfl_sync(0xffffffffu, s.hi, r);
    plo = __shfl_sync(0xffffffffu, s.lo, r);
    px = __shfl_sync(0xffffffffu, s.x, r);
    const int c = warp_count_ge(lhi, llo, count, phi, plo, lane);
    if (c >= K && c <= max_keep) { found = true; break; }
    if (c < K) { if (r == 31) break; r = min(31, r + max(2, r >> 1)); }
    else { if (r == 0) break; r = r - 1; }  // pivot so low that nothing would be freed: move up
  }
  if (!found) {
    warp_exact_kth(lhi, llo, count, K, lane, phi, plo);
    // logit of that key: exactly one list entry carries it
    float xx = 0.f;
    bool mine = false;
    for (int i = lane; i < count; i += 32)
      if (lhi[i] == phi && llo[i] == plo) { xx = lx[i]; mine = true; }
    const uint32_t owner = __ballot_sync(0xffffffffu, mine);
    px = __shfl_sync(0xffffffffu, xx, owner ? (__ffs(owner) - 1) : 0);
  }
  x_pivot = px;
  if (!compact) return count;
  // in-place stable compaction (writes trail reads)
  int w = 0;
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    uint32_t hi = 0, lo = 0;
    float x = 0.f;
    bool keep = false;
    if (i < count) {
      hi = lhi[i]; lo = llo[i]; x = lx[i];
      keep = key_ge(hi, lo, phi, plo);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) {
      const int pos = w + __popc(b & ((1u << lane) - 1u));
      lhi[pos] = hi; llo[pos] = lo; lx[pos] = x;
    }
    w += __popc(b);
    __syncwarp();
  }
  x_pivot = px;
  return w;
}

// Descending bitonic sort of a[0..npad) (shared memory, npad a power of two) by ONE warp: no block barriers.
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t* a, int npad, int lane) {
  const int half = npad >> 1;
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int c = lane; c < half; c += 32) {
        const int i = ((c & ~(j - 1)) << 1) | (c & (j - 1));
        const int q = i | j;
        const uint64_t x = a[i], y = a[q];
        const bool first_block = ((i & k) == 0);
        if (first_block ? (x < y) : (x > y)) { a[i] = y; a[q] = x; }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(kStreamThreads, 2) decode_stream_kernel(const DecodeParams p, const StreamGeom g,
                                                                          uint32_t* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ StreamCtl ctl;

  constexpr int E = Group<T>::kElems;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.W, H = p.H, K = p.K;
  const int HW = H * W;
  const int S = g.cluster;
  const int rank = (S > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int prob = blockIdx.x / S;
  int b, plane0, nplanes;
  if (MODE == kModeMain) { b = prob; plane0 = 0; nplanes = p.C; }
  else { b = prob / p.C; plane0 = prob - b * p.C; nplanes = 1; }

  // centre rows owned by this CTA, cut into chunks of CR rows
  const int ys = static_cast<int>((static_cast<long long>(rank) * H) / S);
  const int ye = static_cast<int>((static_cast<long long>(rank + 1) * H) / S);
  const int CR = g.chunk_rows;
  const int cpp = (ye - ys + CR - 1) / CR;  // chunks per plane
  const int NQ = nplanes * cpp;

  // shared carve-up
  unsigned char* ring = smem;
  size_t o = static_cast<size_t>(kStages) * g.stage_bytes;
  uint64_t* qkey = reinterpret_cast<uint64_t*>(smem + o);            o += static_cast<size_t>(kQCap) * 8;
  float* qx = reinterpret_cast<float*>(smem + o);                    o += static_cast<size_t>(kQCap) * 4;
  uint32_t* lhi = reinterpret_cast<uint32_t*>(smem + o);             o += static_cast<size_t>(g.list_cap) * 4;
  uint32_t* llo = reinterpret_cast<uint32_t*>(smem + o);             o += static_cast<size_t>(g.list_cap) * 4;
  float* lx = reinterpret_cast<float*>(smem + o);                    o += static_cast<size_t>(g.list_cap) * 4;
  uint16_t* wl_all = reinterpret_cast<uint16_t*>(smem + o);          o += static_cast<size_t>(kScanWarps) * kWorkList * 2;
  o = (o + 15) & ~size_t(15);
  uint2* pk_all = reinterpret_cast<uint2*>(smem + o);                o += static_cast<size_t>(kScanWarps) * kPeakList * 8;
  uint64_t* best = reinterpret_cast<uint64_t*>(smem + o);            // [K] local top-K, read by cluster peers (S > 1)

  if (tid == 0) RTM3D_TL(0, globaltimer_ns());
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&ctl.full[s]), 1);
      mbar_init(smem_u32(&ctl.empty[s]), kScanWarps);
    }
    ctl.q_tail = 0; ctl.q_head = 0; ctl.done = 0;
    ctl.t_filter = p.t0;
    ctl.list_count = 0; ctl.best_count = 0; ctl.misc[0] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kQCap; i += kStreamThreads) qkey[i] = 0ull;
  __syncthreads();
  if (tid == 0) RTM3D_TL(1, globaltimer_ns());

  const unsigned char* gbase = reinterpret_cast<const unsigned char*>(p.hm);
  const size_t plane_bytes = static_cast<size_t>(HW) * sizeof(T);

  if (warp == kProdWarp) {
    // ================================ producer ================================
    if (lane == 0) {
      int pl = 0, j = 0;
      for (int q = 0; q < NQ; ++q) {
        const int s = q % kStages;
        if (q >= kStages) mbar_wait(smem_u32(&ctl.empty[s]), ((q / kStages) - 1) & 1, status, 0xE0000001u);
        const int c_lo = ys + j * CR, c_hi = min(c_lo + CR, ye);
        const int top = max(c_lo - 1, 0), bot = min(c_hi + 1, H);   // rows [top, bot) incl. the halo rows
        const uint32_t bytes = static_cast<uint32_t>(bot - top) * g.row_bytes;
        const unsigned char* src = gbase + (static_cast<size_t>(b) * p.C + plane0 + pl) * plane_bytes +
                                   static_cast<size_t>(top) * g.row_bytes;
        const uint32_t bar = smem_u32(&ctl.full[s]);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(ring + static_cast<size_t>(s) * g.stage_bytes), src, bytes, bar);
        if (++j == cpp) { j = 0; ++pl; }
      }
    }
  } else if (warp < kScanWarps) {
    // ================================ scanners ================================
    // pass 1 : groups -> per-warp worklist of groups holding a logit >= running threshold (one vote per 32 groups)
    // pass 2A: worklist -> logit-domain 3x3 test per hit pixel (9 LDS, no transcendental); survivors go to a per-warp
    //          peak list as (flat index, logit).  The exact sigmoid-domain compare is only needed when a neighbour is
    //          within the collapse distance of the centre (rare) and is done right there.
    // pass 2B: the peak list, dense (one peak per lane): sigmoid, score threshold, warp-aggregated push to the selector.
    //          Its entries do not reference the ring, so it runs whenever 32 peaks have piled up, not per chunk.
    const int gpr = W / E;  // 16-byte groups per row
    uint16_t* wl = wl_all + warp * kWorkList;
    uint2* pk = pk_all + warp * kPeakList;
    int pkc = 0;            // peaks waiting in pk[] (warp-uniform)
    int pl = 0, j = 0;      // plane within the problem, chunk within the plane
    int n_flush = 0;        // instrumentation: worklist passes of this warp

    auto pass2b = [&](bool all) {
      // consume full batches of 32 (or everything when `all`)
      int done_n = 0;
      const float tnow = ctl.t_filter;
      while (pkc - done_n >= (all ? 1 : 32)) {
        const int i = done_n + lane;
        bool ok = false;
        float xc = 0.f, sc = 0.f;
        uint32_t flat = 0;
        if (i < pkc) {
          const uint2 en = pk[i];
          flat = en.x;
          xc = __uint_as_float(en.y);
          if (xc >= tnow) {
            sc = sigmoid_ref(xc);
            ok = (MODE == kModeMain) ? (sc > p.thresh) : (sc > 0.0f);
          }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (bal) {
          // push (multi-producer, single consumer): one reservation per warp, payload first, key last
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(const_cast<uint32_t*>(&ctl.q_tail), static_cast<uint32_t>(__popc(bal)));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (ok) {
            const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
            if (slot - ctl.q_head >= static_cast<uint32_t>(kQCap)) {
              const long long t0 = clock64();
              while (slot - ctl.q_head >= static_cast<uint32_t>(kQCap)) {
                __nanosleep(64);
                if (clock64() - t0 > 4000000000LL) { if (status) atomicExch(status, 0xE0000004u); __threadfence_system(); __trap(); }
              }
            }
            qx[slot & (kQCap - 1)] = xc;
            __threadfence_block();
            *reinterpret_cast<volatile uint64_t*>(&qkey[slot & (kQCap - 1)]) = make_key(sc, flat);
          }
        }
        done_n += 32;
      }
      __syncwarp();
      if (done_n >= pkc) { pkc = 0; }
      else if (done_n > 0) {
        // keep the (< 32) leftovers at the front
        const int rest = pkc - done_n;
        uint2 en = make_uint2(0u, 0u);
        if (lane < rest) en = pk[done_n + lane];
        __syncwarp();
        if (lane < rest) pk[lane] = en;
        __syncwarp();
        pkc = rest;
      }
    };

    for (int q = 0; q < NQ; ++q) {
      const int s = q % kStages;
      mbar_wait(smem_u32(&ctl.full[s]), (q / kStages) & 1, status, 0xE0000002u);
      if (q == 0 && tid == 0) RTM3D_TL(2, globaltimer_ns());
      const int c_lo = ys + j * CR, c_hi = min(c_lo + CR, ye);
      const int top = max(c_lo - 1, 0);       // image row of the stage's first row
      const unsigned char* stage = ring + static_cast<size_t>(s) * g.stage_bytes;
      const int g_begin = (c_lo - top) * gpr;
      const int g_end = (c_hi - top) * gpr;
      const uint32_t plane_flat = (MODE == kModeMain) ? static_cast<uint32_t>(plane0 + pl) * HW : 0u;
      float tf = ctl.t_filter;
      int wlc = 0;

      auto pass2a = [&]() {
        __syncwarp();
        tf = ctl.t_filter;
        for (int i0 = 0; i0 < wlc; i0 += 32) {
          ++n_flush;
          const int i = i0 + lane;
          uint32_t mask = 0;
          int gi = 0;
          if (i < wlc) {
            gi = wl[i];
            float v[E];
            Group<T>::load(stage + static_cast<size_t>(gi) * 16, v);
#pragma unroll
            for (int e = 0; e < E; ++e) mask |= (v[e] >= tf) ? (1u << e) : 0u;
          }
          const int rl = static_cast<int>(__umulhi(static_cast<uint32_t>(gi), g.gpr_magic));  // row within the stage
          const int c0 = (gi - rl * gpr) * E;                                                // first column
          const int y = top + rl;                                                            // image row
          const unsigned char* rmid = stage + static_cast<size_t>(rl) * g.row_bytes;
          const unsigned char* rup = rmid - g.row_bytes;   // resident whenever y > 0
          const unsigned char* rdn = rmid + g.row_bytes;   // resident whenever y + 1 < H
          const bool has_up = y > 0, has_dn = y + 1 < H;
          // four elements at a time (bounds what one step can add to the peak list): test, then every lane appends its
          // survivors behind those of the lower lanes
          for (int part = 0; part < E; part += 4) {
            uint32_t surv = 0;           // bit e: element e survived
            uint32_t m2 = mask & (0xFu << part);
            while (m2) {
              const int e = __ffs(m2) - 1;
              m2 &= m2 - 1;
              const int x = c0 + e;
              const float xc = ring_elem<T>(rmid, x);
              const bool has_l = x > 0, has_r = x + 1 < W;
              // neighbours (logits); -inf where the image ends = max_pool2d's implicit padding
              const int xl = has_l ? x - 1 : x, xr = has_r ? x + 1 : x;
              float nb[8];
              nb[0] = has_l ? ring_elem<T>(rmid, xl) : -INFINITY;
              nb[1] = has_r ? ring_elem<T>(rmid, xr) : -INFINITY;
              nb[2] = (has_up && has_l) ? ring_elem<T>(rup, xl) : -INFINITY;
              nb[3] = has_up ? ring_elem<T>(rup, x) : -INFINITY;
              nb[4] = (has_up && has_r) ? ring_elem<T>(rup, xr) : -INFINITY;
              nb[5] = (has_dn && has_l) ? ring_elem<T>(rdn, xl) : -INFINITY;
              nb[6] = has_dn ? ring_elem<T>(rdn, x) : -INFINITY;
              nb[7] = (has_dn && has_r) ? ring_elem<T>(rdn, xr) : -INFINITY;
              const float mx = fmaxf(fmaxf(fmaxf(nb[0], nb[1]), fmaxf(nb[2], nb[3])), fmaxf(fmaxf(nb[4], nb[5]), fmaxf(nb[6], nb[7])));
              // a neighbour this much larger is larger after the sigmoid too (logits <= 2 do not collapse that far)
              if (xc <= kSatKnee && xc >= kDenormKnee && mx > xc + kTieTol) continue;
              if (neighbour_needs_exact(mx, xc)) {
                // rare: a neighbour within the collapse distance (or in the saturated / denormal range)
                const float sc = sigmoid_ref(xc);
                bool peak = true;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  if (neighbour_needs_exact(nb[k], xc) && sigmoid_ref(nb[k]) > sc) peak = false;
                if (!peak) continue;
              }
              surv |= 1u << e;
            }
            __syncwarp();
            // exclusive prefix of the survivor counts over the lanes
            const int mine = __popc(surv);
            int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, incl, d);
              if (lane >= d) incl += t;
            }
            int pos = pkc + incl - mine;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            while (surv) {
              const int e = __ffs(surv) - 1;
              surv &= surv - 1;
              const int x = c0 + e;
              pk[pos++] = make_uint2(plane_flat + static_cast<uint32_t>(y) * W + x, __float_as_uint(ring_elem<T>(rmid, x)));
            }
            pkc += total;
            __syncwarp();
            if (pkc > kPeakList - 32 * 4) pass2b(false);
          }
        }
        __syncwarp();
        wlc = 0;
      };

      // one scan iteration over U*32 groups starting at group `first` (lane-interleaved)
      auto scan_it = [&](int first, auto u_tag) {
        constexpr int U = decltype(u_tag)::value;
        const int base = first + lane;
        const unsigned char* src = stage + static_cast<size_t>(base) * 16;
        bool hit[U];
        bool any = false;
        if (first + 32 * U <= g_end) {
          // full iteration: no bounds predicates, loads at immediate offsets from one base
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float v[E];
            Group<T>::load(src + u * 512, v);
            float m = v[0];
#pragma unroll
            for (int e = 1; e < E; ++e) m = fmaxf(m, v[e]);
            hit[u] = (m >= tf);
            any |= hit[u];
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int gi = base + 32 * u;
            float m = -INFINITY;
            if (gi < g_end) {
              float v[E];
              Group<T>::load(src + u * 512, v);
              m = v[0];
#pragma unroll
              for (int e = 1; e < E; ++e) m = fmaxf(m, v[e]);
            }
            hit[u] = (gi < g_end) && (m >= tf);  // (-inf >= -inf holds: padded lanes must not hit when tf = -inf)
            any |= hit[u];
          }
        }
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t bal = __ballot_sync(0xffffffffu, hit[u]);
            if (bal) {
              if (hit[u]) wl[wlc + __popc(bal & ((1u << lane) - 1u))] = static_cast<uint16_t>(base + 32 * u);
              wlc += __popc(bal);
            }
          }
        }
      };

      if (q == 0 && !(g.debug & 1)) {
        // Warm-up: the threshold is still the score threshold (or -inf), under which most pixels "hit".  Scan the
        // first chunk in small rounds, hand the peaks to the selector at once, and wait (once, bounded) for the first
        // threshold as soon as enough candidates are out, instead of pushing a whole chunk through the slow path.
        const uint32_t t0_bits = __float_as_uint(p.t0);
        const uint32_t enough = static_cast<uint32_t>(max(K + 32, 64));
        const int n_it = (g_end - g_begin + 31) / 32;
        for (int it = warp; it < n_it; it += kScanWarps) {
          if (__float_as_uint(ctl.t_filter) == t0_bits && ctl.q_tail >= enough) {
            const long long t0 = clock64();
            while (__float_as_uint(ctl.t_filter) == t0_bits && clock64() - t0 < 40000) __nanosleep(32);
          }
          tf = ctl.t_filter;
          scan_it(g_begin + it * 32, std::integral_constant<int, 1>{});
          if (wlc) { pass2a(); pass2b(true); }
        }
      } else {
        const int per_it = 32 * kUnroll;
        const int n_it = (g_end - g_begin + per_it - 1) / per_it;
        for (int it = warp; it < n_it; it += kScanWarps) {
          tf = ctl.t_filter;
          scan_it(g_begin + it * per_it, std::integral_constant<int, kUnroll>{});
          if (wlc > kWorkList - per_it) pass2a();
        }
        if (wlc) pass2a();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctl.empty[s]));   // chunks are self-contained: the stage is free again
      if (++j == cpp) { j = 0; ++pl; }
    }
    pass2b(true);
    __threadfence_block();
    __syncwarp();
    if (tid == 0) { RTM3D_TL(3, globaltimer_ns()); RTM3D_TL(12, static_cast<unsigned long long>(n_flush)); }
    if (lane == 0) atomicAdd(const_cast<uint32_t*>(&ctl.done), 1u);
  } else if (warp == kSelWarp) {
    // ================================ selector ================================
    int count = 0;
    uint32_t head = 0;
    float tcur = p.t0;
    const int cap = g.list_cap;
    int next_update = max(K + 32, 64);   // list size at which the threshold is next re-derived
    int n_prune = 0;
    const long long t0 = clock64();
    while (true) {
      const uint32_t tail = ctl.q_tail;
      const bool finished = (ctl.done == static_cast<uint32_t>(kScanWarps));
      const uint32_t pending = (finished ? ctl.q_tail : tail) - head;
      if (pending == 0) {
        if (finished) break;
        __nanosleep(100);
        if (clock64() - t0 > 8000000000LL) { if (status) atomicExch(status, 0xE0000005u); __threadfence_system(); __trap(); }
        continue;
      }
      if (pending < 32u && !finished) { __nanosleep(100); continue; }   // batch up: this warp is the serial resource
      const uint32_t avail = min(pending, 32u);
      const uint32_t slot = (head + lane) & (kQCap - 1);
      uint64_t key = 0;
      if (static_cast<uint32_t>(lane) < avail) key = *reinterpret_cast<volatile uint64_t*>(&qkey[slot]);
      const uint32_t ready = __ballot_sync(0xffffffffu, key != 0ull);
      const int take = (ready == 0xffffffffu) ? 32 : (__ffs(~ready) - 1);
      if (take == 0) continue;
      bool keep = false;
      float x = 0.f;
      if (lane < take) {
        x = qx[slot];
        *reinterpret_cast<volatile uint64_t*>(&qkey[slot]) = 0ull;
        keep = (x >= tcur);   // pushed before the threshold last rose: already out
      }
      const uint32_t kb = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = count + __popc(kb & ((1u << lane) - 1u));
        lhi[pos] = static_cast<uint32_t>(key >> 32);
        llo[pos] = static_cast<uint32_t>(key);
        lx[pos] = x;
      }
      count += __popc(kb);
      __threadfence_block();
      __syncwarp();
      head += take;
      if (lane == 0) ctl.q_head = head;
      if (count >= next_update || count > cap - 32) {
        float xp;
        ++n_prune;
        const bool must_shrink = count > cap - 32 - K;
        count = warp_prune(lhi, llo, lx, count, K, must_shrink ? max(count - 64 - K, K) : count, true, lane, xp);
        const float tn = filter_from_kth_logit(xp);
        if (tn > tcur) { tcur = tn; if (lane == 0) ctl.t_filter = tn; }
        next_update = min(cap - 32, count + K);
      }
    }
    if (lane == 0) { RTM3D_TL(4, globaltimer_ns()); RTM3D_TL(9, static_cast<unsigned long long>(n_prune)); RTM3D_TL(11, static_cast<unsigned long long>(head)); }
    // The scan is over (every bulk copy has been consumed), so the ring is free: sort the survivors there.
    const int keep_max = max(min(2 * K, kWarpSortMax), K);
    if (count > keep_max) {
      float xp;
      count = warp_prune(lhi, llo, lx, count, K, keep_max, true, lane, xp);
    }
    if (lane == 0) { RTM3D_TL(5, globaltimer_ns()); RTM3D_TL(10, static_cast<unsigned long long>(count)); }
    if (lane == 0) ctl.list_count = static_cast<uint32_t>(count);
  }

  // ================================ final: merge + epilogue (all threads) ================================
  __syncthreads();
  if (tid == 0) RTM3D_TL(6, globaltimer_ns());
  uint64_t* fin = reinterpret_cast<uint64_t*>(ring);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(ring + static_cast<size_t>(g.fin_cap) * 8);
  int n = static_cast<int>(ctl.list_count);
  if (n <= kWarpSortMax && !(g.debug & 2)) {
    // Each scanner warp sorts 32 survivors in registers (shuffles, no barriers); a key's final rank is its position in
    // its own run plus, for every other run, the number of larger keys there (binary search).  Keys are distinct.
    uint64_t* runs = fin + kWarpSortMax;
    if (warp < kScanWarps) {
      const int idx = warp * 32 + lane;
      SampleKey sk;
      sk.hi = idx < n ? lhi[idx] : 0u;
      sk.lo = idx < n ? llo[idx] : 0u;
      sk.x = 0.f;
      if (warp * 32 < n) warp_sort_desc(sk, lane);
      runs[idx] = (static_cast<uint64_t>(sk.hi) << 32) | sk.lo;
    }
    __syncthreads();
    if (warp < kScanWarps) {
      const uint64_t key = runs[warp * 32 + lane];
      if (key != 0ull) {
        int rank = lane;
        for (int v = 0; v * 32 < n; ++v) {
          if (v == warp) continue;
          const uint64_t* r = runs + v * 32;
          int lo = 0, hi = 32;
#pragma unroll
          for (int it = 0; it < 6; ++it) {   // 33 possible answers (0..32): six halvings
            const int mid = (lo + hi) >> 1;
            if (lo < hi) { if (r[mid] > key) lo = mid + 1; else hi = mid; }
          }
          rank += lo;
        }
        fin[rank] = key;
      }
    }
    __syncthreads();
  } else {
    // large K: more survivors than the rank merge is sized for
    const int npad = next_pow2(max(n, 1));
    for (int i = tid; i < npad; i += kStreamThreads) fin[i] = (i < n) ? ((static_cast<uint64_t>(lhi[i]) << 32) | llo[i]) : 0ull;
    __syncthreads();
    block_bitonic_sort_desc(fin, npad);
  }
  int have = min(n, K);
  if (S > 1) {
    for (int i = tid; i < have; i += kStreamThreads) best[i] = fin[i];
    if (tid == 0) ctl.best_count = static_cast<uint32_t>(have);
    cluster_sync_all();
    if (rank == 0) {
      // pull the peers' local top-K lists through distributed shared memory, then sort the union
      n = have;
      for (int r = 1; r < S; ++r) {
        const uint32_t rc = ld_dsmem_u32(map_to_rank(smem_u32(&ctl.best_count), r));
        const uint32_t rb = map_to_rank(smem_u32(best), r);
        for (int i = tid; i < static_cast<int>(rc); i += kStreamThreads) fin[n + i] = ld_dsmem_u64(rb + 8u * i);
        n += static_cast<int>(rc);
      }
      const int npad = next_pow2(max(n, 1));
      for (int i = n + tid; i < npad; i += kStreamThreads) fin[i] = 0ull;
      __syncthreads();
      block_bitonic_sort_desc(fin, npad);
      have = min(n, K);
    }
  }
  if (tid == 0) RTM3D_TL(7, globaltimer_ns());
  if (rank == 0) block_emit<T, MODE>(p, b, plane0, fin, have, scratch);
  if (g.timeline) { __syncthreads(); if (tid == 0) RTM3D_TL(8, globaltimer_ns()); }
  if (S > 1) cluster_sync_all();  // peers keep their shared memory alive until rank 0 has read it
}

// ---------------------------------------------------------------------------------------------------------------
static bool make_geom(const DecodeParams& p, int dtype, int mode, int cluster_override, StreamGeom& g) {
  const int es = dtype == 0 ? 4 : 2;
  const int E = 16 / es;
  if (p.W % E != 0) return false;
  if (reinterpret_cast<uintptr_t>(p.hm) % 16 != 0) return false;
  const int row_bytes = p.W * es;
  if (row_bytes > 16384) return false;
  g.row_bytes = row_bytes;
  g.list_cap = 2 * p.K + 256;
  if (g.list_cap < 768) g.list_cap = 768;
  // cluster size: enough CTAs to cover the chip about twice, rows permitting
  const long long nprob = (mode == kModeMain) ? p.B : static_cast<long long>(p.B) * p.C;
  int S = 1;
  if (cluster_override > 0) S = cluster_override;
  else while (S < kMaxCluster && nprob * S < 296 && p.H / (2 * S) >= 8) S *= 2;
  if (S > kMaxCluster) S = kMaxCluster;
  while (S > 1 && p.H < S) S /= 2;
  g.cluster = S;
  // final buffer (aliases the ring): the padded survivor list or the union of the cluster's top-K lists, then scratch
  const int keep_max = (2 * p.K < kWarpSortMax ? 2 * p.K : kWarpSortMax) > p.K ? (2 * p.K < kWarpSortMax ? 2 * p.K : kWarpSortMax) : p.K;
  const int own = keep_max <= kWarpSortMax ? 2 * kWarpSortMax : next_pow2(keep_max);
  const int uni = next_pow2(S * p.K);
  const size_t fin_need = static_cast<size_t>(own > uni ? own : uni);
  const size_t fin_bytes = fin_need * 8 + (3 * static_cast<size_t>(p.K) + 8) * 4;
  g.fin_cap = static_cast<int>(fin_need);
  // chunk geometry: centre rows per chunk so that a stage is about (kSmemBudget - fixed parts) / kStages bytes, evened out over the strip
  const int strip = (p.H + S - 1) / S;   // most rows a CTA owns
  const size_t other = static_cast<size_t>(kQCap) * 12 + static_cast<size_t>(g.list_cap) * 12 +
                       static_cast<size_t>(kScanWarps) * (kWorkList * 2 + kPeakList * 8) + static_cast<size_t>(p.K) * 8 + 64;
  const int stage_target = other + 3 * 4096 < static_cast<size_t>(kSmemBudget) ? static_cast<int>((kSmemBudget - other) / kStages) : 4096;
  int cr = stage_target / row_bytes - 2;
  if (cr < 1) cr = 1;
  if (static_cast<size_t>(kStages) * (cr + 2) * row_bytes < fin_bytes)
    cr = static_cast<int>((fin_bytes + static_cast<size_t>(kStages) * row_bytes - 1) / (static_cast<size_t>(kStages) * row_bytes)) - 2;
  if (cr < 1) cr = 1;
  if (cr > strip) cr = strip;
  const int nchunks = (strip + cr - 1) / cr;
  if (static_cast<size_t>(kStages) * ((strip + nchunks - 1) / nchunks + 2) * row_bytes >= fin_bytes) cr = (strip + nchunks - 1) / nchunks;
  g.chunk_rows = cr;
  g.stage_bytes = (cr + 2) * row_bytes;
  const size_t ring_bytes = static_cast<size_t>(kStages) * g.stage_bytes;
  if (ring_bytes < fin_bytes) return false;
  {
    const unsigned gpr = static_cast<unsigned>(p.W / E);
    g.gpr_magic = static_cast<unsigned>((0x100000000ULL + gpr - 1) / gpr);
  }
  if (static_cast<size_t>(g.stage_bytes) / 16 > 65535) return false;
  size_t o = ring_bytes + static_cast<size_t>(kQCap) * 12 + static_cast<size_t>(g.list_cap) * 12 +
             static_cast<size_t>(kScanWarps) * kWorkList * 2;
  o = (o + 15) & ~size_t(15);
  o += static_cast<size_t>(kScanWarps) * kPeakList * 8;
  o += static_cast<size_t>(p.K) * 8;
  g.smem = static_cast<unsigned>(o);
  return o <= 220 * 1024;
}

void debug_set_timeline(unsigned long long* ptr) { g_timeline = ptr; }

bool stream_eligible(const DecodeParams& p, int dtype, int mode) {
  StreamGeom g{};
  return make_geom(p, dtype, mode, 0, g);
}

template <typename T, int MODE>
static int launch_stream_t(const DecodeParams& p, const StreamGeom& g, uint32_t* status, cudaStream_t s) {
  auto kern = decode_stream_kernel<T, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g.smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long nprob = (MODE == kModeMain) ? p.B : static_cast<long long>(p.B) * p.C;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(nprob * g.cluster));
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(g.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, p, g, status);
  return static_cast<int>(e);
}

int launch_stream(const DecodeParams& p, int dtype, int mode, cudaStream_t s) {
  StreamGeom g{};
  if (!make_geom(p, dtype, mode, p.cluster_override, g)) return -1000;
  g.timeline = g_timeline;
  g.debug = p.debug;
  uint32_t* status = p.status;
  if (dtype == 0)
    return mode == kModeMain ? launch_stream_t<float, kModeMain>(p, g, status, s) : launch_stream_t<float, kModeKpt>(p, g, status, s);
  return mode == kModeMain ? launch_stream_t<__nv_bfloat16, kModeMain>(p, g, status, s)
                           : launch_stream_t<__nv_bfloat16, kModeKpt>(p, g, status, s);
}

}  // namespace rtm3d
