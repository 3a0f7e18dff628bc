// Plane-resident scan kernel for sm_100a: the streaming half of the decode.
//
// Work unit = one STRIP: a block of whole rows of one heat-map plane that fits the shared-memory ring (a 96x320 fp32
// plane is ONE strip; larger planes are cut into a few strips).  For every strip the kernel writes a short candidate list
// that provably contains the strip's K best 3x3 peaks (utils/model_utils.py:17-26 applied to the sigmoid of
// models/model.py:85,107, ordered by (score desc, index asc) as torch.topk on CUDA orders them, models/model.py:90,112),
// or all of its peaks when there are fewer.  The select kernels (select.cu) merge the lists of a selection problem --
// the C planes of an image (flat top-K over C*H*W, models/model.py:87-98) or one keypoint plane (:109-114) -- and sort.
//
// One persistent CTA per SM, strips handed out by an atomic counter (no static assignment, no tail imbalance):
//   producer (1 lane)   streams a strip as chunks of `slot_rows` rows, one cp.async.bulk (SASS UBLKCP) per chunk, into a ring
//                       of S slots; the rows of a strip lie in consecutive slots, so the strip is RESIDENT and halo-free
//                       once its last chunk has landed.  Slots go back to the producer when the strip is finished; the
//                       other S - n_chunks slots keep the copy engine busy meanwhile.
//   16 scan warps       pass 1, chunk by chunk as they land: one LDS.128 per 16-byte group, the group's maximum stays in a
//                       REGISTER (16 per lane cover a strip).  Then a logit threshold T is picked from an order statistic
//                       of the lane maxima (the smallest of the maxima of small lane groups: no state from earlier planes
//                       or launches), pass 2 runs over the registers only, and the few groups with a pixel >= T get the exact
//                       3x3 peak test; peaks go to the strip's list in global memory as (logit, index) keys -- the sigmoid of
//                       the ~270 listed pixels is evaluated by the select kernel, not here.
//   verification        every peak with logit >= T is in the list, every other pixel has score <= s(T) (the sigmoid is
//                       monotone: tests/test_sigmoid_gpu.py sweeps all 2^32 inputs).  So the list holds the K best keys as
//                       soon as K of its pixels have a score > s(T) STRICTLY, which logit > T + collapse distance guarantees.
//                       If not (T was too high), T is lowered and pass 2 repeated for the new pixels only -- the strip is
//                       still resident, nothing is re-read from HBM.
//   exact path          when the list overflows (plateaus, saturated or heavily tied maps), the K-th best (score, index) key
//                       is found by an MSB-first radix select over the resident strip and exactly the K best are collected.
// The result never depends on T, on scheduling or on earlier launches.
#include <cuda_runtime.h>

#include <cmath>

#include "common.cuh"
#include "params.h"
#include "scan_common.cuh"

namespace rtm3d {

constexpr int kScanWarps = 15;                        // scan warps; warp kScanWarps is the producer (16 warps: 128 registers each)
constexpr int kScanConsumers = kScanWarps * 32;
constexpr int kScanThreads = kScanConsumers + 32;
constexpr int kMaxChunks = 8;                         // chunks (ring slots) per strip
constexpr int kGroupRegs = 16;                        // registers of group maxima per lane: register r of scan thread t holds the
                                                      // maximum of the strip's 16-byte group r * 480 + t
constexpr int kChunkGroups = 1024;                    // groups per chunk at most (16 KB)
constexpr int kSelRegs = 12;                          // register rows the first threshold is computed from
constexpr int kMaxSlots = 16;
constexpr int kWlEntries = 64;                        // per-warp worklist (group indices waiting for dense batches)
constexpr int kDeepenFactor = 4;

struct ScanGeom {
  int S;                 // ring slots
  int slot_rows, slot_bytes, row_bytes, gpr;
  unsigned gpr_magic;    // ceil(2^32 / gpr)
  unsigned slot_bytes_magic;   // ceil(2^32 / slot_bytes)
  int Sp;                // strips per plane; strip s owns rows [(s*H)/Sp, ((s+1)*H)/Sp)
  int list_cap;          // keys per candidate list
  int rank_j;            // first attempt: every warp takes the rank_j-th largest of its 32 lane maxima (over the first kSelRegs
                         // register rows), T = the smallest of the warps' values; 0 = bisection (exact count per warp, slower)
  int target0;           // first attempt (bisection) and base of the deepening: groups per warp with a pixel >= T aimed at
  int n_strips;
  int grid;
  int debug;             // 1: first threshold forced too high (deepening path), 2: tiny list (exact path), 3: both
  unsigned smem;
};

struct StripDesc {       // written by the producer, read by the scan warps
  int strip;             // strip id = output slot; -1 = no more work
  int slot0;             // ring slot of the strip's first chunk
  int n_chunks;
  int top_halo;          // 1 when the image row above the strip is loaded in front of it
  int rows;              // owned rows
  int y0;                // first owned image row
  uint32_t flat_base;    // added to y*W+x: c*H*W for main plane c (one flat top-K per image), 0 for keypoint planes
  int is_main;
};

struct __align__(16) ScanCtl {
  unsigned long long full_a[kMaxSlots];   // producer -> scan warps, per strip in flight (queue entry): the chunks that hold the first
                                          // kSelRegs register rows have landed ...
  unsigned long long full_b[kMaxSlots];   // ... the rest of the strip has landed
  unsigned long long empty[kMaxSlots];    // scan warps -> producer: slot free
  StripDesc queue[kMaxSlots];             // the producer is never more than S strips ahead
  float t_warp[kScanWarps];
  uint32_t t_sub[kScanWarps];             // the warps' order statistics (ordered bits of the float)
  uint32_t n_list[2];                     // keys appended to the strip's list (runs past the capacity when the list is full)
  uint32_t n_strict[2];                   // ... of which provably score > s(T)
  uint32_t hist[256];                     // exact path: digit histogram
  uint32_t sel[4];                        // exact path: digit, keys still needed, total
  unsigned short wl[kScanWarps][kWlEntries];
};

enum ScanStat { kSsStrips = 0, kSsDeepen, kSsExact, kSsListKeys, kSsBatches, kSsWaitFirst, kSsPass1, kSsSelect, kSsPass2, kSsVerify, kSsFlush,
                kSsProdWait, kSsProdTotal, kSsSlots };
#ifdef RTM3D_DEV
#define SCAN_CLK() clock64()
#else
#define SCAN_CLK() 0ll
#endif

__device__ __forceinline__ void scan_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kScanConsumers) : "memory"); }

// Distance above a logit x that guarantees a strictly larger fp32 sigmoid (and below: a strictly smaller one): logits <= 2
// closer than 2^-14 may collapse to one fp32 value; above 2 the collapse distance grows like 2^-24 e^x (2^-12 e^x leaves a
// factor 4096); beyond 8 (scores within 3.4e-4 of 1.0) and in the denormal range nothing is promised -- exact sigmoids decide.
__device__ __forceinline__ float collapse_tol(float x) {
  if (!(x >= kDenormKnee) || x > 8.0f) return INFINITY;
  return (x <= kSatKnee) ? kTieTol : 2.44140625e-04f * __expf(x);
}

struct StripView {       // what the batches need to find a pixel of the resident strip
  const unsigned char* ring;
  uint32_t ring_bytes;
  uint32_t strip_off;    // byte offset of the strip's first loaded row in the ring: the rows of a strip lie in consecutive slots
                         // and a slot holds whole rows without padding, so the strip is ONE contiguous byte range (mod ring size)
  int row_bytes, gpr;
  unsigned gpr_magic;    // ceil(2^32 / gpr)
  int H, W;
  int top_halo, y0;
  uint32_t flat_base;
};

// The 3x3 neighbourhood of group `gidx` (strip-level index of a 16-byte group): v = its pixels, nb = the largest of the 8
// neighbours of each pixel (-inf outside the image: max_pool2d's implicit padding); also the image row and the group column.
template <typename T>
__device__ __forceinline__ void load_group_window(const StripView& a, int gidx, float (&v)[Grp<T>::E], float (&nb)[Grp<T>::E], int& y, int& c4) {
  constexpr int E = Grp<T>::E;
  const int ry = static_cast<int>(__umulhi(static_cast<uint32_t>(gidx), a.gpr_magic));
  c4 = gidx - ry * a.gpr;
  y = a.y0 + ry;
  // ring offsets of the group in its own row, the row above and the row below (the strip wraps at the end of the ring)
  const uint32_t rb = static_cast<uint32_t>(a.row_bytes);
  uint32_t oc = a.strip_off + static_cast<uint32_t>(ry + a.top_halo) * rb;
  if (oc >= a.ring_bytes) oc -= a.ring_bytes;
  const uint32_t ou = (oc >= rb) ? oc - rb : oc + (a.ring_bytes - rb);
  uint32_t od = oc + rb;
  if (od >= a.ring_bytes) od -= a.ring_bytes;
  const uint32_t gx = static_cast<uint32_t>(c4) * 16u;
  const unsigned char* gp = a.ring + oc + gx;
  Grp<T>::load(gp, v);
  const bool has_up = y > 0, has_dn = y + 1 < a.H, has_l = c4 > 0, has_r = c4 + 1 < a.gpr;
  float up[E + 2], dn[E + 2];
  load_window_row<T>(a.ring + ou + gx, has_up, has_l, has_r, up);
  load_window_row<T>(a.ring + od + gx, has_dn, has_l, has_r, dn);
  const float ml = has_l ? Grp<T>::elem(gp, -1) : -INFINITY;
  const float mr = has_r ? Grp<T>::elem(gp, E) : -INFINITY;
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const float au = fmaxf(fmaxf(up[i], up[i + 1]), up[i + 2]);
    const float ad = fmaxf(fmaxf(dn[i], dn[i + 1]), dn[i + 2]);
    const float left = (i == 0) ? ml : v[i - 1];
    const float right = (i == E - 1) ? mr : v[i + 1];
    nb[i] = fmaxf(fmaxf(au, ad), fmaxf(left, right));
  }
}

// Is the pixel with logit xc a 3x3 peak in the SIGMOID domain, given its largest neighbour xn?  Decided on the logits where
// that is provably the same (a neighbour further above than the collapse distance is larger after the sigmoid too; one that
// is not above at all is not larger, the sigmoid being monotone); the exact sigmoids are compared only in between.
__device__ __forceinline__ bool is_peak(float xc, float xn) {
  if (!(xn > xc)) return true;                                     // equal-valued neighbours are all kept
  if (xn > xc + collapse_tol(xc)) return false;
  return !(sigmoid_cold(xn) > sigmoid_cold(xc));
}

enum BatchMode { kBmHist = 1, kBmCollect = 2 };

struct ExactArgs {
  StripView view;
  float lim, t_lo;
  int mode;
  unsigned long long* out;      // the strip's list in global memory
  uint32_t* n_list;
  int cap;
  unsigned long long prefix, mask, kth;
  int shift;
  uint32_t* hist;
};

// Exact path, one dense batch: (score, index) keys of the group's peaks are counted in the digit histogram of the radix
// select, or collected when they are among the K best.
template <typename T>
static __device__ __noinline__ void exact_batch(const ExactArgs& a, int gidx, int lane) {
  constexpr int E = Grp<T>::E;
  uint32_t pm = 0;
  float v[E], nb[E];
  int y = 0, c4 = 0;
  if (gidx >= 0) {
    load_group_window<T>(a.view, gidx, v, nb, y, c4);
#pragma unroll
    for (int i = 0; i < E; ++i)
      if (v[i] >= a.t_lo) pm |= 1u << i;
  }
  const uint32_t lt = (1u << lane) - 1u;
  while (__any_sync(0xffffffffu, pm != 0u)) {
    bool cand = false;
    unsigned long long key = 0ull;
    if (pm != 0u) {
      const int i = __ffs(pm) - 1;
      pm &= pm - 1;
      float xc = v[0], xn = nb[0];
#pragma unroll
      for (int j = 1; j < E; ++j) { if (j == i) { xc = v[j]; xn = nb[j]; } }
      const float sc = sigmoid_cold(xc);
      cand = sc > a.lim && is_peak(xc, xn);
      if (cand) key = make_key(sc, a.view.flat_base + static_cast<uint32_t>(y) * a.view.W + static_cast<uint32_t>(c4 * E + i));
    }
    if (a.mode == kBmHist) {
      if (cand && (key & a.mask) == a.prefix) atomicAdd(&a.hist[static_cast<uint32_t>(key >> a.shift) & 255u], 1u);
    } else {
      cand = cand && key >= a.kth;
      const uint32_t bal = __ballot_sync(0xffffffffu, cand);
      if (bal != 0u) {
        uint32_t base = 0;
        if (lane == __ffs(bal) - 1) base = atomicAdd(a.n_list, static_cast<uint32_t>(__popc(bal)));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        const uint32_t pos = base + static_cast<uint32_t>(__popc(bal & lt));
        if (cand && pos < static_cast<uint32_t>(a.cap)) a.out[pos] = key;
      }
    }
  }
}

// A threshold T with at least `target` of the warp's group maxima >= T (largest such T on a 16-bit grid of the float
// order); -inf when the warp has fewer than `target` groups.
static __device__ __noinline__ float warp_bisect_threshold(const float (&m)[kGroupRegs], int target, int lane) {
  uint32_t res = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 16; --bit) {
    const uint32_t cand = res | (1u << bit);
    const float cf = f32_unord(cand);
    int c = 0;
#pragma unroll
    for (int r = 0; r < kGroupRegs; ++r) c += (m[r] >= cf) ? 1 : 0;     // NaN patterns compare false
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= target) res = cand;
  }
  if (res < 0x00800000u) return -INFINITY;                                // below every finite float
  return f32_unord(res);
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads, 1) scan_planes_kernel(const ScanParams p, const ScanGeom g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ ScanCtl ctl;
  constexpr int E = Grp<T>::E;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = g.S;
  unsigned char* ring = smem;

  if (tid == 0) {
    for (int s = 0; s < kMaxSlots; ++s) {
      pl::mbar_init(pl::smem_u32(&ctl.full_a[s]), 1);
      pl::mbar_init(pl::smem_u32(&ctl.full_b[s]), 1);
      pl::mbar_init(pl::smem_u32(&ctl.empty[s]), 1);
    }
    ctl.n_list[0] = 0u; ctl.n_list[1] = 0u;
    ctl.n_strict[0] = 0u; ctl.n_strict[1] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#ifdef RTM3D_DEV
    if (p.stats) p.stats[64 + 2 * blockIdx.x] = pl::globaltimer_ns();
#endif
  }
  __syncthreads();

  const int n_strips = g.n_strips;
  const int HW = p.H * p.W;
  const size_t plane_bytes = static_cast<size_t>(HW) * sizeof(T);
  const int n_main_planes = p.C > 0 ? p.B * p.C : 0;

  if (warp == kScanWarps) {
    // ================================ producer ================================
    if (lane == 0) {
      const unsigned long long policy = pl::l2_policy_evict_first();
      const uint32_t total = static_cast<uint32_t>(n_strips) + gridDim.x;   // the counter wraps to 0 after the launch's last fetch
      uint32_t slot = 0, par = 0, issued = 0, seq = 0;
      long long pw = 0;
      const long long pt0 = SCAN_CLK();
      uint32_t next = atomicInc(p.queue, total - 1u);
      while (true) {
        const uint32_t idx = next;
        const bool end = idx >= static_cast<uint32_t>(n_strips);
        if (!end) next = atomicInc(p.queue, total - 1u);          // (the reply is not needed before the next strip)
        StripDesc d;
        d.strip = end ? -1 : static_cast<int>(idx);
        d.slot0 = static_cast<int>(slot);
        const unsigned char* src = nullptr;
        int n_loaded = 0;
        if (!end) {
          const int plane = static_cast<int>(idx) / g.Sp, s = static_cast<int>(idx) - plane * g.Sp;
          d.y0 = (s * p.H) / g.Sp;
          const int y1 = ((s + 1) * p.H) / g.Sp;
          const int top = max(d.y0 - 1, 0), bot = min(y1 + 1, p.H);
          d.top_halo = d.y0 - top;
          d.rows = y1 - d.y0;
          n_loaded = bot - top;
          d.n_chunks = (n_loaded + g.slot_rows - 1) / g.slot_rows;
          d.is_main = plane < n_main_planes ? 1 : 0;
          if (d.is_main) {
            d.flat_base = static_cast<uint32_t>(plane % p.C) * static_cast<uint32_t>(HW);
            src = reinterpret_cast<const unsigned char*>(p.hm_main) + static_cast<size_t>(plane) * plane_bytes;
          } else {
            d.flat_base = 0u;
            src = reinterpret_cast<const unsigned char*>(p.hm_kpt) + static_cast<size_t>(plane - n_main_planes) * plane_bytes;
          }
          src += static_cast<size_t>(top) * g.row_bytes;
        } else {
          d.n_chunks = 0; d.top_halo = 0; d.rows = 0; d.y0 = 0; d.flat_base = 0u; d.is_main = 0;
        }
        // two barriers per strip: the scan warps wait once for the part the first threshold is computed from and once for
        // the rest, not once per chunk
        const uint32_t q = seq & (kMaxSlots - 1);
        const uint32_t bar_a = pl::smem_u32(&ctl.full_a[q]), bar_b = pl::smem_u32(&ctl.full_b[q]);
        int chunks_a = 0;
        if (!end) {
          const uint32_t halo_bytes = static_cast<uint32_t>(d.top_halo) * g.row_bytes;
          const uint32_t bytes_sel = halo_bytes + static_cast<uint32_t>(min(kSelRegs * kScanConsumers, d.rows * g.gpr)) * 16u;
          chunks_a = min(d.n_chunks, static_cast<int>(__umulhi(bytes_sel + g.slot_bytes - 1u, g.slot_bytes_magic)));
        }
        const int nch = end ? 1 : d.n_chunks;
        for (int c = 0; c < nch; ++c) {
          if (issued >= static_cast<uint32_t>(S)) {
            const long long w0 = SCAN_CLK();
            pl::mbar_wait(pl::smem_u32(&ctl.empty[slot]), par ^ 1u, p.status, 0xE2000001u, 32);
            pw += SCAN_CLK() - w0;
          }
          if (c == 0) {
            ctl.queue[q] = d;                                      // (made visible to the scan warps by the barrier operation below)
            if (end) {
              pl::mbar_arrive(bar_a);
            } else {
              const uint32_t total = static_cast<uint32_t>(n_loaded) * g.row_bytes;
              const uint32_t bytes_a = min(static_cast<uint32_t>(chunks_a) * g.slot_bytes, total);
              pl::mbar_arrive_expect_tx(bar_a, bytes_a);
              if (total > bytes_a) pl::mbar_arrive_expect_tx(bar_b, total - bytes_a); else pl::mbar_arrive(bar_b);
            }
          }
          if (!end) {
            const int r0 = c * g.slot_rows;
            const uint32_t bytes = static_cast<uint32_t>(min(g.slot_rows, n_loaded - r0)) * g.row_bytes;
            pl::bulk_g2s(pl::smem_u32(ring + static_cast<size_t>(slot) * g.slot_bytes), src + static_cast<size_t>(r0) * g.row_bytes, bytes,
                         c < chunks_a ? bar_a : bar_b, policy);
          }
          ++issued;
          if (++slot == static_cast<uint32_t>(S)) { slot = 0; par ^= 1u; }
        }
        if (end) break;
        ++seq;
      }
#ifdef RTM3D_DEV
      if (p.stats) {
        atomicAdd(&p.stats[kSsProdWait], static_cast<unsigned long long>(pw));
        atomicAdd(&p.stats[kSsProdTotal], static_cast<unsigned long long>(SCAN_CLK() - pt0));
      }
#endif
      (void)pw; (void)pt0;
    }
  } else {
    // ================================ scan warps ================================
    uint32_t seq = 0;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned short* wl = ctl.wl[warp];
    long long st_deepen = 0, st_exact = 0, st_keys = 0, st_strips = 0;
    long long ph[6] = {0, 0, 0, 0, 0, 0};
    StripView view;
    view.ring = ring; view.ring_bytes = static_cast<uint32_t>(S) * g.slot_bytes; view.row_bytes = g.row_bytes;
    view.gpr = g.gpr; view.gpr_magic = g.gpr_magic; view.H = p.H; view.W = p.W;
    const int lane_g = warp * 32 + lane;
    while (true) {
      const long long c0 = SCAN_CLK();
      const uint32_t q = seq & (kMaxSlots - 1), qpar = (seq / kMaxSlots) & 1u;
      pl::mbar_wait(pl::smem_u32(&ctl.full_a[q]), qpar, p.status, 0xE2000002u, 32);
      const long long c1 = SCAN_CLK();
      const StripDesc d = ctl.queue[q];
      if (d.strip < 0) break;
      const int buf = static_cast<int>(seq & 1u);
      uint32_t* n_list = &ctl.n_list[buf];
      uint32_t* n_strict = &ctl.n_strict[buf];
      const int cap = (g.debug & 2) ? min(g.list_cap, p.K + 8) : g.list_cap;
      unsigned long long* out = p.cand + static_cast<size_t>(d.strip) * g.list_cap;
      view.strip_off = static_cast<uint32_t>(d.slot0) * g.slot_bytes;
      view.top_halo = d.top_halo; view.y0 = d.y0; view.flat_base = d.flat_base;
      const float t_floor = d.is_main ? p.t0 : -INFINITY;       // x < t0  =>  sigmoid(x) <= thresh
      const int G = d.rows * g.gpr;                               // owned 16-byte groups of the strip

      // ---- pass 1: group maxima into registers as the copies land.  The threshold is chosen after the first kSelRegs
      //      register rows -- an order statistic of a sample is as good as one of the whole strip, and it is verified anyway --
      //      so that choosing it overlaps the arrival of the strip's last chunks.
      float m[kGroupRegs];
      float tw = INFINITY;
      {
        const uint32_t halo_bytes = static_cast<uint32_t>(d.top_halo) * g.row_bytes;
        // ring offset of this thread's first group (threads beyond the strip's last group read that one: never used)
        uint32_t lane_off = view.strip_off + halo_bytes + static_cast<uint32_t>(min(lane_g, G - 1)) * 16u;
        if (lane_off >= view.ring_bytes) lane_off -= view.ring_bytes;
        const int n_full = G / kScanConsumers;
        const int n_mine = n_full + ((lane_g < G - n_full * kScanConsumers) ? 1 : 0);           // register rows in which this thread has a group
        auto load_reg = [&](int r) {
          // (branch-free: a register row without a group of this thread re-reads the thread's first group)
          const bool mine = r < n_mine;
          uint32_t o = lane_off + (mine ? static_cast<uint32_t>(r) * (kScanConsumers * 16u) : 0u);
          if (o >= view.ring_bytes) o -= view.ring_bytes;
          float v[E];
          Grp<T>::load(ring + o, v);
          float mx = v[0];
#pragma unroll
          for (int i = 1; i < E; ++i) mx = fmaxf(mx, v[i]);
          return mine ? mx : -INFINITY;
        };
#pragma unroll
        for (int r = 0; r < kSelRegs; ++r) m[r] = load_reg(r);
        if (g.rank_j > 0) {
          // ---- threshold of the first attempt: every warp takes the rank_j-th largest of its 32 lane maxima (rank_j rounds of
          //      warp maximum + knock-out), T = the smallest of the warps' values
          float lm = m[0];
#pragma unroll
          for (int r = 1; r < kSelRegs; ++r) lm = fmaxf(lm, m[r]);
          uint32_t o = (lm > -INFINITY) ? f32_ord(lm) : 0u;
          uint32_t mx = 0u;
#pragma unroll 1
          for (int it = 0; it < g.rank_j; ++it) {
            mx = __reduce_max_sync(0xffffffffu, o);
            o = (o == mx) ? 0u : o;                               // (equal maxima go together: the rank counts distinct values)
          }
          if (lane == 0) ctl.t_sub[warp] = mx;
          scan_bar_sync();
          const uint32_t tmin = __reduce_min_sync(0xffffffffu, ctl.t_sub[min(lane, kScanWarps - 1)]);
          tw = tmin < 0x00800000u ? -INFINITY : f32_unord(tmin);
        }
        pl::mbar_wait(pl::smem_u32(&ctl.full_b[q]), qpar, p.status, 0xE2000003u, 32);
#pragma unroll
        for (int r = kSelRegs; r < kGroupRegs; ++r) m[r] = load_reg(r);
      }
      const long long c2 = SCAN_CLK();

      // CTA-wide minimum of the warps' thresholds
      auto cta_threshold = [&](float tw) {
        if (lane == 0) ctl.t_warp[warp] = tw;
        scan_bar_sync();
        float t = ctl.t_warp[0];
#pragma unroll
        for (int w = 1; w < kScanWarps; ++w) t = fminf(t, ctl.t_warp[w]);
        return t;
      };
      // the groups (all registers) with a pixel >= t_lo, compacted into the warp's worklist: returns their number (nothing
      // is written when they do not fit: the caller then goes register by register)
      auto compact_all = [&](float t_lo) {
        const float tl = fmaxf(t_lo, -3.402823466e+38f);          // m >= -FLT_MAX <=> m > -inf: a register without a group, or a
        uint32_t hm = 0u;                                          // group of zero-score pixels, never counts
#pragma unroll
        for (int r = 0; r < kGroupRegs; ++r)
          if (m[r] >= tl) hm |= 1u << r;
        // one round per "next hit of every lane": positions by ballot (a lane rarely holds more than two hits)
        int total = 0;
        while (true) {
          const uint32_t bal = __ballot_sync(0xffffffffu, hm != 0u);
          if (bal == 0u) break;
          if (hm != 0u) {
            const int r = __ffs(hm) - 1;
            hm &= hm - 1;
            const int pos = total + __popc(bal & lt);
            if (pos < kWlEntries) wl[pos] = static_cast<unsigned short>(r * kScanConsumers + lane_g);
          }
          total += __popc(bal);
        }
        __syncwarp();
        return total;
      };
      // ... of ONE register row (at most 32 groups): the slow paths
      auto compact_one = [&](float t_lo, int r) {
        float mr = m[0];
#pragma unroll
        for (int q = 1; q < kGroupRegs; ++q) if (q == r) mr = m[q];
        const bool hit = mr >= t_lo && mr > -INFINITY;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (hit) wl[__popc(bal & lt)] = static_cast<unsigned short>(r * kScanConsumers + lane_g);
        __syncwarp();
        return __popc(bal);
      };

      int target = g.target0;
      if (g.rank_j == 0) tw = cta_threshold(warp_bisect_threshold(m, target, lane));
      if (g.debug & 1) tw = INFINITY;
      float t_lo = fmaxf(tw, t_floor), t_hi = INFINITY;
      if (tid == 0) { ctl.n_list[buf ^ 1] = 0u; ctl.n_strict[buf ^ 1] = 0u; }   // every warp has left the previous strip
      const long long c3 = SCAN_CLK();
      long long c4 = 0;

      // ---- pass 2 (+ deepening until the list provably holds the K best)
      bool overflow = false, raised = false;
      while (true) {
        // a listed pixel is counted as strictly above s(T) when its logit exceeds T by the collapse distance; for T beyond
        // the range where that distance is known, the exact sigmoids are compared
        const float t_strict = t_lo + collapse_tol(t_lo);
        const bool exact_strict = t_lo > 8.0f;
        const float s_t = exact_strict ? sigmoid_cold(t_lo) : 0.f;
        int my_strict = 0;
        int step = kGroupRegs;
#pragma unroll 1
        for (int r0 = 0; r0 < kGroupRegs;) {
          const int total = (step == 1) ? compact_one(t_lo, r0) : compact_all(t_lo);
          if (total > kWlEntries) { step = 1; continue; }            // (very low thresholds: register by register, <= 32 groups each)
          // dense batches of 32 groups: exact peak test of the pixels in [t_lo, t_hi), (logit, index) keys to the list --
          // one reservation per batch (prefix sum over the lanes' peak counts, one shared-memory atomic)
#pragma unroll 1
          for (int b0 = 0; b0 < total; b0 += 32) {
#ifdef RTM3D_DEV
            if (g.debug & 4) break;
#endif
            const int gi = (b0 + lane < total) ? static_cast<int>(wl[b0 + lane]) : -1;
            uint32_t pk = 0u, ps = 0u;              // peaks of this lane's group, and those counted as strict
            float v[E], nb[E];
            int y = 0, c4i = 0;
            if (gi >= 0) {
              load_group_window<T>(view, gi, v, nb, y, c4i);
#pragma unroll
              for (int i = 0; i < E; ++i) {
                if (v[i] >= t_lo && v[i] < t_hi && is_peak(v[i], nb[i])) {
                  pk |= 1u << i;
                  bool st = v[i] > t_strict;
                  if (exact_strict) st = sigmoid_cold(v[i]) > s_t;
                  if (st) ps |= 1u << i;
                }
              }
            }
            const int n_new = __reduce_add_sync(0xffffffffu, __popc(pk));
            if (n_new != 0) {
              uint32_t base = 0u;
              if (lane == 0) base = atomicAdd(n_list, static_cast<uint32_t>(n_new));
              base = __shfl_sync(0xffffffffu, base, 0);
              const uint32_t fl0 = d.flat_base + static_cast<uint32_t>(y) * p.W + static_cast<uint32_t>(c4i * E);
              while (true) {                                        // one round per "next peak of every lane"
                const uint32_t bal = __ballot_sync(0xffffffffu, pk != 0u);
                if (bal == 0u) break;
                if (pk != 0u) {
                  const int i = __ffs(pk) - 1;
                  pk &= pk - 1;
                  float xc = v[0];
#pragma unroll
                  for (int j = 1; j < E; ++j) if (j == i) xc = v[j];
                  const uint32_t pos = base + static_cast<uint32_t>(__popc(bal & lt));
                  if (pos < static_cast<uint32_t>(cap))
                    out[pos] = (static_cast<unsigned long long>(f32_ord(xc)) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - (fl0 + i));
                }
                base += static_cast<uint32_t>(__popc(bal));
              }
              my_strict += __reduce_add_sync(0xffffffffu, __popc(ps));
            }
          }
          __syncwarp();
          r0 += step;
        }
        if (lane == 0 && my_strict != 0) atomicAdd(n_strict, static_cast<uint32_t>(my_strict));
        if (c4 == 0) c4 = SCAN_CLK();
        scan_bar_sync();
        const uint32_t appended = *reinterpret_cast<volatile uint32_t*>(n_list);
        const uint32_t strict = *reinterpret_cast<volatile uint32_t*>(n_strict);
        if (appended > static_cast<uint32_t>(cap)) {
          // the list is full.  With K strict keys in it T was simply too low (the order statistic is noisy): start over ONCE
          // with a threshold aimed at ~2.3 K groups; otherwise (ties, plateaus, saturation) the exact path decides.
          if (!raised && strict >= static_cast<uint32_t>(p.K) && t_hi == INFINITY) {
            raised = true;
            st_deepen += 1;
            const float tn = cta_threshold(warp_bisect_threshold(m, g.target0, lane));   // (the barrier: every warp has read the counters)
            if (tn > t_lo) {
              if (tid == 0) { *n_list = 0u; *n_strict = 0u; }
              t_lo = tn;
              scan_bar_sync();
              continue;
            }
          }
          overflow = true;
          break;
        }
        const bool complete = !(t_lo > t_floor);                      // every candidate of the strip is in the list
        if (complete || strict >= static_cast<uint32_t>(p.K)) break;
        // ---- T was too high: lower it (4x more groups per attempt) and examine the pixels in [new T, old T)
        st_deepen += 1;
        target *= kDeepenFactor;
        float tn = (target >= kGroupRegs * 32) ? -INFINITY : warp_bisect_threshold(m, target, lane);
        tn = fmaxf(cta_threshold(tn), t_floor);
        if (!(tn < t_lo)) tn = t_floor;                               // no progress possible on this grid: take everything
        t_hi = t_lo;
        t_lo = tn;
        // every key listed so far has a logit >= the old T: all of them are strictly above s(new T) when the old T exceeds
        // the new T by the collapse distance (else the old count stays: those keys were strict before and still are)
        if (tid == 0 && t_hi > t_lo + collapse_tol(t_lo)) atomicAdd(n_strict, appended - strict);
      }
      if (overflow) {
        // ---- exact path: radix select of the K-th best (score, index) key over the resident strip
        st_exact += 1;
        const uint32_t strict = *reinterpret_cast<volatile uint32_t*>(n_strict);
        ExactArgs a;
        a.view = view;
        a.lim = d.is_main ? p.thresh : 0.0f;
        a.t_lo = (strict >= static_cast<uint32_t>(p.K)) ? t_lo : t_floor;   // the K best lie at or above it
        a.out = out; a.n_list = n_list; a.cap = cap; a.hist = ctl.hist;
        a.prefix = 0ull; a.mask = 0ull; a.kth = 0ull; a.shift = 0;
        auto scan_exact = [&](int mode) {
          a.mode = mode;
#pragma unroll 1
          for (int r = 0; r < kGroupRegs; ++r) {
            const int total = compact_one(a.t_lo, r);
            if (total > 0) exact_batch<T>(a, lane < total ? static_cast<int>(wl[lane]) : -1, lane);
            __syncwarp();
          }
        };
        uint32_t need = static_cast<uint32_t>(p.K);
        bool all = false;
#pragma unroll 1
        for (int pass = 0; pass < 8; ++pass) {
          a.shift = 56 - 8 * pass;
          if (tid < 256) ctl.hist[tid] = 0u;
          scan_bar_sync();
          scan_exact(kBmHist);
          scan_bar_sync();
          if (warp == 0) {
            // lane l owns digits 255-8l .. 248-8l (descending)
            uint32_t c[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = ctl.hist[255 - 8 * lane - q]; s += c[q]; }
            uint32_t incl = s;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
              const uint32_t vv = __shfl_up_sync(0xffffffffu, incl, dd);
              if (lane >= dd) incl += vv;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t excl = incl - s;
            if (excl < need && incl >= need) {
              uint32_t run = excl;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (run < need && run + c[q] >= need) { ctl.sel[0] = 255 - 8 * lane - q; ctl.sel[1] = need - run; }
                run += c[q];
              }
            }
            if (lane == 0) ctl.sel[2] = total;
          }
          scan_bar_sync();
          if (ctl.sel[2] < need) { all = true; break; }               // fewer than K candidates at all (first pass only)
          a.prefix |= static_cast<unsigned long long>(ctl.sel[0]) << a.shift;
          a.mask |= 0xFFull << a.shift;
          need = ctl.sel[1];
        }
        if (tid == 0) *n_list = 0u;
        scan_bar_sync();
        a.kth = all ? 0ull : a.prefix;                                // keys are distinct: exactly K of them are >= the K-th
        scan_exact(kBmCollect);
        scan_bar_sync();
      }

      // ---- the strip is done: slots back to the producer, the list's length (and key kind) to global memory
      const long long c5 = SCAN_CLK();
      if (tid == 0) {
        int s = d.slot0;
        for (int c = 0; c < d.n_chunks; ++c) {
          pl::mbar_arrive(pl::smem_u32(&ctl.empty[s]));
          if (++s == S) s = 0;
        }
        const uint32_t n_out = min(*reinterpret_cast<volatile uint32_t*>(n_list), static_cast<uint32_t>(cap));
        p.cand_count[d.strip] = n_out | (overflow ? kCandScoreKeys : 0u);
        st_keys += n_out;
        st_strips += 1;
      }
      ++seq;
      const long long c6 = SCAN_CLK();
      ph[0] += c1 - c0; ph[1] += c2 - c1; ph[2] += c3 - c2; ph[3] += c4 - c3; ph[4] += c5 - c4; ph[5] += c6 - c5;
    }
#ifdef RTM3D_DEV
    if (p.stats && lane == 0 && warp == 3) {
      for (int i = 0; i < 6; ++i) atomicAdd(&p.stats[kSsWaitFirst + i], static_cast<unsigned long long>(ph[i]));
    }
    if (p.stats && tid == 0) {
      atomicAdd(&p.stats[kSsStrips], static_cast<unsigned long long>(st_strips));
      atomicAdd(&p.stats[kSsListKeys], static_cast<unsigned long long>(st_keys));
    }
    if (p.stats && tid == 32) {
      atomicAdd(&p.stats[kSsDeepen], static_cast<unsigned long long>(st_deepen));
      atomicAdd(&p.stats[kSsExact], static_cast<unsigned long long>(st_exact));
    }
#endif
    (void)ph; (void)st_deepen; (void)st_exact; (void)st_keys; (void)st_strips;
  }
#ifdef RTM3D_DEV
  __syncthreads();
  if (p.stats && tid == 0) p.stats[64 + 2 * blockIdx.x + 1] = pl::globaltimer_ns();
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// Host side: geometry, launch.

static int scan_sm_count() {
  // per device: a process may drive several GPUs
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

int scan_list_cap(int K) {
  int cap = 4 * K;
  if (cap < 1024) cap = 1024;
  return (cap + 31) & ~31;
}

bool make_scan_geom(int B, int C, int Cv, int H, int W, int K, int dtype, int strips_override, const void* hm_main, const void* hm_kpt,
                    ScanGeom& g) {
  const int es = dtype == 0 ? 4 : 2;
  const int E = 16 / es;
  if (W % E != 0) return false;
  if ((hm_main && reinterpret_cast<uintptr_t>(hm_main) % 16 != 0) || (hm_kpt && reinterpret_cast<uintptr_t>(hm_kpt) % 16 != 0)) return false;
  g.row_bytes = W * es;
  g.gpr = g.row_bytes / 16;
  if (g.gpr > kChunkGroups) return false;                  // one row must fit a chunk
  if (g.gpr < 2) return false;                             // a one-group row: ceil(2^32 / gpr) does not fit the 32-bit magic (generic kernels)
  g.slot_rows = kChunkGroups / g.gpr;
  if (g.slot_rows > H) g.slot_rows = H;
  g.slot_bytes = g.slot_rows * g.row_bytes;
  g.gpr_magic = static_cast<unsigned>((0x100000000ULL + g.gpr - 1) / g.gpr);
  g.slot_bytes_magic = static_cast<unsigned>((0x100000000ULL + g.slot_bytes - 1) / g.slot_bytes);
  g.list_cap = scan_list_cap(K);
  const size_t budget = 232448 - sizeof(ScanCtl) - 256;    // 227 KB per block, static control block included
  if (3ull * g.slot_bytes > budget) return false;
  int S = static_cast<int>(budget / g.slot_bytes);
  if (S > kMaxSlots) S = kMaxSlots;
  // a resident strip may take all but three slots of the ring (the others keep the copy engine busy meanwhile)
  int max_chunks = S - 3;
  if (max_chunks > kMaxChunks) max_chunks = kMaxChunks;
  if (max_chunks < 1) return false;
  const int max_loaded = max_chunks * g.slot_rows;
  int Sp = 1;
  if (H > max_loaded) {
    if (max_loaded < 3) return false;
    Sp = (H + (max_loaded - 2) - 1) / (max_loaded - 2);
  }
  if (strips_override > 0) {
    if (strips_override > H) return false;
    const int rows_hi = (H + strips_override - 1) / strips_override;
    if (rows_hi + (strips_override > 1 ? 2 : 0) > max_loaded) return false;
    Sp = strips_override;
  }
  // even strips: the longest owns ceil(H / Sp) rows
  while (Sp > 1 && (H + Sp - 1) / Sp + 2 > max_loaded) ++Sp;
  g.S = S;
  g.Sp = Sp;
  const int rows_hi = (H + Sp - 1) / Sp;
  const double G = static_cast<double>(rows_hi) * g.gpr;                 // groups of a strip
  if (G > static_cast<double>(kGroupRegs) * kScanConsumers) return false;   // (cannot happen: a strip is at most kMaxChunks chunks)
  // First attempt: every warp takes the j-th largest of its 32 lane maxima over the first kSelRegs register rows (each lane:
  // an interleaved sample of the strip), T = the smallest of the 15 warps' values.  Simulated over i.i.d. maps this leaves
  // ~30 j / f groups with a pixel >= T (f = sampled fraction of the strip) whatever the strip size; j = K f / 15 aims at ~2 K
  // groups (K = 100, f = 3/4: j = 5, mean 202, 0.01 % .. 99.99 % = 96 .. 448, fewer than K in 3 strips of 10 000 -- those are
  // deepened, nothing is lost).  Unusual shapes (tiny strips, K > 360) use the bisection.
  g.target0 = static_cast<int>((2.3 * K + 16.0) / kScanWarps) + 1;
  const int n_regs = static_cast<int>((G + kScanConsumers - 1) / kScanConsumers);
  g.rank_j = 0;
  {
    const double f = static_cast<double>(n_regs < kSelRegs ? n_regs : kSelRegs) / n_regs;
    const int j = static_cast<int>(K * f / 15.0 + 0.5);
    if (j >= 2 && j <= 24 && G >= 2.0 * kScanConsumers) g.rank_j = j;
  }
  g.n_strips = B * (C + Cv) * Sp;
  g.smem = static_cast<unsigned>(static_cast<size_t>(S) * g.slot_bytes);
  g.debug = 0;
  return true;
}

bool scan_eligible(int B, int C, int Cv, int H, int W, int K, int dtype, const void* hm_main, const void* hm_kpt) {
  ScanGeom g{};
  return make_scan_geom(B, C, Cv, H, W, K, dtype, 0, hm_main, hm_kpt, g);
}

int scan_strips_per_plane(int H, int W, int K, int dtype, int strips_override) {
  ScanGeom g{};
  if (!make_scan_geom(1, 1, 0, H, W, K, dtype, strips_override, nullptr, nullptr, g)) return 0;
  return g.Sp;
}

int scan_max_strips_per_plane(int H, int W, int K) {
  // the strips per plane the kernel chooses by itself, over both element types (for the workspace size)
  int m = 1;
  for (int dt = 0; dt < 2; ++dt) {
    const int s = scan_strips_per_plane(H, W, K, dt, 0);
    if (s > m) m = s;
  }
  return m;
}

template <typename T>
static int launch_scan_t(const ScanParams& p, const ScanGeom& g, cudaStream_t s) {
  auto kern = scan_planes_kernel<T>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - static_cast<int>(sizeof(ScanCtl)) - 256);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  kern<<<g.grid, kScanThreads, g.smem, s>>>(p, g);
  return static_cast<int>(cudaGetLastError());
}

// Returns -1000 when the shape is not eligible (caller falls back to the generic path).
int launch_scan(const ScanParams& p, int dtype, int strips_override, int max_ctas, int debug, cudaStream_t s, int* strips_per_plane,
                int* list_cap) {
  ScanGeom g{};
  if (!make_scan_geom(p.B, p.C, p.Cv, p.H, p.W, p.K, dtype, strips_override, p.hm_main, p.hm_kpt, g)) return -1000;
  const int sms = scan_sm_count();
  if (sms <= 0) return -1000;
  g.grid = g.n_strips < sms ? g.n_strips : sms;
  if (max_ctas > 0 && g.grid > max_ctas) g.grid = max_ctas;
  g.debug = debug;
  if (strips_per_plane) *strips_per_plane = g.Sp;
  if (list_cap) *list_cap = g.list_cap;
  return dtype == 0 ? launch_scan_t<float>(p, g, s) : launch_scan_t<__nv_bfloat16>(p, g, s);
}

}  // namespace rtm3d
