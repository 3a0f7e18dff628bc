// Post-selection kernels: keypoint grouping (Tier B, models/model.py:134-162) and closed-form 3D recovery (Tier C).
#include "common.cuh"
#include "params.h"
#include "../../include/rtm3d_decode.h"
#include "postproc.h"
#include "tier_math.cuh"
#include "select_common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace rtm3d {

// ---------------------------------------------------------------------------------------------------------------
// Tier A epilogue (models/model.py:47-50 gather + sub-pixel add, :63-73 regress / scale / 2D box) for the rows the
// plane-streaming kernel selected: one thread per (detection, vertex), wide over the whole batch so that the scattered
// 4-byte gathers of the regression planes are latency-hidden by occupancy instead of stalling the streaming kernel.
// Rows >= counts[b] are zero-filled (cls = -1).
template <typename T>
__global__ void __launch_bounds__(256) epilogue_main_kernel(const EpiMainParams p, int vp, int vp_shift) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = idx >> vp_shift, v = idx & (vp - 1);
  const int V = p.n_vert, K = p.K, HW = p.H * p.W;
  const bool row_ok = j < K;
  const size_t row = static_cast<size_t>(b) * K + (row_ok ? j : 0);
  const bool valid = row_ok && j < p.counts[b];
  const bool vert = v < V;
  float vx = 0.f, vy = 0.f, mx = 0.f, my = 0.f;
  int c = -1;
  if (valid) {
    const int flat = p.flat[row];
    c = flat / HW;
    const int rem = flat - c * HW;
    const int yi = rem / p.W;
    const int xi = rem - yi * p.W;
    const T* off2 = reinterpret_cast<const T*>(p.off2) + static_cast<size_t>(b) * 2 * HW;
    const T* off = reinterpret_cast<const T*>(p.off) + static_cast<size_t>(b) * 2 * V * HW;
    const float r0 = to_f32(off2[rem]);
    const float r1 = to_f32(off2[HW + rem]);
    float ox = 0.f, oy = 0.f;
    if (vert) {
      ox = to_f32(off[static_cast<size_t>(2 * v) * HW + rem]);
      oy = to_f32(off[static_cast<size_t>(2 * v + 1) * HW + rem]);
    }
    mx = subpixel(xi, r0);
    my = subpixel(yi, r1);
    vx = regress_coord(p.down, ox, mx);
    vy = regress_coord(p.down, oy, my);
  }
  float lo_x = (valid && vert) ? vx : INFINITY, hi_x = (valid && vert) ? vx : -INFINITY;
  float lo_y = (valid && vert) ? vy : INFINITY, hi_y = (valid && vert) ? vy : -INFINITY;
  for (int d = 1; d < vp; d <<= 1) {
    lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, d));
    hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, d));
    lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, d));
    hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, d));
  }
  if (!row_ok) return;
  if (vert) {
    float* vout = p.verts + (row * V + v) * 2;
    vout[0] = valid ? vx : 0.f;
    vout[1] = valid ? vy : 0.f;
  }
  if (v == 0) {
    p.cls[row] = c;
    p.proj[row * 2 + 0] = valid ? scale_coord(p.down, mx) : 0.f;
    p.proj[row * 2 + 1] = valid ? scale_coord(p.down, my) : 0.f;
    p.bbox[row * 4 + 0] = valid ? lo_x : 0.f;
    p.bbox[row * 4 + 1] = valid ? lo_y : 0.f;
    p.bbox[row * 4 + 2] = valid ? hi_x : 0.f;
    p.bbox[row * 4 + 3] = valid ? hi_y : 0.f;
  }
}

int launch_epilogue_main(const EpiMainParams& p, int dtype, cudaStream_t s) {
  int vp = 1, sh = 0;
  while (vp < p.n_vert) { vp <<= 1; ++sh; }
  const int threads = p.K * vp;
  dim3 grid((threads + 255) / 256, p.B);
  if (dtype == 0) epilogue_main_kernel<float><<<grid, 256, 0, s>>>(p, vp, sh);
  else epilogue_main_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p, vp, sh);
  return static_cast<int>(cudaGetLastError());
}

// Tier B epilogue: index split + sub-pixel add (models/model.py:113-114 and the commented :55-57), one thread per candidate.
template <typename T>
__global__ void __launch_bounds__(256) epilogue_kpt_kernel(const EpiKptParams p) {
  const size_t n = static_cast<size_t>(p.B) * p.Cv * p.K;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (row >= n) return;
  const int HW = p.H * p.W;
  const int b = static_cast<int>(row / (static_cast<size_t>(p.Cv) * p.K));
  const int flat = p.kflat[row];
  const int yi = flat / p.W;
  const int xi = flat - yi * p.W;
  const T* off2 = reinterpret_cast<const T*>(p.voff2) + static_cast<size_t>(b) * 2 * HW;
  const float r0 = to_f32(off2[flat]);
  const float r1 = to_f32(off2[HW + flat]);
  p.kxy[row * 2 + 0] = subpixel(xi, r0);
  p.kxy[row * 2 + 1] = subpixel(yi, r1);
}

int launch_epilogue_kpt(const EpiKptParams& p, int dtype, cudaStream_t s) {
  const size_t n = static_cast<size_t>(p.B) * p.Cv * p.K;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  if (dtype == 0) epilogue_kpt_kernel<float><<<grid, 256, 0, s>>>(p);
  else epilogue_kpt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Fused post kernel of rtm3d_decode_fused: for one image, (1) Tier B epilogue of its Cv*K candidates (kept in shared
// memory for the grouping), (2) per (detection n, channel k): the Tier A gathers / regress and the nearest-candidate
// search of _group_vertexs_kf, (3) per detection: 2D box, class, centre.  Arithmetic and association order are those of
// epilogue_main_kernel, epilogue_kpt_kernel and group_vertices_kernel (bit-identical results).
constexpr int kPostThreads = 256;
// A cluster of four CTAs per image, each with a quarter of the detections: small CTAs spread evenly over the SMs and a
// CTA's (detection, channel) pairs fit one round of its threads.  Every CTA needs all of the image's candidates: each
// computes a quarter and stores it into the shared memory of all four (distributed shared memory).
constexpr int kPostSplit = 4;
template <typename T>
__global__ void __cluster_dims__(kPostSplit, 1, 1) __launch_bounds__(kPostThreads) post_fused_kernel(const PostFusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / kPostSplit, half = static_cast<int>(cluster.block_rank()), tid = threadIdx.x;
  const int K = p.K, Cv = p.Cv, V = p.n_vert, HW = p.H * p.W;
  const int KP = K + 1;                                              // padded row: channels start in different banks
  const int per = (K + kPostSplit - 1) / kPostSplit;                 // detections per CTA
  const int n0 = half * per, n1 = min(K, n0 + per);
  float* s_xy = reinterpret_cast<float*>(smem_raw);                 // [Cv][KP][2] candidate positions
  float* s_v = s_xy + static_cast<size_t>(Cv) * KP * 2;              // [per*V*2]  scaled regressed vertices (Tier A)
  float* s_m = s_v + static_cast<size_t>(per) * V * 2;               // [per*2]    unscaled centre (mx, my)
  const int n_det = p.counts[b];
  const T* voff2 = reinterpret_cast<const T*>(p.voff2) + static_cast<size_t>(b) * 2 * HW;
  const T* off2 = reinterpret_cast<const T*>(p.off2) + static_cast<size_t>(b) * 2 * HW;
  const T* off = reinterpret_cast<const T*>(p.off) + static_cast<size_t>(b) * 2 * V * HW;
  // ---- (1) candidates: index split + sub-pixel add (models/model.py:113-114, :55-57)
  const int cand_per = (Cv * K + kPostSplit - 1) / kPostSplit;
  float2* peer_xy[kPostSplit];
#pragma unroll
  for (int r = 0; r < kPostSplit; ++r) peer_xy[r] = reinterpret_cast<float2*>(cluster.map_shared_rank(s_xy, r));
  cluster.sync();          // every CTA of the cluster has started: its shared memory may be written by its peers from here on
  for (int i = half * cand_per + tid; i < min(Cv * K, (half + 1) * cand_per); i += kPostThreads) {
    const size_t row = static_cast<size_t>(b) * Cv * K + i;
    const int flat = p.kflat[row];
    const int yi = flat / p.W, xi = flat - yi * p.W;
    const float x = subpixel(xi, to_f32(voff2[flat]));
    const float y = subpixel(yi, to_f32(voff2[HW + flat]));
    const int k = i / K, j = i - k * K;
#pragma unroll
    for (int r = 0; r < kPostSplit; ++r) peer_xy[r][k * KP + j] = make_float2(x, y);
    p.kxy[row * 2] = x; p.kxy[row * 2 + 1] = y;
  }
  // ---- centres of the detections (models/model.py:48-50)
  for (int n = n0 + tid; n < n1; n += kPostThreads) {
    float mx = 0.f, my = 0.f;
    if (n < n_det) {
      const int rem = p.flat[static_cast<size_t>(b) * K + n] % HW;
      const int yi = rem / p.W, xi = rem - yi * p.W;
      mx = subpixel(xi, to_f32(off2[rem]));
      my = subpixel(yi, to_f32(off2[HW + rem]));
    }
    s_m[2 * (n - n0)] = mx; s_m[2 * (n - n0) + 1] = my;
  }
  cluster.sync();          // every CTA of the image has all candidates (and nobody writes into a peer after this)
  // ---- (2) per (detection, channel): vertex regress (models/model.py:63-69) + nearest candidate (:144-161)
  const int KC = (Cv > V ? Cv : V);                                   // channels that need work per detection
  for (int w = tid; w < (n1 - n0) * KC; w += kPostThreads) {
    const int nl = w / KC, k = w - nl * KC, n = n0 + nl;
    const bool valid = n < n_det;
    float ox = 0.f, oy = 0.f, mx = 0.f, my = 0.f;
    if (valid) {
      const int rem = p.flat[static_cast<size_t>(b) * K + n] % HW;
      mx = s_m[2 * nl]; my = s_m[2 * nl + 1];
      if (k < V) {
        ox = to_f32(off[static_cast<size_t>(2 * k) * HW + rem]);
        oy = to_f32(off[static_cast<size_t>(2 * k + 1) * HW + rem]);
      }
    }
    const float vx = valid ? regress_coord(p.down, ox, mx) : 0.f;
    const float vy = valid ? regress_coord(p.down, oy, my) : 0.f;
    if (k < V) {
      s_v[(nl * V + k) * 2] = vx; s_v[(nl * V + k) * 2 + 1] = vy;
      float* vout = p.verts + ((static_cast<size_t>(b) * K + n) * V + k) * 2;
      vout[0] = vx; vout[1] = vy;
    }
    if (k < Cv) {
      const size_t row = (static_cast<size_t>(b) * K + n) * Cv + k;
      if (!valid) {
        p.kpt_proj[row * 2] = 0.f; p.kpt_proj[row * 2 + 1] = 0.f;
        p.kpt_score[row] = 0.f;
        p.kpt_j[row] = -1;
        if (p.verts_cv) { p.verts_cv[row * 2] = 0.f; p.verts_cv[row * 2 + 1] = 0.f; }
      } else {
        const float* cand = s_xy + static_cast<size_t>(k) * KP * 2;
        const int bj = nearest_candidate(cand, K, mx, my, ox, oy);
        p.kpt_proj[row * 2] = scale_coord(p.down, cand[2 * bj]);
        p.kpt_proj[row * 2 + 1] = scale_coord(p.down, cand[2 * bj + 1]);
        p.kpt_score[row] = p.kscore[(static_cast<size_t>(b) * Cv + k) * K + bj];
        p.kpt_j[row] = bj;
        if (p.verts_cv) { p.verts_cv[row * 2] = vx; p.verts_cv[row * 2 + 1] = vy; }
      }
    }
  }
  __syncthreads();
  // ---- (3) per detection: class, centre, 2D box over the regressed vertices (models/model.py:70-73)
  for (int n = n0 + tid; n < n1; n += kPostThreads) {
    const size_t row = static_cast<size_t>(b) * K + n;
    const int nl = n - n0;
    const bool valid = n < n_det;
    float lo_x = 0.f, lo_y = 0.f, hi_x = 0.f, hi_y = 0.f;
    int c = -1;
    if (valid) {
      c = p.flat[row] / HW;
      lo_x = lo_y = INFINITY; hi_x = hi_y = -INFINITY;
      for (int v = 0; v < V; ++v) {
        const float vx = s_v[(nl * V + v) * 2], vy = s_v[(nl * V + v) * 2 + 1];
        lo_x = fminf(lo_x, vx); hi_x = fmaxf(hi_x, vx);
        lo_y = fminf(lo_y, vy); hi_y = fmaxf(hi_y, vy);
      }
    }
    p.cls[row] = c;
    p.proj[row * 2] = valid ? scale_coord(p.down, s_m[2 * nl]) : 0.f;
    p.proj[row * 2 + 1] = valid ? scale_coord(p.down, s_m[2 * nl + 1]) : 0.f;
    p.bbox[row * 4 + 0] = lo_x; p.bbox[row * 4 + 1] = lo_y; p.bbox[row * 4 + 2] = hi_x; p.bbox[row * 4 + 3] = hi_y;
  }
}

// Scattered 4-byte gathers from the regression planes: every scalar is its own 32-byte sector of an NCHW plane.  Without a
// hint the L2 fetches the whole 128-byte line from DRAM for it (measured: 122 B per gather); .L2::64B halves that.  The value is
// used once: L1::no_allocate keeps the L1 lines for the loads that are reused (measured: -1.6 us per cfg4 launch).
template <typename T> __device__ __forceinline__ float gather_ld(const T* p);
template <> __device__ __forceinline__ float gather_ld<float>(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <> __device__ __forceinline__ float gather_ld<__nv_bfloat16>(const __nv_bfloat16* p) {
  unsigned short u;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u16 %0, [%1];" : "=h"(u) : "l"(p));
  return __uint_as_float(static_cast<uint32_t>(u) << 16);
}

// Words [s0, s0 + n_words) of this rank's gather buffer to the same place in every other rank's buffer.  The buffers are
// symmetric (same offsets, same base alignment): source and destination share their low address bits, so the 16-byte body is
// aligned at both ends.  All NT threads of the block; stores are posted (nobody waits for them here).
template <int NT>
__device__ __forceinline__ void push_words(int32_t* const* peers, int n_peers, int self, size_t s0, int n_words) {
  const int32_t* src = peers[self] + s0;
  const int to16 = static_cast<int>(((16u - (reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) >> 2);
  const int head = to16 < n_words ? to16 : n_words;
  const int body4 = (n_words - head) >> 2;
  const int tail0 = head + body4 * 4;
  const int tid = threadIdx.x;
  if (tid < head) {
    const int32_t v = src[tid];
    for (int r = 0; r < n_peers; ++r) if (r != self) peers[r][s0 + tid] = v;
  }
  for (int q4 = tid; q4 < body4; q4 += NT) {
    const int i = head + q4 * 4;
    const int4 v = *reinterpret_cast<const int4*>(src + i);
    for (int r = 0; r < n_peers; ++r) if (r != self) *reinterpret_cast<int4*>(peers[r] + s0 + i) = v;
  }
  if (tid < n_words - tail0) {
    const int32_t v = src[tail0 + tid];
    for (int r = 0; r < n_peers; ++r) if (r != self) peers[r][s0 + tail0 + tid] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused select + post kernel of rtm3d_decode_fused behind the scan kernel.  A cluster of four CTAs per image:
//   (S) the image's selection problems are sorted by different CTAs at the same time -- CTA 0: the main problem (its C*Sp
//       lists), then the centres of the detections (models/model.py:48-50); CTAs 1..3: the keypoint planes (three each for
//       Cv = 9), each followed by the sub-pixel positions of its K candidates (models/model.py:113-114, :55-57).  Everything
//       the other CTAs need (candidate positions, detections) is stored into the shared memory of all four.
//   (2) per (detection, channel): vertex regress (models/model.py:63-69) + nearest candidate (:144-161),
//   (3) per detection: class, centre, 2D box (:70-73) -- as in post_fused_kernel, same arithmetic and association order.
struct __align__(8) PostDet { int flat; float mx, my; };
#ifdef RTM3D_DEV
#define SP_MARK(i) do { if (sp.stats && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); sp.stats[1024 + (static_cast<size_t>(blockIdx.x) * 8 + (i))] = t_; } } while (0)
#else
#define SP_MARK(i) do { } while (0)
#endif
constexpr int kSpThreads = 128;              // 4 warps: 8 CTAs per SM, the 4 B CTAs of a 256-image batch are resident at once
constexpr int kSpWarps = kSpThreads / 32;

template <typename T>
__global__ void __cluster_dims__(kPostSplit, 1, 1) __launch_bounds__(kSpThreads, 8) select_post_kernel(const SelectPostParams sp, int ns, int sort_bytes, int list_slots) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_n;
  __shared__ int s_count;
  __shared__ int s_cnt[kFastLists];
  cg::cluster_group cluster = cg::this_cluster();
  const PostFusedParams& p = sp.post;
  const int b = blockIdx.x / kPostSplit, rank = static_cast<int>(cluster.block_rank()), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K, Cv = p.Cv, V = p.n_vert, HW = p.H * p.W;
  const int KP = K + 1;                                              // padded row: channels start in different banks
  const int per = (K + kPostSplit - 1) / kPostSplit;                 // detections per CTA
  const int n0 = rank * per, n1 = min(K, n0 + per);
  // shared memory: sort area (fast path: [list_slots][kFastPad] sorted lists + [kFastProblems][kFastPad] merged +
  // [kFastProblems][kFastPad] filler indices; general path: [ns] sort buffer) | candidate positions | vertices | detections
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_top = s_keys + list_slots * kFastPad;
  uint32_t* s_fill = reinterpret_cast<uint32_t*>(s_top + kFastProblems * kFastPad);
  float* s_xy = reinterpret_cast<float*>(smem_raw + sort_bytes);      // [Cv][KP][2] candidate positions
  float* s_v = s_xy + static_cast<size_t>(Cv) * KP * 2;              // [per*V*2]  scaled regressed vertices (Tier A)
  PostDet* s_det = reinterpret_cast<PostDet*>(s_v + static_cast<size_t>(per) * V * 2);   // [K] detections: flat index, unscaled centre
  const T* voff2 = reinterpret_cast<const T*>(p.voff2) + static_cast<size_t>(b) * 2 * HW;
  const T* off2 = reinterpret_cast<const T*>(p.off2) + static_cast<size_t>(b) * 2 * HW;
  const T* off = reinterpret_cast<const T*>(p.off) + static_cast<size_t>(b) * 2 * V * HW;
  float2* peer_xy[kPostSplit];
  PostDet* peer_det[kPostSplit];
  int* peer_count[kPostSplit];
#pragma unroll
  for (int r = 0; r < kPostSplit; ++r) {
    peer_xy[r] = reinterpret_cast<float2*>(cluster.map_shared_rank(s_xy, r));
    peer_det[r] = cluster.map_shared_rank(s_det, r);
    peer_count[r] = cluster.map_shared_rank(&s_count, r);
  }
  SP_MARK(0);
  if (sp.push_step != 0u) {
    // deferred gather: the previous batch's rows of this CTA's block (and, by the image's first CTA, its count word) go to the
    // other ranks now; they drain over NVLink while this kernel sorts
    const int words_ = 9 + 2 * V;
    const size_t per_img_ = static_cast<size_t>(K) * words_ + 1;
    const size_t img0_ = (static_cast<size_t>(sp.wire_rank) * p.B + b) * per_img_;
    push_words<kSpThreads>(sp.push_peers, sp.n_peers, sp.wire_rank, img0_ + static_cast<size_t>(n0) * words_, max(0, n1 - n0) * words_);
    if (rank == 0 && tid == 0) {
      const int32_t c = sp.push_peers[sp.wire_rank][img0_ + static_cast<size_t>(K) * words_];
      for (int r = 0; r < sp.n_peers; ++r) if (r != sp.wire_rank) sp.push_peers[r][img0_ + static_cast<size_t>(K) * words_] = c;
    }
  }
  cluster.sync();          // every CTA of the cluster has started: its shared memory may be written by its peers from here on
  SP_MARK(1);

  // ---- (S) the selection problems of this CTA: rank 0 the main problem (C*Sp lists), ranks 1..3 the keypoint planes
  //      kc = rank-1, rank+2, ... (Sp lists each)
  const int n_kpt_ctas = kPostSplit - 1;
  const int n_prob = rank == 0 ? 1 : (Cv - (rank - 1) + n_kpt_ctas - 1) / n_kpt_ctas;     // (<= 0: nothing to do)
  const int lists_per_prob = rank == 0 ? p.C * sp.Sp : sp.Sp;
  const int n_lists = n_prob > 0 ? n_prob * lists_per_prob : 0;
  auto first_list = [&](int pr) {                                    // first strip of this CTA's pr-th problem
    return rank == 0 ? b * p.C * sp.Sp : (p.B * p.C + b * Cv + (rank - 1) + pr * n_kpt_ctas) * sp.Sp;
  };
  // fast path: every list of the CTA fits the register sort
  bool fast = K <= kFastPad && n_lists <= list_slots && n_lists <= kFastLists && n_prob <= kFastProblems;
  {
    // (lists far beyond the usual ~2 K keys -- ties, plateaus -- go the general way: the register sort takes them in rounds)
    int too_long = 0;
    if (fast && tid < n_lists) {
      const int pr = tid / lists_per_prob;
      too_long = (sp.cand_count[first_list(pr) + (tid - pr * lists_per_prob)] & ~kCandScoreKeys) > static_cast<uint32_t>(4 * kFastKeys);
    }
    fast = fast && !__syncthreads_or(too_long);
  }
  if (fast) {
    // one warp per list: (logit, index) -> (score, index) keys (the sigmoid of models/model.py:85,107 for the listed pixels;
    // pixels at or below the score floor drop out), register sort, the kFastPad best to shared memory
    for (int t = warp; t < n_lists; t += kSpWarps) {
      const int pr = t / lists_per_prob;
      const int cnt = warp_sort_list(sp.cand, sp.cand_count, sp.list_cap, first_list(pr) + (t - pr * lists_per_prob),
                                     rank == 0 ? sp.thresh : 0.0f, K, s_keys + t * kFastPad, lane);
      if (lane == 0) s_cnt[t] = cnt;
    }
    __syncthreads();
    SP_MARK(2);
    // problems with several lists: rank-merge (a key's rank = its position in its own list + the larger keys of the others)
    if (lists_per_prob > 1) {
      for (int pr = 0; pr < n_prob; ++pr)
        block_merge_lists<kSpThreads>(s_keys + pr * lists_per_prob * kFastPad, s_cnt + pr * lists_per_prob, lists_per_prob, K, s_top + pr * kFastPad);
      __syncthreads();
    }
  }
  SP_MARK(3);
  // Tier A rows of the image from the main problem's K best keys: score, index, count; the detections (index + centre,
  // models/model.py:48-50) go to all four CTAs
  auto emit_main = [&](const uint64_t* top, int have) {
    for (int j = tid; j < K; j += kSpThreads) {
      const size_t row = static_cast<size_t>(b) * K + j;
      const bool valid = j < have;                                   // every listed key has score > thresh (models/model.py:91)
      const int fl = valid ? static_cast<int>(key_flat(top[j])) : -1;
      sp.score[row] = valid ? key_score(top[j]) : 0.f;
      sp.flat[row] = fl;
      PostDet dt{fl, 0.f, 0.f};
      if (valid) {
        const int rem = fl % HW;
        const int yi = rem / p.W, xi = rem - yi * p.W;
        dt.mx = subpixel(xi, gather_ld<T>(off2 + rem));
        dt.my = subpixel(yi, gather_ld<T>(off2 + HW + rem));
      }
#pragma unroll
      for (int r = 0; r < kPostSplit; ++r) peer_det[r][j] = dt;
    }
    if (tid == 0) {
      sp.counts[b] = have;
#pragma unroll
      for (int r = 0; r < kPostSplit; ++r) *peer_count[r] = have;
    }
  };
  // rows have..K-1 of a keypoint plane: 0.0-score fillers = the lowest flat indices that are not among the plane's
  // positive-score peaks (what a top-K over the zero-filled peak map returns); with fewer than K valid keys the
  // lists held ALL peaks.  One thread walks the indices in order (rare: fewer than K positive peaks in a whole plane).
  auto find_fillers = [&](const uint64_t* top, int have, uint32_t* fill) {
    int r = have;
    for (int i = 0; i < HW && r < K; ++i) {
      bool taken = false;
      for (int q = 0; q < have; ++q) taken |= key_flat(top[q]) == static_cast<uint32_t>(i);
      if (!taken) fill[r++] = static_cast<uint32_t>(i);
    }
  };
  // one Tier B candidate row: score, index, sub-pixel position (models/model.py:113-114, :55-57), to all four CTAs
  auto emit_kpt_row = [&](int kc, int j, const uint64_t* top, int have, const uint32_t* fill) {
    const size_t row = (static_cast<size_t>(b) * Cv + kc) * K + j;
    const int fl = static_cast<int>(j < have ? key_flat(top[j]) : fill[j]);
    sp.kscore[row] = j < have ? key_score(top[j]) : 0.0f;
    sp.kflat[row] = fl;
    const int yi = fl / p.W, xi = fl - yi * p.W;
    const float x = subpixel(xi, gather_ld<T>(voff2 + fl));
    const float y = subpixel(yi, gather_ld<T>(voff2 + HW + fl));
#pragma unroll
    for (int r = 0; r < kPostSplit; ++r) peer_xy[r][kc * KP + j] = make_float2(x, y);
    p.kxy[row * 2] = x; p.kxy[row * 2 + 1] = y;
  };
  if (fast) {
    auto prob_top = [&](int pr) { return lists_per_prob > 1 ? s_top + pr * kFastPad : s_keys + pr * kFastPad; };
    auto prob_have = [&](int pr) {
      int total = 0;
      for (int o = 0; o < lists_per_prob; ++o) total += s_cnt[pr * lists_per_prob + o];
      return min(K, total);
    };
    if (rank == 0) {
      emit_main(prob_top(0), prob_have(0));
    } else {
      if (tid < n_prob && prob_have(tid) < K) find_fillers(prob_top(tid), prob_have(tid), s_fill + tid * kFastPad);
      __syncthreads();
      // every row of every plane of this CTA in one pass; a thread issues the sub-pixel gathers of ALL its rows before it
      // finishes the first one: one DRAM round trip per thread instead of one per row
      constexpr int kEmitPre = 3;
      for (int base = 0; base < n_prob * K; base += kEmitPre * kSpThreads) {
        int fl[kEmitPre];
        float gx[kEmitPre], gy[kEmitPre], sc[kEmitPre];
#pragma unroll
        for (int u = 0; u < kEmitPre; ++u) {
          const int idx = base + u * kSpThreads + tid;
          fl[u] = -1;
          if (idx < n_prob * K) {
            const int pr = idx / K, j = idx - pr * K, have = prob_have(pr);
            const uint64_t* top = prob_top(pr);
            fl[u] = static_cast<int>(j < have ? key_flat(top[j]) : s_fill[pr * kFastPad + j]);
            sc[u] = j < have ? key_score(top[j]) : 0.0f;
            gx[u] = gather_ld<T>(voff2 + fl[u]);
            gy[u] = gather_ld<T>(voff2 + HW + fl[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < kEmitPre; ++u) {
          const int idx = base + u * kSpThreads + tid;
          if (fl[u] < 0) continue;
          const int pr = idx / K, j = idx - pr * K, kc = (rank - 1) + pr * n_kpt_ctas;
          const size_t row = (static_cast<size_t>(b) * Cv + kc) * K + j;
          sp.kscore[row] = sc[u];
          sp.kflat[row] = fl[u];
          const int yi = fl[u] / p.W, xi = fl[u] - yi * p.W;
          const float x = subpixel(xi, gx[u]);
          const float y = subpixel(yi, gy[u]);
#pragma unroll
          for (int r = 0; r < kPostSplit; ++r) peer_xy[r][kc * KP + j] = make_float2(x, y);
          p.kxy[row * 2] = x; p.kxy[row * 2 + 1] = y;
        }
      }
    }
  } else {
    for (int pr = 0; pr < n_prob; ++pr) {
      const int have = block_select_sorted<kSpThreads>(sp.cand, sp.cand_count, sp.list_cap, first_list(pr), lists_per_prob, K,
                                                       rank == 0 ? sp.thresh : 0.0f, s_keys, ns, &s_n);
      if (rank == 0) {
        emit_main(s_keys, have);
      } else {
        uint32_t* fill = reinterpret_cast<uint32_t*>(s_keys + next_pow2(K));      // behind the K best keys of the sort buffer
        if (have < K) {
          if (tid == 0) find_fillers(s_keys, have, fill);
          __syncthreads();
        }
        for (int j = tid; j < K; j += kSpThreads) emit_kpt_row((rank - 1) + pr * n_kpt_ctas, j, s_keys, have, fill);
      }
      __syncthreads();                                                // the sort buffer is reused by the next problem
    }
  }
  SP_MARK(4);
  cluster.sync();          // every CTA of the image has all candidates and detections (and nobody writes into a peer after this)
  SP_MARK(5);
  const int n_det = s_count;
  // ---- (2) per (detection, channel): vertex regress (models/model.py:63-69) + nearest candidate (:144-161)
  const int KC = (Cv > V ? Cv : V);                                   // channels that need work per detection
  const int n_items = (n1 - n0) * KC;
  // the vertex offsets of this thread's first rounds are gathered up front: one DRAM round trip for all of them
  constexpr int kPre = 2;
  float pox[kPre], poy[kPre];
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    const int w = tid + u * kSpThreads;
    pox[u] = 0.f; poy[u] = 0.f;
    if (w < n_items) {
      const int nl = w / KC, k = w - nl * KC, n = n0 + nl;
      if (n < n_det && k < V) {
        const int rem = s_det[n].flat % HW;
        pox[u] = gather_ld<T>(off + static_cast<size_t>(2 * k) * HW + rem);
        poy[u] = gather_ld<T>(off + static_cast<size_t>(2 * k + 1) * HW + rem);
      }
    }
  }
  for (int w = tid, u = 0; w < n_items; w += kSpThreads, ++u) {
    const int nl = w / KC, k = w - nl * KC, n = n0 + nl;
    const bool valid = n < n_det;
    float ox = 0.f, oy = 0.f, mx = 0.f, my = 0.f;
    if (valid) {
      const PostDet dt = s_det[n];
      const int rem = dt.flat % HW;
      mx = dt.mx; my = dt.my;
      if (k < V) {
        if (u < kPre) {
          ox = u == 0 ? pox[0] : pox[1]; oy = u == 0 ? poy[0] : poy[1];
        } else {
          ox = gather_ld<T>(off + static_cast<size_t>(2 * k) * HW + rem);
          oy = gather_ld<T>(off + static_cast<size_t>(2 * k + 1) * HW + rem);
        }
      }
    }
    const float vx = valid ? regress_coord(p.down, ox, mx) : 0.f;
    const float vy = valid ? regress_coord(p.down, oy, my) : 0.f;
    if (k < V) {
      s_v[(nl * V + k) * 2] = vx; s_v[(nl * V + k) * 2 + 1] = vy;
      float* vout = p.verts + ((static_cast<size_t>(b) * K + n) * V + k) * 2;
      vout[0] = vx; vout[1] = vy;
    }
    if (k < Cv) {
      const size_t row = (static_cast<size_t>(b) * K + n) * Cv + k;
      if (!valid) {
        p.kpt_proj[row * 2] = 0.f; p.kpt_proj[row * 2 + 1] = 0.f;
        p.kpt_score[row] = 0.f;
        p.kpt_j[row] = -1;
        if (p.verts_cv) { p.verts_cv[row * 2] = 0.f; p.verts_cv[row * 2 + 1] = 0.f; }
      } else {
        const float* cand = s_xy + static_cast<size_t>(k) * KP * 2;
        const int bj = nearest_candidate(cand, K, mx, my, ox, oy);
        p.kpt_proj[row * 2] = scale_coord(p.down, cand[2 * bj]);
        p.kpt_proj[row * 2 + 1] = scale_coord(p.down, cand[2 * bj + 1]);
        p.kpt_score[row] = sp.kscore[(static_cast<size_t>(b) * Cv + k) * K + bj];
        p.kpt_j[row] = bj;
        if (p.verts_cv) { p.verts_cv[row * 2] = vx; p.verts_cv[row * 2 + 1] = vy; }
      }
    }
  }
  __syncthreads();
  SP_MARK(6);
  // ---- (3) per detection: class, centre, 2D box over the regressed vertices (models/model.py:70-73)
  const int words = 9 + 2 * V;                                       // wire row: cls | score | proj 2 | verts 2V | bbox 4 | flat
  int32_t* s_wire = reinterpret_cast<int32_t*>(s_keys);             // [per][words]: the sort area is free by now
  for (int n = n0 + tid; n < n1; n += kSpThreads) {
    const size_t row = static_cast<size_t>(b) * K + n;
    const int nl = n - n0;
    const bool valid = n < n_det;
    float lo_x = 0.f, lo_y = 0.f, hi_x = 0.f, hi_y = 0.f;
    int c = -1;
    if (valid) {
      c = s_det[n].flat / HW;
      lo_x = lo_y = INFINITY; hi_x = hi_y = -INFINITY;
      for (int v = 0; v < V; ++v) {
        const float vx = s_v[(nl * V + v) * 2], vy = s_v[(nl * V + v) * 2 + 1];
        lo_x = fminf(lo_x, vx); hi_x = fmaxf(hi_x, vx);
        lo_y = fminf(lo_y, vy); hi_y = fmaxf(hi_y, vy);
      }
    }
    const float px = valid ? scale_coord(p.down, s_det[n].mx) : 0.f, py = valid ? scale_coord(p.down, s_det[n].my) : 0.f;
    p.cls[row] = c;
    p.proj[row * 2] = px;
    p.proj[row * 2 + 1] = py;
    p.bbox[row * 4 + 0] = lo_x; p.bbox[row * 4 + 1] = lo_y; p.bbox[row * 4 + 2] = hi_x; p.bbox[row * 4 + 3] = hi_y;
    if (sp.n_peers > 0) {
      int32_t* wr = s_wire + nl * words;
      wr[0] = c;
      wr[1] = __float_as_int(valid ? sp.score[row] : 0.f);
      wr[2] = __float_as_int(px); wr[3] = __float_as_int(py);
      for (int v = 0; v < 2 * V; ++v) wr[4 + v] = __float_as_int(s_v[nl * V * 2 + v]);
      wr[4 + 2 * V] = __float_as_int(lo_x); wr[5 + 2 * V] = __float_as_int(lo_y);
      wr[6 + 2 * V] = __float_as_int(hi_x); wr[7 + 2 * V] = __float_as_int(hi_y);
      wr[8 + 2 * V] = valid ? s_det[n].flat : -1;
    }
  }
  if (sp.n_peers > 0) {
    // ---- (4) the path's one exchange, fused: this CTA's wire rows go straight into every peer's gather
    //      buffer (posted stores over NVLink; the rows of image b of rank r start at word ((r*B + b) * (K*words + 1))
    __syncthreads();
    const size_t per_img = static_cast<size_t>(K) * words + 1;
    const size_t img0 = (static_cast<size_t>(sp.wire_rank) * p.B + b) * per_img;
    const int n_words = (n1 - n0) * words;
    // (16-byte stores for the aligned body of the CTA's block of rows, scalar head and tail: remote stores are packets)
    const size_t s0 = img0 + static_cast<size_t>(n0) * words;
    for (int r = 0; r < sp.n_peers; ++r) {
      if (sp.deferred && r != sp.wire_rank) continue;                // (the others get these rows at the start of the next launch)
      int32_t* dst = sp.wire_peers[r] + s0;
      const int to16 = static_cast<int>(((16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) >> 2);   // words up to the next 16-byte boundary
      const int head = to16 < n_words ? to16 : n_words;
      const int body4 = (n_words - head) >> 2;
      const int tail0 = head + body4 * 4;
      if (tid < head) dst[tid] = s_wire[tid];
      for (int q4 = tid; q4 < body4; q4 += kSpThreads) {
        const int i = head + q4 * 4;
        *reinterpret_cast<int4*>(dst + i) = make_int4(s_wire[i], s_wire[i + 1], s_wire[i + 2], s_wire[i + 3]);
      }
      if (tid < n_words - tail0) dst[tail0 + tid] = s_wire[tail0 + tid];
      if (rank == 0 && tid == 0) sp.wire_peers[r][img0 + static_cast<size_t>(K) * words] = n_det;
    }
    const uint32_t flag_step = sp.deferred ? sp.push_step : sp.step_id;      // (deferred: the batch pushed at the start)
    if (flag_step != 0u) {
      // arrival flag: every CTA makes its remote stores visible system-wide and counts itself; the last one of the launch
      // then tells every peer that ALL rows of this rank's batch have landed (release pattern: rows, fence, flag)
      __syncthreads();
      if (tid == 0) {
        __threadfence_system();
        const uint32_t prev = atomicAdd(sp.done_counter, 1u);
        if (prev == gridDim.x - 1u) {
          *sp.done_counter = 0u;                                     // clean for the next launch
          __threadfence_system();
          for (int r = 0; r < sp.n_peers; ++r) {
            volatile uint32_t* fl = reinterpret_cast<volatile uint32_t*>((sp.deferred ? sp.push_peers[r] : sp.wire_peers[r]) + sp.flag_offset);
            fl[sp.wire_rank] = flag_step;
          }
        }
      }
    }
  }
  SP_MARK(7);
}

static int select_post_ns(int K) {
  int ns = 4 * next_pow2(K);
  return ns < 1024 ? 1024 : ns;
}
static int select_post_list_slots(int C, int Cv, int Sp) {
  const int kpt = ((Cv + kPostSplit - 2) / (kPostSplit - 1)) * Sp, main = C * Sp;
  const int need = kpt > main ? kpt : main;
  return need > kFastLists ? kFastLists : need;
}
static size_t select_post_sort_bytes(int C, int Cv, int Sp, int K) {
  const size_t fast = static_cast<size_t>(select_post_list_slots(C, Cv, Sp) + kFastProblems) * kFastPad * 8 + static_cast<size_t>(kFastProblems) * kFastPad * 4;
  const size_t general = static_cast<size_t>(select_post_ns(K)) * 8;
  return fast > general ? fast : general;
}
static size_t select_post_smem_sp(int C, int Cv, int Sp, int K, int n_vert) {
  const size_t per = static_cast<size_t>((K + kPostSplit - 1) / kPostSplit);
  return select_post_sort_bytes(C, Cv, Sp, K) + (static_cast<size_t>(Cv) * (K + 1) * 2 + per * n_vert * 2) * sizeof(float) +
         static_cast<size_t>(K) * sizeof(PostDet) + 16;
}
size_t select_post_smem(int Cv, int K, int n_vert) { return select_post_smem_sp(kFastLists, Cv, 1, K, n_vert); }   // upper bound

int launch_select_post(const SelectPostParams& p, int dtype, cudaStream_t s) {
  const size_t smem = select_post_smem_sp(p.post.C, p.post.Cv, p.Sp, p.post.K, p.post.n_vert);
  const unsigned grid = static_cast<unsigned>(p.post.B) * kPostSplit;
  const int ns = select_post_ns(p.post.K);
  const int sort_bytes = static_cast<int>(select_post_sort_bytes(p.post.C, p.post.Cv, p.Sp, p.post.K));
  const int list_slots = select_post_list_slots(p.post.C, p.post.Cv, p.Sp);
  static bool attr_set[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int di = dtype == 0 ? 0 : 1;
  if (dev < 0 || dev >= 64 || !attr_set[dev][di]) {
    cudaError_t e = dtype == 0 ? cudaFuncSetAttribute(select_post_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                               : cudaFuncSetAttribute(select_post_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (dev >= 0 && dev < 64) attr_set[dev][di] = true;
  }
  if (dtype == 0) select_post_kernel<float><<<grid, kSpThreads, smem, s>>>(p, ns, sort_bytes, list_slots);
  else select_post_kernel<__nv_bfloat16><<<grid, kSpThreads, smem, s>>>(p, ns, sort_bytes, list_slots);
  return static_cast<int>(cudaGetLastError());
}

size_t post_fused_smem(int Cv, int K, int n_vert) {
  const size_t per = static_cast<size_t>((K + kPostSplit - 1) / kPostSplit);
  return (static_cast<size_t>(Cv) * (K + 1) * 2 + per * n_vert * 2 + per * 2) * sizeof(float);
}

int launch_post_fused(const PostFusedParams& p, int dtype, cudaStream_t s) {
  const size_t smem = post_fused_smem(p.Cv, p.K, p.n_vert);
  const unsigned grid = static_cast<unsigned>(p.B) * kPostSplit;
  cudaError_t e;
  if (dtype == 0) {
    e = ensure_dynamic_smem<post_fused_kernel<float>>(smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    post_fused_kernel<float><<<grid, kPostThreads, smem, s>>>(p);
  } else {
    e = ensure_dynamic_smem<post_fused_kernel<__nv_bfloat16>>(smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    post_fused_kernel<__nv_bfloat16><<<grid, kPostThreads, smem, s>>>(p);
  }
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// _group_vertexs_kf (models/model.py:134-162).  One CTA per image; thread per (detection n, channel k).
//   rel  = v_kj - m_n                 (:147)      diff = rel - off_kn        (:149)
//   dist = diff_x^2 + diff_y^2        (:150)      j*   = argmin_j, first minimal index (:151)
// m_n and off_kn are recomputed from the flat peak index exactly as the Tier A epilogue does (models/model.py:47-50).
template <typename T>
__global__ void __launch_bounds__(256) group_vertices_kernel(const GroupParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_xy = reinterpret_cast<float*>(smem_raw);            // [Cv*K*2]
  const int b = blockIdx.x;
  const int K = p.K, Cv = p.Cv, HW = p.H * p.W;
  const int n_det = p.counts[b];
  const float* kxy = p.kxy + static_cast<size_t>(b) * Cv * K * 2;
  const float* kscore = p.kscore + static_cast<size_t>(b) * Cv * K;
  for (int i = threadIdx.x; i < Cv * K * 2; i += blockDim.x) s_xy[i] = kxy[i];
  __syncthreads();
  const T* off2 = reinterpret_cast<const T*>(p.off2) + static_cast<size_t>(b) * 2 * HW;
  const T* off = reinterpret_cast<const T*>(p.off) + static_cast<size_t>(b) * 2 * p.n_vert * HW;
  for (int w = threadIdx.x; w < K * Cv; w += blockDim.x) {
    const int n = w / Cv, k = w - n * Cv;
    const size_t row = (static_cast<size_t>(b) * K + n) * Cv + k;
    if (n >= n_det) {
      p.kpt_proj[row * 2] = 0.f; p.kpt_proj[row * 2 + 1] = 0.f;
      p.kpt_score[row] = 0.f;
      p.kpt_j[row] = -1;
      if (p.verts_cv) { p.verts_cv[row * 2] = 0.f; p.verts_cv[row * 2 + 1] = 0.f; }
      continue;
    }
    const int flat = p.flat[static_cast<size_t>(b) * K + n];
    const int rem = flat % HW;
    const int yi = rem / p.W, xi = rem - yi * p.W;
    const float mx = subpixel(xi, to_f32(off2[rem]));
    const float my = subpixel(yi, to_f32(off2[HW + rem]));
    float ox = 0.f, oy = 0.f;
    if (k < p.n_vert) {
      ox = to_f32(off[static_cast<size_t>(2 * k) * HW + rem]);
      oy = to_f32(off[static_cast<size_t>(2 * k + 1) * HW + rem]);
    }
    const float* cand = s_xy + static_cast<size_t>(k) * K * 2;
    const int bj = nearest_candidate(cand, K, mx, my, ox, oy);
    p.kpt_proj[row * 2] = scale_coord(p.down, cand[2 * bj]);
    p.kpt_proj[row * 2 + 1] = scale_coord(p.down, cand[2 * bj + 1]);
    p.kpt_score[row] = kscore[static_cast<size_t>(k) * K + bj];
    p.kpt_j[row] = bj;
    if (p.verts_cv) {
      p.verts_cv[row * 2] = regress_coord(p.down, ox, mx);
      p.verts_cv[row * 2 + 1] = regress_coord(p.down, oy, my);
    }
  }
}

int launch_group(const GroupParams& p, int dtype, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(p.Cv) * p.K * 2 * sizeof(float);
  cudaError_t e;
  if (dtype == 0) {
    e = ensure_dynamic_smem<group_vertices_kernel<float>>(smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    group_vertices_kernel<float><<<p.B, 256, smem, s>>>(p);
  } else {
    e = ensure_dynamic_smem<group_vertices_kernel<__nv_bfloat16>>(smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    group_vertices_kernel<__nv_bfloat16><<<p.B, 256, smem, s>>>(p);
  }
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Tier C: closed-form 3D box recovery (NOT in the reference; normative text: oracle/box3d_ref.py).
// Geometry conventions are the reference's: Ry = [[c,0,s],[0,1,0],[-s,0,c]] with |sin|,|cos| < 1e-3 snapped to 0
// (utils/model_utils.py:66-76), corner order x in {+,-} L/2, y in {+,-} H/2, z in {+,-} W/2 nested in that order
// (:80-119), uv = (K X)[:2] / (z + 1e-6) (:147-152), flat-9 row-major camera matrix (datasets/dataset_reader.py:108).
constexpr float kPi = 3.14159265358979323846f;

template <typename T>
__global__ void __launch_bounds__(128) box3d_kernel(const Box3dParams p) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.K) return;
  const size_t row = static_cast<size_t>(b) * p.K + j;
  const bool valid = j < p.counts[b];
  float loc[3] = {0, 0, 0}, dim[3] = {0, 0, 0}, alpha = 0, roty = 0, uv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) uv[i] = 0.f;
  if (valid) {
    const int HW = p.H * p.W;
    const int flat = p.flat[row];
    const int cls = flat / HW;
    const int rem = flat - cls * HW;
    const int yi = rem / p.W, xi = rem - yi * p.W;
    const T* reg = reinterpret_cast<const T*>(p.reg) + static_cast<size_t>(b) * p.Creg * HW + rem;
    float r[14];
#pragma unroll
    for (int c = 0; c < 14; ++c) r[c] = (c < p.Creg) ? to_f32(reg[static_cast<size_t>(c) * HW]) : 0.f;
    const float* cam = p.cam + static_cast<size_t>(b) * 9;
    const float fx = cam[0], cx = cam[2], fy = cam[4], cy = cam[5];
    const bool multibin = (p.mode & 1) != 0;
    const bool sig_sub = (p.mode & 2) != 0;
    const float u = static_cast<float>(xi) + (sig_sub ? sigmoid_ref(r[1]) : r[1]);
    const float v = static_cast<float>(yi) + (sig_sub ? sigmoid_ref(r[2]) : r[2]);
    float z;
    if (multibin) z = 1.0f / (sigmoid_ref(r[0]) + 1e-6f) - 1.0f;
    else z = r[0] * p.depth_sigma + p.depth_mu;
    loc[0] = (u - cx) * z / fx;
    loc[1] = (v - cy) * z / fy;
    loc[2] = z;
#pragma unroll
    for (int d = 0; d < 3; ++d) dim[d] = expf(sigmoid_ref(r[3 + d]) - 0.5f) * p.dim_ref[cls * 3 + d];
    if (multibin) {
      if (r[7] > r[11]) alpha = atan2f(r[8], r[9]) - 0.5f * kPi;
      else alpha = atan2f(r[12], r[13]) + 0.5f * kPi;
      roty = alpha + atan2f(u - cx, fx);
    } else {
      const float nrm = sqrtf(r[6] * r[6] + r[7] * r[7]);
      const float o0 = r[6] / nrm, o1 = r[7] / nrm;
      alpha = atanf(o0 / (o1 + 1e-7f));
      alpha += (o1 >= 0.f) ? -0.5f * kPi : 0.5f * kPi;
      roty = alpha + atanf(loc[0] / (loc[2] + 1e-7f));
    }
    if (roty > kPi) roty -= 2.0f * kPi;
    if (roty < -kPi) roty += 2.0f * kPi;
    float sn = sinf(roty), cs = cosf(roty);
    if (fabsf(sn) < 1e-3f) sn = 0.f;
    if (fabsf(cs) < 1e-3f) cs = 0.f;
    const float hl = 0.5f * dim[2], hh = 0.5f * dim[0], hw = 0.5f * dim[1];  // dims are (h,w,l)
    int q = 0;
#pragma unroll
    for (int i = 1; i >= -1; i -= 2)
#pragma unroll
      for (int jj = 1; jj >= -1; jj -= 2)
#pragma unroll
        for (int k = 1; k >= -1; k -= 2) {
          const float X = cs * (hl * i) + sn * (hw * k) + loc[0];
          const float Y = hh * jj + loc[1];
          const float Z = -sn * (hl * i) + cs * (hw * k) + loc[2];
          const float px = cam[0] * X + cam[1] * Y + cam[2] * Z;
          const float py = cam[3] * X + cam[4] * Y + cam[5] * Z;
          const float pz = cam[6] * X + cam[7] * Y + cam[8] * Z;
          uv[2 * q] = px / (pz + 1e-6f);
          uv[2 * q + 1] = py / (pz + 1e-6f);
          ++q;
        }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) { p.loc[row * 3 + d] = loc[d]; p.dim[row * 3 + d] = dim[d]; }
  p.alpha[row] = alpha;
  p.rot_y[row] = roty;
#pragma unroll
  for (int i = 0; i < 16; ++i) p.corners2d[row * 16 + i] = uv[i];
}

// Wire rows for the NCCL gather of the detections: per image K rows of (9 + 2V) 32-bit words
// (cls | score | proj 2 | verts 2V | bbox 4 | flat) followed by one word holding counts[b].
__global__ void __launch_bounds__(256) pack_wire_kernel(const int64_t* cls, const float* score, const float* proj, const float* verts,
                                                        const float* bbox, const int32_t* flat, const int32_t* counts, int B, int K,
                                                        int V, int32_t* wire) {
  const int words = 9 + 2 * V;
  const size_t per_img = static_cast<size_t>(K) * words + 1;
  const size_t n = static_cast<size_t>(B) * per_img;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_img);
    const int r = static_cast<int>(i - b * per_img);
    int32_t v;
    if (r == K * words) {
      v = counts[b];
    } else {
      const int j = r / words, w = r - j * words;
      const size_t row = static_cast<size_t>(b) * K + j;
      if (w == 0) v = static_cast<int32_t>(cls[row]);
      else if (w == 1) v = __float_as_int(score[row]);
      else if (w < 4) v = __float_as_int(proj[row * 2 + (w - 2)]);
      else if (w < 4 + 2 * V) v = __float_as_int(verts[row * 2 * V + (w - 4)]);
      else if (w < 8 + 2 * V) v = __float_as_int(bbox[row * 4 + (w - 4 - 2 * V)]);
      else v = flat[row];
    }
    wire[i] = v;
  }
}
int launch_pack_wire(const int64_t* cls, const float* score, const float* proj, const float* verts, const float* bbox,
                     const int32_t* flat, const int32_t* counts, int B, int K, int V, int32_t* wire, cudaStream_t s) {
  const size_t n = static_cast<size_t>(B) * (static_cast<size_t>(K) * (9 + 2 * V) + 1);
  unsigned grid = static_cast<unsigned>((n + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  pack_wire_kernel<<<grid, 256, 0, s>>>(cls, score, proj, verts, bbox, flat, counts, B, K, V, wire);
  return static_cast<int>(cudaGetLastError());
}

// The library's sigmoid, element-wise (verification aid: tests sweep every fp32 value through it).
__global__ void sigmoid_kernel(const float* x, float* y, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    y[i] = sigmoid_ref(x[i]);
}
int launch_sigmoid(const float* x, float* y, size_t n, cudaStream_t s) {
  sigmoid_kernel<<<148 * 8, 256, 0, s>>>(x, y, n);
  return static_cast<int>(cudaGetLastError());
}

int launch_box3d(const Box3dParams& p, int dtype, cudaStream_t s) {
  dim3 grid((p.K + 127) / 128, p.B);
  if (dtype == 0) box3d_kernel<float><<<grid, 128, 0, s>>>(p);
  else box3d_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(p);
  return static_cast<int>(cudaGetLastError());
}

// One thread waits until every rank's arrival flag has reached `value` (the rows of that batch have landed in this rank's
// gather buffer).  Bounded: a lost peer surfaces as a launch failure, not as a hung GPU.
__global__ void wait_flags_kernel(const uint32_t* flags, int n, uint32_t value) {
  const long long t0 = clock64();
  for (int r = 0; r < n; ++r) {
    while (static_cast<int32_t>(*reinterpret_cast<const volatile uint32_t*>(flags + r) - value) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000LL) __trap();
    }
  }
  __threadfence_system();
}
// Flush of the deferred gather (behind the last batch): the rows, then -- in a second launch, i.e. with every store of the
// first one complete -- the arrival flag.
struct PushPeers { int32_t* p[8]; };
__global__ void __launch_bounds__(128) push_rows_kernel(PushPeers peers, int n_peers, int rank, int B, int K, int words) {
  const int b = blockIdx.x / kPostSplit, q = blockIdx.x % kPostSplit;
  const int per = (K + kPostSplit - 1) / kPostSplit, n0 = q * per, n1 = min(K, n0 + per);
  const size_t per_img = static_cast<size_t>(K) * words + 1;
  const size_t img0 = (static_cast<size_t>(rank) * B + b) * per_img;
  // (the image's last block takes the count word along)
  const int n_words = max(0, n1 - n0) * words + (q == kPostSplit - 1 ? 1 : 0);
  const size_t s0 = q == kPostSplit - 1 && n1 <= n0 ? img0 + static_cast<size_t>(K) * words : img0 + static_cast<size_t>(n0) * words;
  push_words<128>(peers.p, n_peers, rank, s0, n_words);
}
__global__ void push_flag_kernel(PushPeers peers, int n_peers, int rank, size_t flag_offset, uint32_t step_id) {
  if (threadIdx.x < n_peers) reinterpret_cast<volatile uint32_t*>(peers.p[threadIdx.x] + flag_offset)[rank] = step_id;
}
int launch_push_rows(int32_t* const* peers, int n_peers, int rank, int B, int K, int n_vert, uint32_t step_id, size_t flag_offset, cudaStream_t s) {
  PushPeers pp{};
  for (int r = 0; r < n_peers; ++r) pp.p[r] = peers[r];
  push_rows_kernel<<<static_cast<unsigned>(B) * kPostSplit, 128, 0, s>>>(pp, n_peers, rank, B, K, 9 + 2 * n_vert);
  push_flag_kernel<<<1, 32, 0, s>>>(pp, n_peers, rank, flag_offset, step_id);
  return static_cast<int>(cudaGetLastError());
}

int launch_push_flag(int32_t* const* peers, int n_peers, int rank, uint32_t step_id, size_t flag_offset, cudaStream_t s) {
  PushPeers pp{};
  for (int r = 0; r < n_peers; ++r) pp.p[r] = peers[r];
  push_flag_kernel<<<1, 32, 0, s>>>(pp, n_peers, rank, flag_offset, step_id);
  return static_cast<int>(cudaGetLastError());
}

int launch_wait_flags(const uint32_t* flags, int n, uint32_t value, cudaStream_t s) {
  wait_flags_kernel<<<1, 1, 0, s>>>(flags, n, value);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rtm3d
