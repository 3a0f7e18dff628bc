// Shape-generic decode: strip-per-CTA scan + per-image (or per-plane) merge by the last CTA to arrive.
//
// This is the path for shapes the streaming kernel (decode_stream.cu) does not take (rows not 16-byte aligned,
// W % 4 != 0, ...), and the anchor the streaming kernel is tested against.  One launch, no host sync:
//
//   grid = (nstrips * C, B).  A CTA loads rows [y0-1, y1] of one (image, channel) plane into shared memory (-inf
//   outside the image = max_pool2d's implicit padding), finds the 3x3 peaks in the SIGMOID domain
//   (utils/model_utils.py:17-26 on models/model.py:85), keeps its K best keys (radix select) and publishes them.
//   The last CTA of the image (kModeMain; models/model.py:87-98 is a flat top-K over C*H*W) or of the plane
//   (kModeKpt; :109-114 is per channel) merges the published lists, sorts, and runs the epilogue.
#include "common.cuh"
#include "epilogue.cuh"
#include "params.h"

namespace rtm3d {

constexpr int kGenericThreads = 256;

template <typename T, int MODE>
__global__ void __launch_bounds__(kGenericThreads) decode_generic_kernel(const DecodeParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = p.W, H = p.H, HW = H * W, K = p.K;
  const int R = p.strip_rows;
  const int kpad = next_pow2(K);
  // shared layout: tile f32[(R+2)*W] | list u64[list_cap] | best u64[kpad] | hist u32[264] | scratch u32[3K+8]
  float* tile = reinterpret_cast<float*>(smem_raw);
  size_t o = (static_cast<size_t>(R + 2) * W * sizeof(float) + 15) & ~size_t(15);
  uint64_t* list = reinterpret_cast<uint64_t*>(smem_raw + o);
  o += static_cast<size_t>(p.list_cap) * 8;
  uint64_t* best = reinterpret_cast<uint64_t*>(smem_raw + o);
  o += static_cast<size_t>(kpad) * 8;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + o);
  o += 264 * 4;
  uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw + o);
  __shared__ uint32_t s_count, s_last;

  const int strip = blockIdx.x % p.nstrips;
  const int c = blockIdx.x / p.nstrips;
  const int b = blockIdx.y;
  const int y0 = strip * R;
  const int rows = min(R, H - y0);
  const T* plane = reinterpret_cast<const T*>(p.hm) + (static_cast<size_t>(b) * p.C + c) * HW;

  // ---- load rows y0-1 .. y0+rows into the tile (tile row r <-> image row y0-1+r)
  for (int i = threadIdx.x; i < (rows + 2) * W; i += kGenericThreads) {
    const int r = i / W;
    const int y = y0 - 1 + r;
    tile[i] = (y >= 0 && y < H) ? to_f32(plane[static_cast<size_t>(y) * W + (i - r * W)]) : -INFINITY;
  }
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  // ---- scan: logit-domain prefilter, exact sigmoid-domain peak test for the survivors
  const float t0 = p.t0;
  for (int i = threadIdx.x; i < rows * W; i += kGenericThreads) {
    const int r = i / W;
    const int x = i - r * W;
    const float xc = tile[(r + 1) * W + x];
    if (!(xc >= t0)) continue;
    const float sc = sigmoid_ref(xc);
    if (MODE == kModeMain ? !(sc > p.thresh) : !(sc > 0.0f)) continue;
    bool peak = true;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        if (dy == 1 && dx == 0) continue;
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        const float xn = tile[(r + dy) * W + xx];
        if (neighbour_needs_exact(xn, xc) && sigmoid_ref(xn) > sc) peak = false;
      }
    }
    if (!peak) continue;
    const uint32_t flat = (MODE == kModeMain ? static_cast<uint32_t>(c) * HW : 0u) + static_cast<uint32_t>(y0 + r) * W + x;
    list[atomicAdd(&s_count, 1u)] = make_key(sc, flat);
  }
  __syncthreads();

  // ---- strip-local top-K, publish
  const int n_local = block_select_topk(list, static_cast<int>(s_count), K, best, hist);
  const int units_per_ticket = (MODE == kModeMain) ? p.nstrips * p.C : p.nstrips;
  const int ticket_id = (MODE == kModeMain) ? b : b * p.C + c;
  const size_t unit = (static_cast<size_t>(b) * p.C + c) * p.nstrips + strip;
  for (int i = threadIdx.x; i < n_local; i += kGenericThreads) p.keys[unit * K + i] = best[i];
  if (threadIdx.x == 0) p.key_counts[unit] = n_local;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&p.tickets[ticket_id], 1u);
    s_last = (t == static_cast<uint32_t>(units_per_ticket - 1));
    if (s_last) p.tickets[ticket_id] = 0;  // leave the workspace clean for the next call
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // ---- merge (last CTA of the image / plane): running best + as many published lists as fit, repeatedly
  const size_t unit0 = (MODE == kModeMain) ? static_cast<size_t>(b) * p.C * p.nstrips
                                            : (static_cast<size_t>(b) * p.C + c) * p.nstrips;
  int have = 0;  // keys currently in best[]
  int u = 0;
  while (u < units_per_ticket) {
    int n = 0;
    for (int i = threadIdx.x; i < have; i += kGenericThreads) list[i] = best[i];
    n = have;
    __syncthreads();
    while (u < units_per_ticket) {
      const int cu = static_cast<int>(__ldcg(&p.key_counts[unit0 + u]));
      if (n + cu > p.list_cap) break;
      for (int i = threadIdx.x; i < cu; i += kGenericThreads) list[n + i] = __ldcg(&p.keys[(unit0 + u) * K + i]);
      n += cu;
      ++u;
    }
    __syncthreads();
    have = block_select_topk(list, n, K, best, hist);
  }
  for (int i = have + threadIdx.x; i < kpad; i += kGenericThreads) best[i] = 0ull;
  __syncthreads();
  block_bitonic_sort_desc(best, kpad);
  block_emit<T, MODE>(p, b, c, best, have, scratch);
}

size_t generic_smem_bytes(int W, int K, int strip_rows, int list_cap) {
  size_t o = (static_cast<size_t>(strip_rows + 2) * W * sizeof(float) + 15) & ~size_t(15);
  o += static_cast<size_t>(list_cap) * 8 + static_cast<size_t>(next_pow2(K)) * 8 + 264 * 4 + (3 * static_cast<size_t>(K) + 8) * 4;
  return o;
}

WorkspaceLayout workspace_layout(int B, int C, int H, int W, int K) {
  WorkspaceLayout L{};
  // rows per strip: tile (R+2)*W*4 + list R*W*8 within ~96 KB so two CTAs fit an SM
  const size_t budget = 96 * 1024;
  long r = (static_cast<long>(budget) - 8L * W) / (12L * W);
  if (r < 1) r = 1;
  if (r > H) r = H;
  L.strip_rows = static_cast<int>(r);
  L.nstrips = (H + L.strip_rows - 1) / L.strip_rows;
  const long cap = static_cast<long>(L.strip_rows) * W;
  L.list_cap = static_cast<int>(cap > 2L * K ? cap : 2L * K);
  L.generic_smem = generic_smem_bytes(W, K, L.strip_rows, L.list_cap);
  size_t off = 0;
  L.table_off = off;                                 // fixed position: rtm3d_workspace_init fills it without knowing the shape
  off += static_cast<size_t>(kFilterTableWords) * 4;
  L.tickets_off = off;
  off += ((static_cast<size_t>(B) * (C + 1) * 4 + 255) / 256) * 256;
  L.status_off = off;
  off += 256;
  L.keys_off = off;
  // units: strips of the generic kernel, or items (<= kPlanesMaxSplit strips per plane) of the plane-streaming kernel
  const size_t units = static_cast<size_t>(B) * C * (L.nstrips > kPlanesMaxSplit ? L.nstrips : kPlanesMaxSplit);
  off += units * K * 8;
  off = (off + 255) / 256 * 256;
  L.counts_off = off;
  off += units * 4;
  off = (off + 255) / 256 * 256;
  L.retry_off = off;
  off += units * 4;
  off = (off + 255) / 256 * 256;
  L.guess_off = off;
  off += 256;
  L.flat_off = off;                                  // [B][K] flat peak indices when the caller does not ask for them
  off += static_cast<size_t>(B) * K * 4;
  off = (off + 255) / 256 * 256;
  L.queue_off = off;                                 // strip counter of the scan kernel
  off += 256;
  // candidate lists of the scan kernel: one per strip.  Room for the strips the kernel chooses by itself (either element
  // type) and, while that stays small, for the 8 strips per plane a caller may force.
  L.cand_cap = scan_list_cap(K);
  int sp = scan_max_strips_per_plane(H, W, K);
  if (sp < 8 && static_cast<size_t>(B) * C * 8 * L.cand_cap * 8 <= (32u << 20)) sp = 8;
  if (sp < 2) sp = 2;
  L.cand_strips = B * C * sp;
  L.cand_count_off = off;
  off += (static_cast<size_t>(L.cand_strips) * 4 + 255) / 256 * 256;
  L.cand_off = off;
  off += static_cast<size_t>(L.cand_strips) * L.cand_cap * 8;
  L.total = (off + 255) / 256 * 256;
  return L;
}

template <typename T, int MODE>
static int launch_generic_t(const DecodeParams& p, size_t smem, cudaStream_t s) {
  auto kern = decode_generic_kernel<T, MODE>;
  cudaError_t e = ensure_dynamic_smem<decode_generic_kernel<T, MODE>>(smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid(p.nstrips * p.C, p.B);
  kern<<<grid, kGenericThreads, smem, s>>>(p);
  return static_cast<int>(cudaGetLastError());
}

int launch_generic(const DecodeParams& p, int dtype, int mode, size_t smem, cudaStream_t s) {
  if (dtype == 0) {
    return mode == kModeMain ? launch_generic_t<float, kModeMain>(p, smem, s) : launch_generic_t<float, kModeKpt>(p, smem, s);
  }
  return mode == kModeMain ? launch_generic_t<__nv_bfloat16, kModeMain>(p, smem, s)
                           : launch_generic_t<__nv_bfloat16, kModeKpt>(p, smem, s);
}

}  // namespace rtm3d
