// extern "C" boundary of librtm3d_decode.so (include/rtm3d_decode.h): argument validation, workspace carving,
// launch.  No allocation, no host synchronisation, no state kept after return.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/rtm3d_decode.h"
#include "params.h"
#include "postproc.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(int e, const char* what) {
  if (e == 0) return 0;
  return fail(e, "%s: %s", what, cudaGetErrorString(static_cast<cudaError_t>(e)));
}

size_t elem_size(int dtype) { return dtype == RTM3D_F32 ? 4 : 2; }

int check_shape(int B, int C, int H, int W, int K) {
  if (B < 1 || C < 1 || H < 1 || W < 1 || B > 65535) return fail(RTM3D_ERR_SHAPE, "bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  if (static_cast<double>(C) * H * W >= 2147483648.0) return fail(RTM3D_ERR_SHAPE, "C*H*W must be < 2^31");
  if (W > 16384) return fail(RTM3D_ERR_SHAPE, "W=%d > 16384 unsupported", W);
  if (K < 1 || K > RTM3D_MAX_TOPK) return fail(RTM3D_ERR_TOPK, "K=%d outside [1,%d]", K, RTM3D_MAX_TOPK);
  return 0;
}

// x < t0  =>  computed sigmoid(x) <= thresh (so the strict `score > thresh` of models/model.py:91 fails).
float prefilter_logit(float thresh) {
  if (!(thresh > 0.0f)) return -INFINITY;
  if (thresh >= 1.0f) return INFINITY;
  const double t = thresh;
  const double L = std::log(t / (1.0 - t));
  const double delta = std::ldexp(1.0, -16) / (1.0 - t) + std::ldexp(1.0, -16) * std::fabs(L) + std::ldexp(1.0, -16);
  const double v = L - delta;
  float f = static_cast<float>(v);
  if (static_cast<double>(f) > v) f = std::nextafterf(f, -INFINITY);
  return f;
}

void carve(rtm3d::DecodeParams& p, const rtm3d::WorkspaceLayout& L, void* ws) {
  unsigned char* base = static_cast<unsigned char*>(ws);
  p.tickets = reinterpret_cast<uint32_t*>(base + L.tickets_off);
  p.keys = reinterpret_cast<uint64_t*>(base + L.keys_off);
  p.key_counts = reinterpret_cast<uint32_t*>(base + L.counts_off);
  p.status = reinterpret_cast<uint32_t*>(base + L.status_off);
  p.strip_rows = L.strip_rows;
  p.nstrips = L.nstrips;
  p.list_cap = L.list_cap;
}

// The wide epilogue kernels that follow the plane-streaming kernel (it writes score / flat / counts and kscore / kflat).
int launch_epilogues(const rtm3d::PlaneParams& q, int dtype, unsigned flags, cudaStream_t s) {
  if (flags & RTM3D_FLAG_NO_EPILOGUE) return 0;
  if (q.C > 0) {
    rtm3d::EpiMainParams e{q.flat, q.counts, q.off, q.off2_main, q.B, q.C, q.H, q.W, q.n_vert, q.K, q.down, q.cls, q.proj, q.verts, q.bbox};
    if (int r = cuda_fail(rtm3d::launch_epilogue_main(e, dtype, s), "Tier A epilogue launch")) return r;
  }
  if (q.Cv > 0) {
    rtm3d::EpiKptParams e{q.kflat, q.off2_kpt, q.B, q.Cv, q.H, q.W, q.K, q.kxy};
    if (int r = cuda_fail(rtm3d::launch_epilogue_kpt(e, dtype, s), "Tier B epilogue launch")) return r;
  }
  return 0;
}


// Plane-resident scan kernel + selection: writes score / flat / counts (C > 0) and kscore / kflat (Cv > 0).
// Returns -1000 when the shape is not eligible for the scan kernel (the caller falls back).
// post: when given (rtm3d_decode_fused), the selection and everything after it run in ONE kernel behind the scan kernel.
struct GatherTarget { void* const* peers; int n_peers, rank; uint32_t step_id; size_t flag_offset; int deferred; void* const* prev_peers; uint32_t prev_step; };
int scan_and_select(const rtm3d::PlaneParams& q, const rtm3d::WorkspaceLayout& L, void* ws, int dtype, unsigned flags, cudaStream_t s,
                    const rtm3d::PostFusedParams* post = nullptr, const GatherTarget* gt = nullptr) {
  unsigned char* base = static_cast<unsigned char*>(ws);
  rtm3d::ScanParams sp{};
  sp.hm_main = q.hm_main; sp.hm_kpt = q.hm_kpt;
  sp.B = q.B; sp.C = q.C; sp.Cv = q.Cv; sp.H = q.H; sp.W = q.W; sp.K = q.K;
  sp.thresh = q.thresh; sp.t0 = q.t0;
  sp.cand = reinterpret_cast<unsigned long long*>(base + L.cand_off);
  sp.cand_count = reinterpret_cast<uint32_t*>(base + L.cand_count_off);
  sp.queue = reinterpret_cast<uint32_t*>(base + L.queue_off);
  sp.status = q.status;
#ifdef RTM3D_DEV
  sp.stats = rtm3d::debug_get_stats();
#else
  sp.stats = nullptr;
#endif
  const int strips_override = static_cast<int>((flags >> 8) & 0xFu);
  // the lists of this call must fit the workspace
  const int sp_default = rtm3d::scan_strips_per_plane(q.H, q.W, q.K, dtype, strips_override);
  if (sp_default <= 0 || static_cast<long long>(q.B) * (q.C + q.Cv) * sp_default > L.cand_strips || rtm3d::scan_list_cap(q.K) > L.cand_cap) return -1000;
  // planes that do not fit the ring as ONE strip: every strip has to supply its own K best, which multiplies the candidates;
  // the streaming kernel of round 1 is the faster one there (unless the caller forces a number of strips)
  if (strips_override == 0 && sp_default > 1) return -1000;
  if (rtm3d::select_smem_bytes(q.C, q.Cv, sp_default, rtm3d::scan_list_cap(q.K), q.K) > 200 * 1024) return -1000;
  int Sp = 0, cap = 0;
  const int rc = rtm3d::launch_scan(sp, dtype, strips_override, static_cast<int>((flags >> 16) & 0xFFu), static_cast<int>((flags >> 24) & 0xFu), s, &Sp, &cap);
  if (rc == -1000) return rc;
  if (int e = cuda_fail(rc, "decode (scan kernel) launch")) return e;
  if (flags & RTM3D_FLAG_NO_SELECT) return 0;                       // the caller continues with rtm3d_select_post (bench.py's marks)
  if (post) {
    rtm3d::SelectPostParams f{sp.cand, sp.cand_count, Sp, cap, q.thresh, q.score, q.flat, q.counts, q.kscore, q.kflat, *post, sp.stats, {}, 0, 0, 0u, 0, nullptr, 0, {}, 0u};
    if (gt) {
      for (int r = 0; r < gt->n_peers; ++r) f.wire_peers[r] = static_cast<int32_t*>(gt->peers[r]);
      f.n_peers = gt->n_peers; f.wire_rank = gt->rank;
      f.step_id = gt->step_id; f.flag_offset = gt->flag_offset;
      f.done_counter = reinterpret_cast<uint32_t*>(base + L.queue_off + 64);
      f.deferred = gt->deferred;
      if (gt->deferred && gt->prev_peers && gt->prev_step != 0u) {
        for (int r = 0; r < gt->n_peers; ++r) f.push_peers[r] = static_cast<int32_t*>(gt->prev_peers[r]);
        f.push_step = gt->prev_step;
      }
    }
    return cuda_fail(rtm3d::launch_select_post(f, dtype, s), "decode (select + post kernel) launch");
  }
  rtm3d::SelectParams sel{sp.cand, sp.cand_count, q.B, q.C, q.Cv, q.H, q.W, q.K, Sp, cap, q.thresh, q.score, q.flat, q.counts, q.kscore, q.kflat};
  return cuda_fail(rtm3d::launch_select(sel, s), "decode (select kernel) launch");
}

int dispatch(rtm3d::DecodeParams& p, const rtm3d::WorkspaceLayout& L, int dtype, int mode, unsigned flags,
             cudaStream_t s) {
  if (!(flags & RTM3D_FLAG_FORCE_GENERIC)) {
    rtm3d::PlaneParams q{};
    const bool main = mode == rtm3d::kModeMain;
    q.hm_main = main ? p.hm : nullptr;
    q.hm_kpt = main ? nullptr : p.hm;
    q.off = p.off;
    q.off2_main = main ? p.off2 : nullptr;
    q.off2_kpt = main ? nullptr : p.off2;
    q.B = p.B; q.C = main ? p.C : 0; q.Cv = main ? 0 : p.C; q.H = p.H; q.W = p.W; q.n_vert = p.n_vert; q.K = p.K;
    q.thresh = p.thresh; q.down = p.down; q.t0 = p.t0;
    q.cls = p.cls; q.score = p.score; q.proj = p.proj; q.verts = p.verts; q.bbox = p.bbox; q.flat = p.flat; q.counts = p.counts;
    q.kscore = p.kscore; q.kxy = p.kxy; q.kflat = p.kflat;
    q.tickets = p.tickets; q.keys = p.keys; q.key_counts = p.key_counts; q.status = p.status;
    q.retry = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(p.tickets) - L.tickets_off + L.retry_off);
    q.guess = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(p.tickets) - L.tickets_off + L.guess_off);
    q.ftable = reinterpret_cast<const float*>(reinterpret_cast<unsigned char*>(p.tickets) - L.tickets_off + L.table_off);
    if (main && !q.flat)   // the epilogue needs the flat indices even when the caller does not
      q.flat = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(p.tickets) - L.tickets_off + L.flat_off);
    int rc = -1000;
    if (!(flags & RTM3D_FLAG_LEGACY_PLANES)) {
      rc = scan_and_select(q, L, reinterpret_cast<unsigned char*>(p.tickets) - L.tickets_off, dtype, flags, s);
      if (rc > 0 || (rc < 0 && rc != -1000)) return rc;
    }
    if (rc == -1000) {
      // (the developer bits of the flags address the scan kernel unless the round-1 kernel was asked for)
      rc = rtm3d::launch_planes(q, dtype, static_cast<int>((flags >> 8) & 0xFu), (flags & RTM3D_FLAG_NO_SPECULATION) ? 0 : 1,
                                static_cast<int>((flags >> 16) & 0xFFu), (flags & RTM3D_FLAG_LEGACY_PLANES) ? static_cast<int>((flags >> 24) & 0xFu) : 0, s);   // -1000: shape not eligible
      if (rc != -1000) if (int e = cuda_fail(rc, "decode (plane-streaming kernel) launch")) return e;
    }
    if (rc != -1000) return launch_epilogues(q, dtype, flags, s);
  }
  if (L.generic_smem > 200 * 1024) return fail(RTM3D_ERR_SHAPE, "row too wide for the generic kernel (%zu B smem)", L.generic_smem);
  return cuda_fail(rtm3d::launch_generic(p, dtype, mode, L.generic_smem, s), "decode (generic kernel) launch");
}

}  // namespace

extern "C" {

int rtm3d_abi_version(void) { return RTM3D_ABI_VERSION; }

const char* rtm3d_last_error(void) { return g_err; }

const char* rtm3d_build_info(void) {
  return "librtm3d_decode: host compiler " __VERSION__ ", nvcc "
#define RTM3D_STR2(x) #x
#define RTM3D_STR(x) RTM3D_STR2(x)
      RTM3D_STR(__CUDACC_VER_MAJOR__) "." RTM3D_STR(__CUDACC_VER_MINOR__)
      " sm_100a; kernels: scan_planes (persistent, plane-resident cp.async.bulk ring, verified order-statistic threshold),"
      " select_post (cluster of 4 CTAs per image: register sort, epilogues, grouping, fused gather), select, decode_planes"
      " (round-1 streaming kernel: planes larger than the ring), decode_generic (any shape), fit_box3d, box3d, target encoder,"
      " focal + gather-L1 losses; fp32+bf16 head maps";
}

#ifdef RTM3D_DEV
/* developer instrumentation (tools/scan_stats.py, tools/plane_stats.py): only in `make DEV=1` builds, never in the
   production library, and deliberately absent from include/rtm3d_decode.h */
void rtm3d_debug_set_copy_rows(int rows) { rtm3d::debug_set_copy_rows(rows); }
void rtm3d_debug_set_trace(void* t) { rtm3d::debug_set_trace(static_cast<unsigned long long*>(t)); }
void rtm3d_debug_set_stats(void* dev_u64_16) { rtm3d::debug_set_stats(static_cast<unsigned long long*>(dev_u64_16)); }
#endif

int rtm3d_decode_workspace_bytes(int B, int C, int H, int W, int K, size_t* out_bytes) {
  if (!out_bytes) return fail(RTM3D_ERR_NULL, "out_bytes is NULL");
  if (int e = check_shape(B, C, H, W, K)) return e;
  *out_bytes = rtm3d::workspace_layout(B, C, H, W, K).total;
  return 0;
}

int rtm3d_workspace_init(void* ws, size_t ws_bytes, void* stream) {
  if (!ws) return fail(RTM3D_ERR_NULL, "ws is NULL");
  if (int e = cuda_fail(static_cast<int>(cudaMemsetAsync(ws, 0, ws_bytes, static_cast<cudaStream_t>(stream))), "workspace memset")) return e;
  // the logit bound of every histogram bin (double-precision log, once per workspace instead of inside the kernels)
  if (ws_bytes >= static_cast<size_t>(rtm3d::kFilterTableWords) * 4)
    return cuda_fail(rtm3d::launch_filter_table(static_cast<float*>(ws), static_cast<cudaStream_t>(stream)), "threshold table launch");
  return 0;
}

int rtm3d_decode_main(const void* hm, const void* off, const void* off2, int dtype, int B, int C, int H, int W,
                      int n_vert, int K, float thresh, float down, int64_t* cls, float* score, float* proj,
                      float* verts, float* bbox, int32_t* flat, int32_t* counts, void* ws, size_t ws_bytes,
                      unsigned flags, void* stream) {
  if (!hm || !off || !off2 || !cls || !score || !proj || !verts || !bbox || !counts || !ws)
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, C, H, W, K)) return e;
  if (n_vert < 1 || n_vert > RTM3D_MAX_VERTS) return fail(RTM3D_ERR_SHAPE, "n_vert=%d outside [1,%d]", n_vert, RTM3D_MAX_VERTS);
  if (static_cast<long long>(K) > static_cast<long long>(C) * H * W) return fail(RTM3D_ERR_TOPK, "K=%d > C*H*W", K);
  if (!(thresh >= 0.0f)) return fail(RTM3D_ERR_THRESH, "score threshold must be >= 0 (got %g)", thresh);
  const size_t es = elem_size(dtype);
  if (reinterpret_cast<uintptr_t>(hm) % es || reinterpret_cast<uintptr_t>(off) % es || reinterpret_cast<uintptr_t>(off2) % es ||
      reinterpret_cast<uintptr_t>(cls) % 8 || reinterpret_cast<uintptr_t>(ws) % 256)
    return fail(RTM3D_ERR_ALIGN, "misaligned pointer (maps: element size, cls: 8 B, ws: 256 B)");
  const rtm3d::WorkspaceLayout L = rtm3d::workspace_layout(B, C, H, W, K);
  if (ws_bytes < L.total) return fail(RTM3D_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, L.total);
  rtm3d::DecodeParams p{};
  p.hm = hm; p.off = off; p.off2 = off2;
  p.B = B; p.C = C; p.H = H; p.W = W; p.n_vert = n_vert; p.K = K;
  p.thresh = thresh; p.down = down; p.t0 = prefilter_logit(thresh);
  p.cls = cls; p.score = score; p.proj = proj; p.verts = verts; p.bbox = bbox; p.flat = flat; p.counts = counts;
  carve(p, L, ws);
  return dispatch(p, L, dtype, rtm3d::kModeMain, flags, static_cast<cudaStream_t>(stream));
}

int rtm3d_select_main(const void* hm, int dtype, int B, int C, int H, int W, int K, float thresh, float* score, int32_t* flat,
                      int32_t* counts, void* ws, size_t ws_bytes, unsigned flags, void* stream) {
  if (!hm || !score || !flat || !counts || !ws) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, C, H, W, K)) return e;
  if (static_cast<long long>(K) > static_cast<long long>(C) * H * W) return fail(RTM3D_ERR_TOPK, "K=%d > C*H*W", K);
  if (!(thresh >= 0.0f)) return fail(RTM3D_ERR_THRESH, "score threshold must be >= 0 (got %g)", thresh);
  if (reinterpret_cast<uintptr_t>(hm) % elem_size(dtype) || reinterpret_cast<uintptr_t>(ws) % 256)
    return fail(RTM3D_ERR_ALIGN, "misaligned pointer (map: element size, ws: 256 B)");
  const rtm3d::WorkspaceLayout L = rtm3d::workspace_layout(B, C, H, W, K);
  if (ws_bytes < L.total) return fail(RTM3D_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, L.total);
  rtm3d::DecodeParams p{};
  p.hm = hm; p.off = nullptr; p.off2 = nullptr;                  // no regression maps: the selection stops before the gathers
  p.B = B; p.C = C; p.H = H; p.W = W; p.n_vert = 1; p.K = K;
  p.thresh = thresh; p.down = 1.f; p.t0 = prefilter_logit(thresh);
  p.score = score; p.flat = flat; p.counts = counts;
  carve(p, L, ws);
  return dispatch(p, L, dtype, rtm3d::kModeMain, flags | RTM3D_FLAG_NO_EPILOGUE, static_cast<cudaStream_t>(stream));
}

int rtm3d_decode_main_host(const void* hm_host, const void* off_host, const void* off2_host, int dtype, int B, int C,
                           int H, int W, int n_vert, int K, float thresh, float down, void* dev_hm, int64_t* cls,
                           float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts,
                           int64_t* cls_host, float* score_host, float* proj_host, float* verts_host, float* bbox_host,
                           int32_t* flat_host, int32_t* counts_host, void* ws, size_t ws_bytes, unsigned flags,
                           void* stream) {
  if (!hm_host || !off_host || !off2_host || !dev_hm || !counts_host) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, C, H, W, K)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // regression planes stay in host memory: the epilogue gathers K*(2V+2) scalars per image through the mapped pointer
  void *d_off = nullptr, *d_off2 = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(&d_off, const_cast<void*>(off_host), 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d_off2, const_cast<void*>(off2_host), 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(static_cast<int>(e), "off_host/off2_host must be page-locked mapped host memory: %s", cudaGetErrorString(e));
  }
  const size_t hm_bytes = static_cast<size_t>(B) * C * H * W * elem_size(dtype);
  if (int r = cuda_fail(static_cast<int>(cudaMemcpyAsync(dev_hm, hm_host, hm_bytes, cudaMemcpyHostToDevice, s)), "H2D heat-map")) return r;
  if (int r = rtm3d_decode_main(dev_hm, d_off, d_off2, dtype, B, C, H, W, n_vert, K, thresh, down, cls, score, proj, verts,
                                bbox, flat, counts, ws, ws_bytes, flags, stream))
    return r;
  const size_t n = static_cast<size_t>(B) * K;
  struct { void* dst; const void* src; size_t bytes; } copies[] = {
      {counts_host, counts, static_cast<size_t>(B) * 4}, {cls_host, cls, n * 8}, {score_host, score, n * 4},
      {proj_host, proj, n * 8}, {verts_host, verts, n * static_cast<size_t>(n_vert) * 8}, {bbox_host, bbox, n * 16},
      {flat_host, flat, n * 4}};
  for (auto& c : copies) {
    if (!c.dst || !c.src) continue;
    if (int r = cuda_fail(static_cast<int>(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, s)), "D2H results")) return r;
  }
  return 0;
}

int rtm3d_decode_keypoints(const void* kpt_hm, const void* voff2, int dtype, int B, int Cv, int H, int W, int K,
                           float* kscore, float* kxy, int32_t* kflat, void* ws, size_t ws_bytes, unsigned flags,
                           void* stream) {
  if (!kpt_hm || !voff2 || !kscore || !kxy || !kflat || !ws) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, Cv, H, W, K)) return e;
  if (static_cast<long long>(K) > static_cast<long long>(H) * W) return fail(RTM3D_ERR_TOPK, "K=%d > H*W", K);
  const size_t es = elem_size(dtype);
  if (reinterpret_cast<uintptr_t>(kpt_hm) % es || reinterpret_cast<uintptr_t>(voff2) % es || reinterpret_cast<uintptr_t>(ws) % 256)
    return fail(RTM3D_ERR_ALIGN, "misaligned pointer");
  const rtm3d::WorkspaceLayout L = rtm3d::workspace_layout(B, Cv, H, W, K);
  if (ws_bytes < L.total) return fail(RTM3D_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, L.total);
  rtm3d::DecodeParams p{};
  p.hm = kpt_hm; p.off2 = voff2;
  p.B = B; p.C = Cv; p.H = H; p.W = W; p.n_vert = 0; p.K = K;
  p.thresh = 0.f; p.down = 1.f; p.t0 = -INFINITY;
  p.kscore = kscore; p.kxy = kxy; p.kflat = kflat;
  carve(p, L, ws);
  return dispatch(p, L, dtype, rtm3d::kModeKpt, flags, static_cast<cudaStream_t>(stream));
}

int rtm3d_decode_keypoints_host(const void* kpt_hm_host, const void* voff2_host, int dtype, int B, int Cv, int H, int W,
                                int K, void* dev_kpt, float* kscore, float* kxy, int32_t* kflat, float* kscore_host,
                                float* kxy_host, int32_t* kflat_host, void* ws, size_t ws_bytes, unsigned flags,
                                void* stream) {
  if (!kpt_hm_host || !voff2_host || !dev_kpt) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, Cv, H, W, K)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void* d_voff2 = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(&d_voff2, const_cast<void*>(voff2_host), 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(static_cast<int>(e), "voff2_host must be page-locked mapped host memory: %s", cudaGetErrorString(e));
  }
  const size_t hm_bytes = static_cast<size_t>(B) * Cv * H * W * elem_size(dtype);
  if (int r = cuda_fail(static_cast<int>(cudaMemcpyAsync(dev_kpt, kpt_hm_host, hm_bytes, cudaMemcpyHostToDevice, s)), "H2D keypoint heat-map")) return r;
  if (int r = rtm3d_decode_keypoints(dev_kpt, d_voff2, dtype, B, Cv, H, W, K, kscore, kxy, kflat, ws, ws_bytes, flags, stream)) return r;
  const size_t n = static_cast<size_t>(B) * Cv * K;
  struct { void* dst; const void* src; size_t bytes; } copies[] = {
      {kscore_host, kscore, n * 4}, {kxy_host, kxy, n * 8}, {kflat_host, kflat, n * 4}};
  for (auto& c : copies) {
    if (!c.dst) continue;
    if (int r = cuda_fail(static_cast<int>(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, s)), "D2H keypoint results")) return r;
  }
  return 0;
}

}  // extern "C"

namespace {
int decode_fused_impl(const void* hm, const void* off, const void* off2, const void* kpt_hm, const void* voff2, int dtype,
                      int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down, int64_t* cls,
                      float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts, float* kscore,
                      float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv,
                      void* ws, size_t ws_bytes, unsigned flags, void* stream, const GatherTarget* gt) {
  if (!hm || !off || !off2 || !kpt_hm || !voff2 || !cls || !score || !proj || !verts || !bbox || !flat || !counts || !kscore ||
      !kxy || !kflat || !kpt_proj || !kpt_score || !kpt_j || !ws)
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (Cv < 1) return fail(RTM3D_ERR_SHAPE, "Cv=%d", Cv);
  if (int e = check_shape(B, C, H, W, K)) return e;
  if (int e = check_shape(B, C + Cv, H, W, K)) return e;
  if (n_vert < 1 || n_vert > RTM3D_MAX_VERTS) return fail(RTM3D_ERR_SHAPE, "n_vert=%d outside [1,%d]", n_vert, RTM3D_MAX_VERTS);
  if (static_cast<long long>(K) > static_cast<long long>(H) * W) return fail(RTM3D_ERR_TOPK, "K=%d > H*W", K);
  if (!(thresh >= 0.0f)) return fail(RTM3D_ERR_THRESH, "score threshold must be >= 0 (got %g)", thresh);
  const rtm3d::WorkspaceLayout L = rtm3d::workspace_layout(B, C + Cv, H, W, K);
  if (ws_bytes < L.total) return fail(RTM3D_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, L.total);
  if (reinterpret_cast<uintptr_t>(ws) % 256) return fail(RTM3D_ERR_ALIGN, "workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bool fused = false;
  if (!(flags & RTM3D_FLAG_FORCE_GENERIC)) {
    unsigned char* base = static_cast<unsigned char*>(ws);
    rtm3d::PlaneParams q{};
    q.hm_main = hm; q.hm_kpt = kpt_hm; q.off = off; q.off2_main = off2; q.off2_kpt = voff2;
    q.B = B; q.C = C; q.Cv = Cv; q.H = H; q.W = W; q.n_vert = n_vert; q.K = K;
    q.thresh = thresh; q.down = down; q.t0 = prefilter_logit(thresh);
    q.cls = cls; q.score = score; q.proj = proj; q.verts = verts; q.bbox = bbox; q.flat = flat; q.counts = counts;
    q.kscore = kscore; q.kxy = kxy; q.kflat = kflat;
    q.tickets = reinterpret_cast<uint32_t*>(base + L.tickets_off);
    q.keys = reinterpret_cast<uint64_t*>(base + L.keys_off);
    q.key_counts = reinterpret_cast<uint32_t*>(base + L.counts_off);
    q.status = reinterpret_cast<uint32_t*>(base + L.status_off);
    q.retry = reinterpret_cast<uint32_t*>(base + L.retry_off);
    q.guess = reinterpret_cast<uint32_t*>(base + L.guess_off);
    q.ftable = reinterpret_cast<const float*>(base + L.table_off);
    int rc = -1000;
    if (!(flags & RTM3D_FLAG_LEGACY_PLANES)) {
      // scan kernel, then selection + epilogues + grouping in one kernel -- unless the caller wants the stages separately
      const bool one_kernel = !(flags & (RTM3D_FLAG_NO_EPILOGUE | RTM3D_FLAG_NO_GROUP)) && rtm3d::select_post_smem(Cv, K, n_vert) <= 200 * 1024;
      rtm3d::PostFusedParams f{flat, counts, kflat, kscore, off, off2, voff2, B, C, Cv, H, W, n_vert, K, down,
                               cls, proj, verts, bbox, kxy, kpt_proj, kpt_score, kpt_j, verts_cv};
      if (gt && !one_kernel) return fail(RTM3D_ERR_SHAPE, "rtm3d_decode_fused_gather: the stages cannot be separated");
      rc = scan_and_select(q, L, ws, dtype, flags, s, one_kernel ? &f : nullptr, gt);
      if (rc > 0 || (rc < 0 && rc != -1000)) return rc;
      if (rc == 0 && (one_kernel || (flags & RTM3D_FLAG_NO_SELECT))) return 0;
    }
    if (gt) return fail(RTM3D_ERR_SHAPE, "rtm3d_decode_fused_gather: the shape is not served by the scan + select kernels");
    if (rc == -1000) {
      if (flags & RTM3D_FLAG_NO_SELECT) return fail(RTM3D_ERR_SHAPE, "RTM3D_FLAG_NO_SELECT: the shape is not served by the scan kernel");
      rc = rtm3d::launch_planes(q, dtype, static_cast<int>((flags >> 8) & 0xFu), (flags & RTM3D_FLAG_NO_SPECULATION) ? 0 : 1,
                                static_cast<int>((flags >> 16) & 0xFFu), (flags & RTM3D_FLAG_LEGACY_PLANES) ? static_cast<int>((flags >> 24) & 0xFu) : 0, s);
      if (rc != -1000) if (int e = cuda_fail(rc, "decode (plane-streaming kernel, fused) launch")) return e;
    }
    if (rc != -1000) {
      fused = true;
      // everything after the selection in one kernel, unless the caller wants the stages separately (bench.py's marks)
      if (!(flags & (RTM3D_FLAG_NO_EPILOGUE | RTM3D_FLAG_NO_GROUP)) &&
          rtm3d::post_fused_smem(Cv, K, n_vert) <= 200 * 1024) {
        rtm3d::PostFusedParams f{flat, counts, kflat, kscore, off, off2, voff2, B, C, Cv, H, W, n_vert, K, down,
                                 cls, proj, verts, bbox, kxy, kpt_proj, kpt_score, kpt_j, verts_cv};
        return cuda_fail(rtm3d::launch_post_fused(f, dtype, s), "fused post kernel launch");
      }
      if (int e = launch_epilogues(q, dtype, flags, s)) return e;
    }
  }
  if (!fused) {
    // shapes the plane-streaming kernel does not take: the two separate entry points (generic kernels) on one workspace
    if (int e = rtm3d_decode_main(hm, off, off2, dtype, B, C, H, W, n_vert, K, thresh, down, cls, score, proj, verts, bbox, flat,
                                  counts, ws, ws_bytes, flags, stream))
      return e;
    // the two calls carve the workspace for different plane counts: what the first one left in the region the second
    // one uses for its tickets must be cleared (each call leaves its OWN tickets clean)
    const rtm3d::WorkspaceLayout Lk = rtm3d::workspace_layout(B, Cv, H, W, K);
    // (tickets only: the threshold table in front of them is written once by rtm3d_workspace_init and must survive)
    if (int e = cuda_fail(static_cast<int>(cudaMemsetAsync(static_cast<unsigned char*>(ws) + Lk.tickets_off, 0, Lk.status_off - Lk.tickets_off, s)), "workspace ticket reset")) return e;
    if (int e = rtm3d_decode_keypoints(kpt_hm, voff2, dtype, B, Cv, H, W, K, kscore, kxy, kflat, ws, ws_bytes, flags, stream))
      return e;
    const rtm3d::WorkspaceLayout Lm = rtm3d::workspace_layout(B, C, H, W, K);
    if (int e = cuda_fail(static_cast<int>(cudaMemsetAsync(static_cast<unsigned char*>(ws) + Lm.tickets_off, 0, Lm.status_off - Lm.tickets_off, s)), "workspace ticket reset")) return e;
  }
  if (flags & RTM3D_FLAG_NO_GROUP) return 0;
  return rtm3d_group_vertices(flat, counts, off, off2, dtype, B, H, W, n_vert, K, kscore, kxy, Cv, down, kpt_proj, kpt_score,
                              kpt_j, verts_cv, stream);
}

}  // namespace

extern "C" {

int rtm3d_decode_fused(const void* hm, const void* off, const void* off2, const void* kpt_hm, const void* voff2, int dtype,
                       int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down, int64_t* cls,
                       float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts, float* kscore,
                       float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv,
                       void* ws, size_t ws_bytes, unsigned flags, void* stream) {
  return decode_fused_impl(hm, off, off2, kpt_hm, voff2, dtype, B, C, Cv, H, W, n_vert, K, thresh, down, cls, score, proj, verts, bbox, flat,
                           counts, kscore, kxy, kflat, kpt_proj, kpt_score, kpt_j, verts_cv, ws, ws_bytes, flags, stream, nullptr);
}

int rtm3d_decode_fused_gather(const void* hm, const void* off, const void* off2, const void* kpt_hm, const void* voff2, int dtype,
                              int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down, int64_t* cls,
                              float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts, float* kscore,
                              float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv,
                              void* ws, size_t ws_bytes, unsigned flags, void* const* peer_wire, int n_peers, int rank,
                              unsigned step_id, void* stream) {
  if (!peer_wire || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers) return fail(RTM3D_ERR_SHAPE, "bad gather target (1..8 peers)");
  for (int r = 0; r < n_peers; ++r)
    if (!peer_wire[r] || reinterpret_cast<uintptr_t>(peer_wire[r]) % 4) return fail(RTM3D_ERR_NULL, "peer_wire[%d] is NULL or misaligned", r);
  // the arrival flags live behind the rows: word n_peers*B*(K*(9+2V)+1) of every gather buffer, one word per source rank
  const GatherTarget gt{peer_wire, n_peers, rank, step_id, static_cast<size_t>(n_peers) * B * (static_cast<size_t>(K) * (9 + 2 * n_vert) + 1), 0, nullptr, 0u};
  return decode_fused_impl(hm, off, off2, kpt_hm, voff2, dtype, B, C, Cv, H, W, n_vert, K, thresh, down, cls, score, proj, verts, bbox, flat,
                           counts, kscore, kxy, kflat, kpt_proj, kpt_score, kpt_j, verts_cv, ws, ws_bytes, flags, stream, &gt);
}

int rtm3d_decode_fused_gather_deferred(const void* hm, const void* off, const void* off2, const void* kpt_hm, const void* voff2, int dtype,
                                       int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down, int64_t* cls,
                                       float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts,
                                       float* kscore, float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score, int32_t* kpt_j,
                                       float* verts_cv, void* ws, size_t ws_bytes, unsigned flags, void* const* peer_wire,
                                       void* const* peer_wire_prev, int n_peers, int rank, unsigned prev_step_id, void* stream) {
  if (!peer_wire || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers) return fail(RTM3D_ERR_SHAPE, "bad gather target (1..8 peers)");
  for (int r = 0; r < n_peers; ++r) {
    if (!peer_wire[r] || reinterpret_cast<uintptr_t>(peer_wire[r]) % 4) return fail(RTM3D_ERR_NULL, "peer_wire[%d] is NULL or misaligned", r);
    if (peer_wire_prev && (!peer_wire_prev[r] || reinterpret_cast<uintptr_t>(peer_wire_prev[r]) % 4))
      return fail(RTM3D_ERR_NULL, "peer_wire_prev[%d] is NULL or misaligned", r);
    // the 16-byte body of a pushed block must be aligned at both ends: symmetric buffers share their low address bits
    if (peer_wire_prev && (reinterpret_cast<uintptr_t>(peer_wire_prev[r]) & 15u) != (reinterpret_cast<uintptr_t>(peer_wire_prev[rank]) & 15u))
      return fail(RTM3D_ERR_ALIGN, "gather buffers are not symmetric (different 16-byte phase)");
  }
  const GatherTarget gt{peer_wire, n_peers, rank, 0u, static_cast<size_t>(n_peers) * B * (static_cast<size_t>(K) * (9 + 2 * n_vert) + 1), 1,
                        peer_wire_prev, peer_wire_prev ? prev_step_id : 0u};
  return decode_fused_impl(hm, off, off2, kpt_hm, voff2, dtype, B, C, Cv, H, W, n_vert, K, thresh, down, cls, score, proj, verts, bbox, flat,
                           counts, kscore, kxy, kflat, kpt_proj, kpt_score, kpt_j, verts_cv, ws, ws_bytes, flags, stream, &gt);
}

int rtm3d_push_gather(void* const* peer_wire, int n_peers, int rank, int B, int K, int n_vert, unsigned step_id, void* stream) {
  if (!peer_wire || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers || B < 1 || K < 1 || n_vert < 0) return fail(RTM3D_ERR_SHAPE, "bad gather target");
  int32_t* peers[8] = {};
  for (int r = 0; r < n_peers; ++r) {
    if (!peer_wire[r] || reinterpret_cast<uintptr_t>(peer_wire[r]) % 4) return fail(RTM3D_ERR_NULL, "peer_wire[%d] is NULL or misaligned", r);
    if ((reinterpret_cast<uintptr_t>(peer_wire[r]) & 15u) != (reinterpret_cast<uintptr_t>(peer_wire[rank]) & 15u))
      return fail(RTM3D_ERR_ALIGN, "gather buffers are not symmetric (different 16-byte phase)");
    peers[r] = static_cast<int32_t*>(peer_wire[r]);
  }
  const size_t flag_offset = static_cast<size_t>(n_peers) * B * (static_cast<size_t>(K) * (9 + 2 * n_vert) + 1);
  return cuda_fail(rtm3d::launch_push_rows(peers, n_peers, rank, B, K, n_vert, step_id, flag_offset, static_cast<cudaStream_t>(stream)),
                   "push_gather launch");
}

int rtm3d_signal_gather(void* const* peer_wire, int n_peers, int rank, int B, int K, int n_vert, unsigned step_id, void* stream) {
  if (!peer_wire || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers || B < 1 || K < 1 || n_vert < 0) return fail(RTM3D_ERR_SHAPE, "bad gather target");
  int32_t* peers[8] = {};
  for (int r = 0; r < n_peers; ++r) {
    if (!peer_wire[r] || reinterpret_cast<uintptr_t>(peer_wire[r]) % 4) return fail(RTM3D_ERR_NULL, "peer_wire[%d] is NULL or misaligned", r);
    peers[r] = static_cast<int32_t*>(peer_wire[r]);
  }
  const size_t flag_offset = static_cast<size_t>(n_peers) * B * (static_cast<size_t>(K) * (9 + 2 * n_vert) + 1);
  return cuda_fail(rtm3d::launch_push_flag(peers, n_peers, rank, step_id, flag_offset, static_cast<cudaStream_t>(stream)), "signal_gather launch");
}

int rtm3d_wait_gather(const void* wire, int B, int K, int n_vert, int n_peers, unsigned step_id, void* stream) {
  if (!wire) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (n_peers < 1 || n_peers > 8 || B < 1 || K < 1) return fail(RTM3D_ERR_SHAPE, "bad gather shape");
  const size_t flag_offset = static_cast<size_t>(n_peers) * B * (static_cast<size_t>(K) * (9 + 2 * n_vert) + 1);
  return cuda_fail(rtm3d::launch_wait_flags(static_cast<const uint32_t*>(wire) + flag_offset, n_peers, step_id, static_cast<cudaStream_t>(stream)),
                   "wait_gather launch");
}

int rtm3d_epilogue_main(const int32_t* flat, const int32_t* counts, const void* off, const void* off2, int dtype, int B, int C,
                        int H, int W, int n_vert, int K, float down, int64_t* cls, float* proj, float* verts, float* bbox,
                        void* stream) {
  if (!flat || !counts || !off || !off2 || !cls || !proj || !verts || !bbox) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, C, H, W, K)) return e;
  if (n_vert < 1 || n_vert > RTM3D_MAX_VERTS) return fail(RTM3D_ERR_SHAPE, "n_vert=%d outside [1,%d]", n_vert, RTM3D_MAX_VERTS);
  rtm3d::EpiMainParams e{flat, counts, off, off2, B, C, H, W, n_vert, K, down, cls, proj, verts, bbox};
  return cuda_fail(rtm3d::launch_epilogue_main(e, dtype, static_cast<cudaStream_t>(stream)), "Tier A epilogue launch");
}

int rtm3d_epilogue_keypoints(const int32_t* kflat, const void* voff2, int dtype, int B, int Cv, int H, int W, int K, float* kxy,
                             void* stream) {
  if (!kflat || !voff2 || !kxy) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, Cv, H, W, K)) return e;
  rtm3d::EpiKptParams e{kflat, voff2, B, Cv, H, W, K, kxy};
  return cuda_fail(rtm3d::launch_epilogue_kpt(e, dtype, static_cast<cudaStream_t>(stream)), "Tier B epilogue launch");
}

int rtm3d_decode_fused_host(const void* hm_host, const void* off_host, const void* off2_host, const void* kpt_hm_host,
                            const void* voff2_host, int dtype, int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh,
                            float down, void* dev_hm, void* dev_kpt, int64_t* cls, float* score, float* proj, float* verts,
                            float* bbox, int32_t* flat, int32_t* counts, float* kscore, float* kxy, int32_t* kflat,
                            float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv, void* ws, size_t ws_bytes,
                            unsigned flags, void* stream) {
  if (!hm_host || !off_host || !off2_host || !kpt_hm_host || !voff2_host || !dev_hm || !dev_kpt)
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, C, H, W, K)) return e;
  if (Cv < 1) return fail(RTM3D_ERR_SHAPE, "Cv=%d", Cv);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void *d_off = nullptr, *d_off2 = nullptr, *d_voff2 = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(&d_off, const_cast<void*>(off_host), 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d_off2, const_cast<void*>(off2_host), 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d_voff2, const_cast<void*>(voff2_host), 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(static_cast<int>(e), "the regression maps must be page-locked mapped host memory: %s", cudaGetErrorString(e));
  }
  const size_t px = static_cast<size_t>(B) * H * W * elem_size(dtype);
  if (int r = cuda_fail(static_cast<int>(cudaMemcpyAsync(dev_hm, hm_host, px * C, cudaMemcpyHostToDevice, s)), "H2D main heat-map")) return r;
  if (int r = cuda_fail(static_cast<int>(cudaMemcpyAsync(dev_kpt, kpt_hm_host, px * Cv, cudaMemcpyHostToDevice, s)), "H2D keypoint heat-map")) return r;
  return rtm3d_decode_fused(dev_hm, d_off, d_off2, dev_kpt, d_voff2, dtype, B, C, Cv, H, W, n_vert, K, thresh, down, cls, score, proj,
                            verts, bbox, flat, counts, kscore, kxy, kflat, kpt_proj, kpt_score, kpt_j, verts_cv, ws, ws_bytes, flags,
                            stream);
}

int rtm3d_select_post(const void* off, const void* off2, const void* voff2, int dtype, int B, int C, int Cv, int H, int W, int n_vert,
                      int K, float thresh, float down, int64_t* cls, float* score, float* proj, float* verts, float* bbox,
                      int32_t* flat, int32_t* counts, float* kscore, float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score,
                      int32_t* kpt_j, float* verts_cv, void* ws, size_t ws_bytes, unsigned flags, void* stream) {
  if (!off || !off2 || !voff2 || !cls || !score || !proj || !verts || !bbox || !flat || !counts || !kscore || !kxy || !kflat ||
      !kpt_proj || !kpt_score || !kpt_j || !ws)
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (Cv < 1) return fail(RTM3D_ERR_SHAPE, "Cv=%d", Cv);
  if (int e = check_shape(B, C + Cv, H, W, K)) return e;
  if (n_vert < 1 || n_vert > RTM3D_MAX_VERTS) return fail(RTM3D_ERR_SHAPE, "n_vert=%d outside [1,%d]", n_vert, RTM3D_MAX_VERTS);
  if (!(thresh >= 0.0f)) return fail(RTM3D_ERR_THRESH, "score threshold must be >= 0 (got %g)", thresh);
  const rtm3d::WorkspaceLayout L = rtm3d::workspace_layout(B, C + Cv, H, W, K);
  if (ws_bytes < L.total) return fail(RTM3D_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, L.total);
  const int Sp = rtm3d::scan_strips_per_plane(H, W, K, dtype, static_cast<int>((flags >> 8) & 0xFu));
  if (Sp <= 0 || static_cast<long long>(B) * (C + Cv) * Sp > L.cand_strips || rtm3d::select_post_smem(Cv, K, n_vert) > 200 * 1024)
    return fail(RTM3D_ERR_SHAPE, "shape not served by the scan kernel: no candidate lists in the workspace");
  unsigned char* base = static_cast<unsigned char*>(ws);
  rtm3d::PostFusedParams f{flat, counts, kflat, kscore, off, off2, voff2, B, C, Cv, H, W, n_vert, K, down,
                           cls, proj, verts, bbox, kxy, kpt_proj, kpt_score, kpt_j, verts_cv};
  rtm3d::SelectPostParams q{reinterpret_cast<const unsigned long long*>(base + L.cand_off), reinterpret_cast<const uint32_t*>(base + L.cand_count_off),
                            Sp, rtm3d::scan_list_cap(K), thresh, score, flat, counts, kscore, kflat, f,
#ifdef RTM3D_DEV
                            rtm3d::debug_get_stats()
#else
                            nullptr
#endif
                            , {}, 0, 0, 0u, 0, nullptr, 0, {}, 0u};
  return cuda_fail(rtm3d::launch_select_post(q, dtype, static_cast<cudaStream_t>(stream)), "select + post kernel launch");
}

int rtm3d_post_fused(const int32_t* flat, const int32_t* counts, const int32_t* kflat, const float* kscore, const void* off,
                     const void* off2, const void* voff2, int dtype, int B, int C, int Cv, int H, int W, int n_vert, int K,
                     float down, int64_t* cls, float* proj, float* verts, float* bbox, float* kxy, float* kpt_proj,
                     float* kpt_score, int32_t* kpt_j, float* verts_cv, void* stream) {
  if (!flat || !counts || !kflat || !kscore || !off || !off2 || !voff2 || !cls || !proj || !verts || !bbox || !kxy || !kpt_proj ||
      !kpt_score || !kpt_j)
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, C, H, W, K)) return e;
  if (Cv < 1) return fail(RTM3D_ERR_SHAPE, "Cv=%d", Cv);
  if (n_vert < 1 || n_vert > RTM3D_MAX_VERTS) return fail(RTM3D_ERR_SHAPE, "n_vert=%d outside [1,%d]", n_vert, RTM3D_MAX_VERTS);
  if (rtm3d::post_fused_smem(Cv, K, n_vert) > 200 * 1024) return fail(RTM3D_ERR_SHAPE, "Cv*K too large for one CTA");
  rtm3d::PostFusedParams f{flat, counts, kflat, kscore, off, off2, voff2, B, C, Cv, H, W, n_vert, K, down,
                           cls, proj, verts, bbox, kxy, kpt_proj, kpt_score, kpt_j, verts_cv};
  return cuda_fail(rtm3d::launch_post_fused(f, dtype, static_cast<cudaStream_t>(stream)), "fused post kernel launch");
}

int rtm3d_group_vertices(const int32_t* flat, const int32_t* counts, const void* off, const void* off2, int dtype, int B,
                         int H, int W, int n_vert, int K, const float* kscore, const float* kxy, int Cv, float down,
                         float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv, void* stream) {
  if (!flat || !counts || !off || !off2 || !kscore || !kxy || !kpt_proj || !kpt_score || !kpt_j)
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, Cv, H, W, K)) return e;
  if (n_vert < 1 || n_vert > RTM3D_MAX_VERTS) return fail(RTM3D_ERR_SHAPE, "n_vert=%d", n_vert);
  if (static_cast<size_t>(Cv) * K * 8 > 200 * 1024) return fail(RTM3D_ERR_SHAPE, "Cv*K too large for one CTA");
  rtm3d::GroupParams g{flat, counts, off, off2, B, H, W, n_vert, K, Cv, kscore, kxy, down, kpt_proj, kpt_score, kpt_j, verts_cv};
  return cuda_fail(rtm3d::launch_group(g, dtype, static_cast<cudaStream_t>(stream)), "group_vertices launch");
}

int rtm3d_decode_box3d(const int32_t* flat, const int32_t* counts, const void* reg, int dtype, int B, int C, int H, int W,
                       int Creg, int K, int mode, const float* cam, const float* dim_ref, float depth_mu,
                       float depth_sigma, float* loc, float* dim, float* alpha, float* rot_y, float* corners2d,
                       void* stream) {
  if (!flat || !counts || !reg || !cam || !dim_ref || !loc || !dim || !alpha || !rot_y || !corners2d)
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (dtype != RTM3D_F32 && dtype != RTM3D_BF16) return fail(RTM3D_ERR_DTYPE, "dtype %d", dtype);
  if (int e = check_shape(B, C, H, W, K)) return e;
  const bool multibin = (mode & 1) != 0;
  if (Creg != (multibin ? 14 : 8)) return fail(RTM3D_ERR_SHAPE, "Creg=%d does not match mode %d (8 SMOKE-style, 14 multi-bin)", Creg, mode);
  rtm3d::Box3dParams q{flat, counts, reg, B, C, H, W, Creg, K, mode, cam, dim_ref, depth_mu, depth_sigma, loc, dim, alpha, rot_y, corners2d};
  return cuda_fail(rtm3d::launch_box3d(q, dtype, static_cast<cudaStream_t>(stream)), "box3d launch");
}

int rtm3d_fit_box3d(const float* verts, const int64_t* cls, const int32_t* counts, const float* cam, int cam_per_image,
                    const float* dim_ref, int n_classes, const float* ref_loc, int B, int K, int max_iter, float* loc, float* dim,
                    float* ry, float* fun, int32_t* accept, double* x8, int32_t* iters, void* stream) {
  if (!verts || !cls || !cam || !dim_ref || !ref_loc || !loc || !dim || !ry || !fun || !accept) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (B < 1 || K < 1 || n_classes < 1 || static_cast<long long>(B) * K > (1LL << 30)) return fail(RTM3D_ERR_SHAPE, "bad shape B=%d K=%d", B, K);
  if (max_iter < 1) max_iter = 100;
  rtm3d::BoxFitParams q{verts, cls, counts, cam, cam_per_image ? 1 : 0, dim_ref, {ref_loc[0], ref_loc[1], ref_loc[2]}, B, K, max_iter,
                        loc, dim, ry, fun, accept, x8, iters};
  return cuda_fail(rtm3d::launch_fit_box3d(q, static_cast<cudaStream_t>(stream)), "fit_box3d launch");
}

int rtm3d_encode_main_targets(const float* bbox, const int64_t* cls, const int64_t* img_id, const uint8_t* mask, const uint8_t* noise_mask,
                              int N, int B, int C, int H, int W, float* m_hm, int32_t* m_proj, float* m_off, float* sigma, int32_t* radius,
                              void* stream) {
  if (!m_hm || (N > 0 && (!bbox || !cls || !img_id || !mask || !noise_mask || !m_proj || !m_off || !sigma || !radius)))
    return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (N < 0 || B < 1 || C < 1 || H < 1 || W < 1) return fail(RTM3D_ERR_SHAPE, "bad shape N=%d B=%d C=%d H=%d W=%d", N, B, C, H, W);
  rtm3d::TargetParams q{bbox, cls, img_id, mask, noise_mask, N, B, C, H, W, m_hm, m_proj, m_off, sigma, radius};
  return cuda_fail(rtm3d::launch_encode_targets(q, static_cast<cudaStream_t>(stream)), "encode_targets launch");
}

int rtm3d_focal_loss(const float* logits, const float* target, size_t n, float alpha, float beta, double* acc, float* loss, void* stream) {
  if (!logits || !target || !acc || !loss) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (reinterpret_cast<uintptr_t>(acc) % 8) return fail(RTM3D_ERR_ALIGN, "acc must be 8-byte aligned");
  return cuda_fail(rtm3d::launch_focal_loss(logits, target, n, alpha, beta, acc, loss, static_cast<cudaStream_t>(stream)), "focal_loss launch");
}

int rtm3d_focal_loss_grad(const float* logits, const float* target, size_t n, float alpha, float beta, const double* acc,
                          const float* upstream, float* grad, void* stream) {
  if (!logits || !target || !acc || !grad) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  return cuda_fail(rtm3d::launch_focal_grad(logits, target, n, alpha, beta, acc, upstream, grad, static_cast<cudaStream_t>(stream)), "focal_loss_grad launch");
}

int rtm3d_gather_l1_loss(const float* map, int B, int C, int H, int W, const int64_t* img, const int64_t* x, const int64_t* y,
                         const int32_t* c0, const uint8_t* valid, const float* target, int n, int sigmoid, double* acc, float* loss,
                         void* stream) {
  if (!map || !acc || !loss || (n > 0 && (!img || !x || !y || !valid || !target))) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (B < 1 || C < 2 || H < 1 || W < 1 || n < 0) return fail(RTM3D_ERR_SHAPE, "bad shape B=%d C=%d H=%d W=%d n=%d", B, C, H, W, n);
  const rtm3d::GatherL1Params p{map, B, C, H, W, img, x, y, c0, valid, target, n, sigmoid ? 1 : 0, acc};
  return cuda_fail(rtm3d::launch_gather_l1(p, loss, static_cast<cudaStream_t>(stream)), "gather_l1_loss launch");
}

int rtm3d_gather_l1_loss_grad(const float* map, int B, int C, int H, int W, const int64_t* img, const int64_t* x, const int64_t* y,
                              const int32_t* c0, const uint8_t* valid, const float* target, int n, int sigmoid, const double* acc,
                              const float* upstream, float* grad, void* stream) {
  if (!map || !acc || !grad || (n > 0 && (!img || !x || !y || !valid || !target))) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (B < 1 || C < 2 || H < 1 || W < 1 || n < 0) return fail(RTM3D_ERR_SHAPE, "bad shape B=%d C=%d H=%d W=%d n=%d", B, C, H, W, n);
  const rtm3d::GatherL1Params p{map, B, C, H, W, img, x, y, c0, valid, target, n, sigmoid ? 1 : 0, const_cast<double*>(acc)};
  return cuda_fail(rtm3d::launch_gather_l1_grad(p, upstream, grad, static_cast<cudaStream_t>(stream)), "gather_l1_loss_grad launch");
}

int rtm3d_pack_wire(const int64_t* cls, const float* score, const float* proj, const float* verts, const float* bbox,
                    const int32_t* flat, const int32_t* counts, int B, int K, int n_vert, int32_t* wire, void* stream) {
  if (!cls || !score || !proj || !verts || !bbox || !flat || !counts || !wire) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (B < 1 || K < 1 || n_vert < 1 || n_vert > RTM3D_MAX_VERTS) return fail(RTM3D_ERR_SHAPE, "bad shape B=%d K=%d n_vert=%d", B, K, n_vert);
  return cuda_fail(rtm3d::launch_pack_wire(cls, score, proj, verts, bbox, flat, counts, B, K, n_vert, wire,
                                           static_cast<cudaStream_t>(stream)), "pack_wire launch");
}

int rtm3d_sigmoid_f32(const float* x, float* y, size_t n, void* stream) {
  if (!x || !y) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  return cuda_fail(rtm3d::launch_sigmoid(x, y, n, static_cast<cudaStream_t>(stream)), "sigmoid launch");
}

int rtm3d_threshold_table(float* logit_bound, uint32_t* score_edge_bits, int capacity, int* n_bins, void* stream) {
  if (!n_bins) return fail(RTM3D_ERR_NULL, "n_bins is NULL");
  *n_bins = rtm3d::threshold_table_bins();
  if (!logit_bound && !score_edge_bits) return 0;
  if (!logit_bound || !score_edge_bits) return fail(RTM3D_ERR_NULL, "NULL pointer argument");
  if (capacity < *n_bins) return fail(RTM3D_ERR_WORKSPACE, "capacity %d < %d bins", capacity, *n_bins);
  return cuda_fail(rtm3d::launch_threshold_table(logit_bound, score_edge_bits, static_cast<cudaStream_t>(stream)), "threshold table launch");
}

}  // extern "C"
