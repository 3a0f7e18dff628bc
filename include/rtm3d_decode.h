/*
 * rtm3d_decode.h -- C ABI of librtm3d_decode.so: the B200 (sm_100a) replacement of RTM3D's post-head
 * keypoint-heatmap decoder.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; every entry point below replaces a span of
 * reference Python (cited file:line, relative to hitfeelee/rtm3d) and is what a ctypes binding on the reference
 * side binds (INTEGRATION.md shows that stub).  Plain pointers and sizes only: no torch types cross this boundary.
 *
 * Conventions
 *  - all map pointers are DEVICE pointers to contiguous NCHW tensors (models/nets/header.py:40-46), dtype
 *    RTM3D_F32 or RTM3D_BF16 (bf16 is widened exactly to fp32 on load, then the fp32 pipeline is followed bit for bit);
 *  - outputs are fixed-size [B,K,...] device buffers; rows >= counts[b] are zero (cls/flat = -1);
 *  - rows are ordered (score desc, flat index asc) with flat = c*H*W + y*W + x -- the order torch.topk yields on
 *    CUDA (measured) and the one the reference decoder therefore produces in detect.py;
 *  - the library never allocates, frees, synchronises the host, or keeps pointers after return; work is enqueued
 *    on `stream` (a cudaStream_t passed as void*);
 *  - return 0 on success, a negative RTM3D_ERR_* on bad arguments, a positive cudaError_t on launch failure; the
 *    message is available from rtm3d_last_error() (thread-local).  No exceptions cross the ABI.
 */
#ifndef RTM3D_DECODE_H_
#define RTM3D_DECODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTM3D_ABI_VERSION 1
#define RTM3D_MAX_TOPK 1024
#define RTM3D_MAX_VERTS 16

enum rtm3d_dtype { RTM3D_F32 = 0, RTM3D_BF16 = 1 };

enum rtm3d_error {
  RTM3D_OK = 0,
  RTM3D_ERR_NULL = -1,      /* a required pointer is NULL */
  RTM3D_ERR_SHAPE = -2,     /* B,C,H,W,n_vert out of range */
  RTM3D_ERR_TOPK = -3,      /* K < 1, K > RTM3D_MAX_TOPK or K > C*H*W */
  RTM3D_ERR_ALIGN = -4,     /* a base pointer is not aligned to its element size */
  RTM3D_ERR_WORKSPACE = -5, /* ws_bytes smaller than rtm3d_decode_workspace_bytes() */
  RTM3D_ERR_DTYPE = -6,     /* unknown dtype */
  RTM3D_ERR_THRESH = -7,    /* threshold negative or NaN (0.0 fillers could then pass) */
  RTM3D_ERR_DEVICE = -8     /* no sm_100 device / kernel image unusable on the current device */
};

/* rtm3d_decode_main flags */
#define RTM3D_FLAG_FORCE_GENERIC 1u /* use the shape-generic strip kernels even when the plane-streaming kernel applies */
#define RTM3D_FLAG_NO_SPECULATION 2u /* plane-streaming kernel: never start a plane at the previous plane's threshold */
#define RTM3D_FLAG_NO_GROUP 4u /* rtm3d_decode_fused: stop after the two decodes (the caller runs rtm3d_group_vertices itself) */
#define RTM3D_FLAG_NO_EPILOGUE 8u /* decode entry points: stop after the selection (score, flat, counts / kscore, kflat); the
                                   caller runs rtm3d_epilogue_main / rtm3d_epilogue_keypoints itself */
#define RTM3D_FLAG_LEGACY_PLANES 16u /* use the round-1 plane-streaming kernel (histogram select inside the streaming CTA) instead of
                                       the plane-resident scan kernel + select kernel */
#define RTM3D_FLAG_NO_SELECT 32u /* rtm3d_decode_fused: stop after the scan kernel (candidate lists in the workspace); the caller
                                   continues with rtm3d_select_post on the same workspace */
#define RTM3D_FLAG_MAX_CTAS(n) (((unsigned)(n) & 0xFFu) << 16) /* plane-streaming kernel: at most n CTAs (0 = one per SM) */
#define RTM3D_FLAG_SPLIT(s) (((unsigned)(s) & 0xFu) << 8) /* plane-streaming kernel: force s strips (1,2,4,8) per plane; 0 = auto */
/* bits 24..27: developer timing experiments of the plane-streaming kernel (tools/debug_time.py; results are then WRONG);
 * must be 0 in production.  The rtm3d_debug_* symbols the library also exports are developer instrumentation, not ABI. */

int rtm3d_abi_version(void);
const char* rtm3d_last_error(void);
/* static string: compiler, arch and kernel variants built in */
const char* rtm3d_build_info(void);

/* Bytes of device scratch the decode entry points need for this shape: the scan kernel's per-strip candidate lists
 * (max(1024, 4K) 64-bit keys each) with their lengths, its strip-queue and completion counters, and -- for the shapes the
 * round-1 kernels serve -- the threshold table, keys of the per-strip top-K lists, per-image tickets and remembered
 * thresholds.  The scratch must be initialised ONCE with
 * rtm3d_workspace_init after allocation (zero fill + the table of logit bounds per score-histogram bin, 8 KB at the
 * start of the scratch); every call leaves it clean again (tickets are reset by the CTA that consumes them).  A scratch
 * that was only zeroed still gives identical results -- the kernels then compute the bounds themselves, more slowly.
 * After a call that returned a CUDA error the scratch must be initialised again. */
int rtm3d_decode_workspace_bytes(int B, int C, int H, int W, int K, size_t* out_bytes);
int rtm3d_workspace_init(void* ws, size_t ws_bytes, void* stream);

/*
 * Tier A -- replaces Model.inference for the main branch (models/model.py:29-75), i.e. _obtain_main_proj2d
 * (:77-98), utils/model_utils.py:17-26 nms_hm, _obtain_offset_fr_main (:117-132), the sub-pixel add (:48-50) and
 * the vertex regress / scale / 2D box (:63-73), for the whole batch in one enqueue.
 *
 *   hm     [B,C,H,W]        main_kf logits
 *   off    [B,2*n_vert,H,W] offset_fr_main (channel 2v = dx of vertex v, 2v+1 = dy; raw, no activation)
 *   off2   [B,2,H,W]        main_offset logits (sigmoid applied after the gather)
 * outputs (device):
 *   cls    int64 [B,K]          score f32 [B,K]          proj  f32 [B,K,2]   (x,y in input pixels)
 *   verts  f32  [B,K,n_vert,2]  bbox  f32 [B,K,4] (xmin,ymin,xmax,ymax)
 *   flat   int32 [B,K]  (may be NULL) flat peak index c*H*W + y*W + x
 *   counts int32 [B]    N_b = #{top-K scores > thresh}  (strict >, models/model.py:91)
 */
int rtm3d_decode_main(const void* hm, const void* off, const void* off2, int dtype,
                      int B, int C, int H, int W, int n_vert, int K, float thresh, float down,
                      int64_t* cls, float* score, float* proj, float* verts, float* bbox,
                      int32_t* flat, int32_t* counts,
                      void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * Selection only: the first half of rtm3d_decode_main (models/model.py:77-98 without the gathers): score f32 [B,K],
 * flat int32 [B,K] (c*H*W + y*W + x, -1 beyond counts[b]) and counts int32 [B].  For callers with their own epilogue behind
 * the peaks -- rtm3d_decode_box3d (depth / dimension / orientation regression, BASELINE configs[2]).
 */
int rtm3d_select_main(const void* hm, int dtype, int B, int C, int H, int W, int K, float thresh,
                      float* score, int32_t* flat, int32_t* counts, void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * Same computation for HOST-resident head outputs (the e2e path of bench.py).  `hm_host` is copied to `dev_hm`
 * (device staging of B*C*H*W elements) with one async copy; `off_host` / `off2_host` must be page-locked, mapped
 * host memory (cudaHostAlloc / cudaHostRegister): only the K*(2*n_vert+2) scalars per image that the decode needs
 * are read from them, by the GPU, over PCIe (zero-copy) -- the regression planes are never staged.  Outputs go to
 * device buffers as above and are then copied to the `*_host` pointers (page-locked) on the same stream.  The call
 * returns after enqueueing; the caller synchronises the stream.
 */
int rtm3d_decode_main_host(const void* hm_host, const void* off_host, const void* off2_host, int dtype,
                           int B, int C, int H, int W, int n_vert, int K, float thresh, float down,
                           void* dev_hm,
                           int64_t* cls, float* score, float* proj, float* verts, float* bbox,
                           int32_t* flat, int32_t* counts,
                           int64_t* cls_host, float* score_host, float* proj_host, float* verts_host,
                           float* bbox_host, int32_t* flat_host, int32_t* counts_host,
                           void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * Tier B (dormant in the reference) -- _obtain_vertex_proj2d (models/model.py:100-115) plus the commented sub-pixel
 * wiring (:52-60): per image AND per channel top-K over H*W, no threshold (0.0-score fillers in ascending index
 * order when a channel has fewer than K positive peaks).
 *   kpt_hm [B,Cv,H,W], voff2 [B,2,H,W]
 *   kscore f32 [B,Cv,K]   kxy f32 [B,Cv,K,2] (x + sigmoid(voff2[0]), y + sigmoid(voff2[1]); heat-map units, unscaled)
 *   kflat  int32 [B,Cv,K] (y*W + x)
 */
int rtm3d_decode_keypoints(const void* kpt_hm, const void* voff2, int dtype,
                           int B, int Cv, int H, int W, int K,
                           float* kscore, float* kxy, int32_t* kflat,
                           void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * rtm3d_decode_keypoints for HOST-resident maps, the Tier B sibling of rtm3d_decode_main_host: `kpt_hm_host` is copied
 * to the device staging buffer `dev_kpt` (B*Cv*H*W elements), `voff2_host` must be page-locked mapped host memory (only
 * Cv*K*2 scalars per image are read from it, zero-copy), results are copied to the page-locked `*_host` pointers (any
 * of which may be NULL) on the same stream.
 */
int rtm3d_decode_keypoints_host(const void* kpt_hm_host, const void* voff2_host, int dtype,
                                int B, int Cv, int H, int W, int K, void* dev_kpt,
                                float* kscore, float* kxy, int32_t* kflat,
                                float* kscore_host, float* kxy_host, int32_t* kflat_host,
                                void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * Tier A + Tier B in ONE enqueue -- the full Model.inference with the commented keypoint wiring restored
 * (models/model.py:45-62, 68-69): rtm3d_decode_main on (hm, off, off2), rtm3d_decode_keypoints on (kpt_hm, voff2) and
 * rtm3d_group_vertices on their results.  Both heat-maps are streamed by a single launch of the plane-streaming
 * kernel (C + Cv planes per image share the GPU), followed by the grouping kernel.  Arguments as in the three entry
 * points; the workspace must be sized with rtm3d_decode_workspace_bytes(B, C + Cv, H, W, K).
 */
int rtm3d_decode_fused(const void* hm, const void* off, const void* off2, const void* kpt_hm, const void* voff2, int dtype,
                       int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down,
                       int64_t* cls, float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts,
                       float* kscore, float* kxy, int32_t* kflat,
                       float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv,
                       void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * rtm3d_decode_fused for HOST-resident head outputs: both heat-maps are copied to their device staging buffers
 * (`dev_hm` B*C*H*W and `dev_kpt` B*Cv*H*W elements), the three regression maps must be page-locked mapped host memory
 * and are read zero-copy (K*(2V+2) + Cv*K*2 scalars per image).  Results land in the device buffers; copying them back
 * is left to the caller (one D2H per buffer on the same stream).
 */
int rtm3d_decode_fused_host(const void* hm_host, const void* off_host, const void* off2_host, const void* kpt_hm_host,
                            const void* voff2_host, int dtype, int B, int C, int Cv, int H, int W, int n_vert, int K,
                            float thresh, float down, void* dev_hm, void* dev_kpt,
                            int64_t* cls, float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts,
                            float* kscore, float* kxy, int32_t* kflat,
                            float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv,
                            void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * Second halves of rtm3d_decode_main / rtm3d_decode_keypoints, enqueued by them unless RTM3D_FLAG_NO_EPILOGUE is set:
 *   rtm3d_epilogue_main       rows (flat, counts) -> cls, proj, verts, bbox: gather of offset_fr_main / main_offset at the
 *                             integer peak, sub-pixel add, vertex regress, x DOWN_SAMPLE, 2D box (models/model.py:47-50,
 *                             63-73, 117-132)
 *   rtm3d_epilogue_keypoints  kflat -> kxy: index split + sigmoid sub-pixel add (models/model.py:113-114, 55-57)
 * One thread per (detection, vertex) / per candidate over the whole batch.
 */
int rtm3d_epilogue_main(const int32_t* flat, const int32_t* counts, const void* off, const void* off2, int dtype,
                        int B, int C, int H, int W, int n_vert, int K, float down,
                        int64_t* cls, float* proj, float* verts, float* bbox, void* stream);
int rtm3d_epilogue_keypoints(const int32_t* kflat, const void* voff2, int dtype, int B, int Cv, int H, int W, int K,
                             float* kxy, void* stream);

/*
 * Everything that follows the selection of rtm3d_decode_fused in ONE kernel (a cluster of CTAs per image): rtm3d_epilogue_keypoints,
 * rtm3d_epilogue_main and rtm3d_group_vertices with bit-identical results.  rtm3d_decode_fused enqueues it itself unless
 * RTM3D_FLAG_NO_EPILOGUE / RTM3D_FLAG_NO_GROUP ask for the stages separately.
 */
int rtm3d_post_fused(const int32_t* flat, const int32_t* counts, const int32_t* kflat, const float* kscore,
                     const void* off, const void* off2, const void* voff2, int dtype,
                     int B, int C, int Cv, int H, int W, int n_vert, int K, float down,
                     int64_t* cls, float* proj, float* verts, float* bbox, float* kxy,
                     float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv, void* stream);

/*
 * rtm3d_decode_fused + the path's one exchange step (the gather of the fixed-size detections), fused: the
 * CTAs that compute an image's Tier A rows store its wire rows -- K rows of (cls | score | proj 2 | verts 2V | bbox 4 | flat)
 * 32-bit words followed by the image's count, the layout of rtm3d_pack_wire -- straight into the gather buffer of EVERY
 * rank over NVLink.  peer_wire[r] = rank r's gather buffer int32 [n_peers * B, K*(9+2V) + 1] as mapped into THIS process
 * (peer-to-peer / symmetric memory; peer_wire[rank] is the local buffer); this rank's images land in rows
 * [rank*B, (rank+1)*B) of each.  No second kernel, no NCCL call on the data path; the caller synchronises the ranks before it
 * reads a buffer or lets it be overwritten.  Shapes the scan + select kernels do not serve are rejected (RTM3D_ERR_SHAPE).
 */
int rtm3d_decode_fused_gather(const void* hm, const void* off, const void* off2, const void* kpt_hm, const void* voff2, int dtype,
                              int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down,
                              int64_t* cls, float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts,
                              float* kscore, float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score, int32_t* kpt_j,
                              float* verts_cv, void* ws, size_t ws_bytes, unsigned flags,
                              void* const* peer_wire, int n_peers, int rank, unsigned step_id, void* stream);
/*
 * Arrival flags of the fused gather: every gather buffer carries n_peers 32-bit words behind its rows (so a buffer has
 * n_peers*B*(K*(9+2V)+1) + n_peers words).  With step_id != 0 the last CTA of rtm3d_decode_fused_gather's launch stores
 * step_id into word [rank] of every peer's flag array once all rows of the batch are visible there.  rtm3d_wait_gather
 * enqueues a one-thread kernel that returns when all n_peers flags of the LOCAL buffer `wire` have reached step_id
 * (wrap-safe comparison; bounded wait): behind it the batch of every rank can be read.  Step ids must increase.
 */
int rtm3d_wait_gather(const void* wire, int B, int K, int n_vert, int n_peers, unsigned step_id, void* stream);
/*
 * Deferred variant of the fused gather, for a stream of batches over two (or more) slots of the gather buffers.  The rows of
 * a batch leave a rank at the END of its select + post kernel; stored to n_peers ranks there they are one burst over NVLink
 * that nothing overlaps (measured at 8 GPUs: +31 us on a 127 us step).  Here the launch stores its rows into THIS rank's
 * buffer only (peer_wire[rank]) and, at its START, pushes the rows of the previous batch -- read back from
 * peer_wire_prev[rank] -- into peer_wire_prev[r] of every other rank: posted stores that drain while the kernel sorts.  At
 * its end it raises the PREVIOUS batch's arrival flag (prev_step_id) in every rank's peer_wire_prev buffer.  peer_wire_prev
 * NULL (or prev_step_id 0): nothing to push (first batch).  The gather buffers must be symmetric (same 16-byte phase).
 * rtm3d_push_gather flushes the last batch: its rows from peer_wire[rank] to the other ranks (one small kernel), then its
 * flag (a second launch).  rtm3d_wait_gather is unchanged.
 */
int rtm3d_decode_fused_gather_deferred(const void* hm, const void* off, const void* off2, const void* kpt_hm, const void* voff2, int dtype,
                                       int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down,
                                       int64_t* cls, float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts,
                                       float* kscore, float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score, int32_t* kpt_j,
                                       float* verts_cv, void* ws, size_t ws_bytes, unsigned flags,
                                       void* const* peer_wire, void* const* peer_wire_prev, int n_peers, int rank, unsigned prev_step_id,
                                       void* stream);
int rtm3d_push_gather(void* const* peer_wire, int n_peers, int rank, int B, int K, int n_vert, unsigned step_id, void* stream);
/*
 * The arrival flag alone: stores step_id into word [rank] of every gather buffer's flag array (one tiny kernel).  For callers
 * that move the rows themselves -- e.g. with the copy engines (cudaMemcpyAsync to the peer-mapped buffers on a second stream,
 * behind rtm3d_decode_fused_gather_deferred with peer_wire_prev = NULL, which leaves the rows in this rank's buffer): stream
 * order puts the flag behind the copies, and no SM time or load/store bandwidth of the decode kernels is spent on NVLink.
 */
int rtm3d_signal_gather(void* const* peer_wire, int n_peers, int rank, int B, int K, int n_vert, unsigned step_id, void* stream);

/*
 * Second half of rtm3d_decode_fused when it was called with RTM3D_FLAG_NO_SELECT (it then stops after the scan kernel, whose
 * per-strip candidate lists stay in the workspace): merges and sorts the lists of every selection problem -- the flat top-K
 * over C*H*W of models/model.py:87-98 and the per-channel top-K of :109-114 -- and runs everything rtm3d_post_fused does, in
 * ONE kernel (a cluster of four CTAs per image).  Same arguments (and the same `flags`: the strips-per-plane override must
 * match) as the rtm3d_decode_fused call that filled `ws`; the selection outputs score / flat / counts / kscore / kflat are
 * written here.  rtm3d_decode_fused enqueues this kernel itself unless a flag asks for the stages separately.
 */
int rtm3d_select_post(const void* off, const void* off2, const void* voff2, int dtype,
                      int B, int C, int Cv, int H, int W, int n_vert, int K, float thresh, float down,
                      int64_t* cls, float* score, float* proj, float* verts, float* bbox, int32_t* flat, int32_t* counts,
                      float* kscore, float* kxy, int32_t* kflat, float* kpt_proj, float* kpt_score, int32_t* kpt_j,
                      float* verts_cv, void* ws, size_t ws_bytes, unsigned flags, void* stream);

/*
 * Tier B -- _group_vertexs_kf (models/model.py:134-162): for every detection n of rtm3d_decode_main and keypoint
 * channel k, j* = argmin_j ||(v_kj - m_n) - off_kn||^2 over the K candidates (first minimal j), with m_n and off_kn
 * recomputed from `flat` exactly as in Tier A; channels k >= n_vert use a zero offset.
 *   kpt_proj f32 [B,K,Cv,2] = down * v_kj*    kpt_score f32 [B,K,Cv]    kpt_j int32 [B,K,Cv]
 *   verts_cv f32 [B,K,Cv,2] = down * (off_kn + m_n)   (may be NULL)
 */
int rtm3d_group_vertices(const int32_t* flat, const int32_t* counts,
                         const void* off, const void* off2, int dtype,
                         int B, int H, int W, int n_vert, int K,
                         const float* kscore, const float* kxy, int Cv, float down,
                         float* kpt_proj, float* kpt_score, int32_t* kpt_j, float* verts_cv, void* stream);

/*
 * Tier C (NOT in the reference: parity unpinned, the normative text is oracle/box3d_ref.py) -- gather of a
 * depth/dimension/orientation regression map at the Tier A peaks and closed-form 3D box recovery with the
 * reference's geometric conventions (utils/model_utils.py:66-76 rotation_matrix, :80-119 create_corners,
 * :147-152 calc_proj_corners; flat-9 row-major camera matrix, datasets/dataset_reader.py:108).
 *   reg [B,Creg,H,W] with Creg = 8 (SMOKE style: depth 1, sub-pixel 2, dims 3, sin/cos 2; `mode` 0) or
 *   Creg = 14 (multi-bin: depth 1, sub-pixel 2, dims 3, bins 8; `mode` 1)
 *   cam [B,9] (already divided by `down` on entries 0..5 if it is to act on heat-map coordinates)
 *   dim_ref [C,3] rows (h,w,l); depth_ref = (mu_z, sigma_z)
 * outputs: loc f32 [B,K,3], dim f32 [B,K,3] (h,w,l), alpha f32 [B,K], rot_y f32 [B,K], corners2d f32 [B,K,8,2]
 */
int rtm3d_decode_box3d(const int32_t* flat, const int32_t* counts, const void* reg, int dtype,
                       int B, int C, int H, int W, int Creg, int K, int mode,
                       const float* cam, const float* dim_ref, float depth_mu, float depth_sigma,
                       float* loc, float* dim, float* alpha, float* rot_y, float* corners2d, void* stream);

/*
 * Batched 3D-box fit: the step right behind the decoder in detect.py (:71-74), replacing optim_decode_bbox3d
 * (utils/model_utils.py:264-312: one scipy L-BFGS-B run per object).  Per detection the reference's reprojection objective
 * (aimFun :155-177) is minimised from the reference's start point [0, 1, l_ref, h_ref, w_ref] + ref_loc (:290) by
 * Levenberg-Marquardt in double precision (analytic Jacobian, the derivatives of :206-234) and accepted when f < 0.1 (:298).
 *   verts  f32 [B,K,8,2]  v_projs_regress (input pixels)      cls int64 [B,K]      counts int32 [B] (NULL: all K rows)
 *   cam    f32 [B,9] (cam_per_image != 0) or [9]: row-major camera matrix      dim_ref f32 [n_classes,3] (h,w,l)
 *   ref_loc: 3 floats in HOST memory (detect.py:74 passes 0, -0.5, 20)
 * Outputs (rows >= counts[b]: zeros, accept 0):
 *   loc f32 [B,K,3]   dim f32 [B,K,3] = (h,w,l) (:301)   ry f32 [B,K] = atan2(sin, cos) (:299)   fun f32 [B,K]
 *   accept int32 [B,K] = (fun < 0.1)   x8 f64 [B,K,8] raw solution [sin,cos,l,h,w,X,Y,Z] (may be NULL)   iters int32 [B,K] (may be NULL)
 * The objective leaves a two-parameter family of minimisers (rotation-vector length; common scale of dimensions and
 * location): dim / loc are reported for the member with sin^2+cos^2 = 1 whose dimensions are closest to the class prior;
 * fun, accept, ry, the reprojected corners and all ratios of (l,h,w,X,Y,Z) are independent of that choice.
 */
int rtm3d_fit_box3d(const float* verts, const int64_t* cls, const int32_t* counts, const float* cam, int cam_per_image,
                    const float* dim_ref, int n_classes, const float* ref_loc, int B, int K, int max_iter,
                    float* loc, float* dim, float* ry, float* fun, int32_t* accept, double* x8, int32_t* iters, void* stream);

/*
 * Training-side mirror of the decoder.
 * rtm3d_encode_main_targets: the m_hm part of DatasetReader._build_targets (datasets/dataset_reader.py:215-291) with
 *   data_utils.dynamic_radius / gaussian2D (utils/data_utils.py:97-141): per labelled object (mask != 0) the centre of its 2D box
 *   (heat-map units), a Gaussian of the box's radius splatted with max into plane (img_id, cls); the centre of a noise object
 *   is 0.9999 (:262-263).  m_hm f32 [B,C,H,W] is zeroed and filled here; m_proj int32 [N,2], m_off f32 [N,2] (:225-228),
 *   sigma f32 [N], radius int32 [N] are the per-object by-products the reader keeps.
 * rtm3d_focal_loss: FocalLoss.forward (models/nets/module.py:41-68) on sigmoid_hm(logits) (utils/model_utils.py:10-14), the
 *   main heat-map loss of models/rtm3d_loss.py:283: loss f32 [1]; acc = 3 doubles of device scratch (positive sum, negative
 *   sum, number of positives) that rtm3d_focal_loss_grad reads.
 * rtm3d_focal_loss_grad: grad[i] = upstream * d loss / d logits[i] (what autograd yields for the reference); upstream = device
 *   pointer to the incoming gradient of the loss (NULL: 1).
 */
int rtm3d_encode_main_targets(const float* bbox, const int64_t* cls, const int64_t* img_id, const uint8_t* mask,
                              const uint8_t* noise_mask, int N, int B, int C, int H, int W, float* m_hm, int32_t* m_proj,
                              float* m_off, float* sigma, int32_t* radius, void* stream);
int rtm3d_focal_loss(const float* logits, const float* target, size_t n, float alpha, float beta, double* acc, float* loss,
                     void* stream);
int rtm3d_focal_loss_grad(const float* logits, const float* target, size_t n, float alpha, float beta, const double* acc,
                          const float* upstream, float* grad, void* stream);

/*
 * The gather-L1 losses of RTM3DLoss.__call__ (models/rtm3d_loss.py:302-330: offset-from-main, vertex offset, main offset):
 * per entry e the channels c0[e], c0[e]+1 (c0 NULL: 0, 1) of the NCHW f32 map at (img[e], y[e], x[e]), through a sigmoid when
 * `sigmoid` != 0 (:319, :327), against target[e][0..1]; loss = mean |pred - target| over the entries with valid[e] != 0
 * (F.l1_loss, reduction 'mean'; NaN when none is valid, as torch).  The reference permutes the whole map to NHWC first
 * (:302, :316, :324: a full copy per loss); this reads the 2 n scalars.  acc: double[3] scratch (sum, elements, valid entries
 * outside the map -- skipped, torch would raise), kept for the gradient call.
 * rtm3d_gather_l1_loss_grad: grad f32 [B,C,H,W] = d(upstream * loss)/d map (zero-filled, then accumulated); upstream NULL = 1.
 */
int rtm3d_gather_l1_loss(const float* map, int B, int C, int H, int W, const int64_t* img, const int64_t* x, const int64_t* y,
                         const int32_t* c0, const uint8_t* valid, const float* target, int n, int sigmoid, double* acc, float* loss,
                         void* stream);
int rtm3d_gather_l1_loss_grad(const float* map, int B, int C, int H, int W, const int64_t* img, const int64_t* x, const int64_t* y,
                              const int32_t* c0, const uint8_t* valid, const float* target, int n, int sigmoid, const double* acc,
                              const float* upstream, float* grad, void* stream);

/*
 * Packs the Tier A result of a batch into the wire rows of the multi-GPU gather (the path's one collective):
 * wire int32 [B][K*(9+2*n_vert) + 1] = per image K rows of (cls | score | proj 2 | verts 2*n_vert | bbox 4 | flat) as
 * 32-bit patterns, then counts[b].  One launch instead of a chain of torch cat / cast kernels.
 */
int rtm3d_pack_wire(const int64_t* cls, const float* score, const float* proj, const float* verts, const float* bbox,
                    const int32_t* flat, const int32_t* counts, int B, int K, int n_vert, int32_t* wire, void* stream);

/*
 * The library's sigmoid s(x) = 1.0f / (1.0f + expf(-x)) applied element-wise to n device floats: the function every
 * score of this library goes through (models/model.py:48,85,107 `sigmoid_`).  Exposed so that tests can sweep all 2^32
 * inputs: bit-identical to torch's CUDA sigmoid, and monotone (the peak test relies on it).
 */
int rtm3d_sigmoid_f32(const float* x, float* y, size_t n, void* stream);

/*
 * Verification aid for the plane-streaming kernel's running threshold: for every bin b of its score histogram,
 * score_edge_bits[b] = fp32 bits of the bin's lower score edge and logit_bound[b] = the logit T the scan uses for it,
 * with the contract  x < T  =>  rtm3d_sigmoid_f32(x) < edge  (bin 0: no bound, T = -inf).  *n_bins is always set; the
 * device arrays (capacity >= *n_bins entries) may both be NULL to query the size only.
 */
int rtm3d_threshold_table(float* logit_bound, uint32_t* score_edge_bits, int capacity, int* n_bins, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RTM3D_DECODE_H_ */
