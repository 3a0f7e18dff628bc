#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtm3d {

struct GroupParams {
  const int32_t* flat; const int32_t* counts;
  const void* off; const void* off2;
  int B, H, W, n_vert, K, Cv;
  const float* kscore; const float* kxy;
  float down;
  float* kpt_proj; float* kpt_score; int32_t* kpt_j; float* verts_cv;
};

struct Box3dParams {
  const int32_t* flat; const int32_t* counts; const void* reg;
  int B, C, H, W, Creg, K, mode;
  const float* cam; const float* dim_ref;
  float depth_mu, depth_sigma;
  float* loc; float* dim; float* alpha; float* rot_y; float* corners2d;
};

// Tier A epilogue after the plane-streaming kernel: rows (flat, counts) -> cls, proj, verts, bbox.
struct EpiMainParams {
  const int32_t* flat; const int32_t* counts; const void* off; const void* off2;
  int B, C, H, W, n_vert, K;
  float down;
  int64_t* cls; float* proj; float* verts; float* bbox;
};
// Tier B epilogue after the plane-streaming kernel: candidate index -> sub-pixel position.
struct EpiKptParams {
  const int32_t* kflat; const void* voff2;
  int B, Cv, H, W, K;
  float* kxy;
};
int launch_epilogue_main(const EpiMainParams& p, int dtype, cudaStream_t s);
int launch_epilogue_kpt(const EpiKptParams& p, int dtype, cudaStream_t s);
// Everything after the selection of a fused decode in ONE kernel (a cluster of CTAs per image): Tier B epilogue, Tier A epilogue
// and the keypoint grouping.
struct PostFusedParams {
  const int32_t* flat; const int32_t* counts; const int32_t* kflat; const float* kscore;
  const void* off; const void* off2; const void* voff2;
  int B, C, Cv, H, W, n_vert, K;
  float down;
  int64_t* cls; float* proj; float* verts; float* bbox;      // Tier A
  float* kxy;                                                 // Tier B candidates
  float* kpt_proj; float* kpt_score; int32_t* kpt_j; float* verts_cv;   // grouping (verts_cv may be null)
};
// Selection + everything after it in ONE kernel (a cluster of four CTAs per image): the candidate lists of the scan kernel
// are merged, converted to scores and sorted here (select_common.cuh), then the Tier B / Tier A epilogues and the grouping.
struct SelectPostParams {
  const unsigned long long* cand; const uint32_t* cand_count;
  int Sp, list_cap;
  float thresh;
  float* score; int32_t* flat; int32_t* counts; float* kscore; int32_t* kflat;     // the selection (written here)
  PostFusedParams post;                                                             // maps, shapes, outputs of the epilogues
  unsigned long long* stats;                                                        // developer timestamps (nullable)
  // fused gather (rtm3d_decode_fused_gather): the wire rows of this rank's detections are stored straight into every
  // peer's gather buffer (peer-mapped pointers over NVLink) by the CTAs that compute them; n_peers = 0: no gather
  int32_t* wire_peers[8];
  int n_peers, wire_rank;
  // ... and, when step_id != 0, the arrival flag: the last CTA of the launch stores step_id into word [wire_rank] of every
  // peer's flag array behind the rows (the flag array starts at word flag_offset of each gather buffer)
  uint32_t step_id;
  size_t flag_offset;
  uint32_t* done_counter;      // workspace word, 0 between launches
  // deferred gather (rtm3d_decode_fused_gather_deferred): this launch stores its rows into THIS rank's buffer only and, at its
  // start, pushes the rows of the previous batch (push_peers = the previous batch's slot of every gather buffer) to the other
  // ranks -- posted stores that drain while the kernel sorts, instead of a burst at its end; the arrival flag it raises at
  // its end is the previous batch's (push_step; 0 = nothing to push)
  int deferred;
  int32_t* push_peers[8];
  uint32_t push_step;
};
// Batched 3D-box fit behind the decoder (boxfit.cu; utils/model_utils.py:264-312).
struct BoxFitParams {
  const float* verts;      // [B,K,8,2] regressed vertices, input pixels (v_projs_regress)
  const int64_t* cls;      // [B,K]
  const int32_t* counts;   // [B] valid detections per image (nullptr: all K)
  const float* cam;        // [B,9] or [1,9] row-major camera matrix
  int cam_per_image;
  const float* dim_ref;    // [C,3] class priors (h, w, l)
  float ref_loc[3];        // start location (detect.py:74: 0, -0.5, 20)
  int B, K, max_iter;
  float* loc; float* dim; float* ry; float* fun; int32_t* accept;   // [B,K,3] [B,K,3] (h,w,l) [B,K] [B,K] [B,K]
  double* x8;              // [B,K,8] raw solution vector (nullable)
  int32_t* iters;          // [B,K] iterations used (nullable)
};
int launch_fit_box3d(const BoxFitParams& p, cudaStream_t s);
// Training-side mirror (train_side.cu): main heat-map target encoder and focal loss.
struct TargetParams {
  const float* bbox;       // [N,4] 2D boxes x1,y1,x2,y2 in HEAT-MAP units (targets' bbox / DOWN_SAMPLE)
  const int64_t* cls;      // [N]
  const int64_t* img_id;   // [N] image of the batch
  const uint8_t* mask;     // [N] labelled main point
  const uint8_t* noise_mask;   // [N]
  int N, B, C, H, W;
  float* m_hm;             // [B,C,H,W] out (zeroed here)
  int32_t* m_proj;         // [N,2] integer centre (x,y)
  float* m_off;            // [N,2] sub-pixel offset of the centre
  float* sigma;            // [N]
  int32_t* radius;         // [N]
};
int launch_encode_targets(const TargetParams& p, cudaStream_t s);
// gather-L1 losses (models/rtm3d_loss.py:302-330): see train_side.cu
struct GatherL1Params {
  const float* map; int B, C, H, W;
  const int64_t* img; const int64_t* x; const int64_t* y; const int32_t* c0; const uint8_t* valid; const float* target;
  int n, sigmoid;
  double* acc;
};
int launch_gather_l1(const GatherL1Params& p, float* loss, cudaStream_t s);
int launch_gather_l1_grad(const GatherL1Params& p, const float* upstream, float* grad, cudaStream_t s);
int launch_focal_loss(const float* logits, const float* target, size_t n, float alpha, float beta, double* acc, float* loss, cudaStream_t s);
int launch_focal_grad(const float* logits, const float* target, size_t n, float alpha, float beta, const double* acc, const float* upstream,
                      float* grad, cudaStream_t s);
// flush of the deferred gather: the rows of one batch from this rank's buffer to the other ranks, then its arrival flag
int launch_push_rows(int32_t* const* peers, int n_peers, int rank, int B, int K, int n_vert, uint32_t step_id, size_t flag_offset, cudaStream_t s);
int launch_push_flag(int32_t* const* peers, int n_peers, int rank, uint32_t step_id, size_t flag_offset, cudaStream_t s);
int launch_wait_flags(const uint32_t* flags, int n, uint32_t value, cudaStream_t s);
size_t select_post_smem(int Cv, int K, int n_vert);
int launch_select_post(const SelectPostParams& p, int dtype, cudaStream_t s);
size_t post_fused_smem(int Cv, int K, int n_vert);
int launch_post_fused(const PostFusedParams& p, int dtype, cudaStream_t s);
int launch_group(const GroupParams& p, int dtype, cudaStream_t s);
int launch_box3d(const Box3dParams& p, int dtype, cudaStream_t s);
int launch_pack_wire(const int64_t* cls, const float* score, const float* proj, const float* verts, const float* bbox,
                     const int32_t* flat, const int32_t* counts, int B, int K, int V, int32_t* wire, cudaStream_t s);
int launch_sigmoid(const float* x, float* y, size_t n, cudaStream_t s);

}  // namespace rtm3d
