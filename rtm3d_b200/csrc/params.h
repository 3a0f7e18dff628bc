// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rtm3d {

enum Mode { kModeMain = 0, kModeKpt = 1 };

struct DecodeParams {
  // maps (device or mapped-host pointers), NCHW contiguous
  const void* hm;    // [B,C,H,W] heat-map logits (main_kf, or the keypoint heat-map in kModeKpt)
  const void* off;   // [B,2*n_vert,H,W]   (kModeMain)
  const void* off2;  // [B,2,H,W] sub-pixel logits (main_offset in kModeMain, vertex_offset in kModeKpt)
  int B, C, H, W, n_vert, K;
  float thresh, down;
  float t0;          // logit-domain prefilter derived from thresh (x < t0 => sigmoid(x) <= thresh); -inf in kModeKpt
  // generic strip kernels
  int strip_rows, nstrips, list_cap;
  // Tier A outputs
  int64_t* cls; float* score; float* proj; float* verts; float* bbox; int32_t* flat; int32_t* counts;
  // Tier B outputs
  float* kscore; float* kxy; int32_t* kflat;
  // workspace
  uint32_t* tickets;     // [B*C] zero between calls
  uint64_t* keys;        // [B*C*nstrips][K]
  uint32_t* key_counts;  // [B*C*nstrips]
  uint32_t* status;      // watchdog word of the streaming kernel (0 = ok)
};

// Parameters of the persistent plane-streaming kernel (decode_planes.cu).  Either heat-map may be absent (C or Cv = 0).
struct PlaneParams {
  const void* hm_main;   // [B,C,H,W]  main_kf logits            (nullptr when C == 0)
  const void* hm_kpt;    // [B,Cv,H,W] keypoint heat-map logits  (nullptr when Cv == 0)
  const void* off;       // [B,2*n_vert,H,W]
  const void* off2_main; // [B,2,H,W] main_offset
  const void* off2_kpt;  // [B,2,H,W] vertex_offset
  int B, C, Cv, H, W, n_vert, K;
  float thresh, down;
  float t0;              // logit-domain prefilter derived from thresh (x < t0 => sigmoid(x) <= thresh)
  // Tier A outputs
  int64_t* cls; float* score; float* proj; float* verts; float* bbox; int32_t* flat; int32_t* counts;
  // Tier B outputs
  float* kscore; float* kxy; int32_t* kflat;
  // workspace
  uint32_t* tickets;     // [B + B*Cv] zero between calls
  uint64_t* keys;        // [items][K]
  uint32_t* key_counts;  // [items]
  uint32_t* retry;       // [items] zero between calls: items to redo without speculation
  uint32_t* guess;       // [64] boundary bin + 1 remembered per plane index from the previous launch (0 = none)
  const float* ftable;   // [kFilterTableWords] logit bound per histogram bin, written by rtm3d_workspace_init (last word = magic)
  uint32_t* status;      // watchdog word (0 = ok)
  unsigned long long* stats;  // developer counters (nullable; decode_planes.cu StatSlot)
};

struct WorkspaceLayout {
  size_t table_off, tickets_off, status_off, keys_off, counts_off, retry_off, guess_off, flat_off, total;
  int strip_rows, nstrips, list_cap;
  size_t generic_smem;
};

// strip geometry + workspace layout for a shape (host)
WorkspaceLayout workspace_layout(int B, int C, int H, int W, int K);   // C = heat-map planes per image (all segments)

// launchers (return cudaError_t as int)
int launch_generic(const DecodeParams& p, int dtype, int mode, size_t smem, cudaStream_t s);
// persistent plane-streaming kernel; returns -1000 when the shape is not eligible (caller falls back to the generic
// path).  split_override: strips per plane (0 = choose).
int launch_planes(const PlaneParams& p, int dtype, int split_override, int speculate, int max_ctas, int debug, cudaStream_t s);
bool planes_eligible(const PlaneParams& p, int dtype);
constexpr int kPlanesMaxSplit = 8;
constexpr int kFilterTableWords = 2048;        // >= histogram bins + 1; the last word holds kFilterTableMagic
constexpr unsigned kFilterTableMagic = 0x5A17AB1Eu;
int launch_filter_table(float* table, cudaStream_t s);   // fills a workspace's table (rtm3d_workspace_init)
int threshold_table_bins();
void debug_set_stats(unsigned long long* dev_u64_16);
void debug_set_trace(unsigned long long* dev_u64_960);
void debug_set_copy_rows(int rows);   // developer instrumentation, not part of the public ABI
int launch_threshold_table(float* t, uint32_t* edge, cudaStream_t s);

}  // namespace rtm3d
