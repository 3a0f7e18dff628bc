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
  int cluster_override;  // streaming kernel: CTAs per problem (0 = choose)
  int debug;             // developer switches of the streaming kernel (flags bits 16..19)
};

struct WorkspaceLayout {
  size_t tickets_off, status_off, keys_off, counts_off, total;
  int strip_rows, nstrips, list_cap;
  size_t generic_smem;
};

// strip geometry + workspace layout for a shape (host)
WorkspaceLayout workspace_layout(int B, int C, int H, int W, int K);

// launchers (return cudaError_t as int)
int launch_generic(const DecodeParams& p, int dtype, int mode, size_t smem, cudaStream_t s);
// streaming TMA kernel; returns -1000 when the shape is not eligible (caller falls back to the generic path)
int launch_stream(const DecodeParams& p, int dtype, int mode, cudaStream_t s);
bool stream_eligible(const DecodeParams& p, int dtype, int mode);
void debug_set_timeline(unsigned long long* ptr);  // developer instrumentation, not part of the public ABI

}  // namespace rtm3d
