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

// cand_count word of a strip: low 31 bits = keys in the list; top bit set = the keys are (score, index) keys (exact path),
// clear = (ordered logit, index) keys whose sigmoid the select kernel evaluates
constexpr uint32_t kCandScoreKeys = 0x80000000u;

// Parameters of the plane-resident scan kernel (scan_planes.cu).  Either heat-map may be absent (C or Cv = 0).
struct ScanParams {
  const void* hm_main;   // [B,C,H,W]  main_kf logits            (nullptr when C == 0)
  const void* hm_kpt;    // [B,Cv,H,W] keypoint heat-map logits  (nullptr when Cv == 0)
  int B, C, Cv, H, W, K;
  float thresh;
  float t0;              // logit-domain prefilter derived from thresh (x < t0 => sigmoid(x) <= thresh)
  unsigned long long* cand;   // [strips][list_cap] candidate keys per strip (unordered)
  uint32_t* cand_count;       // [strips]
  uint32_t* queue;            // [1] strip counter, 0 between launches (wraps back to 0 by itself)
  uint32_t* status;           // watchdog word (0 = ok)
  unsigned long long* stats;  // developer counters (nullable)
};

// Selection after the scan (select.cu): merge the strip lists of every selection problem, exact top-K, sort, write
// score / flat / counts (Tier A) and kscore / kflat (Tier B, with the 0.0-score fillers).
struct SelectParams {
  const unsigned long long* cand; const uint32_t* cand_count;
  int B, C, Cv, H, W, K, Sp, list_cap;
  float thresh;
  float* score; int32_t* flat; int32_t* counts;      // Tier A (C > 0)
  float* kscore; int32_t* kflat;                      // Tier B (Cv > 0)
};

struct WorkspaceLayout {
  size_t table_off, tickets_off, status_off, keys_off, counts_off, retry_off, guess_off, flat_off, queue_off, cand_off, cand_count_off, total;
  int cand_strips, cand_cap;     // capacity of the scan kernel's candidate lists
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
// plane-resident scan kernel + selection (scan_planes.cu, select.cu); launch_scan returns -1000 when the shape is not eligible
int launch_scan(const ScanParams& p, int dtype, int strips_override, int max_ctas, int debug, cudaStream_t s, int* strips_per_plane,
                int* list_cap);
bool scan_eligible(int B, int C, int Cv, int H, int W, int K, int dtype, const void* hm_main, const void* hm_kpt);
int scan_max_strips_per_plane(int H, int W, int K);
int scan_strips_per_plane(int H, int W, int K, int dtype, int strips_override);   // 0 = not eligible
int scan_list_cap(int K);
size_t select_smem_bytes(int C, int Cv, int Sp, int list_cap, int K);
int launch_select(const SelectParams& p, cudaStream_t s);
constexpr int kPlanesMaxSplit = 8;
constexpr int kFilterTableWords = 2048;        // >= histogram bins + 1; the last word holds kFilterTableMagic
constexpr unsigned kFilterTableMagic = 0x5A17AB1Eu;
int launch_filter_table(float* table, cudaStream_t s);   // fills a workspace's table (rtm3d_workspace_init)
int threshold_table_bins();
void debug_set_stats(unsigned long long* dev_u64_16);
unsigned long long* debug_get_stats();
void debug_set_trace(unsigned long long* dev_u64_960);
void debug_set_copy_rows(int rows);   // developer instrumentation, not part of the public ABI
int launch_threshold_table(float* t, uint32_t* edge, cudaStream_t s);

}  // namespace rtm3d
