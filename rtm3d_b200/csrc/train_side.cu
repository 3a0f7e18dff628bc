// Training-side mirror of the decoder: the main heat-map TARGET encoder and the penalty-reduced focal loss.
//
//   encode_targets_kernel   datasets/dataset_reader.py:215-291 (_build_targets, the m_hm part) with utils/data_utils.py:97-141
//                           (dynamic_radius :122-125, _compute_gaussian_radius :97-119, gaussian2D :128-141): per labelled
//                           object the centre of its 2D box on the heat-map grid, a Gaussian of that box's radius splatted
//                           with MAX into plane (image, class) -- the reference's per-object Python/numpy loop.  One warp per
//                           object; max is order independent, so an atomic max on the (positive) float bits reproduces the
//                           loop's result whatever the order.
//   focal_reduce_kernel     models/nets/module.py:41-68 (FocalLoss.forward) on utils/model_utils.py:10-14 (sigmoid_hm: sigmoid
//   focal_grad_kernel       clamped to [1e-4, 1 - 1e-4]), as called at models/rtm3d_loss.py:283: one streaming pass over logits +
//                           targets for the loss (sums in double), one more for the gradient w.r.t. the logits (what autograd
//                           computes for the reference).  HBM-bound elementwise work: float4 loads, grid = SMs x 8.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "postproc.h"

namespace rtm3d {

// data_utils._compute_gaussian_radius (:97-119), min_overlap = 0.7, for one box (heat-map units), in double like numpy
__device__ __forceinline__ double gaussian_radius_f64(double x1, double y1, double x2, double y2) {
  const double height = ceil(y2 - y1), width = ceil(x2 - x1), ov = 0.7;
  const double b1 = height + width, c1 = width * height * (1 - ov) / (1 + ov);
  const double r1 = (b1 + sqrt(b1 * b1 - 4 * c1)) / 2;
  const double b2 = 2 * (height + width), c2 = (1 - ov) * width * height;
  const double r2 = (b2 + sqrt(b2 * b2 - 16 * c2)) / 2;
  const double a3 = 4 * ov, b3 = -2 * ov * (height + width), c3 = (ov - 1) * width * height;
  const double r3 = (b3 + sqrt(b3 * b3 - 4 * a3 * c3)) / 2;
  return fmin(r1, fmin(r2, r3));
}

__global__ void __launch_bounds__(128) encode_targets_kernel(const TargetParams p) {
  const int obj = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (obj >= p.N) return;
  const double x1 = p.bbox[obj * 4 + 0], y1 = p.bbox[obj * 4 + 1], x2 = p.bbox[obj * 4 + 2], y2 = p.bbox[obj * 4 + 3];
  const double cx = (x1 + x2) / 2, cy = (y1 + y2) / 2;                 // data_utils.bbox_center
  const int px = static_cast<int>(cx), py = static_cast<int>(cy);      // astype(np.long): truncation toward zero (:225)
  const double rad = gaussian_radius_f64(x1, y1, x2, y2);
  const double sigma = (2 * rad + 1) / 6;                              // dynamic_radius (:123-125)
  const int R = static_cast<int>(ceil(rad));
  if (lane == 0) {
    p.m_proj[obj * 2 + 0] = px; p.m_proj[obj * 2 + 1] = py;
    p.m_off[obj * 2 + 0] = static_cast<float>(cx - px); p.m_off[obj * 2 + 1] = static_cast<float>(cy - py);
    p.sigma[obj] = static_cast<float>(sigma); p.radius[obj] = R;
  }
  if (!p.mask[obj]) return;                                            // (:266: only labelled main points are splatted)
  const int cls = static_cast<int>(p.cls[obj]), img = static_cast<int>(p.img_id[obj]);
  if (cls < 0 || cls >= p.C || img < 0 || img >= p.B) return;
  const bool noise = p.noise_mask[obj] != 0;
  const int side = 2 * R + 1;
  float* plane = p.m_hm + (static_cast<size_t>(img) * p.C + cls) * p.H * p.W;
  for (int t = lane; t < side * side; t += 32) {
    const int dy = t / side - R, dx = t - (t / side) * side - R;
    const int x = px + dx, y = py + dy;
    if (x < 0 || x >= p.W || y < 0 || y >= p.H) continue;
    double v = exp(-1.0 * static_cast<double>(dx * dx + dy * dy) / (2 * (sigma * sigma)));   // gaussian2D (:138-139)
    if (noise && t == (side * side) / 2) v = 0.9999;                   // (:262-263: the centre of a noise object)
    // m_hm = max(m_hm, kernel) (:272-273): values are positive floats, their bit patterns order like the values
    atomicMax(reinterpret_cast<int*>(plane + static_cast<size_t>(y) * p.W + x), __float_as_int(static_cast<float>(v)));
  }
}

int launch_encode_targets(const TargetParams& p, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(p.m_hm, 0, static_cast<size_t>(p.B) * p.C * p.H * p.W * sizeof(float), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (p.N > 0) encode_targets_kernel<<<(p.N * 32 + 127) / 128, 128, 0, s>>>(p);
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamped_sigmoid(float x) {               // model_utils.sigmoid_hm (:10-14)
  const float s = 1.0f / (1.0f + expf(-x));
  return fminf(fmaxf(s, 1e-4f), 1.0f - 1e-4f);
}

// acc[0] = sum of the positive terms, acc[1] = sum of the negative terms, acc[2] = number of positives (all double)
__global__ void __launch_bounds__(256) focal_reduce_kernel(const float* __restrict__ logits, const float* __restrict__ target, size_t n,
                                                          float alpha, float beta, double* acc) {
  double pos = 0.0, neg = 0.0, cnt = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float p = clamped_sigmoid(logits[i]), t = target[i];
    if (t == 1.0f) {                                                      // positive_index = target.eq(1)
      pos += static_cast<double>(logf(p) * powf(1.0f - p, alpha));
      cnt += 1.0;
    } else if (t < 1.0f) {                                                // negative_index = target.lt(1)
      neg += static_cast<double>(logf(1.0f - p) * powf(p, alpha) * powf(1.0f - t, beta));
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    pos += __shfl_xor_sync(0xffffffffu, pos, d);
    neg += __shfl_xor_sync(0xffffffffu, neg, d);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  __shared__ double sh[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = pos; sh[1][warp] = neg; sh[2][warp] = cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
    atomicAdd(&acc[threadIdx.x], v);
  }
}

// loss = -(pos + neg) / num_pos, or -neg when there is no positive (module.py:61-66); written as float
__global__ void focal_finish_kernel(const double* acc, float* loss) {
  const double num = acc[2];
  *loss = static_cast<float>(num == 0.0 ? -acc[1] : -(acc[0] + acc[1]) / num);
}

// d loss / d logit, scaled by `upstream` (the gradient of the caller's scalar w.r.t. the loss)
__global__ void __launch_bounds__(256) focal_grad_kernel(const float* __restrict__ logits, const float* __restrict__ target, size_t n, float alpha,
                                                        float beta, const double* acc, const float* upstream, float* grad) {
  const double num = acc[2];
  const float scale = (upstream ? *upstream : 1.0f) * static_cast<float>(num == 0.0 ? -1.0 : -1.0 / num);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float x = logits[i], t = target[i];
    const float s = 1.0f / (1.0f + expf(-x));
    float g = 0.f;
    if (s > 1e-4f && s < 1.0f - 1e-4f) {                                   // inside the clamp: dp/dx = p (1 - p); outside: 0
      const float p = s, q = 1.0f - s;
      float dldp = 0.f;
      if (t == 1.0f) {
        if (num != 0.0) dldp = powf(q, alpha) / p - alpha * powf(q, alpha - 1.0f) * logf(p);
      } else if (t < 1.0f) {
        dldp = powf(1.0f - t, beta) * (alpha * powf(p, alpha - 1.0f) * logf(q) - powf(p, alpha) / q);
      }
      g = scale * dldp * p * q;
    }
    grad[i] = g;
  }
}

int launch_focal_loss(const float* logits, const float* target, size_t n, float alpha, float beta, double* acc, float* loss, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(acc, 0, 3 * sizeof(double), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  focal_reduce_kernel<<<148 * 8, 256, 0, s>>>(logits, target, n, alpha, beta, acc);
  focal_finish_kernel<<<1, 1, 0, s>>>(acc, loss);
  return static_cast<int>(cudaGetLastError());
}
int launch_focal_grad(const float* logits, const float* target, size_t n, float alpha, float beta, const double* acc, const float* upstream,
                      float* grad, cudaStream_t s) {
  focal_grad_kernel<<<148 * 8, 256, 0, s>>>(logits, target, n, alpha, beta, acc, upstream, grad);
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// The three gather-L1 losses of RTM3DLoss.__call__ (models/rtm3d_loss.py:302-330): two channels (c0, c0+1) of an NCHW map
// gathered at (image, y, x) per entry, optional sigmoid, mean absolute error against a target pair over the valid entries.
// The reference permutes the WHOLE map to NHWC (a full copy) before indexing it; here only the 2 n scalars are read.
//   acc[0] = sum |pred - target|, acc[1] = number of elements (2 per valid entry), acc[2] = valid entries with (image, y, x)
//   outside the map (skipped; torch would raise)
__device__ __forceinline__ bool gather_l1_entry(const GatherL1Params& p, int e, size_t& idx, float& t0, float& t1) {
  if (!p.valid[e]) return false;
  const long long img = p.img[e], x = p.x[e], y = p.y[e];
  const int c0 = p.c0 ? p.c0[e] : 0;
  if (img < 0 || img >= p.B || x < 0 || x >= p.W || y < 0 || y >= p.H || c0 < 0 || c0 + 1 >= p.C) { idx = ~static_cast<size_t>(0); return true; }
  idx = ((static_cast<size_t>(img) * p.C + c0) * p.H + static_cast<size_t>(y)) * p.W + static_cast<size_t>(x);
  t0 = p.target[2 * static_cast<size_t>(e)]; t1 = p.target[2 * static_cast<size_t>(e) + 1];
  return true;
}
__global__ void __launch_bounds__(256) gather_l1_reduce_kernel(const GatherL1Params p) {
  double sum = 0.0, cnt = 0.0, bad = 0.0;
  const size_t HW = static_cast<size_t>(p.H) * p.W;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n; e += gridDim.x * blockDim.x) {
    size_t idx; float t0 = 0.f, t1 = 0.f;
    if (!gather_l1_entry(p, e, idx, t0, t1)) continue;
    if (idx == ~static_cast<size_t>(0)) { bad += 1.0; continue; }
    float p0 = p.map[idx], p1 = p.map[idx + HW];
    if (p.sigmoid) { p0 = 1.0f / (1.0f + expf(-p0)); p1 = 1.0f / (1.0f + expf(-p1)); }
    sum += static_cast<double>(fabsf(__fsub_rn(p0, t0))) + static_cast<double>(fabsf(__fsub_rn(p1, t1)));
    cnt += 2.0;
  }
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    bad += __shfl_xor_sync(0xffffffffu, bad, d);
  }
  __shared__ double sh[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = sum; sh[1][warp] = cnt; sh[2][warp] = bad; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
    if (v != 0.0) atomicAdd(&p.acc[threadIdx.x], v);
  }
}
// F.l1_loss(..., reduction='mean'): sum / count (NaN for an empty selection, as torch)
__global__ void gather_l1_finish_kernel(const double* acc, float* loss) { *loss = static_cast<float>(acc[0] / acc[1]); }

// d loss / d map, accumulated (atomicAdd: several entries may share a pixel) into a zeroed gradient map
__global__ void __launch_bounds__(256) gather_l1_grad_kernel(const GatherL1Params p, const float* upstream, float* grad) {
  const size_t HW = static_cast<size_t>(p.H) * p.W;
  const float scale = (upstream ? *upstream : 1.0f) / static_cast<float>(p.acc[1]);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n; e += gridDim.x * blockDim.x) {
    size_t idx; float t0 = 0.f, t1 = 0.f;
    if (!gather_l1_entry(p, e, idx, t0, t1) || idx == ~static_cast<size_t>(0)) continue;
    float p0 = p.map[idx], p1 = p.map[idx + HW], d0 = 1.f, d1 = 1.f;
    if (p.sigmoid) {
      p0 = 1.0f / (1.0f + expf(-p0)); p1 = 1.0f / (1.0f + expf(-p1));
      d0 = p0 * (1.0f - p0); d1 = p1 * (1.0f - p1);
    }
    const float s0 = p0 > t0 ? 1.f : (p0 < t0 ? -1.f : 0.f), s1 = p1 > t1 ? 1.f : (p1 < t1 ? -1.f : 0.f);   // sign(pred - target)
    if (s0 != 0.f) atomicAdd(grad + idx, scale * s0 * d0);
    if (s1 != 0.f) atomicAdd(grad + idx + HW, scale * s1 * d1);
  }
}

int launch_gather_l1(const GatherL1Params& p, float* loss, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(p.acc, 0, 3 * sizeof(double), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int blocks = p.n > 0 ? (p.n + 255) / 256 : 1;
  gather_l1_reduce_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, s>>>(p);
  gather_l1_finish_kernel<<<1, 1, 0, s>>>(p.acc, loss);
  return static_cast<int>(cudaGetLastError());
}
int launch_gather_l1_grad(const GatherL1Params& p, const float* upstream, float* grad, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(grad, 0, static_cast<size_t>(p.B) * p.C * p.H * p.W * sizeof(float), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int blocks = p.n > 0 ? (p.n + 255) / 256 : 1;
  gather_l1_grad_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, s>>>(p, upstream, grad);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rtm3d
