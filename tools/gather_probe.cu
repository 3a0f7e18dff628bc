// Developer probe: DRAM bytes fetched per random 4-byte gather on B200 for different load flavours
// (run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum`).  Not part of the library.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int MODE>
__global__ void gather_kernel(const float* __restrict__ a, const unsigned* __restrict__ idx, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = a + idx[i];
  float v;
  if (MODE == 0) asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 1) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 2) asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 6) asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 7) asm volatile("ld.global.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 8) asm volatile("ld.global.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else asm volatile("ld.global.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p));
  out[i] = v;
}
__global__ void flush_kernel(float* f, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f[i] = 1.f;
}

template <int MODE>
void run(const char* name, const float* a, const unsigned* idx, float* out, int n, float* fl, size_t nfl) {
  flush_kernel<<<148 * 8, 256>>>(fl, nfl);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  gather_kernel<MODE><<<(n + 255) / 256, 256>>>(a, idx, out, n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("mode %d %-28s %8.2f us  (%d gathers)  err=%s\n", MODE, name, ms * 1e3f, n, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const size_t N = 160ull * 1000 * 1000;  // 640 MB of floats
  const int n = 921600;
  const int gran = argc > 1 ? atoi(argv[1]) : 0;
  if (gran) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("set L2 fetch granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), g);
  } else {
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("default L2 fetch granularity %zu\n", g);
  }
  float *a, *out, *fl; unsigned* idx;
  cudaMalloc(&a, N * 4); cudaMalloc(&out, n * 4); cudaMalloc(&idx, n * 4);
  const size_t nfl = 80ull * 1000 * 1000;
  cudaMalloc(&fl, nfl * 4);
  cudaMemset(a, 0, N * 4);
  std::vector<unsigned> h(n);
  unsigned long long s = 88172645463325252ull;
  for (int i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (unsigned)(s % N); }
  cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
  run<0>("ld.global", a, idx, out, n, fl, nfl);
  run<1>("ld.global.nc", a, idx, out, n, fl, nfl);
  run<2>("ld.global.L2::64B", a, idx, out, n, fl, nfl);
  run<3>("ld.global.nc.L2::64B", a, idx, out, n, fl, nfl);
  run<4>("ld.global.L1::no_allocate", a, idx, out, n, fl, nfl);
  run<5>("ld.global.cg", a, idx, out, n, fl, nfl);
  run<6>("ld.global.cv", a, idx, out, n, fl, nfl);
  run<7>("L1::no_allocate.L2::64B", a, idx, out, n, fl, nfl);
  run<8>("ld.global.L2::128B", a, idx, out, n, fl, nfl);
  run<9>("L1/L2 evict_first", a, idx, out, n, fl, nfl);
  cudaDeviceSynchronize();
  return 0;
}
