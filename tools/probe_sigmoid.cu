// Probe kernel: the sigmoid formula the decode kernels use, for bit-comparison against torch's CUDA sigmoid.
#include <cuda_runtime.h>
extern "C" __global__ void k_sig(const float* __restrict__ x, float* __restrict__ y, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = 1.0f / (1.0f + expf(-x[i]));
}
extern "C" int probe_sigmoid(const float* x, float* y, long n, void* stream) {
    k_sig<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    return (int)cudaGetLastError();
}
