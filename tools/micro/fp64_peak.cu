// Micro-benchmark: sustained DFMA / FFMA rate of this GPU (register-resident chains), to calibrate the
// arithmetic ceiling of the in-tile sweeps.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_kernel(T *out, int iters, T a, T b) {
  T x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = x[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T>
double run(int blocks, int threads, int iters) {
  T *out;
  cudaMalloc(&out, sizeof(T) * blocks * threads);
  fma_kernel<T><<<blocks, threads>>>(out, 10, (T)1.0000001, (T)1e-9);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  fma_kernel<T><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(out);
  return 2.0 * 16.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int wpsm : {4, 8, 12, 16, 32, 64}) {
    int threads = 128, blocks = sms * wpsm * 32 / threads;
    printf("warps/SM=%2d  fp64 %.2f TFLOP/s   fp32 %.2f TFLOP/s\n", wpsm, run<double>(blocks, threads, 4096), run<float>(blocks, threads, 4096));
  }
  return 0;
}
