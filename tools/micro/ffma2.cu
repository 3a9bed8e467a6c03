// Microbenchmark: FP32 FMA issue rate, scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, ILP = 8;
__global__ void k_scalar(float* out, float a, float b) {
  float x[2 * ILP];
  for (int i = 0; i < 2 * ILP; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0;
  for (int i = 0; i < 2 * ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
  unsigned long long x[ILP], aa, bb;
  float2 av = make_float2(a, a), bv = make_float2(b, b);
  aa = *reinterpret_cast<unsigned long long*>(&av);
  bb = *reinterpret_cast<unsigned long long*>(&bv);
  for (int i = 0; i < ILP; ++i) {
    float2 v = make_float2(threadIdx.x + 2 * i, threadIdx.x + 2 * i + 1);
    x[i] = *reinterpret_cast<unsigned long long*>(&v);
  }
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  float s = 0;
  for (int i = 0; i < ILP; ++i) {
    float2 v = *reinterpret_cast<float2*>(&x[i]);
    s += v.x + v.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_scalar<<<148 * 4, 512>>>(out, 1.0001f, 0.5f);
      else k_packed<<<148 * 4, 512>>>(out, 1.0001f, 0.5f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 4 * 512 * ITERS * 2 * ILP;
      printf("%s: %.3f ms  %.1f TFMA/s  (%.1f TFLOP/s)\n", mode ? "FFMA2 " : "FFMA  ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
