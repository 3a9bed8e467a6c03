// Microbenchmark: tcgen05.ld throughput / latency and MUFU.EX2 throughput per SM on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// mode 0: one load then wait (latency); mode 1: `depth` loads in flight then wait (throughput)
template <int X>
__global__ void tmem_kernel(int iters, int depth, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t r[4][X];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (d < depth) ld<X>(base + ((i * 4 + d) * X) % (512 - X), r[d]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (d < depth) acc ^= r[d][0] ^ r[d][X - 1];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

__global__ void mufu_kernel(int iters, long long* out, float* sink) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3f + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0;
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 1.2345f) sink[0] = s;
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 1024 * 8);
  cudaMalloc(&sink, 64);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    for (int depth : {1, 2, 4}) {
      long long h = 0;
      tmem_kernel<32><<<1, warps * 32>>>(iters, depth, out, sink);
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * depth * 32 * 32 * 4 * warps;
      printf("tcgen05.ld x32: %2d warps depth %d: %.1f cyc/iter, %.1f B/cyc/SM\n", warps, depth, double(h) / iters, bytes / h);
      tmem_kernel<16><<<1, warps * 32>>>(iters, depth, out, sink);
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      const double bytes16 = double(iters) * depth * 16 * 32 * 4 * warps;
      printf("tcgen05.ld x16: %2d warps depth %d: %.1f cyc/iter, %.1f B/cyc/SM\n", warps, depth, double(h) / iters, bytes16 / h);
    }
  }
  for (int warps : {1, 4, 8, 16}) {
    long long h = 0;
    mufu_kernel<<<1, warps * 32>>>(iters, out, (float*)sink);
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("ex2: %2d warps: %.2f ex2/cyc/SM\n", warps, double(iters) * 8 * 32 * warps / h);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
