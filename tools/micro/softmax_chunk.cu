// Microbenchmark of the attention softmax inner loop on sm_100a: per 32-column chunk
//   tcgen05.ld x32 (next chunk in flight) -> p = ex2(s*c - mc) -> row sum -> bf16 pack -> tcgen05.st x16
// with each stage switchable, to find which one costs what for a single warp per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_chunk softmax_chunk.cu && ./softmax_chunk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

template <bool LD, bool EX, bool ST>
__global__ void __launch_bounds__(384, 1) kern(int units, float c, float mc, long long* out, float* sink, int cw) {
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(cw * 32) : "memory");
  }
  __syncthreads();
  if (warp >= cw) {            // spinning warps: one lane polls the barrier with a clock watchdog, like mbar_wait_wd
    if ((threadIdx.x & 31) == 0) {
      const long long t0 = clock64();
      while (!try_wait(&bar, 0)) {
        if (clock64() - t0 > 4000000000ll) __trap();
      }
    }
    return;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 208;
  float ls0 = 0.f, ls1 = 0.f;
  uint32_t sa[32], sb[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) sa[i] = sb[i] = __float_as_uint(threadIdx.x * 1e-3f + i * 0.01f);
  auto chunk = [&](const uint32_t* s, int ch) {
    uint32_t pk[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float e0 = fmaf(__uint_as_float(s[2 * e]), c, -mc), e1 = fmaf(__uint_as_float(s[2 * e + 1]), c, -mc);
      if (EX) {
        e0 = ex2f(e0);
        e1 = ex2f(e1);
      }
      ls0 += e0;
      ls1 += e1;
      pk[e] = pack(e0, e1);
    }
    if (ST) st16(base + ch * 16, pk);
    else if (pk[3] == 0x12345u) sink[1] = 1.f;
  };
  asm volatile("bar.sync 1, %0;" ::"r"(cw * 32));
  const long long t0 = clock64();
#pragma unroll 1
  for (int u = 0; u < units; ++u) {
    if (LD) ld32(base, sa);
#pragma unroll 1
    for (int ch = 0; ch < 6; ch += 2) {
      if (LD) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        ld32(base + (ch + 1) * 32, sb);
      }
      chunk(sa, ch);
      if (LD) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        ld32(base + (ch + 2) * 32, sa);
      }
      chunk(sb, ch + 1);
    }
    if (LD) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (ST) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("bar.sync 1, %0;" ::"r"(cw * 32));
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (ls0 + ls1 == 1.2345f) sink[0] = ls0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <bool LD, bool EX, bool ST>
void run(const char* name, int warps, long long* out, float* sink, int spin = 0) {
  const int units = 500;
  kern<LD, EX, ST><<<1, (warps + spin) * 32>>>(units, 0.25f, 3.0f, out, sink, warps);
  long long h = 0;
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %2d warps: %7.1f cycles per 32-column chunk (per warp)\n", name, warps, double(h) / units / 6);
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 64);
  cudaMalloc(&sink, 64);
  for (int warps : {4, 8}) {
    run<true, true, true>("ld + ex2 + st", warps, out, sink);
    run<true, false, true>("ld + st (no ex2)", warps, out, sink);
    run<true, true, false>("ld + ex2 (no st)", warps, out, sink);
    run<false, true, true>("ex2 + st (no ld)", warps, out, sink);
    run<false, true, false>("ex2 only", warps, out, sink);
    run<false, false, false>("fma + add + pack only", warps, out, sink);
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
