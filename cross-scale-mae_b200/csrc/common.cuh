// Shared device/host helpers for the csmae_b200 kernels (sm_100a only).
//
// Everything here is plain CUDA + inline PTX: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld), bf16 packing, warp reductions and the
// error plumbing behind the C-ABI (include/csmae_b200.h).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// ---------------------------------------------------------------------------------------------
// host side: error reporting across the C-ABI (no exceptions cross the boundary)
// ---------------------------------------------------------------------------------------------
void csm_set_error(const char* fmt, ...);

#define CSM_OK 0
#define CSM_ERR_ARG (-1)
#define CSM_ERR_CUDA (-2)
#define CSM_ERR_DEVICE (-3)

#define CSM_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      csm_set_error(__VA_ARGS__);                \
      return CSM_ERR_ARG;                        \
    }                                            \
  } while (0)

#define CSM_CHECK_LAUNCH(name)                                                       \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      csm_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));         \
      return CSM_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

static inline int csm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Cached cuTensorMapEncodeTiled over a row-major matrix [outer, inner] (leading dimension ld elements) of bf16
// (elem_bytes 2) or f32 (4), boxes {box_inner, box_outer}, 128- or 64-byte swizzle (gemm_tcgen05.cu).
int csm_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer, uint32_t elem_bytes, uint32_t swizzle_bytes);

#include <cstdlib>
#ifdef __CUDACC__
// Launch with programmatic dependent launch (PDL): the grid may be scheduled while the previous kernel of the
// stream is still draining; every kernel launched this way calls csm::pdl_wait() before it touches global
// memory, so only its launch latency and prologue overlap the predecessor's tail (~600 dependent launches/step).
template <typename... KArgs, typename... Args>
inline cudaError_t csm_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // CSMAE_PDL=0 (development switch): plain stream-ordered launches; griddepcontrol.* are then no-ops
  static const bool pdl = [] {
    const char* e = getenv("CSMAE_PDL");
    return e == nullptr || e[0] != '0';
  }();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
namespace csm {

// PDL: block until the preceding kernel of the stream has completed and its writes are visible / allow the
// next kernel of the stream to be scheduled (it blocks in its own pdl_wait until this grid has finished).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Round-to-nearest-even to bf16 precision, result kept as fp32.  Integer arithmetic on purpose: the scalar
// cvt (F2F.BF16) runs on the XU pipe at a quarter of the MUFU rate on sm_100 and was the binding pipe of the GEMM
// epilogues that round every accumulator (ncu: XU 85 % busy).  Inf stays Inf; NaN payloads are not preserved.
__device__ __forceinline__ float bf16_round(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return __uint_as_float(u & 0xFFFF0000u);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

// Two values rounded to bf16 at once: one packed cvt (F2FP, not an XU op) + two logic ops.
__device__ __forceinline__ void bf16_round_pair(float& a, float& b) {
  const float2 r = unpack_bf16x2(pack_bf16x2(a, b));
  a = r.x;
  b = r.y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Exact-erf GELU (nn.GELU()) and its derivative in one evaluation:
//   gelu(x) = x * Phi(x),  gelu'(x) = Phi(x) + x * phi(x),  phi = exp(-x^2/2)/sqrt(2 pi).
// Phi by Abramowitz-Stegun 26.2.17 (|abs err| < 7.5e-8, far inside the bf16 the results are stored in):
//   1 - Phi(|x|) = phi(|x|) * (b1 t + ... + b5 t^5),  t = 1 / (1 + 0.2316419 |x|)
// 15 FP32 ops + 2 MUFU (rcp, ex2) for both values -- the GEMM epilogue that calls this has a budget of
// K/32 instructions per output element before it, not the tensor pipe, bounds the kernel.
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& gp) {
  const float ax = fabsf(x);
  // (the epilogue is instruction-issue bound: one MUFU.RCP beats a Newton iteration on the FMA pipe)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.2316419f, 1.0f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044f));      // exp(-x^2/2)
  // coefficients pre-multiplied by 1/sqrt(2 pi)
  float poly = fmaf(t, 0.53070271f, -0.72657601f);
  poly = fmaf(poly, t, 0.71070688f);
  poly = fmaf(poly, t, -0.14224836f);
  poly = fmaf(poly, t, 0.12741480f);
  const float q = poly * t * e;                  // 1 - Phi(|x|)
  const float cdf = x >= 0.f ? 1.0f - q : q;
  g = x * cdf;
  gp = fmaf(x * e, 0.39894228040f, cdf);
}

// Two GELU / GELU' evaluations at once on packed FP32 pairs (FFMA2 / FMUL2 / FADD2, sm_100): same arithmetic as
// gelu_and_grad, half the FP32-pipe instructions, so the MUFU, logic and shared-memory instructions of the fc1
// epilogue find free issue slots.
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void gelu_and_grad2(float x0, float x1, float& g0, float& g1, float& gp0, float& gp1) {
  const unsigned long long X = f2_pack(x0, x1);
  const unsigned long long AX = f2_pack(fabsf(x0), fabsf(x1));
  float d0, d1, t0, t1, a0, a1, e0, e1;
  f2_unpack(f2_fma(AX, f2_pack(0.2316419f, 0.2316419f), f2_pack(1.0f, 1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  f2_unpack(f2_mul(f2_mul(X, X), f2_pack(-0.72134752044f, -0.72134752044f)), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const unsigned long long T = f2_pack(t0, t1), E = f2_pack(e0, e1);
  unsigned long long poly = f2_fma(T, f2_pack(0.53070271f, 0.53070271f), f2_pack(-0.72657601f, -0.72657601f));
  poly = f2_fma(poly, T, f2_pack(0.71070688f, 0.71070688f));
  poly = f2_fma(poly, T, f2_pack(-0.14224836f, -0.14224836f));
  poly = f2_fma(poly, T, f2_pack(0.12741480f, 0.12741480f));
  const unsigned long long Q = f2_mul(f2_mul(poly, T), E);                       // 1 - Phi(|x|)
  // cdf = x >= 0 ? 1 - q : q  ==  0.5 + copysign(0.5 - q, x)
  float h0, h1;
  f2_unpack(f2_fma(Q, f2_pack(-1.0f, -1.0f), f2_pack(0.5f, 0.5f)), h0, h1);
  h0 = __uint_as_float(__float_as_uint(h0) ^ (__float_as_uint(x0) & 0x80000000u));
  h1 = __uint_as_float(__float_as_uint(h1) ^ (__float_as_uint(x1) & 0x80000000u));
  const unsigned long long CDF = f2_add(f2_pack(h0, h1), f2_pack(0.5f, 0.5f));
  f2_unpack(f2_mul(X, CDF), g0, g1);
  f2_unpack(f2_fma(f2_mul(X, E), f2_pack(0.39894228040f, 0.39894228040f), CDF), gp0, gp1);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: the warp sleeps in hardware instead
      : "memory");                                          // of spinning through the issue slots of its scheduler
  return ok != 0;
}
// The wait loop as ONE asm block (scoped labels): try_wait (which suspends the warp up to the hint), branch -- nothing
// else per wake-up.  A third of all instructions the attention backward issued were the compiler's version of this
// loop (select, compare, counter, watchdog clock) run by warps that are waiting anyway, in the issue slots of the
// warps that are not.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n"
      "MBAR_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra MBAR_WAIT_DONE;\n\t"
      "bra MBAR_WAIT_LOOP;\n"
      "MBAR_WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// Variants over a 32-bit shared-window address.  In a register-starved loop the compiler re-derives a generic pointer's
// shared address at every use (cvta of the dynamic shared-memory base, the 1024-byte alignment: ~8 instructions);
// a base the compiler cannot see through (csm::pin) plus immediate offsets costs none.
__device__ __forceinline__ uint32_t pin(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
// pin() is transparent to ptxas (a PTX mov); a value read back through a warp shuffle of the own lane is not: what it
// produces stays in its register instead of being re-derived (one SHFL, for values computed once per work item)
__device__ __forceinline__ uint32_t pin_hard(uint32_t v) {
  uint32_t r, lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v), "r"(lane));
  return r;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_wait_bounded_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1, P2;\n\t"
      ".reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n"
      "MBAR_WAITA_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "@P1 bra MBAR_WAITA_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 P2, cnt, 0x4000000;\n\t"
      "@P2 bra MBAR_WAITA_LOOP;\n"
      "MBAR_WAITA_DONE:\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity), "r"(0x989680u)
      : "memory");
  if (!ok) __trap();
}
// the same with a bound on the number of wake-ups: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1, P2;\n\t"
      ".reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n"
      "MBAR_WAITB_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "@P1 bra MBAR_WAITB_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 P2, cnt, 0x4000000;\n\t"
      "@P2 bra MBAR_WAITB_LOOP;\n"
      "MBAR_WAITB_DONE:\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  if (!ok) __trap();
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (bytes). c0 = inner coord, c1 = outer.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with f32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor for tcgen05.mma, 128-byte swizzle (layout_type 2), sm_100 version 1.
//   start address, LBO and SBO are encoded in 16-byte units.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// The same descriptor as two 32-bit halves.  The MMA issuer is ONE thread whose instruction stream sits on the critical
// path of every short product (a 128 x 32 x 16 MMA retires in 16 cycles): stepping a descriptor through a tile must
// be a single 32-bit add on the low word (start address, 16-byte units; it cannot carry out of its 14-bit field for
// offsets inside the 227 KB of shared memory) instead of re-deriving the 64-bit value with shifts and masks per MMA.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t umma_desc_hi_sw128(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
}
// D[tmem] (+)= A[smem] * B[smem], descriptors passed as {lo, hi} words (cta_group::1)
__device__ __forceinline__ void umma_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> f32, dense. major: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4)                                   // D format f32
         | (1u << 7)                                 // A format bf16
         | (1u << 10)                                // B format bf16
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(n >> 3) << 17)
         | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace csm
#endif  // __CUDACC__
