// LayerNorm forward / backward over the fp32 residual stream (timm Block norm1/norm2 and decoder_norm,
// eps = 1e-6: models_mae/MAE_ViT_Baseline.py:43-45,160-188,292), bias-gradient column sums and the
// fp32 -> bf16 weight cast.  HBM-bound warp-per-row kernels with 16-byte accesses.
#include "common.cuh"

namespace {
using namespace csm;

constexpr int LN_MAX_VEC = 8;  // float4 per lane -> D <= 1024

__device__ __forceinline__ void red_add_v4(float* addr, const float4 v) {   // 16-byte aligned address
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// One warp per row, VEC float4 per lane (D <= VEC * 128).  Statistics in fp32 (two-pass: mean, then centred
// variance), output rounded once.  Each warp walks rows with a stride and keeps the NEXT row's loads in flight
// while it reduces and writes the current one (the kernel is a chain of ~1 us memory latencies otherwise).
template <int VEC>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int rows, int D, float eps) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * warps;
  int row = blockIdx.x * warps + (threadIdx.x >> 5);
  pdl_wait();
  pdl_trigger();
  if (row >= rows) return;
  const float inv_d = 1.0f / static_cast<float>(D);
  float4 cur[VEC], nxt[VEC];
  auto load = [&](float4 (&v)[VEC], int r) {
    const float* xr = x + static_cast<size_t>(r) * D;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int i = (k * 32 + lane) * 4;
      v[k] = i < D ? *reinterpret_cast<const float4*>(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load(cur, row);
  for (; row < rows; row += stride) {
    const bool more = row + stride < rows;
    if (more) load(nxt, row + stride);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) s += cur[k].x + cur[k].y + cur[k].z + cur[k].w;
    const float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int i = (k * 32 + lane) * 4;
      if (i < D) {
        const float a = cur[k].x - mean, b = cur[k].y - mean, c = cur[k].z - mean, d = cur[k].w - mean;
        ss += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int i = (k * 32 + lane) * 4;
      if (i < D) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + i);
        const float4 b = *reinterpret_cast<const float4*>(beta + i);
        float4 o;
        o.x = (cur[k].x - mean) * rstd * g.x + b.x;
        o.y = (cur[k].y - mean) * rstd * g.y + b.y;
        o.z = (cur[k].z - mean) * rstd * g.z + b.z;
        o.w = (cur[k].w - mean) * rstd * g.w + b.w;
        if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * D + i) = o;
        if (out_bf16 != nullptr) {
          uint2 pk;
          pk.x = pack_bf16x2(o.x, o.y);
          pk.y = pack_bf16x2(o.z, o.w);
          *reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * D + i) = pk;
        }
      }
    }
    if (more) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) cur[k] = nxt[k];
    }
  }
}

// dy = float(dy_bf16) + dy2_f32 (either may be null).  dx_ln = rstd * (g - mean(g) - xhat * mean(g * xhat)),
// g = dy * gamma.  dres_out = (dres_in ? dres_in : 0) + dx_ln, plus a bf16 copy for the following GEMMs.
// Column sums accumulated on the way: dgamma += sum_r dy*xhat, dbeta += sum_r dy and (optional)
// dcolsum += sum_r bf16(dres_out) -- the bias gradient of the Linear whose output gradient dres_bf16 is
// (attn.proj / mlp.fc2: their dY IS the residual-stream gradient), so no separate column-sum pass.
//
// Layout: a row is spread over TPR = D/4 threads (one float4 each), a CTA handles blockDim/TPR rows per
// iteration; the two row statistics cross the row's warps through shared memory (one __syncthreads per
// iteration, double-buffered).  Every thread owns 4 fixed columns, so the three column sums cost 12
// registers and the kernel runs at full occupancy (HBM-latency hiding by many resident rows).
__host__ __device__ constexpr int ln_tile_threads(int tpr) { return (tpr == 96 || tpr == 192) ? 384 : (tpr == 256 ? 512 : 256); }

template <int TPR>
__global__ void __launch_bounds__(ln_tile_threads(TPR))
layernorm_bwd_tile_kernel(const __nv_bfloat16* __restrict__ dy_bf16, const float* __restrict__ dy2,
                          const float* __restrict__ x, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                          const float* __restrict__ dres_in, float* __restrict__ dres_out,
                          __nv_bfloat16* __restrict__ dres_bf16, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows) {
  constexpr int D = TPR * 4;
  constexpr int WPR = TPR / 32;                       // warps per row
  constexpr int THREADS = ln_tile_threads(TPR);
  constexpr int RPI = THREADS / TPR;                  // rows per iteration
  static_assert(THREADS % TPR == 0 && TPR % 32 == 0, "a row must be a whole number of warps");
  __shared__ float2 s_part[2][RPI][WPR];
  __shared__ float4 s_col[3][RPI > 1 ? RPI - 1 : 1][TPR];
  const int rsub = threadIdx.x / TPR;
  const int ct = threadIdx.x % TPR;
  const int wir = ct >> 5;                            // warp index inside the row
  const int lane = threadIdx.x & 31;
  const int col = ct * 4;
  pdl_wait();
  pdl_trigger();
  const float4 gm = *reinterpret_cast<const float4*>(gamma + col);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, ac = ag;

  struct RowIn {
    float4 xv, d, rin;
    float mean, rstd;
  };
  auto load_row = [&](int row) {
    RowIn r;
    r.xv = make_float4(0.f, 0.f, 0.f, 0.f);
    r.d = r.xv;
    r.rin = r.xv;
    r.mean = 0.f;
    r.rstd = 0.f;
    if (row < rows) {
      const size_t base = static_cast<size_t>(row) * D + col;
      r.xv = *reinterpret_cast<const float4*>(x + base);
      if (dy_bf16 != nullptr) {
        const uint2 ev = *reinterpret_cast<const uint2*>(dy_bf16 + base);
        const float2 e0 = unpack_bf16x2(ev.x), e1 = unpack_bf16x2(ev.y);
        r.d = make_float4(e0.x, e0.y, e1.x, e1.y);
      }
      if (dy2 != nullptr) {
        const float4 d2 = *reinterpret_cast<const float4*>(dy2 + base);
        r.d.x += d2.x; r.d.y += d2.y; r.d.z += d2.z; r.d.w += d2.w;
      }
      if (dres_in != nullptr) r.rin = *reinterpret_cast<const float4*>(dres_in + base);
      r.mean = mean_in[row];
      r.rstd = rstd_in[row];
    }
    return r;
  };

  int buf = 0;
  const int step = gridDim.x * RPI;
  int row = blockIdx.x * RPI + rsub;
  RowIn cur = load_row(row);
  for (int row0 = blockIdx.x * RPI; row0 < rows; row0 += step, row += step, buf ^= 1) {
    // software pipeline: the next row's loads are in flight while this row is reduced and written
    const RowIn nxt = load_row(row0 + step < rows ? row + step : rows);
    const bool valid = row < rows;
    const float4 xv = cur.xv, d = cur.d, rin = cur.rin;
    const float mean = cur.mean, rstd = cur.rstd;
    const float4 xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd,
                                  (xv.w - mean) * rstd);
    const float4 g = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
    float c1 = g.x + g.y + g.z + g.w;
    float c2 = g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
    c1 = warp_sum(c1);
    c2 = warp_sum(c2);
    if (lane == 0) s_part[buf][rsub][wir] = make_float2(c1, c2);
    __syncthreads();
    c1 = 0.f; c2 = 0.f;
#pragma unroll
    for (int w = 0; w < WPR; ++w) {
      const float2 pr = s_part[buf][rsub][w];
      c1 += pr.x; c2 += pr.y;
    }
    c1 *= (1.0f / D);
    c2 *= (1.0f / D);
    if (valid) {
      const size_t base = static_cast<size_t>(row) * D + col;
      float4 o;
      o.x = rstd * (g.x - c1 - xh.x * c2) + rin.x;
      o.y = rstd * (g.y - c1 - xh.y * c2) + rin.y;
      o.z = rstd * (g.z - c1 - xh.z * c2) + rin.z;
      o.w = rstd * (g.w - c1 - xh.w * c2) + rin.w;
      *reinterpret_cast<float4*>(dres_out + base) = o;
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      if (dres_bf16 != nullptr) *reinterpret_cast<uint2*>(dres_bf16 + base) = pk;
      const float2 r0 = unpack_bf16x2(pk.x), r1 = unpack_bf16x2(pk.y);
      ac.x += r0.x; ac.y += r0.y; ac.z += r1.x; ac.w += r1.y;
      ag.x += d.x * xh.x; ag.y += d.y * xh.y; ag.z += d.z * xh.z; ag.w += d.w * xh.w;
      ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
    }
    cur = nxt;
  }
  // fold the RPI row slots, then one vector atomic per thread and sum
  if (RPI > 1) {
    if (rsub > 0) {
      s_col[0][rsub - 1][ct] = ag;
      s_col[1][rsub - 1][ct] = ab;
      s_col[2][rsub - 1][ct] = ac;
    }
    __syncthreads();
    if (rsub == 0) {
#pragma unroll
      for (int r = 0; r < RPI - 1; ++r) {
        const float4 a = s_col[0][r][ct], b = s_col[1][r][ct], c = s_col[2][r][ct];
        ag.x += a.x; ag.y += a.y; ag.z += a.z; ag.w += a.w;
        ab.x += b.x; ab.y += b.y; ab.z += b.z; ab.w += b.w;
        ac.x += c.x; ac.y += c.y; ac.z += c.z; ac.w += c.w;
      }
    }
  }
  if (rsub == 0) {
    red_add_v4(dgamma + col, ag);
    red_add_v4(dbeta + col, ab);
    if (dcolsum != nullptr) red_add_v4(dcolsum + col, ac);
  }
}

// Same maths, deeper memory pipeline: every thread stages ITS OWN slice of the next LN_NST-1 rows in shared memory
// with cp.async (no thread reads another thread's slice, so the only CTA barrier stays the one for the row
// statistics), which keeps ~3 rows x 40 B per thread in flight without holding them in registers, and handles VPT
// float4 column groups per thread so the shuffles / barrier / address arithmetic are amortised over more bytes.
constexpr int LN_NST = 4;

__device__ __forceinline__ void cp_async16(void* s, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* s, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* s, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}

__host__ __device__ constexpr int ln_pipe_threads(int tpr) { return tpr == 96 ? 192 : (tpr == 192 ? 384 : 256); }
__host__ __device__ constexpr size_t ln_pipe_smem(int tpr, int vpt) {
  return static_cast<size_t>(LN_NST) * vpt * ln_pipe_threads(tpr) * 40 + static_cast<size_t>(LN_NST) * (ln_pipe_threads(tpr) / 32) * 8;
}

template <int TPR, int VPT>
__global__ void __launch_bounds__(ln_pipe_threads(TPR))
layernorm_bwd_pipe_kernel(const __nv_bfloat16* __restrict__ dy_bf16, const float* __restrict__ dy2,
                          const float* __restrict__ x, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                          const float* __restrict__ dres_in, float* __restrict__ dres_out,
                          __nv_bfloat16* __restrict__ dres_bf16, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows) {
  constexpr int D = TPR * 4 * VPT;
  constexpr int WPR = TPR / 32;
  constexpr int THREADS = ln_pipe_threads(TPR);
  constexpr int RPI = THREADS / TPR;
  constexpr int NW = THREADS / 32;
  static_assert(THREADS % TPR == 0 && TPR % 32 == 0, "a row must be a whole number of warps");
  extern __shared__ __align__(16) unsigned char ln_smem[];
  float4* s_x = reinterpret_cast<float4*>(ln_smem);                       // [NST][VPT][THREADS]
  float4* s_r = s_x + LN_NST * VPT * THREADS;
  uint2* s_d = reinterpret_cast<uint2*>(s_r + LN_NST * VPT * THREADS);
  float2* s_stat = reinterpret_cast<float2*>(s_d + LN_NST * VPT * THREADS);  // [NST][NW]
  __shared__ float2 s_part[2][RPI][WPR];
  const int tid = threadIdx.x;
  const int rsub = tid / TPR;
  const int ct = tid % TPR;
  const int wir = ct >> 5;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  pdl_wait();
  pdl_trigger();
  float4 gm[VPT], ag[VPT], ab[VPT], ac[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    gm[v] = *reinterpret_cast<const float4*>(gamma + (v * TPR + ct) * 4);
    ag[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[v] = ag[v];
    ac[v] = ag[v];
  }
  const bool has_d = dy_bf16 != nullptr, has_r = dres_in != nullptr;

  auto issue = [&](int st, int row) {
    if (row < rows) {
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const size_t off = static_cast<size_t>(row) * D + (v * TPR + ct) * 4;
        const int si = (st * VPT + v) * THREADS + tid;
        cp_async16(s_x + si, x + off);
        if (has_r) cp_async16(s_r + si, dres_in + off);
        if (has_d) cp_async8(s_d + si, dy_bf16 + off);
      }
      if (lane == 0) {
        cp_async4(&s_stat[st * NW + warp].x, mean_in + row);
        cp_async4(&s_stat[st * NW + warp].y, rstd_in + row);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int step = gridDim.x * RPI;
  int row = blockIdx.x * RPI + rsub;
#pragma unroll
  for (int s = 0; s < LN_NST - 1; ++s) issue(s, row + s * step);
  int st = 0, buf = 0;
  for (int row0 = blockIdx.x * RPI; row0 < rows; row0 += step, row += step, buf ^= 1) {
    issue(st == 0 ? LN_NST - 1 : st - 1, row + (LN_NST - 1) * step);
    asm volatile("cp.async.wait_group %0;" ::"n"(LN_NST - 1) : "memory");
    __syncwarp();
    const bool valid = row < rows;
    float4 xh[VPT], d[VPT], g[VPT], rin[VPT];
    float mean = 0.f, rstd = 0.f;
    if (valid) {
      const float2 ms = s_stat[st * NW + warp];
      mean = ms.x;
      rstd = ms.y;
    }
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int si = (st * VPT + v) * THREADS + tid;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      d[v] = xv;
      rin[v] = xv;
      if (valid) {
        xv = s_x[si];
        if (has_r) rin[v] = s_r[si];
        if (has_d) {
          const uint2 ev = s_d[si];
          const float2 e0 = unpack_bf16x2(ev.x), e1 = unpack_bf16x2(ev.y);
          d[v] = make_float4(e0.x, e0.y, e1.x, e1.y);
        }
        if (dy2 != nullptr) {
          const float4 d2 = *reinterpret_cast<const float4*>(dy2 + static_cast<size_t>(row) * D + (v * TPR + ct) * 4);
          d[v].x += d2.x; d[v].y += d2.y; d[v].z += d2.z; d[v].w += d2.w;
        }
      }
      xh[v] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      g[v] = make_float4(d[v].x * gm[v].x, d[v].y * gm[v].y, d[v].z * gm[v].z, d[v].w * gm[v].w);
      c1 += g[v].x + g[v].y + g[v].z + g[v].w;
      c2 += g[v].x * xh[v].x + g[v].y * xh[v].y + g[v].z * xh[v].z + g[v].w * xh[v].w;
    }
    c1 = warp_sum(c1);
    c2 = warp_sum(c2);
    if (WPR > 1) {
      if (lane == 0) s_part[buf][rsub][wir] = make_float2(c1, c2);
      __syncthreads();
      c1 = 0.f; c2 = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) {
        const float2 pr = s_part[buf][rsub][w];
        c1 += pr.x; c2 += pr.y;
      }
    }
    c1 *= (1.0f / D);
    c2 *= (1.0f / D);
    if (valid) {
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const size_t base = static_cast<size_t>(row) * D + (v * TPR + ct) * 4;
        float4 o;
        o.x = rstd * (g[v].x - c1 - xh[v].x * c2) + rin[v].x;
        o.y = rstd * (g[v].y - c1 - xh[v].y * c2) + rin[v].y;
        o.z = rstd * (g[v].z - c1 - xh[v].z * c2) + rin[v].z;
        o.w = rstd * (g[v].w - c1 - xh[v].w * c2) + rin[v].w;
        *reinterpret_cast<float4*>(dres_out + base) = o;
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        if (dres_bf16 != nullptr) *reinterpret_cast<uint2*>(dres_bf16 + base) = pk;
        const float2 r0 = unpack_bf16x2(pk.x), r1 = unpack_bf16x2(pk.y);
        ac[v].x += r0.x; ac[v].y += r0.y; ac[v].z += r1.x; ac[v].w += r1.y;
        ag[v].x += d[v].x * xh[v].x; ag[v].y += d[v].y * xh[v].y; ag[v].z += d[v].z * xh[v].z; ag[v].w += d[v].w * xh[v].w;
        ab[v].x += d[v].x; ab[v].y += d[v].y; ab[v].z += d[v].z; ab[v].w += d[v].w;
      }
    }
    st = st == LN_NST - 1 ? 0 : st + 1;
  }
  // fold the RPI row slots through the (now idle) staging area, then one vector atomic per thread and sum
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (RPI > 1) {
    __syncthreads();
    float4* s_col = s_x;                                    // [3][RPI-1][VPT][TPR]
    if (rsub > 0) {
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        s_col[((0 * (RPI - 1) + rsub - 1) * VPT + v) * TPR + ct] = ag[v];
        s_col[((1 * (RPI - 1) + rsub - 1) * VPT + v) * TPR + ct] = ab[v];
        s_col[((2 * (RPI - 1) + rsub - 1) * VPT + v) * TPR + ct] = ac[v];
      }
    }
    __syncthreads();
    if (rsub == 0) {
#pragma unroll
      for (int r = 0; r < RPI - 1; ++r) {
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          const float4 a = s_col[((0 * (RPI - 1) + r) * VPT + v) * TPR + ct];
          const float4 b = s_col[((1 * (RPI - 1) + r) * VPT + v) * TPR + ct];
          const float4 c = s_col[((2 * (RPI - 1) + r) * VPT + v) * TPR + ct];
          ag[v].x += a.x; ag[v].y += a.y; ag[v].z += a.z; ag[v].w += a.w;
          ab[v].x += b.x; ab[v].y += b.y; ab[v].z += b.z; ab[v].w += b.w;
          ac[v].x += c.x; ac[v].y += c.y; ac[v].z += c.z; ac[v].w += c.w;
        }
      }
    }
  }
  if (rsub == 0) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int col = (v * TPR + ct) * 4;
      red_add_v4(dgamma + col, ag[v]);
      red_add_v4(dbeta + col, ab[v]);
      if (dcolsum != nullptr) red_add_v4(dcolsum + col, ac[v]);
    }
  }
}

// Generic-D fallback (D not a multiple of 128, D <= 1024): one warp per row, register-resident row.
__global__ void layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy_bf16, const float* __restrict__ dy2,
                                     const float* __restrict__ x, const float* __restrict__ mean_in,
                                     const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                                     const float* __restrict__ dres_in, float* __restrict__ dres_out,
                                     __nv_bfloat16* __restrict__ dres_bf16, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows, int D) {
  extern __shared__ float s_red[];  // [3][warps][D]
  const int warps = blockDim.x >> 5;
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float4 ag[LN_MAX_VEC], ab[LN_MAX_VEC], ac[LN_MAX_VEC];
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ac[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * warps + w; row < rows; row += gridDim.x * warps) {
    const size_t base = static_cast<size_t>(row) * D;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[LN_MAX_VEC], g[LN_MAX_VEC];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int i = (k * 32 + lane) * 4;
      if (i < D) {
        const float4 xv = *reinterpret_cast<const float4*>(x + base + i);
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dy_bf16 != nullptr) {
          const uint2 ev = *reinterpret_cast<const uint2*>(dy_bf16 + base + i);
          const float2 e0 = unpack_bf16x2(ev.x), e1 = unpack_bf16x2(ev.y);
          d = make_float4(e0.x, e0.y, e1.x, e1.y);
        }
        if (dy2 != nullptr) {
          const float4 d2 = *reinterpret_cast<const float4*>(dy2 + base + i);
          d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
        }
        const float4 gm = *reinterpret_cast<const float4*>(gamma + i);
        xh[k] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        g[k] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        c1 += g[k].x + g[k].y + g[k].z + g[k].w;
        c2 += g[k].x * xh[k].x + g[k].y * xh[k].y + g[k].z * xh[k].z + g[k].w * xh[k].w;
        ag[k].x += d.x * xh[k].x; ag[k].y += d.y * xh[k].y; ag[k].z += d.z * xh[k].z; ag[k].w += d.w * xh[k].w;
        ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
      }
    }
    c1 = warp_sum(c1) / D;
    c2 = warp_sum(c2) / D;
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int i = (k * 32 + lane) * 4;
      if (i < D) {
        float4 o;
        o.x = rstd * (g[k].x - c1 - xh[k].x * c2);
        o.y = rstd * (g[k].y - c1 - xh[k].y * c2);
        o.z = rstd * (g[k].z - c1 - xh[k].z * c2);
        o.w = rstd * (g[k].w - c1 - xh[k].w * c2);
        if (dres_in != nullptr) {
          const float4 r = *reinterpret_cast<const float4*>(dres_in + base + i);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dres_out + base + i) = o;
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        if (dres_bf16 != nullptr) *reinterpret_cast<uint2*>(dres_bf16 + base + i) = pk;
        const float2 r0 = unpack_bf16x2(pk.x), r1 = unpack_bf16x2(pk.y);
        ac[k].x += r0.x; ac[k].y += r0.y; ac[k].z += r1.x; ac[k].w += r1.y;
      }
    }
  }
  float* sg = s_red;
  float* sb = s_red + warps * D;
  float* sc = s_red + 2 * warps * D;
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    const int i = (k * 32 + lane) * 4;
    if (i < D) {
      *reinterpret_cast<float4*>(sg + w * D + i) = ag[k];
      *reinterpret_cast<float4*>(sb + w * D + i) = ab[k];
      *reinterpret_cast<float4*>(sc + w * D + i) = ac[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int ww = 0; ww < warps; ++ww) {
      a += sg[ww * D + i];
      b += sb[ww * D + i];
      c += sc[ww * D + i];
    }
    atomicAdd(dgamma + i, a);
    atomicAdd(dbeta + i, b);
    if (dcolsum != nullptr) atomicAdd(dcolsum + i, c);
  }
}

// db[n] += sum over rows r (r % skip_period != 0 when skip_period > 0) of dy[r, n]
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db, int rows, int N,
                                   int skip_period) {
  __shared__ float s[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + tx * 8;
  pdl_wait();
  pdl_trigger();
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (col < N) {
    const int stride = gridDim.y * 8;
    for (int r0 = blockIdx.y * 8 + ty; r0 < rows; r0 += 4 * stride) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {     // four independent 16-byte loads in flight per thread
        const int r = r0 + u * stride;
        const bool ok = r < rows && !(skip_period > 0 && (r % skip_period) == 0);
        v[u] = ok ? *reinterpret_cast<const uint4*>(dy + static_cast<size_t>(r) * N + col) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t wv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(wv[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;  // 256 threads, one column each
  if (blockIdx.x * 256 + c < N) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += s[k][c];
    atomicAdd(db + blockIdx.x * 256 + c, a);
  }
}

struct CastEntry {
  const float* src;
  __nv_bfloat16* dst;
  long long n;
};

// Multi-tensor fp32 -> bf16 cast (master weights -> tensor-core operands): blockIdx.y selects the tensor.
__global__ void cast_multi_kernel(const CastEntry* __restrict__ table) {
  const CastEntry e = table[blockIdx.y];
  const long long n4 = e.n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(e.src + i * 4);
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(e.dst + i * 4) = pk;
  }
  if (blockIdx.x == 0 && threadIdx.x < (e.n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    e.dst[i] = __float2bfloat16_rn(e.src[i]);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i * 4) = pk;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

}  // namespace

template <int VEC>
cudaError_t launch_ln_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16, float* out_f32,
                          float* mean, float* rstd, int rows, int D, float eps, cudaStream_t stream) {
  static int ctas_per_sm = 0, num_sms = 0;
  if (ctas_per_sm == 0) {
    int occ = 0, dev = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_fwd_kernel<VEC>, 256, 0) != cudaSuccess || occ < 1)
      occ = 1;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms < 1) num_sms = 148;
    ctas_per_sm = occ;
  }
  int grid = csm_cdiv(rows, 8);
  const int cap = num_sms * ctas_per_sm;            // one resident wave; warps stride over the remaining rows
  if (grid > cap) grid = cap;
  return csm_launch_pdl(layernorm_fwd_kernel<VEC>, dim3(grid), dim3(256), 0, stream, x, gamma, beta,
                        reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32, mean, rstd, rows, D, eps);
}

extern "C" int csm_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16,
                                 float* out_f32, float* mean, float* rstd, int rows, int D, float eps,
                                 cudaStream_t stream) {
  CSM_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0 && D <= LN_MAX_VEC * 128,
                "csm_layernorm_fwd: D must be a multiple of 4 and <= %d (rows=%d D=%d)", LN_MAX_VEC * 128, rows, D);
  cudaError_t le;
  if (D <= 256) le = launch_ln_fwd<2>(x, gamma, beta, out_bf16, out_f32, mean, rstd, rows, D, eps, stream);
  else if (D <= 512) le = launch_ln_fwd<4>(x, gamma, beta, out_bf16, out_f32, mean, rstd, rows, D, eps, stream);
  else if (D <= 768) le = launch_ln_fwd<6>(x, gamma, beta, out_bf16, out_f32, mean, rstd, rows, D, eps, stream);
  else le = launch_ln_fwd<8>(x, gamma, beta, out_bf16, out_f32, mean, rstd, rows, D, eps, stream);
  if (le != cudaSuccess) {
    csm_set_error("layernorm_fwd: launch failed: %s", cudaGetErrorString(le));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

template <int TPR>
void launch_ln_bwd_tile(const void* dy_bf16, const float* dy2_f32, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* dres_in, float* dres_out, void* dres_bf16, float* dgamma,
                        float* dbeta, float* dcolsum, int rows, int num_sms, cudaStream_t stream) {
  constexpr int THREADS = ln_tile_threads(TPR);
  constexpr int RPI = THREADS / TPR;
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_bwd_tile_kernel<TPR>, THREADS, 0) != cudaSuccess ||
        occ < 1)
      occ = 1;
    ctas_per_sm = occ;
  }
  int grid = csm_cdiv(rows, RPI);
  const int cap = num_sms * ctas_per_sm;               // exactly one resident wave
  if (grid > cap) grid = cap;
  csm_launch_pdl(layernorm_bwd_tile_kernel<TPR>, dim3(grid), dim3(THREADS), 0, stream,
                 reinterpret_cast<const __nv_bfloat16*>(dy_bf16), dy2_f32, x, mean, rstd, gamma, dres_in, dres_out,
                 reinterpret_cast<__nv_bfloat16*>(dres_bf16), dgamma, dbeta, dcolsum, rows);
}

template <int TPR, int VPT>
void launch_ln_bwd_pipe(const void* dy_bf16, const float* dy2_f32, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* dres_in, float* dres_out, void* dres_bf16, float* dgamma,
                        float* dbeta, float* dcolsum, int rows, int num_sms, cudaStream_t stream) {
  constexpr int THREADS = ln_pipe_threads(TPR);
  constexpr int RPI = THREADS / TPR;
  constexpr size_t SMEM = ln_pipe_smem(TPR, VPT);
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    cudaFuncSetAttribute(layernorm_bwd_pipe_kernel<TPR, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    cudaFuncSetAttribute(layernorm_bwd_pipe_kernel<TPR, VPT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_bwd_pipe_kernel<TPR, VPT>, THREADS, SMEM) !=
            cudaSuccess || occ < 1)
      occ = 1;
    ctas_per_sm = occ;
  }
  int grid = csm_cdiv(rows, RPI);
  const int cap = num_sms * ctas_per_sm;               // exactly one resident wave
  if (grid > cap) grid = cap;
  csm_launch_pdl(layernorm_bwd_pipe_kernel<TPR, VPT>, dim3(grid), dim3(THREADS), SMEM, stream,
                 reinterpret_cast<const __nv_bfloat16*>(dy_bf16), dy2_f32, x, mean, rstd, gamma, dres_in, dres_out,
                 reinterpret_cast<__nv_bfloat16*>(dres_bf16), dgamma, dbeta, dcolsum, rows);
}

extern "C" int csm_layernorm_bwd(const void* dy_bf16, const float* dy2_f32, const float* x, const float* mean,
                                 const float* rstd, const float* gamma, const float* dres_in, float* dres_out,
                                 void* dres_bf16, float* dgamma, float* dbeta, float* dcolsum, int rows, int D,
                                 int num_sms, cudaStream_t stream) {
  CSM_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0 && D <= LN_MAX_VEC * 128,
                "csm_layernorm_bwd: D must be a multiple of 4 and <= %d (rows=%d D=%d)", LN_MAX_VEC * 128, rows, D);
  CSM_CHECK_ARG(dy_bf16 != nullptr || dy2_f32 != nullptr, "csm_layernorm_bwd: no incoming gradient");
  if (num_sms <= 0) num_sms = 148;
#define CSM_LN_TILE(TPR)                                                                                         \
  launch_ln_bwd_tile<TPR>(dy_bf16, dy2_f32, x, mean, rstd, gamma, dres_in, dres_out, dres_bf16, dgamma, dbeta, \
                          dcolsum, rows, num_sms, stream)
#define CSM_LN_PIPE(TPR, VPT)                                                                                    \
  launch_ln_bwd_pipe<TPR, VPT>(dy_bf16, dy2_f32, x, mean, rstd, gamma, dres_in, dres_out, dres_bf16, dgamma,    \
                               dbeta, dcolsum, rows, num_sms, stream)
  // the production widths stage rows with cp.async (deeper memory pipeline); other widths keep the register pipeline
  if (D == 512 || D == 768 || D == 1024) {
    if (D == 512) CSM_LN_PIPE(64, 2);
    else if (D == 768) CSM_LN_PIPE(96, 2);
    else CSM_LN_PIPE(128, 2);
    CSM_CHECK_LAUNCH("layernorm_bwd");
    return CSM_OK;
  }
  switch (D) {
    case 128: CSM_LN_TILE(32); break;
    case 256: CSM_LN_TILE(64); break;
    case 384: CSM_LN_TILE(96); break;
    default: {
      const int wpb = 8;
      int grid = csm_cdiv(rows, wpb);
      if (grid > num_sms * 2) grid = num_sms * 2;
      const size_t smem = static_cast<size_t>(3) * wpb * D * sizeof(float);
      static bool configured = false;
      if (!configured) {
        cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             3 * wpb * LN_MAX_VEC * 128 * (int)sizeof(float));
        configured = true;
      }
      layernorm_bwd_kernel<<<grid, wpb * 32, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy_bf16), dy2_f32,
                                                             x, mean, rstd, gamma, dres_in, dres_out,
                                                             reinterpret_cast<__nv_bfloat16*>(dres_bf16), dgamma,
                                                             dbeta, dcolsum, rows, D);
    }
  }
#undef CSM_LN_TILE
#undef CSM_LN_PIPE
  CSM_CHECK_LAUNCH("layernorm_bwd");
  return CSM_OK;
}

extern "C" int csm_colsum_bf16(const void* dy_bf16, float* db, int rows, int N, int skip_period, int num_sms,
                               cudaStream_t stream) {
  CSM_CHECK_ARG(rows > 0 && N > 0 && N % 8 == 0, "csm_colsum_bf16: N must be a multiple of 8 (rows=%d N=%d)", rows, N);
  if (num_sms <= 0) num_sms = 148;
  const int gx = csm_cdiv(N, 256);
  int gy = csm_cdiv(6 * num_sms, gx);       // ~6 CTAs of 256 threads per SM: enough 16-byte loads in flight for HBM
  const int max_gy = csm_cdiv(rows, 8);
  if (gy > max_gy) gy = max_gy;
  cudaError_t le = csm_launch_pdl(colsum_bf16_kernel, dim3(gx, gy), dim3(256), 0, stream,
                                  reinterpret_cast<const __nv_bfloat16*>(dy_bf16), db, rows, N, skip_period);
  if (le != cudaSuccess) {
    csm_set_error("colsum_bf16: launch failed: %s", cudaGetErrorString(le));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

// table: device array of {const float* src, bf16* dst, int64 n} (24 bytes per entry), every pointer 16B aligned.
extern "C" int csm_cast_multi(const void* table_dev, int num_tensors, int blocks_per_tensor, cudaStream_t stream) {
  CSM_CHECK_ARG(num_tensors > 0 && blocks_per_tensor > 0, "csm_cast_multi: empty table");
  cast_multi_kernel<<<dim3(blocks_per_tensor, num_tensors), 256, 0, stream>>>(
      reinterpret_cast<const CastEntry*>(table_dev));
  CSM_CHECK_LAUNCH("cast_multi");
  return CSM_OK;
}

extern "C" int csm_cast_f32_bf16(const float* src, void* dst_bf16, long long n, cudaStream_t stream) {
  CSM_CHECK_ARG(n > 0, "csm_cast_f32_bf16: empty tensor");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), n);
  CSM_CHECK_LAUNCH("cast_f32_bf16");
  return CSM_OK;
}
