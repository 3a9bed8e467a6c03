// Loss heads of Cross-Scale MAE as HBM-bound reduction kernels (forward and backward):
//   * masked per-patch pixel MSE      models_mae/MAE_ViT_Shared.py:24-39,97-120,269-290
//   * cross-scale decoder MSE         models_mae/MAE_ViT_MsLdCeCd.py:57-59 (target NOT detached)
//   * BatchNorm1d over the patch axis models_mae/MLP.py:4-10 (channel = patch index, train mode)
//   * NT-Xent contrastive loss        util/contrast_loss.py:17-41,71-101; MAE_ViT_MsLdCeCd.py:62-69
// The image is read in place in patch order (no materialised target), every kernel accumulates its
// scalar with one atomicAdd per CTA/warp, and nothing here synchronises with the host.
#include "common.cuh"

namespace {
using namespace csm;

__device__ __forceinline__ float block_sum(float v, float* s_buf) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) s_buf[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += s_buf[i];
  return t;
}

// ---------------------------------------------------------------------------------------------
// pixel reconstruction loss.  One warp per (image, patch); only masked patches do work.
// pred_full is the decoder_pred output for ALL Sd = L+1 rows (row 0 = cls, ignored).
// mode 0: loss_sum += mean_e (bf16(pred - bf16(img)))^2            (forward)
// mode 1: dpred = g * coef * 2 * diff (masked) / 0 (unmasked, cls)  (backward; coef = 1 / (P * #masked))
// ---------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void recon_loss_kernel(const __nv_bfloat16* __restrict__ pred_full, const float* __restrict__ imgs,
                                  const float* __restrict__ mask, float* __restrict__ loss_sum,
                                  __nv_bfloat16* __restrict__ dpred, const float* __restrict__ gptr, float coef,
                                  int nimg, int C, int H, int p, int L, int norm_pix) {
  const int warps = blockDim.x >> 5;
  const int item = blockIdx.x * warps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int Sd = L + 1;
  if (item >= nimg * Sd) return;
  const int n = item / Sd, tok = item % Sd;
  const int P = C * p * p;
  const size_t prow = static_cast<size_t>(item) * P;
  const bool active = tok > 0 && mask[static_cast<size_t>(n) * L + tok - 1] != 0.f;
  if (!active) {
    if (BWD)
      for (int i = lane * 8; i < P; i += 256) *reinterpret_cast<uint4*>(dpred + prow + i) = make_uint4(0, 0, 0, 0);
    return;
  }
  const int G = H / p;
  const int l = tok - 1, ph = l / G, pw = l % G;
  const float* ibase = imgs + static_cast<size_t>(n) * C * H * H + static_cast<size_t>(ph) * p * H + pw * p;
  float mean = 0.f, inv_std = 1.f;
  if (norm_pix) {
    float s = 0.f;
    for (int i = lane; i < P; i += 32) {
      const int px = i % p, cy = i / p;
      s += ibase[(static_cast<size_t>(cy / p) * H + (cy % p)) * H + px];
    }
    mean = warp_sum(s) / P;
    float ss = 0.f;
    for (int i = lane; i < P; i += 32) {
      const int px = i % p, cy = i / p;
      const float d = ibase[(static_cast<size_t>(cy / p) * H + (cy % p)) * H + px] - mean;
      ss += d * d;
    }
    inv_std = rsqrtf(warp_sum(ss) / (P - 1) + 1.0e-6f);
  }
  const float g = BWD ? gptr[0] * coef * 2.0f : 0.f;
  float acc = 0.f;
  if (p == 16 && C <= 4) {
    // 16 x 16 patches (every config of BASELINE.json): a warp covers two pixel rows per step, no index divisions,
    // all 8 x C steps unrolled so the loads of the whole patch are in flight together
    const int px = lane & 15, half = lane >> 4;
    float pix[4][8], prd[4][8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int py = 2 * k + half;
          pix[c][k] = ibase[(static_cast<size_t>(c) * H + py) * H + px];
          prd[c][k] = __bfloat162float(pred_full[prow + (py * 16 + px) * C + c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int py = 2 * k + half;
          float diff;
          if (norm_pix) diff = prd[c][k] - (pix[c][k] - mean) * inv_std;
          else diff = bf16_round(prd[c][k] - bf16_round(pix[c][k]));
          if (BWD) dpred[prow + (py * 16 + px) * C + c] = __float2bfloat16_rn(g * diff);
          else acc += diff * diff;
        }
      }
    }
  } else
  // i runs in image order (c, py, px) so pixel reads are contiguous; e is the (py, px, c) slot of pred
  for (int i = lane; i < P; i += 32) {
    const int px = i % p, cy = i / p;
    const int py = cy % p, c = cy / p;
    const float pix = ibase[(static_cast<size_t>(c) * H + py) * H + px];
    const int e = (py * p + px) * C + c;
    const float pr = __bfloat162float(pred_full[prow + e]);
    float diff;
    if (norm_pix) diff = pr - (pix - mean) * inv_std;
    else diff = bf16_round(pr - bf16_round(pix));
    if (BWD) dpred[prow + e] = __float2bfloat16_rn(g * diff);
    else acc += diff * diff;
  }
  if (!BWD) {
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(loss_sum, acc / P);
  }
}

// ---------------------------------------------------------------------------------------------
// cross-scale decoder loss: sum over (n, l >= 1, d) of (float(cp) - tgt)^2   (cls rows skipped)
// backward: d_cp = bf16(g * coef * 2 * diff), d_tgt = -g * coef * 2 * diff (f32); cls rows zero.
// ---------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void cross_mse_kernel(const __nv_bfloat16* __restrict__ cp, const float* __restrict__ tgt,
                                 float* __restrict__ loss_sum, __nv_bfloat16* __restrict__ d_cp,
                                 float* __restrict__ d_tgt, const float* __restrict__ gptr, float coef, int rows,
                                 int Sd, int Dd) {
  __shared__ float s_buf[32];
  const float g = BWD ? gptr[0] * coef * 2.0f : 0.f;
  float acc = 0.f;
  const long long total4 = static_cast<long long>(rows) * Dd / 4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total4;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = idx * 4;
    const int row = static_cast<int>(e / Dd);
    const bool is_cls = (row % Sd) == 0;
    if (is_cls) {
      if (BWD) {
        *reinterpret_cast<uint2*>(d_cp + e) = make_uint2(0, 0);
        *reinterpret_cast<float4*>(d_tgt + e) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    const uint2 cv = *reinterpret_cast<const uint2*>(cp + e);
    const float2 c0 = unpack_bf16x2(cv.x), c1 = unpack_bf16x2(cv.y);
    const float4 t = *reinterpret_cast<const float4*>(tgt + e);
    const float d0 = c0.x - t.x, d1 = c0.y - t.y, d2 = c1.x - t.z, d3 = c1.y - t.w;
    if (BWD) {
      uint2 pk;
      pk.x = pack_bf16x2(g * d0, g * d1);
      pk.y = pack_bf16x2(g * d2, g * d3);
      *reinterpret_cast<uint2*>(d_cp + e) = pk;
      *reinterpret_cast<float4*>(d_tgt + e) = make_float4(-g * d0, -g * d1, -g * d2, -g * d3);
    } else {
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
  if (!BWD) {
    acc = block_sum(acc, s_buf);
    if (threadIdx.x == 0) atomicAdd(loss_sum, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d(L) over [N, L, Hp] with channel = patch index l; rows are laid out n*Sd + 1 + l
// (cls rows of the predictor GEMM output are ignored / zeroed).  One CTA per channel.
// ---------------------------------------------------------------------------------------------
// (512-thread CTAs, two per SM and all 196 channels in one wave, were measured equal: a channel's three dependent
// passes bound the CTA, not the wave count; the next step would be several CTAs per channel)
__global__ void bn_patch_fwd_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                    float* __restrict__ running_mean, float* __restrict__ running_var, int N, int Sd,
                                    int Hp, float eps, float momentum, int training) {
  __shared__ float s_buf[32];
  const int l = blockIdx.x;
  const int per_row = Hp / 8;
  const int total = N * per_row;
  float mean, rstd;
  if (training) {
  float s = 0.f;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = idx / per_row, c = (idx % per_row) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(h + (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s += f.x + f.y;
    }
  }
  const float cnt = static_cast<float>(N) * Hp;
  mean = block_sum(s, s_buf) / cnt;
  float ss = 0.f;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = idx / per_row, c = (idx % per_row) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(h + (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      ss += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
    }
  }
  const float var = block_sum(ss, s_buf) / cnt;
  rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    mean_out[l] = mean;
    rstd_out[l] = rstd;
    if (running_mean != nullptr) {
      running_mean[l] = (1.f - momentum) * running_mean[l] + momentum * mean;
      running_var[l] = (1.f - momentum) * running_var[l] + momentum * var * (cnt / (cnt - 1.f));
    }
  }
  } else {  // eval mode: normalise with the running statistics (kept for the backward, which then uses
            // dh = gamma * rstd * dy: the statistics are constants)
    mean = running_mean[l];
    rstd = rsqrtf(running_var[l] + eps);
    if (threadIdx.x == 0) {
      mean_out[l] = mean;
      rstd_out[l] = rstd;
    }
  }
  const float a = rstd * gamma[l], b = beta[l] - mean * rstd * gamma[l];
#pragma unroll 4
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = idx / per_row, c = (idx % per_row) * 8;
    const size_t off = (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c;
    const uint4 v = *reinterpret_cast<const uint4*>(h + off);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      o[j] = pack_bf16x2(fmaxf(bf16_round(f.x * a + b), 0.f), fmaxf(bf16_round(f.y * a + b), 0.f));
    }
    *reinterpret_cast<uint4*>(out + off) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  // cls row of image n: zero (channel 0's CTA does it)
  if (l == 0) {
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int n = idx / per_row, c = (idx % per_row) * 8;
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(n) * Sd * Hp + c) = make_uint4(0, 0, 0, 0);
    }
  }
}

// dy = d_out * (out > 0);  dh = gamma * rstd * (dy - mean(dy) - xhat * mean(dy * xhat))
__global__ void bn_patch_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ out,
                                    const __nv_bfloat16* __restrict__ d_out, const float* __restrict__ gamma,
                                    const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                    __nv_bfloat16* __restrict__ dh, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int N, int Sd, int Hp, int training) {
  __shared__ float s_buf[32];
  const int l = blockIdx.x;
  const int per_row = Hp / 8;
  const int total = N * per_row;
  const float mean = mean_in[l], rstd = rstd_in[l];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = idx / per_row, c = (idx % per_row) * 8;
    const size_t off = (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c;
    const uint4 hv = *reinterpret_cast<const uint4*>(h + off);
    const uint4 ov = *reinterpret_cast<const uint4*>(out + off);
    const uint4 dv = *reinterpret_cast<const uint4*>(d_out + off);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 hf = unpack_bf16x2(hw[j]), of = unpack_bf16x2(ow[j]), df = unpack_bf16x2(dw[j]);
      const float dy0 = of.x > 0.f ? df.x : 0.f, dy1 = of.y > 0.f ? df.y : 0.f;
      s1 += dy0 + dy1;
      s2 += dy0 * (hf.x - mean) * rstd + dy1 * (hf.y - mean) * rstd;
    }
  }
  s1 = block_sum(s1, s_buf);
  s2 = block_sum(s2, s_buf);
  if (threadIdx.x == 0) {
    dgamma[l] = s2;
    dbeta[l] = s1;
  }
  const float cnt = static_cast<float>(N) * Hp;
  // eval mode (running statistics are constants): dh = gamma * rstd * dy, no batch-statistic terms
  const float m1 = training ? s1 / cnt : 0.f, m2 = training ? s2 / cnt : 0.f, gr = gamma[l] * rstd;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = idx / per_row, c = (idx % per_row) * 8;
    const size_t off = (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c;
    const uint4 hv = *reinterpret_cast<const uint4*>(h + off);
    const uint4 ov = *reinterpret_cast<const uint4*>(out + off);
    const uint4 dv = *reinterpret_cast<const uint4*>(d_out + off);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 hf = unpack_bf16x2(hw[j]), of = unpack_bf16x2(ow[j]), df = unpack_bf16x2(dw[j]);
      const float dy0 = of.x > 0.f ? df.x : 0.f, dy1 = of.y > 0.f ? df.y : 0.f;
      r[j] = pack_bf16x2(gr * (dy0 - m1 - (hf.x - mean) * rstd * m2), gr * (dy1 - m1 - (hf.y - mean) * rstd * m2));
    }
    *reinterpret_cast<uint4*>(dh + off) = make_uint4(r[0], r[1], r[2], r[3]);
  }
  if (l == 0) {
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int n = idx / per_row, c = (idx % per_row) * 8;
      *reinterpret_cast<uint4*>(dh + static_cast<size_t>(n) * Sd * Hp + c) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same BatchNorm with a channel spread over a cluster of BN_CL CTAs (training mode): every CTA keeps its
// N / BN_CL rows of the channel in shared memory (ONE global read), the per-channel sums cross the cluster through
// distributed shared memory, and the three passes of the single-CTA kernels become passes over shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int BN_CL = 8;

__device__ __forceinline__ uint32_t bn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void bn_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// sum over the cluster's CTAs of the float at the same shared-memory location
__device__ __forceinline__ float bn_cluster_sum(const float* slot) {
  float t = 0.f;
#pragma unroll
  for (uint32_t r = 0; r < BN_CL; ++r) {
    uint32_t a;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(slot)), "r"(r));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    t += v;
  }
  return t;
}

__global__ void __cluster_dims__(BN_CL, 1, 1) __launch_bounds__(256)
bn_patch_fwd_cluster_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ gamma,
                            const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                            float* __restrict__ running_mean, float* __restrict__ running_var, int N, int Sd, int Hp,
                            float eps, float momentum) {
  extern __shared__ uint4 s_rows[];                   // [rows of this CTA][Hp / 8]
  __shared__ float s_buf[32];
  __shared__ float s_x[2];
  const int l = blockIdx.x / BN_CL;
  const int part = static_cast<int>(bn_cluster_rank());
  const int rows_per = (N + BN_CL - 1) / BN_CL;
  const int n0 = part * rows_per;
  const int nrows = max(0, min(N, n0 + rows_per) - n0);
  const int per_row = Hp / 8;
  const int total = nrows * per_row;
  float s = 0.f;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = n0 + idx / per_row, c = (idx % per_row) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(h + (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c);
    s_rows[idx] = v;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s += f.x + f.y;
    }
  }
  s = block_sum(s, s_buf);
  if (threadIdx.x == 0) s_x[0] = s;
  bn_cluster_sync();
  const float cnt = static_cast<float>(N) * Hp;
  const float mean = bn_cluster_sum(&s_x[0]) / cnt;
  float ss = 0.f;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const uint4 v = s_rows[idx];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      ss += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
    }
  }
  ss = block_sum(ss, s_buf);
  if (threadIdx.x == 0) s_x[1] = ss;
  bn_cluster_sync();
  const float var = bn_cluster_sum(&s_x[1]) / cnt;
  const float rstd = rsqrtf(var + eps);
  if (part == 0 && threadIdx.x == 0) {
    mean_out[l] = mean;
    rstd_out[l] = rstd;
    if (running_mean != nullptr) {
      running_mean[l] = (1.f - momentum) * running_mean[l] + momentum * mean;
      running_var[l] = (1.f - momentum) * running_var[l] + momentum * var * (cnt / (cnt - 1.f));
    }
  }
  const float a = rstd * gamma[l], b = beta[l] - mean * rstd * gamma[l];
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = n0 + idx / per_row, c = (idx % per_row) * 8;
    const uint4 v = s_rows[idx];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      o[j] = pack_bf16x2(fmaxf(bf16_round(f.x * a + b), 0.f), fmaxf(bf16_round(f.y * a + b), 0.f));
    }
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (l == 0) {                                       // cls rows of this CTA's images: zero
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int n = n0 + idx / per_row, c = (idx % per_row) * 8;
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(n) * Sd * Hp + c) = make_uint4(0, 0, 0, 0);
    }
  }
  bn_cluster_sync();                                  // no CTA leaves while a peer may still read its s_x
}

__global__ void __cluster_dims__(BN_CL, 1, 1) __launch_bounds__(256)
bn_patch_bwd_cluster_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ out,
                            const __nv_bfloat16* __restrict__ d_out, const float* __restrict__ gamma,
                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                            __nv_bfloat16* __restrict__ dh, float* __restrict__ dgamma, float* __restrict__ dbeta,
                            int N, int Sd, int Hp) {
  extern __shared__ uint4 s_rows[];                   // [2][rows of this CTA][Hp / 8]: h, dy = d_out * (out > 0)
  __shared__ float s_buf[32];
  __shared__ float s_x[2];
  const int l = blockIdx.x / BN_CL;
  const int part = static_cast<int>(bn_cluster_rank());
  const int rows_per = (N + BN_CL - 1) / BN_CL;
  const int n0 = part * rows_per;
  const int nrows = max(0, min(N, n0 + rows_per) - n0);
  const int per_row = Hp / 8;
  const int total = nrows * per_row;
  uint4* s_h = s_rows;
  uint4* s_dy = s_rows + rows_per * per_row;
  const float mean = mean_in[l], rstd = rstd_in[l];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = n0 + idx / per_row, c = (idx % per_row) * 8;
    const size_t off = (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c;
    const uint4 hv = *reinterpret_cast<const uint4*>(h + off);
    const uint4 ov = *reinterpret_cast<const uint4*>(out + off);
    const uint4 dv = *reinterpret_cast<const uint4*>(d_out + off);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t dyw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 hf = unpack_bf16x2(hw[j]), of = unpack_bf16x2(ow[j]), df = unpack_bf16x2(dw[j]);
      const float dy0 = of.x > 0.f ? df.x : 0.f, dy1 = of.y > 0.f ? df.y : 0.f;
      s1 += dy0 + dy1;
      s2 += dy0 * (hf.x - mean) * rstd + dy1 * (hf.y - mean) * rstd;
      dyw[j] = pack_bf16x2(dy0, dy1);                 // exact: dy is d_out or 0
    }
    s_h[idx] = hv;
    s_dy[idx] = make_uint4(dyw[0], dyw[1], dyw[2], dyw[3]);
  }
  s1 = block_sum(s1, s_buf);
  s2 = block_sum(s2, s_buf);
  if (threadIdx.x == 0) {
    s_x[0] = s1;
    s_x[1] = s2;
  }
  bn_cluster_sync();
  s1 = bn_cluster_sum(&s_x[0]);
  s2 = bn_cluster_sum(&s_x[1]);
  if (part == 0 && threadIdx.x == 0) {
    dgamma[l] = s2;
    dbeta[l] = s1;
  }
  const float cnt = static_cast<float>(N) * Hp;
  const float m1 = s1 / cnt, m2 = s2 / cnt, gr = gamma[l] * rstd;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = n0 + idx / per_row, c = (idx % per_row) * 8;
    const uint4 hv = s_h[idx], dv = s_dy[idx];
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 hf = unpack_bf16x2(hw[j]), df = unpack_bf16x2(dw[j]);
      r[j] = pack_bf16x2(gr * (df.x - m1 - (hf.x - mean) * rstd * m2), gr * (df.y - m1 - (hf.y - mean) * rstd * m2));
    }
    *reinterpret_cast<uint4*>(dh + (static_cast<size_t>(n) * Sd + 1 + l) * Hp + c) = make_uint4(r[0], r[1], r[2], r[3]);
  }
  if (l == 0) {
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int n = n0 + idx / per_row, c = (idx % per_row) * 8;
      *reinterpret_cast<uint4*>(dh + static_cast<size_t>(n) * Sd * Hp + c) = make_uint4(0, 0, 0, 0);
    }
  }
  bn_cluster_sync();
}

// ---------------------------------------------------------------------------------------------
// NT-Xent.  feat[n] = mean over tokens 1.. of the un-normed encoder output, z = normalize(feat),
// cos re-normalises z (CosineSimilarity), E = exp(cos / tau), positives (i, i +- B),
// loss = mean_i -log(E_pos / (sum_{j not in {i, pos}} E_ij + eps)).
// ---------------------------------------------------------------------------------------------
__global__ void token_mean_normalize_kernel(const float* __restrict__ x, float* __restrict__ zhat,
                                            float* __restrict__ fnorm, int Se, int D) {
  __shared__ float s_buf[32];
  const int n = blockIdx.x;
  const float* xr = x + static_cast<size_t>(n) * Se * D;
  float local[4];  // D <= 4 * blockDim.x
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * blockDim.x;
    float a = 0.f;
    if (i < D) {
#pragma unroll 8
      for (int t = 1; t < Se; ++t) a += xr[static_cast<size_t>(t) * D + i];
      a /= static_cast<float>(Se - 1);
    }
    local[k] = a;
    ss += a * a;
  }
  const float norm = sqrtf(block_sum(ss, s_buf));
  const float d1 = fmaxf(norm, 1e-12f);           // F.normalize eps
  float ss2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    local[k] /= d1;
    ss2 += local[k] * local[k];
  }
  const float n2 = fmaxf(sqrtf(block_sum(ss2, s_buf)), 1e-8f);  // CosineSimilarity eps
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * blockDim.x;
    if (i < D) zhat[static_cast<size_t>(n) * D + i] = local[k] / n2;
  }
  if (threadIdx.x == 0) fnorm[n] = d1 * n2;
}

// Same result with the tokens spread over RL row lanes of D/4 float4 column groups (blockDim = RL * D/4, a multiple
// of 32): every thread has its ~12 independent 16-byte loads in flight at once instead of walking 49 tokens per
// column (27 us -> latency of one round).
__global__ void token_mean_normalize_v4_kernel(const float* __restrict__ x, float* __restrict__ zhat,
                                               float* __restrict__ fnorm, int Se, int D) {
  extern __shared__ float4 s_acc[];                 // [RL][D/4]
  __shared__ float s_buf[32];
  const int n = blockIdx.x;
  const int vec = D >> 2;
  const int cg = threadIdx.x % vec, rl = threadIdx.x / vec, RL = blockDim.x / vec;
  const float* xr = x + static_cast<size_t>(n) * Se * D + cg * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int t = 1 + rl; t < Se; t += RL) {
    const float4 v = *reinterpret_cast<const float4*>(xr + static_cast<size_t>(t) * D);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  s_acc[rl * vec + cg] = a;
  __syncthreads();
  float ss = 0.f;
  if (rl == 0) {
    for (int k = 1; k < RL; ++k) {
      const float4 v = s_acc[k * vec + cg];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const float inv = 1.0f / static_cast<float>(Se - 1);
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  const float norm = sqrtf(block_sum(ss, s_buf));
  const float d1 = fmaxf(norm, 1e-12f);           // F.normalize eps
  float ss2 = 0.f;
  if (rl == 0) {
    a.x /= d1; a.y /= d1; a.z /= d1; a.w /= d1;
    ss2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  const float n2 = fmaxf(sqrtf(block_sum(ss2, s_buf)), 1e-8f);  // CosineSimilarity eps
  if (rl == 0)
    *reinterpret_cast<float4*>(zhat + static_cast<size_t>(n) * D + cg * 4) =
        make_float4(a.x / n2, a.y / n2, a.z / n2, a.w / n2);
  if (threadIdx.x == 0) fnorm[n] = d1 * n2;
}

// One CTA (4 warps) per row i.  E row kept in smem.  mode fwd: neg[i], loss_sum += loss_i / (2B).
__global__ void ntxent_fwd_kernel(const float* __restrict__ zhat, float* __restrict__ neg_out,
                                  float* __restrict__ loss_sum, int B, int D, float inv_tau, float eps) {
  extern __shared__ float s_e[];  // [2B]
  const int i = blockIdx.x, n2 = 2 * B;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* zi = zhat + static_cast<size_t>(i) * D;
  for (int j0 = w * 4; j0 < n2; j0 += nw * 4) {        // four independent dot products in flight per warp
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 4; k < D; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(zi + k);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < n2) {
          const float4 b = *reinterpret_cast<const float4*>(zhat + static_cast<size_t>(j0 + u) * D + k);
          d[u] += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float du = warp_sum(d[u]);
      if (lane == 0 && j0 + u < n2) s_e[j0 + u] = __expf(du * inv_tau);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int pos = (i + B) % n2;
    float neg = 0.f;
    for (int j = 0; j < n2; ++j)
      if (j != i && j != pos) neg += s_e[j];
    neg_out[i] = neg;
    atomicAdd(loss_sum, -logf(s_e[pos] / (neg + eps)) / static_cast<float>(n2));
  }
}

// d_feat[i] = g * (v - (zhat_i . v) zhat_i) / fnorm_i,  v = (1/tau) * sum_j (G_ij + G_ji) zhat_j
__global__ void ntxent_bwd_kernel(const float* __restrict__ zhat, const float* __restrict__ fnorm,
                                  const float* __restrict__ neg, const float* __restrict__ gptr,
                                  float* __restrict__ d_feat, int B, int D, float inv_tau, float eps) {
  extern __shared__ float s_w[];  // [2B] weights, then [32] reduction scratch
  const int i = blockIdx.x, n2 = 2 * B;
  float* s_buf = s_w + n2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* zi = zhat + static_cast<size_t>(i) * D;
  const int pos_i = (i + B) % n2;
  const float inv_n = 1.f / static_cast<float>(n2);
  for (int j0 = w * 4; j0 < n2; j0 += nw * 4) {
    float dd[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 4; k < D; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(zi + k);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < n2) {
          const float4 b = *reinterpret_cast<const float4*>(zhat + static_cast<size_t>(j0 + u) * D + k);
          dd[u] += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
    const int j = j0 + u;
    const float d = warp_sum(dd[u]);
    if (lane == 0 && j < n2) {
      const float e = __expf(d * inv_tau);
      float gij = 0.f, gji = 0.f;
      if (j == pos_i) {
        gij = -1.f;   // j is i's positive, and i is j's positive
        gji = -1.f;
      } else if (j != i) {
        gij = e / (neg[i] + eps);
        gji = e / (neg[j] + eps);
      }
      s_w[j] = (gij + gji) * inv_n * inv_tau;
    }
    }
  }
  __syncthreads();
  float v[4], dot = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = threadIdx.x + k * blockDim.x;
    float a = 0.f;
    if (c < D) {
#pragma unroll 8
      for (int j = 0; j < n2; ++j) a += s_w[j] * zhat[static_cast<size_t>(j) * D + c];
      dot += a * zi[c];
    }
    v[k] = a;
  }
  dot = block_sum(dot, s_buf);
  const float sc = gptr[0] / fnorm[i];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = threadIdx.x + k * blockDim.x;
    if (c < D) d_feat[static_cast<size_t>(i) * D + c] = sc * (v[k] - dot * zi[c]);
  }
}

}  // namespace

extern "C" int csm_recon_loss_fwd(const void* pred_full_bf16, const float* imgs, const float* mask, float* loss_sum,
                                  int nimg, int C, int H, int p, int L, int norm_pix, cudaStream_t stream) {
  CSM_CHECK_ARG(nimg > 0 && H % p == 0 && (H / p) * (H / p) == L, "csm_recon_loss_fwd: bad geometry H=%d p=%d L=%d", H,
                p, L);
  const int items = nimg * (L + 1), wpb = 8;
  recon_loss_kernel<false><<<csm_cdiv(items, wpb), wpb * 32, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(pred_full_bf16), imgs, mask, loss_sum, nullptr, nullptr, 0.f, nimg, C, H,
      p, L, norm_pix);
  CSM_CHECK_LAUNCH("recon_loss_fwd");
  return CSM_OK;
}

extern "C" int csm_recon_loss_bwd(const void* pred_full_bf16, const float* imgs, const float* mask, void* dpred_bf16,
                                  const float* grad_scalar, float coef, int nimg, int C, int H, int p, int L,
                                  int norm_pix, cudaStream_t stream) {
  CSM_CHECK_ARG(nimg > 0 && H % p == 0 && (H / p) * (H / p) == L, "csm_recon_loss_bwd: bad geometry H=%d p=%d L=%d", H,
                p, L);
  CSM_CHECK_ARG((C * p * p) % 8 == 0, "csm_recon_loss_bwd: patch dim must be a multiple of 8");
  const int items = nimg * (L + 1), wpb = 8;
  recon_loss_kernel<true><<<csm_cdiv(items, wpb), wpb * 32, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(pred_full_bf16), imgs, mask, nullptr,
      reinterpret_cast<__nv_bfloat16*>(dpred_bf16), grad_scalar, coef, nimg, C, H, p, L, norm_pix);
  CSM_CHECK_LAUNCH("recon_loss_bwd");
  return CSM_OK;
}

// loss = sum_i acc[i] * coef[i]: the scalar the step returns, from the accumulator array every loss kernel adds into
// (n <= 32 terms; one warp)
__global__ void loss_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ coef,
                                     float* __restrict__ loss, int n) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x;
  float v = lane < n ? acc[lane] * coef[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) *loss = v;
}

extern "C" int csm_loss_finalize(const float* loss_terms, const float* coefs, float* loss, int n, cudaStream_t stream) {
  CSM_CHECK_ARG(n > 0 && n <= 32, "csm_loss_finalize: 1..32 terms (n=%d)", n);
  cudaError_t e = csm_launch_pdl(loss_finalize_kernel, dim3(1), dim3(32), 0, stream, loss_terms, coefs, loss, n);
  if (e != cudaSuccess) {
    csm_set_error("csm_loss_finalize: launch failed: %s", cudaGetErrorString(e));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

extern "C" int csm_cross_mse_fwd(const void* cp_bf16, const float* tgt, float* loss_sum, int rows, int Sd, int Dd,
                                 cudaStream_t stream) {
  CSM_CHECK_ARG(rows > 0 && Dd % 4 == 0, "csm_cross_mse_fwd: bad sizes rows=%d Dd=%d", rows, Dd);
  int blocks = csm_cdiv(static_cast<long long>(rows) * Dd / 4, 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  cross_mse_kernel<false><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(cp_bf16), tgt, loss_sum,
                                                      nullptr, nullptr, nullptr, 0.f, rows, Sd, Dd);
  CSM_CHECK_LAUNCH("cross_mse_fwd");
  return CSM_OK;
}

extern "C" int csm_cross_mse_bwd(const void* cp_bf16, const float* tgt, void* d_cp_bf16, float* d_tgt,
                                 const float* grad_scalar, float coef, int rows, int Sd, int Dd, cudaStream_t stream) {
  CSM_CHECK_ARG(rows > 0 && Dd % 4 == 0, "csm_cross_mse_bwd: bad sizes rows=%d Dd=%d", rows, Dd);
  int blocks = csm_cdiv(static_cast<long long>(rows) * Dd / 4, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cross_mse_kernel<true><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(cp_bf16), tgt, nullptr,
                                                     reinterpret_cast<__nv_bfloat16*>(d_cp_bf16), d_tgt, grad_scalar,
                                                     coef, rows, Sd, Dd);
  CSM_CHECK_LAUNCH("cross_mse_bwd");
  return CSM_OK;
}

extern "C" int csm_bn_patch_fwd(const void* h_bf16, const float* gamma, const float* beta, void* out_bf16, float* mean,
                                float* rstd, float* running_mean, float* running_var, int N, int L, int Hp, float eps,
                                float momentum, int training, cudaStream_t stream) {
  CSM_CHECK_ARG(N > 0 && L > 0 && Hp % 8 == 0, "csm_bn_patch_fwd: bad sizes N=%d L=%d Hp=%d", N, L, Hp);
  CSM_CHECK_ARG(training || (running_mean != nullptr && running_var != nullptr),
                "csm_bn_patch_fwd: eval mode needs running statistics");
  const size_t cl_smem = static_cast<size_t>((N + BN_CL - 1) / BN_CL) * Hp * 2;
  if (training && N >= BN_CL && cl_smem <= 96 * 1024) {
    static size_t configured = 0;
    if (cl_smem > configured) {
      cudaFuncSetAttribute(bn_patch_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl_smem);
      configured = cl_smem;
    }
    bn_patch_fwd_cluster_kernel<<<L * BN_CL, 256, cl_smem, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(h_bf16), gamma, beta, reinterpret_cast<__nv_bfloat16*>(out_bf16), mean,
        rstd, running_mean, running_var, N, L + 1, Hp, eps, momentum);
    CSM_CHECK_LAUNCH("bn_patch_fwd");
    return CSM_OK;
  }
  bn_patch_fwd_kernel<<<L, 1024, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(h_bf16), gamma, beta,
                                             reinterpret_cast<__nv_bfloat16*>(out_bf16), mean, rstd, running_mean,
                                             running_var, N, L + 1, Hp, eps, momentum, training);
  CSM_CHECK_LAUNCH("bn_patch_fwd");
  return CSM_OK;
}

extern "C" int csm_bn_patch_bwd(const void* h_bf16, const void* out_bf16, const void* d_out_bf16, const float* gamma,
                                const float* mean, const float* rstd, void* dh_bf16, float* dgamma, float* dbeta,
                                int N, int L, int Hp, int training, cudaStream_t stream) {
  CSM_CHECK_ARG(N > 0 && L > 0 && Hp % 8 == 0, "csm_bn_patch_bwd: bad sizes N=%d L=%d Hp=%d", N, L, Hp);
  const size_t cl_smem = static_cast<size_t>((N + BN_CL - 1) / BN_CL) * Hp * 2 * 2;
  if (training && N >= BN_CL && cl_smem <= 96 * 1024) {
    static size_t configured = 0;
    if (cl_smem > configured) {
      cudaFuncSetAttribute(bn_patch_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl_smem);
      configured = cl_smem;
    }
    bn_patch_bwd_cluster_kernel<<<L * BN_CL, 256, cl_smem, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(h_bf16), reinterpret_cast<const __nv_bfloat16*>(out_bf16),
        reinterpret_cast<const __nv_bfloat16*>(d_out_bf16), gamma, mean, rstd, reinterpret_cast<__nv_bfloat16*>(dh_bf16),
        dgamma, dbeta, N, L + 1, Hp);
    CSM_CHECK_LAUNCH("bn_patch_bwd");
    return CSM_OK;
  }
  bn_patch_bwd_kernel<<<L, 1024, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(h_bf16), reinterpret_cast<const __nv_bfloat16*>(out_bf16),
      reinterpret_cast<const __nv_bfloat16*>(d_out_bf16), gamma, mean, rstd, reinterpret_cast<__nv_bfloat16*>(dh_bf16),
      dgamma, dbeta, N, L + 1, Hp, training);
  CSM_CHECK_LAUNCH("bn_patch_bwd");
  return CSM_OK;
}

extern "C" int csm_ntxent_fwd(const float* enc_out, float* zhat, float* fnorm, float* neg, float* loss_sum, int B,
                              int Se, int D, float tau, float eps, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && Se >= 2 && D > 0 && D <= 1024 && D % 4 == 0, "csm_ntxent_fwd: bad sizes B=%d Se=%d D=%d", B, Se,
                D);
  const int vec = D / 4;
  int RL = 1024 / vec;
  if (RL > 4) RL = 4;
  while (RL > 1 && (RL * vec) % 32 != 0) --RL;
  if ((RL * vec) % 32 == 0 && (reinterpret_cast<uintptr_t>(enc_out) & 15) == 0)
    token_mean_normalize_v4_kernel<<<2 * B, RL * vec, static_cast<size_t>(RL) * vec * sizeof(float4), stream>>>(
        enc_out, zhat, fnorm, Se, D);
  else
    token_mean_normalize_kernel<<<2 * B, 256, 0, stream>>>(enc_out, zhat, fnorm, Se, D);
  CSM_CHECK_LAUNCH("token_mean_normalize");
  // 16 warps per row: 2B = 128 CTAs leave most SMs with one CTA, so the latency of the 2B dot products per row is
  // hidden inside the CTA (4 warps: 26 us)
  ntxent_fwd_kernel<<<2 * B, 512, 2 * B * sizeof(float), stream>>>(zhat, neg, loss_sum, B, D, 1.f / tau, eps);
  CSM_CHECK_LAUNCH("ntxent_fwd");
  return CSM_OK;
}

extern "C" int csm_ntxent_bwd(const float* zhat, const float* fnorm, const float* neg, const float* grad_scalar,
                              float* d_feat, int B, int D, float tau, float eps, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && D > 0 && D <= 1024, "csm_ntxent_bwd: bad sizes B=%d D=%d", B, D);
  ntxent_bwd_kernel<<<2 * B, 512, (2 * B + 32) * sizeof(float), stream>>>(zhat, fnorm, neg, grad_scalar, d_feat, B, D,
                                                                         1.f / tau, eps);
  CSM_CHECK_LAUNCH("ntxent_bwd");
  return CSM_OK;
}
