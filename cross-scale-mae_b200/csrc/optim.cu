// Fused AdamW step over every parameter tensor in ONE launch (row f1 of SURVEY.md 8f): decoupled weight decay,
// moment updates, bias-corrected update and -- for the GEMM weights -- the refresh of the bf16 shadow copy the
// tcgen05 kernels consume, so no separate fp32 -> bf16 cast pass runs before the next forward.
// Replaces the reference's torch.optim.AdamW(param_groups, lr, betas=(0.9, 0.95)) (main_pretrain.py:426-427); the
// arithmetic follows torch's single-tensor AdamW (decay, lerp, addcmul, sqrt / sqrt(bc2) + eps, addcdiv) so that
// results agree to fp32 rounding.  HBM-bound: 16 B read + 12 B (+2 B shadow) written per parameter.
#include "common.cuh"

namespace {
using namespace csm;

struct AdamEntry {          // 64 bytes, filled by the host each step (pointers cached, hyper-parameters refreshed)
  float* p;
  const float* g;
  float* m;
  float* v;
  __nv_bfloat16* w16;       // bf16 shadow of a GEMM weight, or nullptr
  long long n;
  float lr, wd, beta1, beta2;
  float eps, bc1, bc2_sqrt, grad_scale;   // grad_scale: multiplies the gradient first (1 = none)
};
static_assert(sizeof(AdamEntry) == 80, "host packs 80-byte entries");

constexpr int ADAM_CHUNK = 8192;          // elements per work item: 256 threads x 8 float4

// ctl (nullable, device): {gradient multiplier, found_inf}.  The multiplier is the GradScaler's 1 / scale (times the
// clip coefficient), read from DEVICE memory so that no host synchronisation sits between backward and step; a nonzero
// found_inf skips the whole update (what GradScaler.step does on the host after an .item(), util/misc.py:325).
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const AdamEntry* __restrict__ table, const int2* __restrict__ chunks, int num_chunks,
                   const float* __restrict__ ctl) {
  float ctl_mul = 1.0f;
  if (ctl != nullptr) {
    if (ctl[1] != 0.0f) return;
    ctl_mul = ctl[0];
  }
  for (int ci = blockIdx.x; ci < num_chunks; ci += gridDim.x) {
    const int2 ch = chunks[ci];
    const AdamEntry e = table[ch.x];
    const long long base = static_cast<long long>(ch.y) * ADAM_CHUNK;
    const long long end = base + ADAM_CHUNK < e.n ? base + ADAM_CHUNK : e.n;
    const float decay = 1.0f - e.lr * e.wd;
    const float step_size = e.lr / e.bc1;
    const float omb1 = 1.0f - e.beta1, omb2 = 1.0f - e.beta2;
    if ((e.n & 3) == 0) {
      for (long long i = base + threadIdx.x * 4; i < end; i += 256 * 4) {
        float4 p = *reinterpret_cast<const float4*>(e.p + i);
        float4 g = *reinterpret_cast<const float4*>(e.g + i);
        float4 m = *reinterpret_cast<const float4*>(e.m + i);
        float4 v = *reinterpret_cast<const float4*>(e.v + i);
        float* pp = reinterpret_cast<float*>(&p);
        float* gp = reinterpret_cast<float*>(&g);
        float* mp = reinterpret_cast<float*>(&m);
        float* vp = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float gg = gp[k] * (e.grad_scale * ctl_mul);
          pp[k] *= decay;
          mp[k] = mp[k] + omb1 * (gg - mp[k]);                 // lerp_(grad, 1 - beta1)
          vp[k] = vp[k] * e.beta2 + omb2 * gg * gg;            // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
          const float denom = sqrtf(vp[k]) / e.bc2_sqrt + e.eps;
          pp[k] -= step_size * (mp[k] / denom);                // addcdiv_(exp_avg, denom, -step_size)
        }
        *reinterpret_cast<float4*>(e.p + i) = p;
        *reinterpret_cast<float4*>(e.m + i) = m;
        *reinterpret_cast<float4*>(e.v + i) = v;
        if (e.w16 != nullptr) {
          uint2 pk;
          pk.x = pack_bf16x2(pp[0], pp[1]);
          pk.y = pack_bf16x2(pp[2], pp[3]);
          *reinterpret_cast<uint2*>(e.w16 + i) = pk;
        }
      }
    } else {
      for (long long i = base + threadIdx.x; i < end; i += 256) {
        const float gg = e.g[i] * (e.grad_scale * ctl_mul);
        float p = e.p[i] * decay;
        const float m = e.m[i] + omb1 * (gg - e.m[i]);
        const float v = e.v[i] * e.beta2 + omb2 * gg * gg;
        p -= step_size * (m / (sqrtf(v) / e.bc2_sqrt + e.eps));
        e.p[i] = p;
        e.m[i] = m;
        e.v[i] = v;
        if (e.w16 != nullptr) e.w16[i] = __float2bfloat16_rn(p);
      }
    }
  }
}

// sum of squares of a flat fp32 buffer (global gradient norm in one pass instead of one torch.norm per parameter,
// util/misc.py:338-355); out[0] accumulates
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float s_buf[8];
  float acc = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_buf[w];
    atomicAdd(out, t);
  }
}

// One pass over the flat gradient buffer for the AMP step (util/misc.py:314-326: unscale_, get_grad_norm_, the
// inf check inside GradScaler.step): stats[0] += sum((x * mul)^2) with mul = *mul_dev (1 when null), stats[1] = 1 if any
// element is not finite.  The gradients themselves are not rewritten -- csm_adamw_multi applies the multiplier.
__global__ void __launch_bounds__(256)
grad_stats_kernel(const float* __restrict__ x, long long n, const float* __restrict__ mul_dev, float* __restrict__ stats) {
  __shared__ float s_buf[8];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const float mul = mul_dev != nullptr ? *mul_dev : 1.0f;
  float acc = 0.f;
  bool bad = false;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    const float a = v.x * mul, b = v.y * mul, c = v.z * mul, d = v.w * mul;
    acc += a * a + b * b + c * c + d * d;
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    acc += (v * mul) * (v * mul);
    bad |= !isfinite(v);
  }
  if (bad) s_bad = 1;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_buf[w];
    atomicAdd(stats, t);
    if (s_bad) stats[1] = 1.0f;
  }
}

// The scalar part of the AMP step on the device (one thread): from stats = {sum of squared UNSCALED gradients,
// found_inf} and the scaler state = {scale, growth_tracker, 1 / scale} it writes the gradient norm, the control pair
// of csm_adamw_multi {1 / scale * clip coefficient, found_inf} and the next scale (GradScaler.update:
// backoff on inf, growth after `interval` clean steps).  clip <= 0: no clipping (util/misc.py:322-323); clip > 0:
// torch.nn.utils.clip_grad_norm_'s coefficient min(1, clip / (norm + 1e-6)) (util/misc.py:317-320).
__global__ void amp_update_kernel(const float* __restrict__ stats, float* __restrict__ state, float* __restrict__ ctl,
                                  float* __restrict__ norm_out, float clip, float growth, float backoff, int interval) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float norm = sqrtf(stats[0]);
  const float found = stats[1];
  float coef = 1.0f;
  if (clip > 0.f) coef = fminf(clip / (norm + 1e-6f), 1.0f);
  ctl[0] = state[2] * coef;
  ctl[1] = found;
  norm_out[0] = norm;
  float scale = state[0], tracker = state[1];
  if (found != 0.f) {
    scale *= backoff;
    tracker = 0.f;
  } else {
    tracker += 1.f;
    if (tracker >= static_cast<float>(interval)) {
      const float grown = scale * growth;
      if (isfinite(grown)) scale = grown;       // torch._amp_update_scale_ keeps the scale when growth would overflow
      tracker = 0.f;
    }
  }
  state[0] = scale;
  state[1] = tracker;
  state[2] = static_cast<float>(1.0 / static_cast<double>(scale));
}

}  // namespace

extern "C" int csm_amp_update(const float* stats, float* state, float* ctl, float* norm_out, float clip, float growth,
                              float backoff, int interval, cudaStream_t stream) {
  CSM_CHECK_ARG(stats && state && ctl && norm_out, "csm_amp_update: null pointer");
  CSM_CHECK_ARG(interval > 0 && growth >= 1.f && backoff > 0.f && backoff <= 1.f, "csm_amp_update: bad scaler constants");
  amp_update_kernel<<<1, 32, 0, stream>>>(stats, state, ctl, norm_out, clip, growth, backoff, interval);
  CSM_CHECK_LAUNCH("amp_update");
  return CSM_OK;
}

extern "C" int csm_grad_stats_f32(const float* x, long long n, const float* mul_dev, float* stats, int num_sms,
                                  cudaStream_t stream) {
  CSM_CHECK_ARG(n > 0, "csm_grad_stats_f32: empty buffer");
  CSM_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "csm_grad_stats_f32: buffer must be 16-byte aligned");
  if (num_sms <= 0) num_sms = 148;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > num_sms * 8) blocks = num_sms * 8;
  if (blocks < 1) blocks = 1;
  grad_stats_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, n, mul_dev, stats);
  CSM_CHECK_LAUNCH("grad_stats_f32");
  return CSM_OK;
}

extern "C" int csm_adamw_multi(const void* table_dev, const void* chunks_dev, int num_chunks, const float* ctl_dev,
                               int num_sms, cudaStream_t stream) {
  CSM_CHECK_ARG(num_chunks > 0, "csm_adamw_multi: nothing to update");
  if (num_sms <= 0) num_sms = 148;
  int grid = num_sms * 8;
  if (grid > num_chunks) grid = num_chunks;
  adamw_multi_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const AdamEntry*>(table_dev),
                                               reinterpret_cast<const int2*>(chunks_dev), num_chunks, ctl_dev);
  CSM_CHECK_LAUNCH("adamw_multi");
  return CSM_OK;
}

extern "C" int csm_sumsq_f32(const float* x, long long n, float* out, int num_sms, cudaStream_t stream) {
  CSM_CHECK_ARG(n > 0, "csm_sumsq_f32: empty buffer");
  CSM_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "csm_sumsq_f32: buffer must be 16-byte aligned");
  if (num_sms <= 0) num_sms = 148;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > num_sms * 8) blocks = num_sms * 8;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, n, out);
  CSM_CHECK_LAUNCH("sumsq_f32");
  return CSM_OK;
}
