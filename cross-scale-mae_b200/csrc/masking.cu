// Random masking, kept-patch gather and the token (un)shuffle kernels of the MAE encoder/decoder.
//
// Reference semantics (paths relative to the upstream repo):
//   models_mae/MAE_ViT_Shared.py:57-84     random_masking (noise -> argsort -> argsort -> keep / mask)
//   models_mae/MAE_ViT_Baseline.py:245-256 patch-embed + pos-embed, masking, cls concat
//   models_mae/MAE_ViT_Baseline.py:270-283 decoder_embed output + mask tokens, un-shuffle, + decoder pos-embed
//
// All of this is HBM-/latency-bound integer and copy work: coalesced, 16-byte vectorised accesses,
// no tensor cores.  Index outputs are bit-exact against a stable argsort.
#include "common.cuh"

#include <algorithm>

namespace {
using namespace csm;

// One CTA per image row.  rank[i] = #{j : noise[j] < noise[i] or (noise[j] == noise[i] and j < i)}
// is the position of element i in the stable ascending sort, i.e. ids_restore[i].
__global__ void random_masking_kernel(const float* __restrict__ noise, int L, int keep,
                                      long long* __restrict__ ids_restore, int* __restrict__ ids_shuffle,
                                      float* __restrict__ mask) {
  extern __shared__ float s_noise[];
  const int n = blockIdx.x;
  const float* row = noise + static_cast<size_t>(n) * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s_noise[i] = row[i];
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float v = s_noise[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const float u = s_noise[j];
      rank += (u < v) || (u == v && j < i);
    }
    ids_restore[static_cast<size_t>(n) * L + i] = rank;
    ids_shuffle[static_cast<size_t>(n) * L + rank] = i;
    mask[static_cast<size_t>(n) * L + i] = rank >= keep ? 1.0f : 0.0f;
  }
}

// A[(n*Se + 1 + j), (c, py, px)] = bf16(img[n, c, h*p + py, w*p + px]) for the j-th kept patch
// (h, w) = divmod(ids_shuffle[n, j], G); row n*Se + 0 (the cls slot) is zero.
// One warp per output row; 16 consecutive pixels (64 B) per (c, py).
__global__ void patch_gather_kernel(const float* __restrict__ imgs, const int* __restrict__ ids_shuffle,
                                    __nv_bfloat16* __restrict__ out, int nimg, int C, int H, int p, int L,
                                    int keep) {
  const int Se = keep + 1;
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nimg * Se) return;
  const int n = row / Se, t = row % Se;
  const int K = C * p * p;
  __nv_bfloat16* o = out + static_cast<size_t>(row) * K;
  if (t == 0) {
    for (int i = lane * 8; i < K; i += 32 * 8) *reinterpret_cast<uint4*>(o + i) = make_uint4(0, 0, 0, 0);
    return;
  }
  const int G = H / p;
  const int patch = ids_shuffle[static_cast<size_t>(n) * L + (t - 1)];
  const int ph = patch / G, pw = patch % G;
  const float* base = imgs + static_cast<size_t>(n) * C * H * H + static_cast<size_t>(ph) * p * H + pw * p;
  // element e = (c*p + py)*p + px; 4 consecutive px per thread
  for (int e = lane * 4; e < K; e += 32 * 4) {
    const int px = e % p, cy = e / p;
    const int py = cy % p, c = cy / p;
    const float4 v = *reinterpret_cast<const float4*>(base + (static_cast<size_t>(c) * H + py) * H + px);
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(o + e) = pk;
  }
}

// x[n, 0] = cls + pos[0];  x[n, 1 + j] = float(emb[n*Se + 1 + j]) + pos[1 + ids_shuffle[n, j]]
__global__ void encoder_assemble_kernel(const __nv_bfloat16* __restrict__ emb, const int* __restrict__ ids_shuffle,
                                        const float* __restrict__ pos, const float* __restrict__ cls,
                                        float* __restrict__ x, int nimg, int L, int keep, int D) {
  const int Se = keep + 1;
  const int row = blockIdx.x;
  const int n = row / Se, t = row % Se;
  float* o = x + static_cast<size_t>(row) * D;
  if (t == 0) {
    for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
      const float4 a = *reinterpret_cast<const float4*>(cls + i);
      const float4 b = *reinterpret_cast<const float4*>(pos + i);
      *reinterpret_cast<float4*>(o + i) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    return;
  }
  const int patch = ids_shuffle[static_cast<size_t>(n) * L + (t - 1)];
  const float* pr = pos + static_cast<size_t>(1 + patch) * D;
  const __nv_bfloat16* e = emb + static_cast<size_t>(row) * D;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    const uint2 ev = *reinterpret_cast<const uint2*>(e + i);
    const float2 e0 = unpack_bf16x2(ev.x), e1 = unpack_bf16x2(ev.y);
    const float4 b = *reinterpret_cast<const float4*>(pr + i);
    *reinterpret_cast<float4*>(o + i) = make_float4(e0.x + b.x, e0.y + b.y, e1.x + b.z, e1.y + b.w);
  }
}

// y[n, 0] = float(demb[n, 0]) + dpos[0]
// y[n, 1 + i] = (r < keep ? float(demb[n, 1 + r]) : mask_token) + dpos[1 + i],  r = ids_restore[n, i]
__global__ void decoder_assemble_kernel(const __nv_bfloat16* __restrict__ demb,
                                        const long long* __restrict__ ids_restore,
                                        const float* __restrict__ mask_token, const float* __restrict__ dpos,
                                        float* __restrict__ y, int L, int keep, int Dd) {
  const int Sd = L + 1, Se = keep + 1;
  const int row = blockIdx.x;
  const int n = row / Sd, t = row % Sd;
  float* o = y + static_cast<size_t>(row) * Dd;
  const float* pr = dpos + static_cast<size_t>(t) * Dd;
  int src = 0;
  if (t > 0) {
    const int r = static_cast<int>(ids_restore[static_cast<size_t>(n) * L + (t - 1)]);
    src = r < keep ? 1 + r : -1;
  }
  if (src >= 0) {
    const __nv_bfloat16* e = demb + (static_cast<size_t>(n) * Se + src) * Dd;
    for (int i = threadIdx.x * 4; i < Dd; i += blockDim.x * 4) {
      const uint2 ev = *reinterpret_cast<const uint2*>(e + i);
      const float2 e0 = unpack_bf16x2(ev.x), e1 = unpack_bf16x2(ev.y);
      const float4 b = *reinterpret_cast<const float4*>(pr + i);
      *reinterpret_cast<float4*>(o + i) = make_float4(e0.x + b.x, e0.y + b.y, e1.x + b.z, e1.y + b.w);
    }
  } else {
    for (int i = threadIdx.x * 4; i < Dd; i += blockDim.x * 4) {
      const float4 a = *reinterpret_cast<const float4*>(mask_token + i);
      const float4 b = *reinterpret_cast<const float4*>(pr + i);
      *reinterpret_cast<float4*>(o + i) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  }
}

// Backward of decoder_assemble.
//   d_demb[n, 0]     = bf16(dy[n, 0])
//   d_demb[n, 1 + r] = bf16(dy[n, 1 + ids_shuffle[n, r]])          r < keep
//   d_mask_token    += sum over masked positions (and images) of dy (second kernel below)
__global__ void decoder_assemble_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ ids_shuffle,
                                            __nv_bfloat16* __restrict__ d_demb, int L, int keep, int Dd) {
  // one CTA per (image, kept token): gather + cast
  const int Sd = L + 1, Se = keep + 1;
  const int n = blockIdx.x / Se, t = blockIdx.x % Se;
  const int src = t == 0 ? 0 : 1 + ids_shuffle[static_cast<size_t>(n) * L + t - 1];
  const float* s = dy + (static_cast<size_t>(n) * Sd + src) * Dd;
  __nv_bfloat16* o = d_demb + static_cast<size_t>(blockIdx.x) * Dd;
  for (int i = threadIdx.x * 4; i < Dd; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(o + i) = pk;
  }
}

// d_mask_token += column sums of the masked rows of all images: grid (128-column groups, row slabs), a warp per
// row with four rows in flight, one atomic per column and CTA (same-address atomics serialise in L2: the first
// version issued 2048 per column and spent 80 us on them)
__global__ void mask_token_grad_kernel(const float* __restrict__ dy, const int* __restrict__ ids_shuffle,
                                       float* __restrict__ d_mask_token, int nimg, int L, int keep, int Dd) {
  __shared__ float4 s_part[8][32];
  const int Sd = L + 1, nm = L - keep;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 128 + lane * 4;
  const int total = nimg * nm;
  const int stride = gridDim.y * 8;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < Dd) {
    for (int m0 = blockIdx.y * 8 + w; m0 < total; m0 += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = m0 + u * stride;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < total) {
          const int n = m / nm, j = m - n * nm;
          const int src = 1 + ids_shuffle[static_cast<size_t>(n) * L + keep + j];
          v[u] = *reinterpret_cast<const float4*>(dy + (static_cast<size_t>(n) * Sd + src) * Dd + col);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
  }
  s_part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && col < Dd) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = s_part[k][lane];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(d_mask_token + col + 0, acc.x);
    atomicAdd(d_mask_token + col + 1, acc.y);
    atomicAdd(d_mask_token + col + 2, acc.z);
    atomicAdd(d_mask_token + col + 3, acc.w);
  }
}

// Gradient entering the encoder output (un-normed: the reference discards encoder_norm,
// MAE_ViT_Baseline.py:264):  d_x[n, t] = float(d_enc_bf16[n, t]) + (t >= 1 ? d_feat[n] / keep : 0)
// where d_feat is the NT-Xent gradient w.r.t. the token-mean feature (MAE_ViT_MsLdCeCd.py:64-65).
// Also emits the bf16 copy the next wgrad/dgrad GEMMs consume.
__global__ void encoder_out_grad_kernel(const __nv_bfloat16* __restrict__ d_enc, const float* __restrict__ d_feat,
                                        float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16, int Se, int D,
                                        float inv_keep) {
  const int row = blockIdx.x;
  const int n = row / Se, t = row % Se;
  const __nv_bfloat16* s = d_enc + static_cast<size_t>(row) * D;
  const float* f = d_feat ? d_feat + static_cast<size_t>(n) * D : nullptr;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    const uint2 ev = *reinterpret_cast<const uint2*>(s + i);
    const float2 e0 = unpack_bf16x2(ev.x), e1 = unpack_bf16x2(ev.y);
    float4 v = make_float4(e0.x, e0.y, e1.x, e1.y);
    if (f != nullptr && t >= 1) {
      const float4 g = *reinterpret_cast<const float4*>(f + i);
      v.x += g.x * inv_keep; v.y += g.y * inv_keep; v.z += g.z * inv_keep; v.w += g.w * inv_keep;
    }
    *reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * D + i) = v;
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dx_bf16 + static_cast<size_t>(row) * D + i) = pk;
  }
}

// d_cls[i] = sum_n dx[n, 0, i]
__global__ void cls_grad_kernel(const float* __restrict__ dx, float* __restrict__ d_cls, int nimg, int Se, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  float acc = 0.f;
  for (int n = 0; n < nimg; ++n) acc += dx[static_cast<size_t>(n) * Se * D + i];
  d_cls[i] = acc;
}

// In-model random-resized-crop of the two-scale models (models_mae/MAE_ViT_MsLd.py:29-35,52): ONE crop box
// (top, left, h, w) for the whole batch, resized to S x S with torchvision's bilinear + antialias resize, i.e.
// ATen's separable triangle filter (aten/src/ATen/native/cuda/UpSampleBilinear2d.cu, _upsample_bilinear2d_aa):
//   scale = in / out; support = max(scale, 1); center = scale * (i + 0.5);
//   taps [xmin, xmin + xsize), xmin = max(int(center - support + 0.5), 0), xsize = min(int(center + support + 0.5), in) - xmin;
//   w_j = max(0, 1 - |(j + xmin - center + 0.5) / max(scale, 1)|), normalised to sum 1;
//   out = sum_y wy * (sum_x wx * src)   (horizontal pass first, fp32).
// The crop area is 25-75 % of the image, so this is (almost always) an up-sampling with 2-3 taps per axis, but
// the general filter is implemented (taps capped at MAX_TAPS per axis; the launcher rejects larger scales).
constexpr int CROP_MAX_TAPS = 8;

__device__ __forceinline__ void aa_taps(int i, float scale, int in_size, int& xmin, int& xsize, float* w) {
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const float center = scale * (static_cast<float>(i) + 0.5f);
  xmin = max(static_cast<int>(center - support + 0.5f), 0);
  xsize = min(static_cast<int>(center + support + 0.5f), in_size) - xmin;
  xsize = min(max(xsize, 0), CROP_MAX_TAPS);
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    const float x = (static_cast<float>(j + xmin) - center + 0.5f) * invscale;
    const float v = fmaxf(0.f, 1.0f - fabsf(x));
    w[j] = v;
    total += v;
  }
  if (total != 0.f) {
    for (int j = 0; j < xsize; ++j) w[j] /= total;
  }
}

// one thread per output pixel, x fastest (coalesced stores; the 2-3 source rows of a warp's pixels are contiguous)
__global__ void resized_crop_kernel(const float* __restrict__ imgs, float* __restrict__ out, int planes, int H, int W,
                                    int top, int left, int h, int w, int S) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= S) return;
  float wx[CROP_MAX_TAPS], wy[CROP_MAX_TAPS];
  int xmin, xsize, ymin, ysize;
  aa_taps(x, static_cast<float>(w) / static_cast<float>(S), w, xmin, xsize, wx);
  aa_taps(y, static_cast<float>(h) / static_cast<float>(S), h, ymin, ysize, wy);
  for (int pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const float* src = imgs + (static_cast<size_t>(pl) * H + top + ymin) * W + left + xmin;
    float acc = 0.f;
    for (int yy = 0; yy < ysize; ++yy) {
      const float* row = src + static_cast<size_t>(yy) * W;
      float t = row[0] * wx[0];
      for (int xx = 1; xx < xsize; ++xx) t += row[xx] * wx[xx];
      acc = yy == 0 ? t * wy[0] : acc + t * wy[yy];
    }
    out[(static_cast<size_t>(pl) * S + y) * S + x] = acc;
  }
}

}  // namespace

// Fixed 2-D sin-cos position table (util/pos_embed.py:16-63), init-time: out[row, :] for row = cls + h*G + w is
// [sin(w*om) | cos(w*om) | sin(h*om) | cos(h*om)], om_k = 10000^(-k / (D/4)) -- the reference's meshgrid is w-first
// (pos_embed.py:24).  Evaluated in fp64 like numpy and rounded once to fp32; the cls row is zero.
namespace {
__global__ void sincos_pos_embed_kernel(float* __restrict__ out, int D, int G, int cls) {
  const int quarter = D / 4;
  const long long total = static_cast<long long>(G) * G * quarter;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % quarter);
    const int pos = static_cast<int>(idx / quarter);
    const double om = 1.0 / pow(10000.0, static_cast<double>(k) / static_cast<double>(quarter));
    const double aw = static_cast<double>(pos % G) * om, ah = static_cast<double>(pos / G) * om;
    float* row = out + static_cast<size_t>(pos + cls) * D;
    row[k] = static_cast<float>(sin(aw));
    row[quarter + k] = static_cast<float>(cos(aw));
    row[2 * quarter + k] = static_cast<float>(sin(ah));
    row[3 * quarter + k] = static_cast<float>(cos(ah));
  }
  if (cls && blockIdx.x == 0)
    for (int c = threadIdx.x; c < D; c += blockDim.x) out[c] = 0.f;
}
}  // namespace

extern "C" int csm_sincos_pos_embed(float* out, int embed_dim, int grid_size, int cls_token, cudaStream_t stream) {
  CSM_CHECK_ARG(out != nullptr && grid_size > 0, "csm_sincos_pos_embed: bad arguments");
  CSM_CHECK_ARG(embed_dim > 0 && embed_dim % 4 == 0, "csm_sincos_pos_embed: embed_dim must be divisible by 4 (got %d)",
                embed_dim);
  const long long total = static_cast<long long>(grid_size) * grid_size * (embed_dim / 4);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  sincos_pos_embed_kernel<<<blocks, 256, 0, stream>>>(out, embed_dim, grid_size, cls_token ? 1 : 0);
  CSM_CHECK_LAUNCH("sincos_pos_embed");
  return CSM_OK;
}

extern "C" int csm_resized_crop(const float* imgs, float* out, int planes, int H, int W, int top, int left, int h,
                                int w, int S, cudaStream_t stream) {
  CSM_CHECK_ARG(planes > 0 && S > 0 && h > 0 && w > 0 && top >= 0 && left >= 0 && top + h <= H && left + w <= W,
                "csm_resized_crop: bad box top=%d left=%d h=%d w=%d in %dx%d", top, left, h, w, H, W);
  CSM_CHECK_ARG(2 * h <= (CROP_MAX_TAPS - 1) * S && 2 * w <= (CROP_MAX_TAPS - 1) * S,
                "csm_resized_crop: down-scaling factor too large for the %d-tap filter (box %dx%d -> %d)", CROP_MAX_TAPS,
                h, w, S);
  dim3 grid(csm_cdiv(S, 128), S, planes < 64 ? planes : 64);
  resized_crop_kernel<<<grid, 128, 0, stream>>>(imgs, out, planes, H, W, top, left, h, w, S);
  CSM_CHECK_LAUNCH("resized_crop");
  return CSM_OK;
}

extern "C" int csm_random_masking(const float* noise, int nimg, int L, int keep, long long* ids_restore,
                                  int* ids_shuffle, float* mask, cudaStream_t stream) {
  CSM_CHECK_ARG(nimg > 0 && L > 0 && keep >= 0 && keep <= L, "csm_random_masking: bad sizes nimg=%d L=%d keep=%d",
                nimg, L, keep);
  CSM_CHECK_ARG(L <= 12288, "csm_random_masking: L=%d exceeds the shared-memory row limit", L);
  const int threads = L >= 512 ? 512 : (L >= 256 ? 256 : 128);
  random_masking_kernel<<<nimg, threads, L * sizeof(float), stream>>>(noise, L, keep, ids_restore, ids_shuffle, mask);
  CSM_CHECK_LAUNCH("random_masking");
  return CSM_OK;
}

extern "C" int csm_patch_gather(const float* imgs, const int* ids_shuffle, void* out_bf16, int nimg, int C, int H,
                                int p, int L, int keep, cudaStream_t stream) {
  CSM_CHECK_ARG(nimg > 0 && H % p == 0 && (H / p) * (H / p) == L, "csm_patch_gather: bad geometry H=%d p=%d L=%d", H,
                p, L);
  CSM_CHECK_ARG(p % 4 == 0 && (C * p * p) % 8 == 0, "csm_patch_gather: patch size must be a multiple of 4 (p=%d)", p);
  CSM_CHECK_ARG((reinterpret_cast<uintptr_t>(imgs) & 15) == 0, "csm_patch_gather: imgs must be 16-byte aligned");
  const int rows = nimg * (keep + 1);
  const int wpb = 8;
  patch_gather_kernel<<<csm_cdiv(rows, wpb), wpb * 32, 0, stream>>>(
      imgs, ids_shuffle, reinterpret_cast<__nv_bfloat16*>(out_bf16), nimg, C, H, p, L, keep);
  CSM_CHECK_LAUNCH("patch_gather");
  return CSM_OK;
}

extern "C" int csm_encoder_assemble(const void* emb_bf16, const int* ids_shuffle, const float* pos, const float* cls,
                                    float* x, int nimg, int L, int keep, int D, cudaStream_t stream) {
  CSM_CHECK_ARG(D % 4 == 0, "csm_encoder_assemble: D must be a multiple of 4 (D=%d)", D);
  encoder_assemble_kernel<<<nimg * (keep + 1), 128, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(emb_bf16), ids_shuffle, pos, cls, x, nimg, L, keep, D);
  CSM_CHECK_LAUNCH("encoder_assemble");
  return CSM_OK;
}

extern "C" int csm_decoder_assemble(const void* demb_bf16, const long long* ids_restore, const float* mask_token,
                                    const float* dpos, float* y, int nimg, int L, int keep, int Dd,
                                    cudaStream_t stream) {
  CSM_CHECK_ARG(Dd % 4 == 0, "csm_decoder_assemble: Dd must be a multiple of 4 (Dd=%d)", Dd);
  decoder_assemble_kernel<<<nimg * (L + 1), 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(demb_bf16),
                                                             ids_restore, mask_token, dpos, y, L, keep, Dd);
  CSM_CHECK_LAUNCH("decoder_assemble");
  return CSM_OK;
}

extern "C" int csm_decoder_assemble_bwd(const float* dy, const int* ids_shuffle, void* d_demb_bf16,
                                        float* d_mask_token, int nimg, int L, int keep, int Dd, cudaStream_t stream) {
  CSM_CHECK_ARG(Dd % 4 == 0, "csm_decoder_assemble_bwd: Dd must be a multiple of 4 (Dd=%d)", Dd);
  decoder_assemble_bwd_kernel<<<nimg * (keep + 1), 128, 0, stream>>>(
      dy, ids_shuffle, reinterpret_cast<__nv_bfloat16*>(d_demb_bf16), L, keep, Dd);
  if (L > keep) {
    const int gx = csm_cdiv(Dd, 128);
    int gy = csm_cdiv(2 * 148, gx);
    const int max_gy = csm_cdiv(nimg * (L - keep), 8);
    if (gy > max_gy) gy = max_gy;
    mask_token_grad_kernel<<<dim3(gx, gy), 256, 0, stream>>>(dy, ids_shuffle, d_mask_token, nimg, L, keep, Dd);
  }
  CSM_CHECK_LAUNCH("decoder_assemble_bwd");
  return CSM_OK;
}

extern "C" int csm_encoder_out_grad(const void* d_enc_bf16, const float* d_feat, float* dx, void* dx_bf16, int nimg,
                                    int Se, int D, cudaStream_t stream) {
  CSM_CHECK_ARG(D % 4 == 0 && Se >= 2, "csm_encoder_out_grad: bad sizes Se=%d D=%d", Se, D);
  encoder_out_grad_kernel<<<nimg * Se, 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(d_enc_bf16), d_feat,
                                                         dx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), Se, D,
                                                         1.0f / static_cast<float>(Se - 1));
  CSM_CHECK_LAUNCH("encoder_out_grad");
  return CSM_OK;
}

extern "C" int csm_cls_grad(const float* dx, float* d_cls, int nimg, int Se, int D, cudaStream_t stream) {
  cls_grad_kernel<<<csm_cdiv(D, 128), 128, 0, stream>>>(dx, d_cls, nimg, Se, D);
  CSM_CHECK_LAUNCH("cls_grad");
  return CSM_OK;
}
