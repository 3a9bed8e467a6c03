// tcgen05 GEMM for the Linear layers of the Cross-Scale MAE hot path (timm Block qkv/proj/fc1/fc2,
// decoder_embed, decoder_pred, patch-embed-as-GEMM, predictor MLP) and their dgrad / wgrad.
//
//   forward : Y[M,N]  = X[M,K] . W[N,K]^T          A K-major,  B K-major
//   dgrad   : dX[M,K] = dY[M,N] . W[N,K]           A K-major,  B MN-major (W is read as stored: no
//                                                  transposed weight copy exists anywhere)
//   wgrad   : dW[N,K] = sum_r dY[r,N] . X[r,K]     A MN-major, B MN-major, reduction over the token
//                                                  rows r, split over blockIdx.z and accumulated
//                                                  with red.global.add.v4.f32
//
// One CTA computes one 128 x BN output tile:
//   warp 0      TMA producer   (cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (accumulator in TMEM)
//   warps 2..5  epilogue       (tcgen05.ld -> registers -> fused bias / GELU / residual / dGELU -> global)
// Two CTAs are co-resident per SM (3-stage ring = 96 KB each) so one CTA's epilogue overlaps the
// other's main loop.
//
// Rounding points follow the reference's CUDA-autocast graph (SURVEY.md 8a'): a Linear's output
// is rounded to bf16 before it is added to the fp32 residual stream; GELU is evaluated in fp32 on
// the bf16-rounded fc1 output and rounded again.
#include "common.cuh"

#include <mutex>
#include <unordered_map>

namespace {

using namespace csm;

constexpr int BM = 128;
constexpr int BK = 64;     // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

enum Epi : int {
  EPI_BF16 = 0,        // out_bf16 = bf16(acc + bias)
  EPI_GELU = 1,        // out_bf16 = h = bf16(acc + bias); out2_bf16 = bf16(gelu(h))
  EPI_RESID = 2,       // out_f32 = aux_f32 + bf16(acc + bias)              (residual stream)
  EPI_DGELU = 3,       // out_bf16 = bf16(bf16(acc) * gelu'(aux_bf16))
  EPI_F32_ATOMIC = 4,  // out_f32 += acc                                    (red.global.add)
  EPI_F32 = 5          // out_f32 = acc + bias
};

struct GemmParams {
  int M, N, K;           // output rows, output cols, reduction length
  int kb_per_split;      // k-blocks handled by one blockIdx.z
  void* out;             // bf16 or f32 [M, ldo]
  void* out2;            // bf16 [M, ldo]      (EPI_GELU)
  const float* bias;     // [N] or nullptr
  const void* aux;       // EPI_DGELU: bf16 pre-activation h; EPI_RESID: f32 residual input
  int ldo;               // leading dimension (elements) of out/out2/aux
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + 1024B alignment slack
};

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
  const int num_kb = kb1 - kb0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        const int k = (kb0 + i) * BK;
        // K-major operand : one box {64 k, rows}.
        // MN-major operand: boxes {64 mn, 64 k}; each 64-wide MN group is its own [64 k][128 B] slab.
        if (!A_MN) {
          tma_load_2d(sa, &tmap_a, &full_bar[s], k, m_tile * BM);
        } else {
#pragma unroll
          for (int g = 0; g < BM / 64; ++g)
            tma_load_2d(sa + g * (BK * 128), &tmap_a, &full_bar[s], m_tile * BM + g * 64, k);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmap_b, &full_bar[s], k, n_tile * BN);
        } else {
#pragma unroll
          for (int g = 0; g < BN / 64; ++g)
            tma_load_2d(sb + g * (BK * 128), &tmap_b, &full_bar[s], n_tile * BN + g * 64, k);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major, 128B swizzle : 8-row groups are 1024 B apart (SBO); a K step of 16 is 32 B inside the row.
          // MN-major, 128B swizzle: 64-wide MN groups are BK*128 B apart (LBO), 8-deep K groups 1024 B (SBO);
          //                         a K step of 16 is 16 rows = 2048 B.
          const uint64_t da = A_MN ? umma_smem_desc_sw128(sa + k * (UMMA_K * 128), BK * 128, 1024)
                                   : umma_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
          const uint64_t db = B_MN ? umma_smem_desc_sw128(sb + k * (UMMA_K * 128), BK * 128, 1024)
                                   : umma_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
          umma_f16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);   // frees the smem slot once these MMAs have read it
      }
      umma_commit(accum_bar);         // accumulator complete
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = m_tile * BM + q * 32 + lane;
    const bool row_ok = row < p.M;
    if (num_kb > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int col0 = n_tile * BN + c * 32;
      if (col0 >= p.N) break;
      uint32_t r[32];
      if (num_kb > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0;
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

      if (EPI != EPI_DGELU && EPI != EPI_F32_ATOMIC) {
        if (p.bias != nullptr) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (col0 + g * 4 + 4 <= p.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + g * 4));
              v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
            }
          }
        }
      }
      if (!row_ok) continue;
      const size_t off = static_cast<size_t>(row) * p.ldo + col0;

      if (EPI == EPI_BF16 || EPI == EPI_GELU || EPI == EPI_DGELU) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (col0 + g * 8 + 8 > p.N) break;
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = v[g * 8 + j];
          if (EPI == EPI_DGELU) {
            const uint4 hv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + off + g * 8);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 h2 = unpack_bf16x2(hw[j]);
              x[2 * j] = bf16_round(x[2 * j]) * gelu_erf_grad(h2.x);
              x[2 * j + 1] = bf16_round(x[2 * j + 1]) * gelu_erf_grad(h2.y);
            }
          }
          uint4 pk;
          pk.x = pack_bf16x2(x[0], x[1]); pk.y = pack_bf16x2(x[2], x[3]);
          pk.z = pack_bf16x2(x[4], x[5]); pk.w = pack_bf16x2(x[6], x[7]);
          *reinterpret_cast<uint4*>(o + g * 8) = pk;
          if (EPI == EPI_GELU) {
            const uint32_t hw[4] = {pk.x, pk.y, pk.z, pk.w};
            uint32_t aw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 h2 = unpack_bf16x2(hw[j]);
              aw[j] = pack_bf16x2(gelu_erf(h2.x), gelu_erf(h2.y));
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + off + g * 8) =
                make_uint4(aw[0], aw[1], aw[2], aw[3]);
          }
        }
      } else if (EPI == EPI_RESID) {
        float* o = reinterpret_cast<float*>(p.out) + off;
        const float* ri = reinterpret_cast<const float*>(p.aux) + off;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (col0 + g * 4 + 4 > p.N) break;
          float4 x = *reinterpret_cast<const float4*>(ri + g * 4);
          x.x += bf16_round(v[g * 4 + 0]); x.y += bf16_round(v[g * 4 + 1]);
          x.z += bf16_round(v[g * 4 + 2]); x.w += bf16_round(v[g * 4 + 3]);
          *reinterpret_cast<float4*>(o + g * 4) = x;
        }
      } else if (EPI == EPI_F32) {
        float* o = reinterpret_cast<float*>(p.out) + off;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (col0 + g * 4 + 4 > p.N) break;
          *reinterpret_cast<float4*>(o + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
      } else {  // EPI_F32_ATOMIC
        float* o = reinterpret_cast<float*>(p.out) + off;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (col0 + g * 4 + 4 > p.N) break;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + g * 4), "f"(v[g * 4]),
                       "f"(v[g * 4 + 1]), "f"(v[g * 4 + 2]), "f"(v[g * 4 + 3])
                       : "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.inner * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.outer * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (k.ld * 31 + k.box_inner * 131 + k.box_outer) + (h << 6) + (h >> 2);
    return h;
  }
};

// bf16 row-major matrix [outer, inner] with leading dimension ld (elements); 128B-swizzled boxes.
int get_tensor_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                   uint32_t box_inner, uint32_t box_outer) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return CSM_OK;
    }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    csm_set_error("cuTensorMapEncodeTiled not available from the driver");
    return CSM_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) {
    csm_set_error("tensor map: base pointer and row pitch must be 16-byte aligned (ptr=%p ld=%llu)", ptr,
                  (unsigned long long)ld);
    return CSM_ERR_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    csm_set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    return CSM_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 8192) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return CSM_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int splits, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gemm_kernel<BN, STAGES, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) {
      csm_set_error("gemm: cudaFuncSetAttribute(smem=%d) failed: %s", L::TOTAL, cudaGetErrorString(e));
      return CSM_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(csm_cdiv(p.N, BN), csm_cdiv(p.M, BM), splits);
  kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(ta, tb, p);
  CSM_CHECK_LAUNCH("gemm_tcgen05");
  return CSM_OK;
}

constexpr int kBN = 128;
constexpr int kStages = 3;

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI (declared in include/csmae_b200.h)
// ---------------------------------------------------------------------------------------------
extern "C" int csm_linear_fwd(const void* x_bf16, const void* w_bf16, const float* bias, void* out, void* aux,
                              int M, int N, int K, int epilogue, cudaStream_t stream) {
  CSM_CHECK_ARG(M > 0 && N > 0 && K > 0, "csm_linear_fwd: empty problem M=%d N=%d K=%d", M, N, K);
  CSM_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "csm_linear_fwd: N and K must be multiples of 8 (N=%d K=%d)", N, K);
  CUtensorMap ta, tb;
  int rc = get_tensor_map(&ta, x_bf16, K, M, K, BK, BM);
  if (rc) return rc;
  rc = get_tensor_map(&tb, w_bf16, K, N, K, BK, kBN);
  if (rc) return rc;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.kb_per_split = csm_cdiv(K, BK);
  p.out = out; p.out2 = aux; p.bias = bias; p.aux = aux; p.ldo = N;
  switch (epilogue) {
    case EPI_BF16: return launch_gemm<kBN, kStages, false, false, EPI_BF16>(ta, tb, p, 1, stream);
    case EPI_GELU:
      CSM_CHECK_ARG(aux != nullptr, "csm_linear_fwd: GELU epilogue needs aux (activation output)");
      return launch_gemm<kBN, kStages, false, false, EPI_GELU>(ta, tb, p, 1, stream);
    case EPI_RESID:
      CSM_CHECK_ARG(aux != nullptr, "csm_linear_fwd: residual epilogue needs aux (f32 residual input)");
      return launch_gemm<kBN, kStages, false, false, EPI_RESID>(ta, tb, p, 1, stream);
    case EPI_F32: return launch_gemm<kBN, kStages, false, false, EPI_F32>(ta, tb, p, 1, stream);
    default:
      csm_set_error("csm_linear_fwd: unsupported epilogue %d", epilogue);
      return CSM_ERR_ARG;
  }
}

extern "C" int csm_linear_dgrad(const void* dy_bf16, const void* w_bf16, void* dx, const void* aux, int M, int N,
                                int K, int epilogue, cudaStream_t stream) {
  // dX[M,K] = dY[M,N] . W[N,K]; reduction over N, W consumed MN-major straight from its [N,K] storage.
  CSM_CHECK_ARG(M > 0 && N > 0 && K > 0, "csm_linear_dgrad: empty problem M=%d N=%d K=%d", M, N, K);
  CSM_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "csm_linear_dgrad: N and K must be multiples of 8 (N=%d K=%d)", N, K);
  CUtensorMap ta, tb;
  int rc = get_tensor_map(&ta, dy_bf16, N, M, N, BK, BM);
  if (rc) return rc;
  rc = get_tensor_map(&tb, w_bf16, K, N, K, 64, BK);
  if (rc) return rc;
  GemmParams p{};
  p.M = M; p.N = K; p.K = N;
  p.kb_per_split = csm_cdiv(N, BK);
  p.out = dx; p.aux = aux; p.ldo = K;
  switch (epilogue) {
    case EPI_BF16: return launch_gemm<kBN, kStages, false, true, EPI_BF16>(ta, tb, p, 1, stream);
    case EPI_DGELU:
      CSM_CHECK_ARG(aux != nullptr, "csm_linear_dgrad: dGELU epilogue needs aux (bf16 pre-activation)");
      return launch_gemm<kBN, kStages, false, true, EPI_DGELU>(ta, tb, p, 1, stream);
    case EPI_F32: return launch_gemm<kBN, kStages, false, true, EPI_F32>(ta, tb, p, 1, stream);
    default:
      csm_set_error("csm_linear_dgrad: unsupported epilogue %d", epilogue);
      return CSM_ERR_ARG;
  }
}

extern "C" int csm_linear_wgrad(const void* dy_bf16, const void* x_bf16, float* dw, int rows, int N, int K,
                                int num_sms, cudaStream_t stream) {
  CSM_CHECK_ARG(rows > 0 && N > 0 && K > 0, "csm_linear_wgrad: empty problem rows=%d N=%d K=%d", rows, N, K);
  CSM_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "csm_linear_wgrad: N and K must be multiples of 8 (N=%d K=%d)", N, K);
  CUtensorMap ta, tb;
  int rc = get_tensor_map(&ta, dy_bf16, N, rows, N, 64, BK);
  if (rc) return rc;
  rc = get_tensor_map(&tb, x_bf16, K, rows, K, 64, BK);
  if (rc) return rc;
  GemmParams p{};
  p.M = N; p.N = K; p.K = rows;
  const int tiles = csm_cdiv(N, BM) * csm_cdiv(K, kBN);
  const int num_kb = csm_cdiv(rows, BK);
  if (num_sms <= 0) num_sms = 148;
  int splits = (2 * num_sms + tiles - 1) / tiles;     // two co-resident CTAs per SM
  splits = splits < 1 ? 1 : splits;
  int max_splits = num_kb / 4;                        // keep >= 4 k-blocks per split
  if (max_splits < 1) max_splits = 1;
  if (splits > max_splits) splits = max_splits;
  p.kb_per_split = csm_cdiv(num_kb, splits);
  splits = csm_cdiv(num_kb, p.kb_per_split);
  p.out = dw; p.ldo = K;
  return launch_gemm<kBN, kStages, true, true, EPI_F32_ATOMIC>(ta, tb, p, splits, stream);
}
