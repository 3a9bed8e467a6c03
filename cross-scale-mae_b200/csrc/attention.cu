// LEGACY mma.sync path, kept only as the A/B baseline of tools/attn_tc_check.py (csm_attention_*_legacy; not
// declared in include/csmae_b200.h and not called by the product).  The product kernels are in attention_tc.cu.
//
// Fused multi-head self-attention forward / backward for the timm Block attention of the MAE
// encoder (S = keep+1, d = 64) and decoder (S = L+1, d = 32): softmax((q k^T) * d^-1/2) v, no mask,
// no dropout (timm 0.4.12 Attention as restated in oracle/timm_shim.py; call sites
// models_mae/MAE_ViT_Baseline.py:160-188).
//
// Layout: qkv is the [rows, 3*Dm] bf16 output of the qkv Linear, row = b*S + s, columns
// [which*Dm + h*d + j]; the output is [rows, Dm] with column h*d + j (== transpose(1,2).reshape).
//
// The S x S score matrix never exists in HBM (the reference materialises it in fp32: 159 MB per
// decoder layer at B=64).  With d = 32/64 and S <= 785 these kernels are bound by the softmax
// arithmetic (one exp2 + ~4 FP32 ops per score element against 2*d MACs), not by the tensor pipe or
// HBM, so they are built around the instruction count per score element:
//   forward : one CTA = the whole K/V of 1..2 heads in shared memory (read from HBM exactly once),
//             one warp per 16-query tile, online softmax over 64-key blocks, exp2 with the
//             d^-1/2 * log2(e) factor folded into one FFMA.
//   backward: ONE pass per head (S <= 256).  One warp per 16-key tile walks the query tiles and forms
//             S^T = K Q^T and dP^T = V dO^T once; P^T and dS^T are already A-operand fragments for
//             dV += P^T dO and dK += dS^T Q; dS is transposed in registers (movmatrix) for
//             dQ += dS K, which is accumulated across the key warps in shared memory and written once
//             (the warps walk the query tiles in rotated order, one barrier per step, so the
//             accumulation needs no atomics -- red.shared.add.f32 compiles to a CAS loop).  Longer sequences (cfg-5's S = 785 decoder) use the two-kernel path
//             further down (dQ pass + dK/dV pass).
// Probabilities / dS are rounded to bf16 for the tensor-core products as in the reference's autocast
// graph; scores and softmax statistics stay in fp32 (the reference rounds the scores to bf16 first --
// this path is strictly more accurate there).  mma.sync.m16n8k16 (bf16, f32 accumulate):
// a tcgen05 version would leave the 128-lane MMA mostly empty at d = 32 and would not remove the
// softmax bound.
//
// The per-row statistic handed from forward to backward is L2 = m * c + log2(sum exp2(s*c - m*c)),
// c = d^-1/2 * log2(e), i.e. the log-sum-exp in the exp2 domain: p = exp2(s * c - L2).
#include "common.cuh"

extern "C" int csm_colsum_bf16(const void* dy_bf16, float* db, int rows, int N, int skip_period, int num_sms,
                               cudaStream_t stream);

namespace {
using namespace csm;


__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Copy `nrows_pad` rows of DH bf16 (row pitch ld_g elements in global) into padded smem rows of
// DH + 8 elements; rows >= nrows_valid are zero-filled.
template <int DH>
__device__ __forceinline__ void load_rows(__nv_bfloat16* s, const __nv_bfloat16* g, int row0, int nrows_valid_total,
                                          int nrows_pad, size_t ld_g) {
  constexpr int CH = DH / 8;  // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < nrows_pad * CH; idx += blockDim.x) {
    const int r = idx / CH, c = idx % CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row0 + r < nrows_valid_total) v = *reinterpret_cast<const uint4*>(g + static_cast<size_t>(row0 + r) * ld_g + c * 8);
    *reinterpret_cast<uint4*>(s + r * (DH + 8) + c * 8) = v;
  }
}

// A fragments (16 rows x DH) from padded smem rows starting at `row`.
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[DH / 16][4], const __nv_bfloat16* s, int row, int lane) {
  const int r = row + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = (lane >> 4) * 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) ldsm_x4(f[kk], smem_u32(s + r * (DH + 8) + kk * 16 + c));
}

// acc[nt] (+)= A(16 x DH) . M[row0 + nt*8 .. +8][0..DH)^T  for NT n-tiles  (B operand: n = smem row, k = column)
template <int DH, int NT>
__device__ __forceinline__ void mma_a_rowsT(float (&acc)[NT][4], const uint32_t (&a)[DH / 16][4],
                                            const __nv_bfloat16* s, int row0, int lane) {
  const int r = (lane & 7) + (lane >> 4) * 8;
  const int c = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4(b, smem_u32(s + (row0 + np * 16 + r) * (DH + 8) + kk * 16 + c));
      mma_bf16(acc[2 * np], a[kk], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[kk], b[2], b[3]);
    }
  }
}

// acc[dn] += P(16 x 16, A fragment) . M[row0 .. row0+16][0..DH)   (B operand: k = smem row, n = column)
template <int DH>
__device__ __forceinline__ void mma_p_rows(float (&acc)[DH / 8][4], const uint32_t (&a)[4], const __nv_bfloat16* s,
                                           int row0, int lane) {
  const int r = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = (lane >> 4) * 8;
#pragma unroll
  for (int dp = 0; dp < DH / 16; ++dp) {
    uint32_t b[4];
    ldsm_x4_trans(b, smem_u32(s + (row0 + r) * (DH + 8) + dp * 16 + c));
    mma_bf16(acc[2 * dp], a, b[0], b[1]);
    mma_bf16(acc[2 * dp + 1], a, b[2], b[3]);
  }
}

// Write a warp's 16 x DH f32 accumulator tile as bf16 rows through its private smem scratch.
template <int DH>
__device__ __forceinline__ void store_tile_bf16(const float (&acc)[DH / 8][4], __nv_bfloat16* scratch,
                                                __nv_bfloat16* g, int row0, int nrows_total, size_t ld_g, int lane) {
  const int gq = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int dn = 0; dn < DH / 8; ++dn) {
    *reinterpret_cast<uint32_t*>(scratch + gq * (DH + 8) + dn * 8 + 2 * t) = pack_bf16x2(acc[dn][0], acc[dn][1]);
    *reinterpret_cast<uint32_t*>(scratch + (gq + 8) * (DH + 8) + dn * 8 + 2 * t) = pack_bf16x2(acc[dn][2], acc[dn][3]);
  }
  __syncwarp();
  constexpr int CH = DH / 8;
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int r = idx / CH, c = idx % CH;
    if (row0 + r < nrows_total)
      *reinterpret_cast<uint4*>(g + static_cast<size_t>(row0 + r) * ld_g + c * 8) =
          *reinterpret_cast<const uint4*>(scratch + r * (DH + 8) + c * 8);
  }
}


__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 ld_shared_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// cooperative copy by `nthreads` threads (index tid) of rows [row0, row0 + nrows_pad) -- see load_rows
template <int DH>
__device__ __forceinline__ void load_rows_part(__nv_bfloat16* s, const __nv_bfloat16* g, int row0, int nrows_valid_total,
                                               int nrows_pad, size_t ld_g, int tid, int nthreads) {
  constexpr int CH = DH / 8;
  for (int idx = tid; idx < nrows_pad * CH; idx += nthreads) {
    const int r = idx / CH, c = idx % CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row0 + r < nrows_valid_total) v = *reinterpret_cast<const uint4*>(g + static_cast<size_t>(row0 + r) * ld_g + c * 8);
    *reinterpret_cast<uint4*>(s + r * (DH + 8) + c * 8) = v;
  }
}

// Same copy with cp.async (16 bytes per request, zero-fill for rows past the end): nothing waits on the
// loads until cp_async_wait_all(), so a CTA exposes one memory latency for everything it stages.
template <int DH>
__device__ __forceinline__ void load_rows_async(__nv_bfloat16* s, const __nv_bfloat16* g, int row0, int nrows_valid_total,
                                                int nrows_pad, size_t ld_g, int tid, int nthreads) {
  constexpr int CH = DH / 8;
  for (int idx = tid; idx < nrows_pad * CH; idx += nthreads) {
    const int r = idx / CH, c = idx % CH;
    const bool ok = row0 + r < nrows_valid_total;
    const __nv_bfloat16* src = g + static_cast<size_t>(ok ? row0 + r : 0) * ld_g + c * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(s + r * (DH + 8) + c * 8)), "l"(src),
                 "r"(ok ? 16 : 0)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// forward: grid (B*H / HPC, ceil(QT / QW)); CTA = HPC heads x QW warps, warp = one 16-query tile
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256, DH == 32 ? 3 : 2)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2,
                int S, int H, int Dm, int HPC, int QW, int QPW, float c) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int LDS = DH + 8;
  const int S16 = (S + 15) & ~15;
  const int QT = S16 >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int hl = warp / QW;                       // head slot inside the CTA
  const int wq = warp % QW;
  const int tph = QW * 32;                        // threads per head slot
  const int bh = blockIdx.x * HPC + hl;
  const int b = bh / H, h = bh % H;
  // per head slot: K [S16][LDS], V [S16][LDS], per-warp Q / output tiles [QPW][16][LDS]
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn) +
                      static_cast<size_t>(hl) * (2 * S16 + QW * QPW * 16) * LDS;
  __nv_bfloat16* sV = sK + S16 * LDS;
  __nv_bfloat16* sQw = sV + S16 * LDS + wq * QPW * 16 * LDS;
  const size_t ld = static_cast<size_t>(3) * Dm;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * S * ld + h * DH;
  __nv_bfloat16* og = out + static_cast<size_t>(b) * S * Dm + h * DH;
  const int tid_h = threadIdx.x - hl * tph;
  const int qt_begin = blockIdx.y * QW * QPW;
  const int qt_end = min(QT, qt_begin + QW * QPW);
  pdl_wait();
  pdl_trigger();
  load_rows_async<DH>(sK, base + Dm, 0, S, S16, ld, tid_h, tph);
  load_rows_async<DH>(sV, base + 2 * Dm, 0, S, S16, ld, tid_h, tph);
  for (int qt = qt_begin + wq, i = 0; qt < qt_end; qt += QW, ++i)
    load_rows_async<DH>(sQw + i * 16 * LDS, base, qt * 16, S, 16, ld, lane, 32);
  cp_async_wait_all();
  __syncthreads();

  for (int qt = qt_begin + wq, qi = 0; qt < qt_end; qt += QW, ++qi) {
    const int q0 = qt * 16;
    __nv_bfloat16* sQ = sQw + qi * 16 * LDS;
    uint32_t qf[DH / 16][4];
    load_a_frags<DH>(qf, sQ, 0, lane);
    float o[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    for (int kb0 = 0; kb0 < S16; kb0 += 64) {
      const int nt16 = min(4, (S16 - kb0) >> 4);     // 16-key tiles in this block (warp-uniform)
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      {
        const int r = (lane & 7) + (lane >> 4) * 8;
        const int cc = ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            if (np < nt16) {
              uint32_t bfr[4];
              ldsm_x4(bfr, smem_u32(sK + (kb0 + np * 16 + r) * LDS + kk * 16 + cc));
              mma_bf16(s[2 * np], qf[kk], bfr[0], bfr[1]);
              mma_bf16(s[2 * np + 1], qf[kk], bfr[2], bfr[3]);
            }
          }
        }
      }
      if (kb0 + nt16 * 16 > S) {                      // only the last 16-key tile holds padded keys
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = kb0 + nt * 8 + 2 * t + (e & 1);
            if (key >= S) s[nt][e] = -INFINITY;
          }
        }
      }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < 2 * nt16) {
          mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
          mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
      }
      float alpha[2], mc[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        const float m_new = fmaxf(m_run[hh], mx[hh]);
        alpha[hh] = ex2_approx((m_run[hh] - m_new) * c);
        m_run[hh] = m_new;
        mc[hh] = m_new * c;
        l_run[hh] *= alpha[hh];
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < 2 * nt16) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = ex2_approx(fmaf(s[nt][e], c, -mc[e >> 1]));
            s[nt][e] = pv;
            l_run[e >> 1] += pv;
          }
        }
      }
#pragma unroll
      for (int dn = 0; dn < DH / 8; ++dn) {
        o[dn][0] *= alpha[0]; o[dn][1] *= alpha[0];
        o[dn][2] *= alpha[1]; o[dn][3] *= alpha[1];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nt16) {
          uint32_t a[4];
          a[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
          a[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
          a[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
          a[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
          mma_p_rows<DH>(o, a, sV, kb0 + j * 16, lane);
        }
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
      l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) {
      o[dn][0] *= inv0; o[dn][1] *= inv0;
      o[dn][2] *= inv1; o[dn][3] *= inv1;
    }
    if (t == 0) {
      if (q0 + gq < S) lse2[static_cast<size_t>(bh) * S + q0 + gq] = m_run[0] * c + log2f(l_run[0]);
      if (q0 + gq + 8 < S) lse2[static_cast<size_t>(bh) * S + q0 + gq + 8] = m_run[1] * c + log2f(l_run[1]);
    }
    store_tile_bf16<DH>(o, sQ, og, q0, S, Dm, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, single pass (S16 <= 256): grid (B*H / HPC); CTA = HPC heads x KT warps, warp = one 16-key tile
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(DH == 64 ? 128 : 512, DH == 64 ? 3 : 1)
attn_bwd_head_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o_fwd,
                     const __nv_bfloat16* __restrict__ d_out, const float* __restrict__ lse2,
                     __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias, int S, int H, int Dm, int HPC,
                     float c, float scale) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int LDS = DH + 8;
  constexpr int LDQ = DH + 8;                     // f32 dQ rows: 64-bit accesses of a half-warp hit 32 distinct banks
  const int S16 = (S + 15) & ~15;
  const int KT = S16 >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int hl = warp / KT;
  const int kt = warp % KT;
  const int tph = KT * 32;
  const int tid_h = threadIdx.x - hl * tph;
  const int bh = blockIdx.x * HPC + hl;
  const int b = bh / H, h = bh % H;
  // per head slot: Q [S16][LDS], dO [S16][LDS], per-warp K and V tiles [16][LDS] each, dQ f32 [S16][LDQ],
  //                {L2, delta} [S16], per-query-tile step counters [16], column-sum scratch [2 KT + 16][DH]
  const size_t slot_bytes = static_cast<size_t>(2 * S16 + KT * 32) * LDS * 2 + static_cast<size_t>(S16) * LDQ * 4 +
                            static_cast<size_t>(S16) * 8 + 64 + static_cast<size_t>(2 * KT + 16) * DH * 4;
  uint8_t* slot = smem_attn + hl * slot_bytes;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(slot);
  __nv_bfloat16* sdO = sQ + S16 * LDS;
  __nv_bfloat16* sKt = sdO + S16 * LDS + kt * 32 * LDS;
  __nv_bfloat16* sVt = sKt + 16 * LDS;
  float* sdQ = reinterpret_cast<float*>(sdO + S16 * LDS + KT * 32 * LDS);
  float* sLD = sdQ + S16 * LDQ;                   // interleaved {L2, delta} per query
  int* sFlag = reinterpret_cast<int*>(sLD + 2 * S16);
  float* sCol = reinterpret_cast<float*>(sFlag + 16);   // [KT][2][DH] dK / dV column sums, then [16][DH] for dQ
  const size_t ld = static_cast<size_t>(3) * Dm;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * S * ld + h * DH;
  const size_t obase = static_cast<size_t>(b) * S * Dm + h * DH;

  pdl_wait();
  pdl_trigger();
  load_rows_async<DH>(sQ, base, 0, S, S16, ld, tid_h, tph);
  load_rows_async<DH>(sdO, d_out + obase, 0, S, S16, Dm, tid_h, tph);
  load_rows_async<DH>(sKt, base + Dm, kt * 16, S, 16, ld, lane, 32);
  load_rows_async<DH>(sVt, base + 2 * Dm, kt * 16, S, 16, ld, lane, 32);
  for (int i = tid_h; i < S16 * LDQ; i += tph) sdQ[i] = 0.f;
  if (tid_h < 16) sFlag[tid_h] = 0;
  // delta[q] = sum_j dO[q, j] * O[q, j]; L2 = +inf for padded queries (p = exp2(-inf) = 0)
  {
    constexpr int TPRW = DH / 8;                  // threads per row (16-byte chunks)
    for (int idx = tid_h; idx < S16 * TPRW; idx += tph) {
      const int r = idx / TPRW, cch = idx % TPRW;
      float part = 0.f;
      if (r < S) {
        const uint4 a = *reinterpret_cast<const uint4*>(d_out + obase + static_cast<size_t>(r) * Dm + cch * 8);
        const uint4 bb = *reinterpret_cast<const uint4*>(o_fwd + obase + static_cast<size_t>(r) * Dm + cch * 8);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
          part += x.x * y.x + x.y * y.y;
        }
      }
      // TPRW (4 or 8) consecutive lanes hold one row
#pragma unroll
      for (int off = TPRW / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (cch == 0)
        *reinterpret_cast<float2*>(sLD + 2 * r) =
            make_float2(r < S ? lse2[static_cast<size_t>(bh) * S + r] : INFINITY, part);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  uint32_t kf[DH / 16][4], vf[DH / 16][4];
  load_a_frags<DH>(kf, sKt, 0, lane);
  load_a_frags<DH>(vf, sVt, 0, lane);
  // the K tile as the B operand of dQ += dS K (k = key, n = feature): loop invariant
  uint32_t kb[DH / 16][4];
  {
    const int r = (lane & 7) + ((lane >> 3) & 1) * 8, cc = (lane >> 4) * 8;
#pragma unroll
    for (int dp = 0; dp < DH / 16; ++dp) ldsm_x4_trans(kb[dp], smem_u32(sKt + r * LDS + dp * 16 + cc));
  }
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  const bool tail_keys = (kt * 16 + 16 > S);       // this warp's tile holds padded keys (warp-uniform)
  const bool key_ok0 = kt * 16 + gq < S, key_ok1 = kt * 16 + gq + 8 < S;

  // per-lane shared-memory addresses; a step only adds its query-tile offset
  const uint32_t rowsT = ((lane & 7) + (lane >> 4) * 8) * LDS + ((lane >> 3) & 1) * 8;   // B operand, n = row
  const uint32_t rowsP = ((lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (lane >> 4) * 8;   // B operand, k = row (.trans)
  const uint32_t aQ = smem_u32(sQ + rowsT), adO = smem_u32(sdO + rowsT);
  const uint32_t pQ = smem_u32(sQ + rowsP), pdO = smem_u32(sdO + rowsP);
  const uint32_t aLD = smem_u32(sLD + 4 * t);          // float2 {L2, delta} of query 2t
  const uint32_t aDQ = smem_u32(sdQ + gq * LDQ + 2 * t);
  const uint32_t aFlag = smem_u32(sFlag);

  // Step i: the warp of key tile kt works on query tile (kt + i) % KT, so within a step no two warps of a
  // head touch the same dQ rows and the accumulation is a plain read-modify-write (no shared atomics --
  // red.shared.add.f32 compiles to a CAS loop).  Tile q was last updated by warp kt+1 in step i-1: a per-tile
  // step counter (release / acquire in shared memory) orders the two, no CTA-wide barrier in the loop.
  int qt = kt;
#pragma unroll 1
  for (int step = 0; step < KT; ++step) {
    const uint32_t qoff = static_cast<uint32_t>(qt) * (16 * LDS * 2);
    float st[2][4], dpt[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t bq[4], bo[4];
      ldsm_x4(bq, aQ + qoff + kk * 32);
      ldsm_x4(bo, adO + qoff + kk * 32);
      mma_bf16(st[0], kf[kk], bq[0], bq[1]);        // S^T  = K Q^T   [16 keys x 16 queries]
      mma_bf16(st[1], kf[kk], bq[2], bq[3]);
      mma_bf16(dpt[0], vf[kk], bo[0], bo[1]);       // dP^T = V dO^T
      mma_bf16(dpt[1], vf[kk], bo[2], bo[3]);
    }
    uint32_t pa[4], da[4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      // {L2, delta} of queries 2t and 2t+1 of this 8-query group
      const uint4 ldv = ld_shared_u4(aLD + static_cast<uint32_t>(qt * 16 + nt * 8) * 8);
      const float l2a = __uint_as_float(ldv.x), dla = __uint_as_float(ldv.y);
      const float l2b = __uint_as_float(ldv.z), dlb = __uint_as_float(ldv.w);
      float pv[4], dsv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = ex2_approx(fmaf(st[nt][e], c, -((e & 1) ? l2b : l2a)));
        pv[e] = p;
        dsv[e] = p * (dpt[nt][e] - ((e & 1) ? dlb : dla));
      }
      if (tail_keys) {
        if (!key_ok0) { pv[0] = pv[1] = 0.f; dsv[0] = dsv[1] = 0.f; }
        if (!key_ok1) { pv[2] = pv[3] = 0.f; dsv[2] = dsv[3] = 0.f; }
      }
      pa[nt * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
      pa[nt * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
      da[nt * 2 + 0] = pack_bf16x2(dsv[0], dsv[1]);
      da[nt * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
    }
#pragma unroll
    for (int dp = 0; dp < DH / 16; ++dp) {
      uint32_t bo[4], bq[4];
      ldsm_x4_trans(bo, pdO + qoff + dp * 32);
      ldsm_x4_trans(bq, pQ + qoff + dp * 32);
      mma_bf16(dv[2 * dp], pa, bo[0], bo[1]);       // dV += P^T dO
      mma_bf16(dv[2 * dp + 1], pa, bo[2], bo[3]);
      mma_bf16(dk[2 * dp], da, bq[0], bq[1]);       // dK += dS^T Q   (x scale at the end)
      mma_bf16(dk[2 * dp + 1], da, bq[2], bq[3]);
    }
    // dQ[16 q, :] += dS[16 q x 16 keys] K[16 keys, :]: transpose the dS^T fragments in registers
    uint32_t dst[4];
    dst[0] = movmatrix_trans(da[0]);
    dst[1] = movmatrix_trans(da[2]);
    dst[2] = movmatrix_trans(da[1]);
    dst[3] = movmatrix_trans(da[3]);
    float dq[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
    for (int dp = 0; dp < DH / 16; ++dp) {
      mma_bf16(dq[2 * dp], dst, kb[dp][0], kb[dp][1]);
      mma_bf16(dq[2 * dp + 1], dst, kb[dp][2], kb[dp][3]);
    }
    // wait until the previous owner of this query tile (step - 1) has published its update
    {
      const uint32_t fa = aFlag + qt * 4;
      uint32_t seen;
      do {
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(fa) : "memory");
      } while (static_cast<int>(seen) < step);
      const uint32_t qlo = aDQ + static_cast<uint32_t>(qt) * (16 * LDQ * 4);
      const uint32_t qhi = qlo + 8 * LDQ * 4;
#pragma unroll
      for (int dn = 0; dn < DH / 8; ++dn) {
        float2 lo = ld_shared_f2(qlo + dn * 32);
        float2 hi = ld_shared_f2(qhi + dn * 32);
        lo.x += dq[dn][0]; lo.y += dq[dn][1];
        hi.x += dq[dn][2]; hi.y += dq[dn][3];
        st_shared_f2(qlo + dn * 32, lo);
        st_shared_f2(qhi + dn * 32, hi);
      }
      __syncwarp();
      if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(fa), "r"(step + 1) : "memory");
    }
    qt = (qt + 1 == KT) ? 0 : qt + 1;
  }
  __syncthreads();      // every dQ tile has received all KT updates
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    dk[i][0] *= scale; dk[i][1] *= scale; dk[i][2] *= scale; dk[i][3] *= scale;
  }
  __nv_bfloat16* dkg = dqkv + static_cast<size_t>(b) * S * ld + Dm + h * DH;
  __nv_bfloat16* dvg = dqkv + static_cast<size_t>(b) * S * ld + 2 * Dm + h * DH;
  // both tiles go out through this warp's private K-tile scratch (its fragments are already in registers)
  store_tile_bf16<DH>(dk, sKt, dkg, kt * 16, S, ld, lane);
  store_tile_bf16<DH>(dv, sKt, dvg, kt * 16, S, ld, lane);
  // qkv.bias gradient = column sums of dqkv over the tokens: taken here from the tiles still on chip instead of
  // re-reading dqkv from HBM (rows past S are exactly zero)
  if (dbias != nullptr) {
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) {
      float k0 = dk[dn][0] + dk[dn][2], k1 = dk[dn][1] + dk[dn][3];
      float v0 = dv[dn][0] + dv[dn][2], v1 = dv[dn][1] + dv[dn][3];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        k0 += __shfl_xor_sync(0xffffffffu, k0, off);
        k1 += __shfl_xor_sync(0xffffffffu, k1, off);
        v0 += __shfl_xor_sync(0xffffffffu, v0, off);
        v1 += __shfl_xor_sync(0xffffffffu, v1, off);
      }
      if (gq == 0) {
        *reinterpret_cast<float2*>(sCol + (kt * 2 + 0) * DH + dn * 8 + 2 * t) = make_float2(k0, k1);
        *reinterpret_cast<float2*>(sCol + (kt * 2 + 1) * DH + dn * 8 + 2 * t) = make_float2(v0, v1);
      }
    }
    const int ngrp = max(1, min(16, tph / DH));
    for (int idx = tid_h; idx < ngrp * DH; idx += tph) {
      const int colq = idx % DH, grp = idx / DH;
      float a = 0.f;
      for (int r = grp; r < S; r += ngrp) a += sdQ[r * LDQ + colq];
      sCol[(2 * KT + grp) * DH + colq] = a * scale;
    }
    __syncthreads();
    for (int idx = tid_h; idx < 3 * DH; idx += tph) {
      const int which = idx / DH, col = idx % DH;
      float a = 0.f;
      if (which == 0) {
        for (int gI = 0; gI < ngrp; ++gI) a += sCol[(2 * KT + gI) * DH + col];
      } else {
        for (int k2 = 0; k2 < KT; ++k2) a += sCol[(k2 * 2 + which - 1) * DH + col];
      }
      atomicAdd(dbias + which * Dm + h * DH + col, a);
    }
  }
  // dQ = scale * sum over key tiles, written once (the last step's barrier already ordered the accumulation)
  {
    __nv_bfloat16* dqg = dqkv + static_cast<size_t>(b) * S * ld + h * DH;
    constexpr int CH = DH / 8;
    for (int idx = tid_h; idx < S * CH; idx += tph) {
      const int r = idx / CH, cch = idx % CH;
      const float4 a = *reinterpret_cast<const float4*>(sdQ + r * LDQ + cch * 8);
      const float4 bb = *reinterpret_cast<const float4*>(sdQ + r * LDQ + cch * 8 + 4);
      uint4 pk;
      pk.x = pack_bf16x2(a.x * scale, a.y * scale); pk.y = pack_bf16x2(a.z * scale, a.w * scale);
      pk.z = pack_bf16x2(bb.x * scale, bb.y * scale); pk.w = pack_bf16x2(bb.z * scale, bb.w * scale);
      *reinterpret_cast<uint4*>(dqg + static_cast<size_t>(r) * ld + cch * 8) = pk;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// long sequences (S16 > 256), two kernels; CTAs of blockDim/32 warps = RPC = 16 * warps rows each (16 warps at
// d = 32, 8 at d = 64: the whole K/V or Q/dO of the head fills most of an SM's shared memory, so the CTA has
// to bring its own occupancy).
// backward, dQ: grid (B*H, ceil(S/64)); K,V of the head + the Q/dO/O rows of this block in smem.
// Also writes delta[bh, s] = sum_j dO[s, j] * O[s, j] for the dK/dV kernel.
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(DH == 32 ? 512 : 256)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o_fwd,
                   const __nv_bfloat16* __restrict__ d_out, const float* __restrict__ lse,
                   float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int S, int H, int Dm, float c,
                   float scale) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int LDS = DH + 8;
  const int S_pad = (S + 31) / 32 * 32;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sV = sK + S_pad * LDS;
  const int rpc = (blockDim.x >> 5) * 16;          // query rows per CTA
  __nv_bfloat16* sQ = sV + S_pad * LDS;
  __nv_bfloat16* sdO = sQ + rpc * LDS;
  __nv_bfloat16* sO = sdO + rpc * LDS;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int q0 = blockIdx.y * rpc;
  const size_t ld = static_cast<size_t>(3) * Dm;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * S * ld + h * DH;
  const size_t obase = static_cast<size_t>(b) * S * Dm + h * DH;
  load_rows<DH>(sK, base + Dm, 0, S, S_pad, ld);
  load_rows<DH>(sV, base + 2 * Dm, 0, S, S_pad, ld);
  load_rows<DH>(sQ, base, q0, S, rpc, ld);
  load_rows<DH>(sdO, d_out + obase, q0, S, rpc, Dm);
  load_rows<DH>(sO, o_fwd + obase, q0, S, rpc, Dm);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int r0 = q0 + warp * 16;
  if (r0 >= S) return;

  // delta: lane pair (2r, 2r+1) reduces row r of this warp's tile
  float dsum = 0.f;
  {
    const int r = warp * 16 + (lane >> 1);
    const int c0 = (lane & 1) * (DH / 2);
#pragma unroll
    for (int c = 0; c < DH / 2; c += 2) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sdO + r * LDS + c0 + c));
      const float2 bb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sO + r * LDS + c0 + c));
      dsum += a.x * bb.x + a.y * bb.y;
    }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    if ((lane & 1) == 0 && r0 + (lane >> 1) < S) delta[static_cast<size_t>(bh) * S + r0 + (lane >> 1)] = dsum;
  }
  float dl[2], ls[2];
  dl[0] = __shfl_sync(0xffffffffu, dsum, 2 * gq);
  dl[1] = __shfl_sync(0xffffffffu, dsum, 2 * (gq + 8));
  ls[0] = r0 + gq < S ? lse[static_cast<size_t>(bh) * S + r0 + gq] : INFINITY;
  ls[1] = r0 + gq + 8 < S ? lse[static_cast<size_t>(bh) * S + r0 + gq + 8] : INFINITY;

  uint32_t qf[DH / 16][4], dof[DH / 16][4];
  load_a_frags<DH>(qf, sQ, warp * 16, lane);
  load_a_frags<DH>(dof, sdO, warp * 16, lane);
  float dq[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

  for (int kb0 = 0; kb0 < S_pad; kb0 += 32) {
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
    mma_a_rowsT<DH, 4>(s, qf, sK, kb0, lane);
    mma_a_rowsT<DH, 4>(dp, dof, sV, kb0, lane);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb0 + nt * 8 + 2 * t + (e & 1);
        const float p = key < S ? ex2_approx(fmaf(s[nt][e], c, -ls[e >> 1])) : 0.f;
        s[nt][e] = p * (dp[nt][e] - dl[e >> 1]) * scale;
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      a[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      a[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      a[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
      mma_p_rows<DH>(dq, a, sK, kb0 + j * 16, lane);
    }
  }
  __nv_bfloat16* dqg = dqkv + static_cast<size_t>(b) * S * ld + h * DH;
  store_tile_bf16<DH>(dq, sQ + warp * 16 * LDS, dqg, r0, S, ld, lane);
}

// ---------------------------------------------------------------------------------------------
// backward, dK/dV: grid (B*H, ceil(S/64)); Q,dO of the head + the K/V rows of this block in smem.
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(DH == 32 ? 512 : 256)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_out,
                    const float* __restrict__ lse, const float* __restrict__ delta,
                    __nv_bfloat16* __restrict__ dqkv, int S, int H, int Dm, float c, float scale) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int LDS = DH + 8;
  const int S_pad = (S + 31) / 32 * 32;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sdO = sQ + S_pad * LDS;
  const int rpc = (blockDim.x >> 5) * 16;          // key rows per CTA
  __nv_bfloat16* sK = sdO + S_pad * LDS;
  __nv_bfloat16* sV = sK + rpc * LDS;
  float* sLse = reinterpret_cast<float*>(sV + rpc * LDS);
  float* sDelta = sLse + S_pad;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int k0 = blockIdx.y * rpc;
  const size_t ld = static_cast<size_t>(3) * Dm;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * S * ld + h * DH;
  const size_t obase = static_cast<size_t>(b) * S * Dm + h * DH;
  load_rows<DH>(sQ, base, 0, S, S_pad, ld);
  load_rows<DH>(sdO, d_out + obase, 0, S, S_pad, Dm);
  load_rows<DH>(sK, base + Dm, k0, S, rpc, ld);
  load_rows<DH>(sV, base + 2 * Dm, k0, S, rpc, ld);
  for (int i = threadIdx.x; i < S_pad; i += blockDim.x) {
    sLse[i] = i < S ? lse[static_cast<size_t>(bh) * S + i] : INFINITY;
    sDelta[i] = i < S ? delta[static_cast<size_t>(bh) * S + i] : 0.f;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane & 3;
  const int kr0 = k0 + warp * 16;
  if (kr0 >= S) return;

  uint32_t kf[DH / 16][4], vf[DH / 16][4];
  load_a_frags<DH>(kf, sK, warp * 16, lane);
  load_a_frags<DH>(vf, sV, warp * 16, lane);
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }

  for (int qb0 = 0; qb0 < S_pad; qb0 += 32) {
    float st[4][4], dpt[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
    mma_a_rowsT<DH, 4>(st, kf, sQ, qb0, lane);
    mma_a_rowsT<DH, 4>(dpt, vf, sdO, qb0, lane);
    uint32_t pa[2][4], da[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float pv[4], dsv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int query = qb0 + nt * 8 + 2 * t + (e & 1);
        const float p = ex2_approx(fmaf(st[nt][e], c, -sLse[query]));   // L2 = +inf for padded queries -> 0
        pv[e] = p;
        dsv[e] = p * (dpt[nt][e] - sDelta[query]) * scale;
      }
      const int j = nt >> 1, hi = (nt & 1) * 2;
      pa[j][hi + 0] = pack_bf16x2(pv[0], pv[1]);
      pa[j][hi + 1] = pack_bf16x2(pv[2], pv[3]);
      da[j][hi + 0] = pack_bf16x2(dsv[0], dsv[1]);
      da[j][hi + 1] = pack_bf16x2(dsv[2], dsv[3]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      mma_p_rows<DH>(dv, pa[j], sdO, qb0 + j * 16, lane);
      mma_p_rows<DH>(dk, da[j], sQ, qb0 + j * 16, lane);
    }
  }
  __nv_bfloat16* dkg = dqkv + static_cast<size_t>(b) * S * ld + Dm + h * DH;
  __nv_bfloat16* dvg = dqkv + static_cast<size_t>(b) * S * ld + 2 * Dm + h * DH;
  // both tiles go out through this warp's private K-row scratch (its fragments are already in registers)
  store_tile_bf16<DH>(dk, sK + warp * 16 * LDS, dkg, kr0, S, ld, lane);
  store_tile_bf16<DH>(dv, sK + warp * 16 * LDS, dvg, kr0, S, ld, lane);
}

template <typename K>
int set_smem(K kern, size_t bytes, size_t* configured, const char* name) {
  if (bytes <= *configured) return CSM_OK;
  // these kernels want as many co-resident CTAs as shared memory allows: ask for the full carve-out
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (bytes > 227 * 1024) {
    csm_set_error("%s: sequence too long for the whole-head-in-shared-memory kernel (%zu bytes needed)", name, bytes);
    return CSM_ERR_ARG;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    csm_set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", name, bytes, cudaGetErrorString(e));
    return CSM_ERR_CUDA;
  }
  *configured = bytes;
  return CSM_OK;
}

template <int DH>
int attn_fwd_launch(const void* qkv, void* out, float* lse, int B, int S, int H, cudaStream_t stream) {
  const int Dm = H * DH;
  const int S16 = (S + 15) & ~15;
  const int QT = S16 >> 4;
  // CTAs of at most 8 warps (so that 2-3 of them co-reside and one CTA's staging overlaps another's math):
  // HPC heads per CTA for short sequences, QPW query tiles per warp for long ones
  int HPC = QT >= 8 ? 1 : (QT >= 4 ? 2 : 4);
  while (HPC > 1 && (B * H) % HPC != 0) --HPC;
  const int wmax = 8 / HPC;
  int QPW = (QT + wmax - 1) / wmax;
  if (QPW > 4) QPW = 4;
  int QW = (QT + QPW - 1) / QPW;
  if (QW > wmax) QW = wmax;
  const size_t smem = static_cast<size_t>(HPC) * (2 * S16 + QW * QPW * 16) * (DH + 8) * 2;
  static size_t cfg_fwd = 0;
  int rc = set_smem(attn_fwd_kernel<DH>, smem, &cfg_fwd, "attention_fwd");
  if (rc) return rc;
  const float c = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
  dim3 grid(B * H / HPC, (QT + QW * QPW - 1) / (QW * QPW));
  cudaError_t le = csm_launch_pdl(attn_fwd_kernel<DH>, grid, dim3(HPC * QW * 32), smem, stream,
                                  reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), lse,
                                  S, H, Dm, HPC, QW, QPW, c);
  if (le != cudaSuccess) {
    csm_set_error("attention_fwd: launch failed: %s", cudaGetErrorString(le));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

template <int DH>
int attn_bwd_launch(const void* qkv, const void* o, const void* d_out, const float* lse, float* delta, void* dqkv,
                    float* dbias, int B, int S, int H, cudaStream_t stream) {
  const int Dm = H * DH;
  const float scale = 1.0f / sqrtf(static_cast<float>(DH));
  const float c = 1.4426950408889634f * scale;
  const int S16 = (S + 15) & ~15;
  if (S16 <= (DH == 64 ? 64 : 256)) {   // d = 64 needs ~170 registers per thread: CTAs of 4 warps, 3 per SM
    const int KT = S16 >> 4;
    const int max_warps = DH == 64 ? 4 : 16;       // the kernel's launch bound
    int HPC = 8 / KT;                               // heads per CTA for short sequences (<= 8 warps)
    if (HPC < 1) HPC = 1;
    if (HPC > 4) HPC = 4;
    while (HPC > 1 && ((B * H) % HPC != 0 || HPC * KT > max_warps)) --HPC;
    const size_t slot = static_cast<size_t>(2 * S16 + KT * 32) * (DH + 8) * 2 + static_cast<size_t>(S16) * (DH + 8) * 4 +
                        static_cast<size_t>(S16) * 8 + 64 + static_cast<size_t>(2 * KT + 16) * DH * 4;
    static size_t cfg_head = 0;
    int rc = set_smem(attn_bwd_head_kernel<DH>, slot * HPC, &cfg_head, "attention_bwd");
    if (rc) return rc;
    cudaError_t le = csm_launch_pdl(attn_bwd_head_kernel<DH>, dim3(B * H / HPC), dim3(HPC * KT * 32), slot * HPC, stream,
                                    reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(o),
                                    reinterpret_cast<const __nv_bfloat16*>(d_out), lse,
                                    reinterpret_cast<__nv_bfloat16*>(dqkv), dbias, S, H, Dm, HPC, c, scale);
    if (le != cudaSuccess) {
      csm_set_error("attention_bwd: launch failed: %s", cudaGetErrorString(le));
      return CSM_ERR_CUDA;
    }
    return CSM_OK;
  }
  CSM_CHECK_ARG(delta != nullptr, "csm_attention_bwd: S=%d needs the delta scratch buffer", S);
  const int S_pad = (S + 31) / 32 * 32;
  const int warps = DH == 32 ? 16 : 8;
  const int rpc = warps * 16;
  dim3 grid(B * H, (S + rpc - 1) / rpc);
  const size_t smem_dq = static_cast<size_t>(2 * S_pad + 3 * rpc) * (DH + 8) * 2;
  static size_t cfg_dq = 0, cfg_dkv = 0;
  int rc = set_smem(attn_bwd_dq_kernel<DH>, smem_dq, &cfg_dq, "attention_bwd_dq");
  if (rc) return rc;
  attn_bwd_dq_kernel<DH><<<grid, warps * 32, smem_dq, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(o),
      reinterpret_cast<const __nv_bfloat16*>(d_out), lse, delta, reinterpret_cast<__nv_bfloat16*>(dqkv), S, H, Dm, c,
      scale);
  CSM_CHECK_LAUNCH("attention_bwd_dq");
  const size_t smem_dkv = static_cast<size_t>(2 * S_pad + 2 * rpc) * (DH + 8) * 2 + static_cast<size_t>(2) * S_pad * 4;
  rc = set_smem(attn_bwd_dkv_kernel<DH>, smem_dkv, &cfg_dkv, "attention_bwd_dkv");
  if (rc) return rc;
  attn_bwd_dkv_kernel<DH><<<grid, warps * 32, smem_dkv, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(d_out), lse, delta,
      reinterpret_cast<__nv_bfloat16*>(dqkv), S, H, Dm, c, scale);
  CSM_CHECK_LAUNCH("attention_bwd_dkv");
  // the long-sequence path has no fused bias sums: one column-sum pass over dqkv
  if (dbias != nullptr) return csm_colsum_bf16(dqkv, dbias, B * S, 3 * Dm, 0, 0, stream);
  return CSM_OK;
}

}  // namespace

extern "C" int csm_attention_fwd_legacy(const void* qkv_bf16, void* out_bf16, float* lse, int B, int S, int H,
                                        int head_dim, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && S > 0 && H > 0, "csm_attention_fwd: bad sizes B=%d S=%d H=%d", B, S, H);
  if (head_dim == 32) return attn_fwd_launch<32>(qkv_bf16, out_bf16, lse, B, S, H, stream);
  if (head_dim == 64) return attn_fwd_launch<64>(qkv_bf16, out_bf16, lse, B, S, H, stream);
  csm_set_error("csm_attention_fwd: head_dim must be 32 or 64 (got %d)", head_dim);
  return CSM_ERR_ARG;
}

extern "C" int csm_attention_bwd_legacy(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16,
                                        const float* lse, float* delta_scratch, void* dqkv_bf16, float* dbias, int B,
                                        int S, int H, int head_dim, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && S > 0 && H > 0, "csm_attention_bwd: bad sizes B=%d S=%d H=%d", B, S, H);
  if (head_dim == 32)
    return attn_bwd_launch<32>(qkv_bf16, out_bf16, d_out_bf16, lse, delta_scratch, dqkv_bf16, dbias, B, S, H, stream);
  if (head_dim == 64)
    return attn_bwd_launch<64>(qkv_bf16, out_bf16, d_out_bf16, lse, delta_scratch, dqkv_bf16, dbias, B, S, H, stream);
  csm_set_error("csm_attention_bwd: head_dim must be 32 or 64 (got %d)", head_dim);
  return CSM_ERR_ARG;
}
