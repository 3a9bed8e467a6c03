// tcgen05 / TMEM / TMA fused multi-head self-attention for the timm Block attention of the MAE encoder
// (S = keep+1, d = 64) and decoder (S = L+1, d = 32): softmax((q k^T) * d^-1/2) v, no mask, no dropout
// (timm 0.4.12 Attention as restated in oracle/timm_shim.py; call sites models_mae/MAE_ViT_Baseline.py:160-188).
//
// Layout: qkv is the [rows, 3*Dm] bf16 output of the qkv Linear, row = b*S + s, columns [which*Dm + h*d + j]; the
// output is [rows, Dm] with column h*d + j (== transpose(1,2).reshape).  The S x S score matrix never exists in HBM.
//
// Forward (attn_fwd_tc_kernel): a persistent CTA per SM walks work items = (image, 64-column head group, 128-query
// tile).  A head group is one d = 64 head or two d = 32 heads: every TMA box is {64 columns, rows} of the qkv matrix
// with the 128-byte swizzle, so the Q / K tiles are K-major UMMA operands (a d = 32 head is two of the four 16-wide
// k-steps of the row) and the V tile is the MN-major B operand of P.V.
//   warp 0     TMA producer: Q tile (2 stages) and K/V blocks (2 stages), prefetching across work items
//   warp 1     TMEM allocator + single-thread MMA issuer:
//                S = Q K^T        tcgen05.mma  M = 128, N = BN <= 208 keys, K = d      -> TMEM (2 buffers)
//                O = P V          tcgen05.mma  M = 128, N = d, K = BN                  -> TMEM
//   warps 4-11 softmax: thread = (query row, half of the key block).  One tcgen05.ld pass brings the thread's half row
//              of S into registers (each score is read from TMEM exactly once), the two halves exchange the row maximum
//              through shared memory, p = exp2(s*c - m*c) goes to shared memory as the bf16 K-major A operand of the
//              P.V product (128-byte swizzle, written conflict-free), O is read back with tcgen05.ld, rescaled in
//              registers when the keys span several blocks (online softmax) and stored as bf16.
// Short sequences (S <= 64, the ViT-B encoder's 50 tokens) pack two images into one 128-row tile: rows / keys
// [0, 64) are image a, [64, 128) image b; the off-diagonal blocks of P are written as zeros.
//
// Probabilities are rounded to bf16 for the tensor-core product as in the reference's autocast graph; scores and
// softmax statistics stay in fp32 (the reference rounds the scores to bf16 first -- this path is strictly more
// accurate there).  The per-row statistic handed to the backward is L2 = m*c + log2(sum exp2(s*c - m*c)),
// c = d^-1/2 * log2(e): p = exp2(s*c - L2).
#include "common.cuh"

#include <cstdio>

namespace {
using namespace csm;

constexpr int AT_SM_WARPS = 8;
constexpr int AT_THREADS = 32 * (4 + AT_SM_WARPS);   // 384: warp group 0 = {TMA, MMA, 2 idle}, groups 1-2 = softmax
constexpr int AT_QBYTES = 128 * 128;                 // one Q tile: 128 rows x 64 bf16
constexpr int AT_MAXCH = 13;                         // 8-column chunks per softmax thread: BN / 2 / 8, BN <= 208
constexpr int AT_XCH_BYTES = 2 * 2 * 2 * 128 * 4;    // {max, sum} x parity x half x row

struct FwdParams {
  int B, S, H, Dm, HG;
  int pack;          // two images per 128-row tile (S <= 64)
  int QT, nb, BN;    // query tiles per image, key blocks, keys per block (multiple of 16)
  int items;
  float c;           // d^-1/2 * log2(e)
  __nv_bfloat16* out;
  float* lse;
};

__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  // bounded wait: a protocol error traps (the launch fails) instead of hanging the device
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("csmae_b200 attention: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int N>
struct IC {
  static constexpr int value = N;
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int DH, int VN>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const FwdParams p) {
  constexpr int NH = 64 / DH;          // heads per 64-column group
  constexpr int KS = DH / 16;          // 16-wide k-steps of one head in the Q / K rows
  constexpr int OC = DH / 2;           // output columns owned by one softmax thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.BN;
  const int kv_bytes = BN * 128;                       // one K (or V) block
  const int pslabs = (BN + 63) >> 6;
  uint8_t* sQ = smem;                                  // [2][16 KB]
  uint8_t* sKV = smem + 2 * AT_QBYTES;                 // [2][K | V]
  uint8_t* sP = sKV + 4 * kv_bytes;                    // [pslabs][128 rows][128 B]
  float* xch = reinterpret_cast<float*>(sP + pslabs * 16384);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xch) + AT_XCH_BYTES);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;       // [2]
  uint64_t* kv_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;        // [2]
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, AT_SM_WARPS);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  // register budget: the softmax threads hold half a score row (up to 104 values) in registers
  // Iteration order inside a work item: heads of the group outermost, key blocks innermost, so one set of online-
  // softmax state is live at a time.  With a single key block both heads share one K/V load; with several blocks the
  // blocks are re-fetched per head (L2 hits).
  const bool shared_kv = (p.nb == 1);
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t ic = 0, kvc = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
          int hg, row_q0, row_q1, row_k0, row_k1;
          if (!p.pack) {
            const int qt = item % p.QT;
            const int r = item / p.QT;
            hg = r % p.HG;
            const int b = r / p.HG;
            row_q0 = b * p.S + qt * 128;
            row_k0 = b * p.S;
            row_q1 = row_k1 = 0;
          } else {
            hg = item % p.HG;
            const int pair = item / p.HG;
            row_q0 = row_k0 = (2 * pair) * p.S;
            row_q1 = row_k1 = (2 * pair + 1) * p.S;
          }
          const int nh = min(NH, p.H - hg * NH);
          const uint32_t qs = ic & 1;
          mbar_wait_wd(&q_empty[qs], ((ic >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qs], AT_QBYTES);
          uint8_t* q = sQ + qs * AT_QBYTES;
          tma_load_2d(q, &tmap_q, &q_full[qs], hg * 64, row_q0);
          if (p.pack) tma_load_2d(q + 8192, &tmap_q, &q_full[qs], hg * 64, row_q1);
          const int nloads = shared_kv ? 1 : nh * p.nb;
          for (int l = 0; l < nloads; ++l, ++kvc) {
            const int j = shared_kv ? 0 : l % p.nb;
            const uint32_t ks = kvc & 1;
            mbar_wait_wd(&kv_empty[ks], ((kvc >> 1) & 1) ^ 1);
            mbar_expect_tx(&kv_full[ks], 2u * kv_bytes);
            uint8_t* k = sKV + ks * 2 * kv_bytes;
            uint8_t* v = k + kv_bytes;
            tma_load_2d(k, &tmap_kv, &kv_full[ks], p.Dm + hg * 64, row_k0 + j * BN);
            tma_load_2d(v, &tmap_kv, &kv_full[ks], 2 * p.Dm + hg * 64, row_k0 + j * BN);
            if (p.pack) {
              tma_load_2d(k + 8192, &tmap_kv, &kv_full[ks], p.Dm + hg * 64, row_k1);
              tma_load_2d(v + 8192, &tmap_kv, &kv_full[ks], 2 * p.Dm + hg * 64, row_k1);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer (one thread) ------------------------------
      if (lane == 0) {
        const uint32_t idesc_s = umma_idesc_bf16(128, BN, 0, 0);
        const uint32_t idesc_o = umma_idesc_bf16(128, VN, 0, 1);
        const uint32_t p_addr = smem_u32(sP);
        const int ksteps = BN >> 4;
        uint32_t it = 0, ic = 0, kvc = 0;
        // the P.V product of an iteration is issued after the NEXT S = Q K^T, so the softmax warps always find their
        // next score tile ready
        bool have_prev = false;
        uint32_t pv_it = 0, pv_ks = 0, pv_qs = 0;
        int pv_hh = 0;
        bool pv_last_kv = false, pv_last_item = false;
        auto issue_pv = [&]() {
          mbar_wait_wd(p_full, pv_it & 1);
          tc_fence_after();
          const uint32_t v_addr =
              smem_u32(sKV + pv_ks * 2 * kv_bytes + kv_bytes) + ((DH == 32 && VN == 32) ? pv_hh * 64 : 0);
          const uint32_t tmem_o = tmem_base + 2 * BN;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = umma_smem_desc_sw128(p_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(v_addr + k * 2048, 8192, 1024);
            umma_f16(tmem_o, da, db, idesc_o, k > 0 ? 1u : 0u);
          }
          umma_commit(o_full);
          if (pv_last_kv) umma_commit(&kv_empty[pv_ks]);
          if (pv_last_item) umma_commit(&q_empty[pv_qs]);
        };
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
          const int hg = p.pack ? item % p.HG : (item / p.QT) % p.HG;
          const int nh = min(NH, p.H - hg * NH);
          const uint32_t qs = ic & 1;
          mbar_wait_wd(&q_full[qs], (ic >> 1) & 1);
          const uint32_t q_addr = smem_u32(sQ + qs * AT_QBYTES);
          uint32_t ks = 0;
          for (int hh = 0; hh < nh; ++hh) {
            for (int j = 0; j < p.nb; ++j, ++it) {
              if (!(shared_kv && hh > 0)) {
                ks = kvc & 1;
                mbar_wait_wd(&kv_full[ks], (kvc >> 1) & 1);
                ++kvc;
              }
              tc_fence_after();
              const uint32_t k_addr = smem_u32(sKV + ks * 2 * kv_bytes);
              const uint32_t tmem_s = tmem_base + (it & 1) * BN;
#pragma unroll
              for (int kk = 0; kk < KS; ++kk) {
                const uint64_t da = umma_smem_desc_sw128(q_addr + (hh * KS + kk) * 32, 16, 1024);
                const uint64_t db = umma_smem_desc_sw128(k_addr + (hh * KS + kk) * 32, 16, 1024);
                umma_f16(tmem_s, da, db, idesc_s, kk > 0 ? 1u : 0u);
              }
              umma_commit(&s_full[it & 1]);
              if (have_prev) issue_pv();
              have_prev = true;
              pv_it = it; pv_ks = ks; pv_qs = qs; pv_hh = hh;
              pv_last_kv = shared_kv ? (hh == nh - 1) : true;
              pv_last_item = (hh == nh - 1) && (j == p.nb - 1);
            }
          }
        }
        if (have_prev) issue_pv();
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------ softmax warps ------------------------------
    const int sw = warp - 4;
    const int half = sw >> 2;                 // which half of the key block
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int BNh = BN >> 1;
    const int nch = BNh >> 3;
    const int colbase = half * BNh;
    const float c = p.c;
    float* xmax = xch;                        // [parity][half][row]
    float* xsum = xch + 512;
    const uint32_t p_row = smem_u32(sP) + row * 128;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);

    float m_run = -INFINITY, l_part = 0.f, alpha_pend = 0.f;
    float o_acc[OC];
#pragma unroll
    for (int i = 0; i < OC; ++i) o_acc[i] = 0.f;
    uint32_t it = 0;
    bool have_prev = false, prev_last = false, prev_valid = false;
    uint32_t prev_it = 0;
    int prev_hh = 0;
    __nv_bfloat16* prev_out = nullptr;
    float* prev_lse = nullptr;

    // read back O of iteration prev_it (its P.V product is complete: the P tile may be rewritten afterwards)
    auto drain = [&]() {
      mbar_wait_wd(o_full, prev_it & 1);
      tc_fence_after();
      uint32_t o[OC];
      const uint32_t oaddr = tmem_base + lane_off + 2 * BN + ((DH == 32 && VN == 64) ? prev_hh * 32 : 0) + half * OC;
      tmem_ld_32x16(oaddr, o);
      if (OC == 32) tmem_ld_32x16(oaddr + 16, o + (OC == 32 ? 16 : 0));
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < OC; ++i) o_acc[i] = fmaf(o_acc[i], alpha_pend, __uint_as_float(o[i]));
      if (prev_last) {
        const float l_tot = l_part + xsum[(prev_it & 1) * 256 + (half ^ 1) * 128 + row];
        if (prev_valid) {
          const float inv = 1.0f / l_tot;
#pragma unroll
          for (int g = 0; g < OC / 8; ++g) {
            uint4 pk;
            pk.x = pack_bf16x2(o_acc[g * 8 + 0] * inv, o_acc[g * 8 + 1] * inv);
            pk.y = pack_bf16x2(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
            pk.z = pack_bf16x2(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
            pk.w = pack_bf16x2(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(prev_out + g * 8) = pk;
          }
          if (half == 0) *prev_lse = m_run * c + log2f(l_tot);
        }
      }
    };

    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int hg, b, srow;
      if (!p.pack) {
        const int qt = item % p.QT;
        const int r = item / p.QT;
        hg = r % p.HG;
        b = r / p.HG;
        srow = qt * 128 + row;
      } else {
        hg = item % p.HG;
        b = 2 * (item / p.HG) + (row >> 6);
        srow = row & 63;
      }
      const int nh = min(NH, p.H - hg * NH);
      const bool row_valid = (srow < p.S) && (b < p.B);
      for (int hh = 0; hh < nh; ++hh) {
        const int h = hg * NH + hh;
        for (int j = 0; j < p.nb; ++j, ++it) {
          const bool last_blk = (j == p.nb - 1);
          const int kvalid = p.pack ? (((row >> 6) == half) ? p.S : 0) : (p.S - j * BN - colbase);
          const uint32_t buf = it & 1;
          mbar_wait_wd(&s_full[buf], (it >> 1) & 1);
          tc_fence_after();
          uint32_t su[AT_MAXCH * 8];
          const uint32_t taddr = tmem_base + lane_off + buf * BN + colbase;
#pragma unroll
          for (int g = 0; g < AT_MAXCH; ++g)
            if (g < nch) tmem_ld_32x8(taddr + g * 8, su + g * 8);
          tmem_ld_wait();
          float pm = -INFINITY;
          if (kvalid >= BNh) {
#pragma unroll
            for (int g = 0; g < AT_MAXCH; ++g) {
              if (g < nch) {
#pragma unroll
                for (int e = 0; e < 8; ++e) pm = fmaxf(pm, __uint_as_float(su[g * 8 + e]));
              }
            }
          } else {
#pragma unroll
            for (int g = 0; g < AT_MAXCH; ++g) {
              if (g < nch) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  if (g * 8 + e >= kvalid) su[g * 8 + e] = 0xff800000u;   // -inf: p = 0
                  pm = fmaxf(pm, __uint_as_float(su[g * 8 + e]));
                }
              }
            }
          }
          const uint32_t par = it & 1;
          xmax[par * 256 + half * 128 + row] = pm;
          softmax_bar();
          pm = fmaxf(pm, xmax[par * 256 + (half ^ 1) * 128 + row]);
          if (have_prev) drain();
          if (j == 0) {
            m_run = -INFINITY;
            l_part = 0.f;
#pragma unroll
            for (int i = 0; i < OC; ++i) o_acc[i] = 0.f;
          }
          const float m_new = fmaxf(m_run, pm);
          const float alpha = ex2f((m_run - m_new) * c);
          m_run = m_new;
          const float mc = m_new * c;
          float lb = 0.f;
#pragma unroll
          for (int g = 0; g < AT_MAXCH; ++g) {
            if (g < nch) {
              float e[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                e[i] = ex2f(fmaf(__uint_as_float(su[g * 8 + i]), c, -mc));
                lb += e[i];
              }
              uint4 pk;
              pk.x = pack_bf16x2(e[0], e[1]);
              pk.y = pack_bf16x2(e[2], e[3]);
              pk.z = pack_bf16x2(e[4], e[5]);
              pk.w = pack_bf16x2(e[6], e[7]);
              const uint32_t gc = static_cast<uint32_t>((colbase >> 3) + g);
              sts_v4(p_row + (gc >> 3) * 16384 + (((gc & 7) ^ sw7) << 4), pk);
            }
          }
          l_part = fmaf(l_part, alpha, lb);
          alpha_pend = alpha;
          if (last_blk) xsum[par * 256 + half * 128 + row] = l_part;
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
          have_prev = true;
          prev_it = it;
          prev_hh = hh;
          prev_last = last_blk;
          prev_valid = row_valid;
          prev_out = p.out + (static_cast<size_t>(b) * p.S + srow) * p.Dm + h * DH + half * OC;
          prev_lse = p.lse + (static_cast<size_t>(b) * p.H + h) * p.S + srow;
        }
      }
    }
    if (have_prev) {
      softmax_bar();
      drain();
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int device_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

template <int DH, int VN>
int attn_fwd_tc_launch(const void* qkv, void* out, float* lse, int B, int S, int H, cudaStream_t stream) {
  const int Dm = H * DH;
  FwdParams p{};
  p.B = B; p.S = S; p.H = H; p.Dm = Dm;
  p.HG = (Dm + 63) / 64;
  p.pack = S <= 64 ? 1 : 0;
  if (p.pack) {
    p.BN = 128; p.nb = 1; p.QT = 1;
    p.items = ((B + 1) / 2) * p.HG;
  } else {
    const int bn_max = DH == 64 ? 192 : 208;
    p.nb = (S + bn_max - 1) / bn_max;
    p.BN = (((S + p.nb - 1) / p.nb) + 15) & ~15;
    p.QT = (S + 127) / 128;
    p.items = B * p.HG * p.QT;
  }
  p.c = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  const int pslabs = (p.BN + 63) / 64;
  const size_t smem = 1024 + 2 * AT_QBYTES + 4 * static_cast<size_t>(p.BN) * 128 + pslabs * 16384 + AT_XCH_BYTES + 128;
  auto kern = attn_fwd_tc_kernel<DH, VN>;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      csm_set_error("attention_fwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return CSM_ERR_CUDA;
    }
    configured = smem;
  }
  CUtensorMap tq, tkv;
  int rc = csm_tensor_map_2d(&tq, qkv, 3ull * Dm, static_cast<uint64_t>(B) * S, 3ull * Dm, 64, p.pack ? 64 : 128, 2, 128);
  if (rc) return rc;
  rc = csm_tensor_map_2d(&tkv, qkv, 3ull * Dm, static_cast<uint64_t>(B) * S, 3ull * Dm, 64, p.pack ? 64 : p.BN, 2, 128);
  if (rc) return rc;
  const int grid = p.items < device_sms() ? p.items : device_sms();
  cudaError_t le = csm_launch_pdl(kern, dim3(grid), dim3(AT_THREADS), smem, stream, tq, tkv, p);
  if (le != cudaSuccess) {
    csm_set_error("attention_fwd: launch failed: %s", cudaGetErrorString(le));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}


// ---------------------------------------------------------------------------------------------
// backward (S <= 256): one CTA walks work items = (image or image pair, 64-column head group); per head it visits the
// 128 x 128 sub-blocks (key block kb outer, query tile i inner) of the score matrix once:
//   MMA   S  = Q_i K_kb^T,  dP = dO_i V_kb^T                         -> TMEM (128 + 128 columns)
//   warps P  = exp2(S*c - L2),  dS = P * (dP - delta)                -> shared memory, bf16, [query][key] tiles
//   MMA   dV_kb += P^T dO_i,  dK_kb += dS^T Q_i,  dQ_i += dS K_kb    -> TMEM accumulators
// The [query][key] tiles of P / dS serve as the MN-major A operand of the dV / dK products and as the K-major A operand
// of dQ; the Q / K / V / dO tiles TMA brought in (128-byte swizzle) are K-major operands of the first two products and
// MN-major B operands of the last three -- nothing is transposed or copied.  dK / dV leave TMEM after the last query
// tile of their key block, dQ after the last key block (x d^-1/2, bf16), so every gradient is written exactly once and
// there are no atomics.  delta = rowsum(dO * O) is recomputed per (head, query tile) from global memory.
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  int B, S, H, Dm, HG;
  int pack;          // two images per 128-row tile (S <= 64)
  int NT;            // 128-row query tiles == 128-key blocks per image (1 or 2)
  int bn_last;       // keys in the last key block (multiple of 16)
  int items;
  float c, scale;
  const __nv_bfloat16* o_fwd;
  const __nv_bfloat16* d_out;
  const float* lse;
  __nv_bfloat16* dqkv;
};

template <int DH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const BwdParams p) {
  constexpr int NH = 64 / DH;
  constexpr int KS = DH / 16;
  constexpr int OC = DH / 2;           // gradient columns owned by one softmax thread
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 256 + 2 * DH, COL_DV = 256 + 3 * DH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NT = p.NT;
  const int tile_bytes = NT * 16384;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + tile_bytes;
  uint8_t* sV = sK + tile_bytes;
  uint8_t* sdO = sV + tile_bytes;
  uint8_t* sP = sdO + tile_bytes;      // [2 slabs of 64 keys][128 query rows][128 B]
  uint8_t* sdS = sP + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 32768);
  uint64_t* ld_full = bars;
  uint64_t* ld_empty = bars + 1;
  uint64_t* sdp_full = bars + 2;
  uint64_t* pds_full = bars + 3;
  uint64_t* pds_free = bars + 4;
  uint64_t* dkv_full = bars + 5;
  uint64_t* dkv_free = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(ld_full, 1);
    mbar_init(ld_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, AT_SM_WARPS);
    mbar_init(pds_free, 1);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_free, AT_SM_WARPS);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, AT_SM_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t ic = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
          const int hg = item % p.HG;
          const int bi = item / p.HG;
          mbar_wait_wd(ld_empty, (ic & 1) ^ 1);
          mbar_expect_tx(ld_full, 4u * tile_bytes);
          const int nbox = p.pack ? 2 : NT;
          for (int t = 0; t < nbox; ++t) {
            const int row = p.pack ? (2 * bi + t) * p.S : bi * p.S + t * 128;
            const int off = p.pack ? t * 8192 : t * 16384;
            tma_load_2d(sQ + off, &tmap_qkv, ld_full, hg * 64, row);
            tma_load_2d(sK + off, &tmap_qkv, ld_full, p.Dm + hg * 64, row);
            tma_load_2d(sV + off, &tmap_qkv, ld_full, 2 * p.Dm + hg * 64, row);
            tma_load_2d(sdO + off, &tmap_do, ld_full, hg * 64, row);
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer (one thread) ------------------------------
      if (lane == 0) {
        const uint32_t idesc_g = umma_idesc_bf16(128, DH, 1, 1);     // dV / dK: A and B MN-major
        const uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);     // dQ: A K-major, B MN-major
        const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), do_addr = smem_u32(sdO);
        const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
        uint32_t n = 0, g = 0, hc = 0, ic = 0;
        // pending sub-block whose gradient products are issued after the NEXT sub-block's S / dP products
        bool pend = false;
        uint32_t pd_n = 0, pd_g = 0, pd_hc = 0;
        int pd_hh = 0, pd_kb = 0, pd_i = 0, pd_bn = 0;
        bool pd_last_group = false, pd_last_head = false, pd_last_item = false;
        auto grads = [&]() {
          if (pd_i == 0) mbar_wait_wd(dkv_free, (pd_g & 1) ^ 1);               // dK / dV of the previous group read out
          if (pd_i == 0 && pd_kb == 0) mbar_wait_wd(dq_free, (pd_hc & 1) ^ 1);  // dQ of the previous head read out
          tc_fence_after();
          const uint32_t hoff = DH == 32 ? pd_hh * 64 : 0;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {           // dV_kb += P^T dO_i   (reduction over the 128 query rows)
            const uint64_t da = umma_smem_desc_sw128(p_addr + kk * 2048, 16384, 1024);
            const uint64_t db = umma_smem_desc_sw128(do_addr + pd_i * 16384 + kk * 2048 + hoff, 8192, 1024);
            umma_f16(tmem_base + COL_DV, da, db, idesc_g, (pd_i > 0 || kk > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {           // dK_kb += dS^T Q_i
            const uint64_t da = umma_smem_desc_sw128(ds_addr + kk * 2048, 16384, 1024);
            const uint64_t db = umma_smem_desc_sw128(q_addr + pd_i * 16384 + kk * 2048 + hoff, 8192, 1024);
            umma_f16(tmem_base + COL_DK, da, db, idesc_g, (pd_i > 0 || kk > 0) ? 1u : 0u);
          }
          const int ksteps = pd_bn >> 4;
          for (int ks = 0; ks < ksteps; ++ks) {      // dQ_i += dS K_kb     (reduction over the keys of the block)
            const uint64_t da = umma_smem_desc_sw128(ds_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(k_addr + pd_kb * 16384 + ks * 2048 + hoff, 8192, 1024);
            umma_f16(tmem_base + COL_DQ + pd_i * DH, da, db, idesc_q, (pd_kb > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(pds_free);
          if (pd_last_group) umma_commit(dkv_full);
          if (pd_last_head) umma_commit(dq_full);
          if (pd_last_item) umma_commit(ld_empty);
        };
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
          const int hg = item % p.HG;
          const int nh = min(NH, p.H - hg * NH);
          mbar_wait_wd(ld_full, ic & 1);
          tc_fence_after();
          for (int hh = 0; hh < nh; ++hh, ++hc) {
            for (int kb = 0; kb < NT; ++kb, ++g) {
              const int bn = (kb == NT - 1) ? p.bn_last : 128;
              const uint32_t idesc_s = umma_idesc_bf16(128, bn, 0, 0);
              for (int i = 0; i < NT; ++i, ++n) {
                if (pend) {
                  mbar_wait_wd(pds_full, pd_n & 1);       // the S / dP tiles have been read, P / dS are in smem
                  tc_fence_after();
                }
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                  const uint32_t ko = (hh * KS + kk) * 32;
                  umma_f16(tmem_base + COL_S, umma_smem_desc_sw128(q_addr + i * 16384 + ko, 16, 1024),
                           umma_smem_desc_sw128(k_addr + kb * 16384 + ko, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
                }
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                  const uint32_t ko = (hh * KS + kk) * 32;
                  umma_f16(tmem_base + COL_DP, umma_smem_desc_sw128(do_addr + i * 16384 + ko, 16, 1024),
                           umma_smem_desc_sw128(v_addr + kb * 16384 + ko, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
                }
                umma_commit(sdp_full);
                if (pend) grads();
                pend = true;
                pd_n = n; pd_g = g; pd_hc = hc; pd_hh = hh; pd_kb = kb; pd_i = i; pd_bn = bn;
                pd_last_group = (i == NT - 1);
                pd_last_head = pd_last_group && (kb == NT - 1);
                pd_last_item = pd_last_head && (hh == nh - 1);
              }
            }
          }
          // the operand tiles are single-buffered: finish this item's products before the next item's loads
          mbar_wait_wd(pds_full, pd_n & 1);
          tc_fence_after();
          grads();
          pend = false;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------ softmax / gradient warps ------------------------------
    const int half = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.c;
    const uint32_t p_row = smem_u32(sP) + row * 128;
    const uint32_t ds_row = smem_u32(sdS) + row * 128;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);
    const size_t ld3 = static_cast<size_t>(3) * p.Dm;
    uint32_t n = 0, g = 0, hc = 0;

    auto store16 = [&](const uint32_t* v, float mul, __nv_bfloat16* dst) {   // OC f32 values -> bf16
#pragma unroll
      for (int q8 = 0; q8 < OC / 8; ++q8) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(v[q8 * 8 + 0]) * mul, __uint_as_float(v[q8 * 8 + 1]) * mul);
        pk.y = pack_bf16x2(__uint_as_float(v[q8 * 8 + 2]) * mul, __uint_as_float(v[q8 * 8 + 3]) * mul);
        pk.z = pack_bf16x2(__uint_as_float(v[q8 * 8 + 4]) * mul, __uint_as_float(v[q8 * 8 + 5]) * mul);
        pk.w = pack_bf16x2(__uint_as_float(v[q8 * 8 + 6]) * mul, __uint_as_float(v[q8 * 8 + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + q8 * 8) = pk;
      }
    };
    auto tmem_ld_oc = [&](uint32_t addr, uint32_t* v) {
      tmem_ld_32x16(addr, v);
      if (OC == 32) tmem_ld_32x16(addr + 16, v + (OC == 32 ? 16 : 0));
    };

    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int hg = item % p.HG;
      const int bi = item / p.HG;
      const int nh = min(NH, p.H - hg * NH);
      // token (query or key) handled by this thread's TMEM lane in tile t
      const int b_img = p.pack ? 2 * bi + (row >> 6) : bi;
      const int tok0 = p.pack ? (row & 63) : row;
      for (int hh = 0; hh < nh; ++hh, ++hc) {
        const int h = hg * NH + hh;
        float Lr[2] = {INFINITY, INFINITY}, dl[2] = {0.f, 0.f};
        for (int kb = 0; kb < NT; ++kb, ++g) {
          const int bn = (kb == NT - 1) ? p.bn_last : 128;
          const int bnh = bn >> 1;
          const int nch = bnh >> 3;
          const int colbase = half * bnh;
          const int kvalid = p.pack ? (((row >> 6) == half) ? p.S : 0) : (p.S - kb * 128 - colbase);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (i < NT) {
              if (kb == 0) {
                // softmax statistic and delta = rowsum(dO * O) of this thread's query row
                const int tok = tok0 + i * 128;
                if (tok < p.S && b_img < p.B) {
                  const size_t grow = static_cast<size_t>(b_img) * p.S + tok;
                  Lr[i] = p.lse[(static_cast<size_t>(b_img) * p.H + h) * p.S + tok];
                  const uint4* dop = reinterpret_cast<const uint4*>(p.d_out + grow * p.Dm + h * DH);
                  const uint4* op = reinterpret_cast<const uint4*>(p.o_fwd + grow * p.Dm + h * DH);
                  float acc = 0.f;
#pragma unroll
                  for (int c8 = 0; c8 < DH / 8; ++c8) {
                    const uint4 a = __ldg(dop + c8), bb = __ldg(op + c8);
                    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(bw[e]);
                      acc = fmaf(x.x, y.x, acc);
                      acc = fmaf(x.y, y.y, acc);
                    }
                  }
                  dl[i] = acc;
                } else {
                  Lr[i] = INFINITY;      // padded query row: p = exp2(-inf) = 0
                  dl[i] = 0.f;
                }
              }
              mbar_wait_wd(sdp_full, n & 1);
              tc_fence_after();
              uint32_t su[64], du[64];
              const uint32_t ts = tmem_base + lane_off + COL_S + colbase;
              const uint32_t td = tmem_base + lane_off + COL_DP + colbase;
#pragma unroll
              for (int q8 = 0; q8 < 8; ++q8) {
                if (q8 < nch) {
                  tmem_ld_32x8(ts + q8 * 8, su + q8 * 8);
                  tmem_ld_32x8(td + q8 * 8, du + q8 * 8);
                }
              }
              tmem_ld_wait();
              if (kvalid < bnh) {
#pragma unroll
                for (int e = 0; e < 64; ++e)
                  if (e >= kvalid) su[e] = 0xff800000u;            // masked key: p = 0, dS = 0
              }
              if (n > 0) mbar_wait_wd(pds_free, (n - 1) & 1);       // the previous P / dS tiles have been consumed
              const float Li = Lr[i], di = dl[i];
#pragma unroll
              for (int q8 = 0; q8 < 8; ++q8) {
                if (q8 < nch) {
                  float pe[8], de[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    pe[e] = ex2f(fmaf(__uint_as_float(su[q8 * 8 + e]), c, -Li));
                    de[e] = pe[e] * (__uint_as_float(du[q8 * 8 + e]) - di);
                  }
                  uint4 pk, dk;
                  pk.x = pack_bf16x2(pe[0], pe[1]); pk.y = pack_bf16x2(pe[2], pe[3]);
                  pk.z = pack_bf16x2(pe[4], pe[5]); pk.w = pack_bf16x2(pe[6], pe[7]);
                  dk.x = pack_bf16x2(de[0], de[1]); dk.y = pack_bf16x2(de[2], de[3]);
                  dk.z = pack_bf16x2(de[4], de[5]); dk.w = pack_bf16x2(de[6], de[7]);
                  const uint32_t gc = static_cast<uint32_t>((colbase >> 3) + q8);
                  const uint32_t off = (gc >> 3) * 16384 + (((gc & 7) ^ sw7) << 4);
                  sts_v4(p_row + off, pk);
                  sts_v4(ds_row + off, dk);
                }
              }
              fence_proxy_async();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(pds_full);
              ++n;
            }
          }
          // ---- dK / dV of this key block are complete after its last query tile ----
          mbar_wait_wd(dkv_full, g & 1);
          tc_fence_after();
          {
            uint32_t vk[OC], vv[OC];
            tmem_ld_oc(tmem_base + lane_off + COL_DK + half * OC, vk);
            tmem_ld_oc(tmem_base + lane_off + COL_DV + half * OC, vv);
            tmem_ld_wait();
            const int tok = tok0 + kb * 128;
            if (tok < p.S && b_img < p.B) {
              __nv_bfloat16* base = p.dqkv + (static_cast<size_t>(b_img) * p.S + tok) * ld3 + h * DH + half * OC;
              store16(vk, p.scale, base + p.Dm);
              store16(vv, 1.0f, base + 2 * p.Dm);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dkv_free);
        }
        // ---- dQ of this head is complete after the last key block ----
        mbar_wait_wd(dq_full, hc & 1);
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i < NT) {
            uint32_t vq[OC];
            tmem_ld_oc(tmem_base + lane_off + COL_DQ + i * DH + half * OC, vq);
            tmem_ld_wait();
            const int tok = tok0 + i * 128;
            if (tok < p.S && b_img < p.B)
              store16(vq, p.scale, p.dqkv + (static_cast<size_t>(b_img) * p.S + tok) * ld3 + h * DH + half * OC);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH>
int attn_bwd_tc_launch(const void* qkv, const void* o, const void* d_out, const float* lse, void* dqkv, int B, int S,
                       int H, cudaStream_t stream) {
  const int Dm = H * DH;
  BwdParams p{};
  p.B = B; p.S = S; p.H = H; p.Dm = Dm;
  p.HG = (Dm + 63) / 64;
  p.pack = S <= 64 ? 1 : 0;
  if (p.pack) {
    p.NT = 1; p.bn_last = 128;
    p.items = ((B + 1) / 2) * p.HG;
  } else {
    p.NT = (S + 127) / 128;
    p.bn_last = ((S - (p.NT - 1) * 128) + 15) & ~15;
    p.items = B * p.HG;
  }
  p.scale = 1.0f / sqrtf(static_cast<float>(DH));
  p.c = 1.4426950408889634f * p.scale;
  p.o_fwd = reinterpret_cast<const __nv_bfloat16*>(o);
  p.d_out = reinterpret_cast<const __nv_bfloat16*>(d_out);
  p.lse = lse;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  const size_t smem = 1024 + 4 * static_cast<size_t>(p.NT) * 16384 + 65536 + 128;
  auto kern = attn_bwd_tc_kernel<DH>;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      csm_set_error("attention_bwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return CSM_ERR_CUDA;
    }
    configured = smem;
  }
  CUtensorMap tq, tdo;
  const uint32_t box_rows = p.pack ? 64 : 128;
  int rc = csm_tensor_map_2d(&tq, qkv, 3ull * Dm, static_cast<uint64_t>(B) * S, 3ull * Dm, 64, box_rows, 2, 128);
  if (rc) return rc;
  rc = csm_tensor_map_2d(&tdo, d_out, Dm, static_cast<uint64_t>(B) * S, Dm, 64, box_rows, 2, 128);
  if (rc) return rc;
  const int grid = p.items < device_sms() ? p.items : device_sms();
  cudaError_t le = csm_launch_pdl(kern, dim3(grid), dim3(AT_THREADS), smem, stream, tq, tdo, p);
  if (le != cudaSuccess) {
    csm_set_error("attention_bwd: launch failed: %s", cudaGetErrorString(le));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

}  // namespace

// variant: 0 = narrow P.V (N = d), 1 = d = 32 heads computed as N = 64 pairs (diagnostic fallback)
extern "C" int csm_attention_fwd_tc(const void* qkv_bf16, void* out_bf16, float* lse, int B, int S, int H, int head_dim,
                                    int variant, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && S > 0 && H > 0, "csm_attention_fwd: bad sizes B=%d S=%d H=%d", B, S, H);
  CSM_CHECK_ARG((H * head_dim) % 8 == 0, "csm_attention_fwd: H * head_dim must be a multiple of 8");
  if (head_dim == 32) {
    if (variant == 1) return attn_fwd_tc_launch<32, 64>(qkv_bf16, out_bf16, lse, B, S, H, stream);
    return attn_fwd_tc_launch<32, 32>(qkv_bf16, out_bf16, lse, B, S, H, stream);
  }
  if (head_dim == 64) return attn_fwd_tc_launch<64, 64>(qkv_bf16, out_bf16, lse, B, S, H, stream);
  csm_set_error("csm_attention_fwd: head_dim must be 32 or 64 (got %d)", head_dim);
  return CSM_ERR_ARG;
}

extern "C" int csm_attention_bwd_tc(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse,
                                    void* dqkv_bf16, int B, int S, int H, int head_dim, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && S > 0 && H > 0, "csm_attention_bwd: bad sizes B=%d S=%d H=%d", B, S, H);
  CSM_CHECK_ARG(S <= 256, "csm_attention_bwd_tc: S=%d > 256 is served by the two-pass kernels", S);
  CSM_CHECK_ARG((H * head_dim) % 8 == 0, "csm_attention_bwd: H * head_dim must be a multiple of 8");
  if (head_dim == 32) return attn_bwd_tc_launch<32>(qkv_bf16, out_bf16, d_out_bf16, lse, dqkv_bf16, B, S, H, stream);
  if (head_dim == 64) return attn_bwd_tc_launch<64>(qkv_bf16, out_bf16, d_out_bf16, lse, dqkv_bf16, B, S, H, stream);
  csm_set_error("csm_attention_bwd: head_dim must be 32 or 64 (got %d)", head_dim);
  return CSM_ERR_ARG;
}
