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

#include <algorithm>

namespace {
using namespace csm;

constexpr int AT_SM_WARPS = 8;
constexpr int AT_THREADS = 32 * (4 + AT_SM_WARPS);   // 384: warp group 0 = {TMA, MMA, 2 idle}, groups 1-2 = softmax
constexpr int AT_QBYTES = 128 * 128;                 // one Q tile: 128 rows x 64 bf16
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

// bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the device
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) { mbar_wait_bounded(bar, parity); }

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

template <int N>
struct IC {
  static constexpr int value = N;
};

// Development builds (CSM_NVCC_EXTRA=-DCSM_ATTN_TIMING, tools/attn_phase.py): cycles spent per phase of the backward
// kernel's sub-block loop, summed over the sub-blocks of one softmax warp (warp 4, lane 0) and of the MMA thread of
// every CTA.  Slot 15 counts sub-blocks.
#ifdef CSM_ATTN_TIMING
__device__ unsigned long long g_attn_phase[16];
#define ATT_T(var) const long long var = clock64()
#define ATT_ACC(slot, t1, t0) acc_t[slot] += (t1) - (t0)
#else
#define ATT_T(var)
#define ATT_ACC(slot, t1, t0)
#endif

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem, bf16 pairs packed along K] * B[smem descriptor as {lo, hi} words]
__device__ __forceinline__ void umma_f16_ts_lh(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// DH: head dim (32: two heads per 64-column group, 64: one).  NG0 / NG1: 16-key groups of the two softmax threads of a
// query row; the key block is BN = 16 * (NG0 + NG1) keys (208 = 96 + 112, so that the 197 keys of the decoder split
// 96 / 101, or 128 = 64 + 64).  PTMEM: the probabilities go back to TMEM (over the score columns they came from) and
// feed the P.V product as its TMEM A operand; otherwise they go through a 128-byte-swizzled shared-memory tile.
template <int DH, int NG0, int NG1, bool PTMEM>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const FwdParams p) {
  constexpr int NH = 64 / DH;          // heads per 64-column group
  constexpr int KS = DH / 16;          // 16-wide k-steps of one head in the Q / K rows
  constexpr int OC = DH / 2;           // output columns owned by one softmax thread
  constexpr int NG = NG0 > NG1 ? NG0 : NG1;
  constexpr int BN = 16 * (NG0 + NG1); // keys per block
  constexpr int BNH0 = 16 * NG0;       // keys of the first half
  constexpr int KV_BYTES = BN * 128;   // one K (or V) block
  constexpr int PSLABS = PTMEM ? 0 : (BN + 63) / 64;
  constexpr uint32_t COL_O = 2 * BN;   // O of key half h at COL_O + h * DH
  static_assert(2 * BN + 2 * DH <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][16 KB]
  uint8_t* sKV = smem + 2 * AT_QBYTES;                 // [2][K | V]
  uint8_t* sP = sKV + 4 * KV_BYTES;                    // [PSLABS][128 rows][128 B]
  float* xch = reinterpret_cast<float*>(sP + PSLABS * 16384);   // {max, sum} x parity x half x row
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

  // Iteration order inside a work item: heads of the group outermost, key blocks innermost.  With a single key block
  // both heads share one K/V load; with several blocks the blocks are re-fetched per head (L2 hits).
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
            mbar_expect_tx(&kv_full[ks], 2u * KV_BYTES);
            uint8_t* k = sKV + ks * 2 * KV_BYTES;
            uint8_t* v = k + KV_BYTES;
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
      // ------------------------------ MMA issuer ------------------------------
      // The whole warp walks the schedule (warp-uniform control flow, so descriptors and loop state live in uniform
      // registers); one elected lane issues.  Every descriptor is base_lo + a constant (see umma_desc_lo).
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
      constexpr uint32_t HI = umma_desc_hi_sw128(1024);
      constexpr uint32_t LBO_K = (16u >> 4) << 16, LBO_MN = (8192u >> 4) << 16;
      constexpr uint32_t OFF_KV = (2u * AT_QBYTES) >> 4, KV16 = KV_BYTES >> 4, OFF_P = OFF_KV + 4 * KV16;
      const uint32_t base_lo = __shfl_sync(0xffffffffu, (smem_u32(smem) & 0x3FFFFu) >> 4, 0);
      const uint32_t p_lo = base_lo + OFF_P + LBO_K;
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
        const uint32_t v_lo = base_lo + OFF_KV + pv_ks * (2 * KV16) + KV16 + (DH == 32 ? pv_hh * 4 : 0) + LBO_MN;
        const uint32_t tmem_p = tmem_base + (pv_it & 1) * BN;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < NG0 + NG1; ++k) {        // each key half accumulates into its own O
            const int hf = k >= NG0 ? 1 : 0, kh = k - hf * NG0;
            const uint32_t tmem_o = tmem_base + COL_O + hf * DH;
            if (PTMEM) {
              umma_f16_ts_lh(tmem_o, tmem_p + hf * BNH0 + kh * 8, v_lo + k * 128, HI, idesc_o, kh > 0 ? 1u : 0u);
            } else {
              umma_f16_lh(tmem_o, p_lo + (k >> 2) * 1024 + (k & 3) * 2, HI, v_lo + k * 128, HI, idesc_o,
                          kh > 0 ? 1u : 0u);
            }
          }
          umma_commit(o_full);
          if (pv_last_kv) umma_commit(&kv_empty[pv_ks]);
          if (pv_last_item) umma_commit(&q_empty[pv_qs]);
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
        const int hg = p.pack ? item % p.HG : (item / p.QT) % p.HG;
        const int nh = min(NH, p.H - hg * NH);
        const uint32_t qs = ic & 1;
        mbar_wait_wd(&q_full[qs], (ic >> 1) & 1);
        const uint32_t q_lo = base_lo + qs * (AT_QBYTES >> 4) + LBO_K;
        uint32_t ks = 0;
        for (int hh = 0; hh < nh; ++hh) {
          for (int j = 0; j < p.nb; ++j, ++it) {
            if (!(shared_kv && hh > 0)) {
              ks = kvc & 1;
              mbar_wait_wd(&kv_full[ks], (kvc >> 1) & 1);
              ++kvc;
            }
            tc_fence_after();
            const uint32_t k_lo = base_lo + OFF_KV + ks * (2 * KV16) + LBO_K;
            const uint32_t tmem_s = tmem_base + (it & 1) * BN;
            const uint32_t ko = hh * KS * 2;            // 32 bytes per k-step, in 16-byte units
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < KS; ++kk)
                umma_f16_lh(tmem_s, q_lo + ko + kk * 2, HI, k_lo + ko + kk * 2, HI, idesc_s, kk > 0 ? 1u : 0u);
              umma_commit(&s_full[it & 1]);
            }
            __syncwarp();
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
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------ softmax warps ------------------------------
    const int half = (warp - 4) >> 2;         // which half of the key block
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int colbase = half * BNH0;
    const int bnh = half ? 16 * NG1 : 16 * NG0;   // keys of this thread's half
    const float c = p.c;
    float* xm = xch;                          // [parity][half][row]
    float* xl = xch + 512;
    const uint32_t p_row = smem_u32(sP) + row * 128;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);

    // online-softmax state of the head in flight; owned by drain()
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[OC];
#pragma unroll
    for (int i = 0; i < OC; ++i) o_acc[i] = 0.f;
    uint32_t it = 0;
    uint32_t have_prev;      // opaque to the compiler: keeps it from peeling the first iteration of the item loops
    asm volatile("mov.u32 %0, 0;" : "=r"(have_prev));
    bool prev_first = false, prev_last = false, prev_valid = false;
    uint32_t prev_it = 0;
    __nv_bfloat16* prev_out = nullptr;
    float* prev_lse = nullptr;

    // Fold the P.V result of iteration prev_it into the running output.  The two key halves were exponentiated
    // against their OWN row maximum (no exchange before the exp): their partial outputs O_h and partial sums l_h are
    // combined here with the factors exp2((m_h - m) * c).
#ifdef CSM_ATTN_TIMING
    long long acc_t[16] = {0};
#endif
    auto drain = [&]() {
      ATT_T(d0);
      mbar_wait_wd(o_full, prev_it & 1);
      tc_fence_after();
      ATT_T(d1);
      ATT_ACC(4, d1, d0);      // wait: P.V product of the previous iteration
      uint32_t o0[OC], o1[OC];
      const uint32_t oaddr = tmem_base + lane_off + COL_O + half * OC;
      tmem_ld_32x16(oaddr, o0);
      tmem_ld_32x16(oaddr + DH, o1);
      if (OC == 32) {
        tmem_ld_32x16(oaddr + 16, o0 + (OC == 32 ? 16 : 0));
        tmem_ld_32x16(oaddr + DH + 16, o1 + (OC == 32 ? 16 : 0));
      }
      const uint32_t xo = (prev_it & 1) * 256 + row;
      const float m0 = xm[xo], m1 = xm[xo + 128], l0 = xl[xo], l1 = xl[xo + 128];
      if (prev_first) {
        m_run = -INFINITY;
        l_run = 0.f;
#pragma unroll
        for (int i = 0; i < OC; ++i) o_acc[i] = 0.f;
      }
      const float m_new = fmaxf(m_run, fmaxf(m0, m1));
      const float a = ex2f((m_run - m_new) * c), b0 = ex2f((m0 - m_new) * c), b1 = ex2f((m1 - m_new) * c);
      l_run = fmaf(l_run, a, fmaf(l0, b0, l1 * b1));
      m_run = m_new;
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < OC; ++i)
        o_acc[i] = fmaf(o_acc[i], a, fmaf(__uint_as_float(o0[i]), b0, __uint_as_float(o1[i]) * b1));
      if (prev_last && prev_valid) {
        const float inv = 1.0f / l_run;
#pragma unroll
        for (int g = 0; g < OC / 8; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(o_acc[g * 8 + 0] * inv, o_acc[g * 8 + 1] * inv);
          pk.y = pack_bf16x2(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
          pk.z = pack_bf16x2(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
          pk.w = pack_bf16x2(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
          *reinterpret_cast<uint4*>(prev_out + g * 8) = pk;
        }
        if (half == 0) *prev_lse = m_run * c + log2f(l_run);
      }
      ATT_T(d2);
      ATT_ACC(5, d2, d1);      // read-back of O, rescale, store
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
#pragma unroll 1
      for (int hh = 0; hh < nh; ++hh) {
        const int h = hg * NH + hh;
#pragma unroll 1
        for (int j = 0; j < p.nb; ++j, ++it) {
          // valid keys of this thread's half (warp-uniform): nfull whole 16-key groups, then one partial group of
          // rem keys (handled by its own 16 registers so that the unrolled loops carry no per-element masking), then
          // groups that only receive zero probabilities
          int kvalid = p.pack ? (((row >> 6) == half) ? p.S : 0) : (p.S - j * BN - colbase);
          kvalid = max(0, min(kvalid, bnh));
          const int nfull = kvalid >> 4;
          const int rem = kvalid & 15;
          const uint32_t buf = it & 1;
          ATT_T(t0);
          mbar_wait_wd(&s_full[buf], (it >> 1) & 1);
          tc_fence_after();
          ATT_T(t1);
          uint32_t su[NG * 16], tu[16];
          const uint32_t taddr = tmem_base + lane_off + buf * BN + colbase;
#pragma unroll
          for (int g = 0; g < NG; ++g)
            if (g < nfull) tmem_ld_32x16(taddr + g * 16, su + g * 16);
          if (rem) tmem_ld_32x16(taddr + nfull * 16, tu);
          tmem_ld_wait();
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (rem) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if (e >= rem) tu[e] = 0xff800000u;                   // -inf: p = 0
              mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(tu[e]));
            }
          }
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (g < nfull) {
#pragma unroll
              for (int e = 0; e < 16; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(su[g * 16 + e]));
            }
          }
          const float m_h = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          const float mc = m_h * c;
          ATT_T(t2);
          if (!PTMEM && have_prev) drain();              // also: the P tile in shared memory is free again
          ATT_T(t3);
          float ls[4] = {0.f, 0.f, 0.f, 0.f};
          auto store_p = [&](int g, const uint32_t* pk) {
            if (PTMEM) {
              tmem_st_32x8(taddr + g * 8, pk);
            } else {
              const uint32_t gc = static_cast<uint32_t>((colbase >> 3) + 2 * g);
              sts_v4(p_row + (gc >> 3) * 16384 + (((gc & 7) ^ sw7) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
              sts_v4(p_row + ((gc + 1) >> 3) * 16384 + ((((gc + 1) & 7) ^ sw7) << 4),
                     make_uint4(pk[4], pk[5], pk[6], pk[7]));
            }
          };
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            uint32_t pk[8];
            if (g < nfull) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float e0 = ex2f(fmaf(__uint_as_float(su[g * 16 + 2 * e]), c, -mc));
                const float e1 = ex2f(fmaf(__uint_as_float(su[g * 16 + 2 * e + 1]), c, -mc));
                ls[e & 3] += e0 + e1;
                pk[e] = pack_bf16x2(e0, e1);
              }
              store_p(g, pk);
            } else if (g * 16 < bnh && (g > nfull || rem == 0)) {
#pragma unroll
              for (int e = 0; e < 8; ++e) pk[e] = 0u;
              store_p(g, pk);
            }
          }
          if (rem) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float e0 = ex2f(fmaf(__uint_as_float(tu[2 * e]), c, -mc));
              const float e1 = ex2f(fmaf(__uint_as_float(tu[2 * e + 1]), c, -mc));
              ls[e & 3] += e0 + e1;
              pk[e] = pack_bf16x2(e0, e1);
            }
            store_p(nfull, pk);
          }
          const uint32_t xo = buf * 256 + half * 128 + row;
          xm[xo] = m_h;
          xl[xo] = (ls[0] + ls[1]) + (ls[2] + ls[3]);
          ATT_T(t4);
          if (PTMEM) {
            if (have_prev) drain();
            ATT_T(t5);
            tmem_st_wait();
            ATT_T(t6);
            ATT_ACC(6, t6, t5);
          } else {
            fence_proxy_async();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
          ATT_T(t7);
          ATT_ACC(0, t1, t0);      // wait: S = Q K^T
          ATT_ACC(1, t2, t1);      // tcgen05.ld of the scores + row maximum
          ATT_ACC(2, t4, t3);      // exp2, row sum, pack, tcgen05.st / st.shared of P
          ATT_ACC(3, t7, t0);      // whole iteration
          have_prev = 1;
          prev_it = it;
          prev_first = (j == 0);
          prev_last = (j == p.nb - 1);
          prev_valid = row_valid;
          prev_out = p.out + (static_cast<size_t>(b) * p.S + srow) * p.Dm + h * DH + half * OC;
          prev_lse = p.lse + (static_cast<size_t>(b) * p.H + h) * p.S + srow;
        }
      }
    }
    if (have_prev) drain();
#ifdef CSM_ATTN_TIMING
    if (warp == 4 && lane == 0) {
      for (int i = 0; i < 8; ++i) atomicAdd(&g_attn_phase[i], static_cast<unsigned long long>(acc_t[i]));
      atomicAdd(&g_attn_phase[15], static_cast<unsigned long long>(it));
    }
#endif
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

template <int DH, int NG0, int NG1, bool PTMEM>
int attn_fwd_tc_launch(const void* qkv, void* out, float* lse, int B, int S, int H, cudaStream_t stream) {
  constexpr int BN = 16 * (NG0 + NG1);
  const int Dm = H * DH;
  FwdParams p{};
  p.B = B; p.S = S; p.H = H; p.Dm = Dm;
  p.HG = (Dm + 63) / 64;
  p.pack = S <= 64 ? 1 : 0;
  p.BN = BN;
  if (p.pack) {
    p.nb = 1; p.QT = 1;
    p.items = ((B + 1) / 2) * p.HG;
  } else {
    p.nb = (S + BN - 1) / BN;
    p.QT = (S + 127) / 128;
    p.items = B * p.HG * p.QT;
  }
  p.c = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  constexpr int PSLABS = PTMEM ? 0 : (BN + 63) / 64;
  const size_t smem = 1024 + 2 * AT_QBYTES + 4 * static_cast<size_t>(BN) * 128 + PSLABS * 16384 + AT_XCH_BYTES + 128;
  auto kern = attn_fwd_tc_kernel<DH, NG0, NG1, PTMEM>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      csm_set_error("attention_fwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return CSM_ERR_CUDA;
    }
    configured = true;
  }
  CUtensorMap tq, tkv;
  int rc = csm_tensor_map_2d(&tq, qkv, 3ull * Dm, static_cast<uint64_t>(B) * S, 3ull * Dm, 64, p.pack ? 64 : 128, 2, 128);
  if (rc) return rc;
  rc = csm_tensor_map_2d(&tkv, qkv, 3ull * Dm, static_cast<uint64_t>(B) * S, 3ull * Dm, 64, p.pack ? 64 : BN, 2, 128);
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
// backward: a persistent CTA per SM walks work items = (image or image pair, 64-column head group, 128-key block).
// The K / V tiles of the block stay in shared memory while the 128-row Q / dO tiles stream past (both double-buffered
// by the TMA producer, across work items too); per (query tile i, head) the 128 x 128 sub-block of the score matrix is
// visited once:
//   MMA   S  = Q_i K^T,  dP = dO_i V^T                               -> TMEM (128 + 128 columns)
//   warps P  = exp2(S*c - L2),  dS = P * (dP - delta)                -> shared memory, bf16, [query][key] tiles
//   MMA   dV += P^T dO_i,  dK += dS^T Q_i,  dQ_i(partial) = dS K     -> TMEM
// The [query][key] tiles of P / dS serve as the MN-major A operand of the dV / dK products and as the K-major A operand
// of dQ; the Q / K / V / dO tiles TMA brought in (128-byte swizzle) are K-major operands of the first two products and
// MN-major B operands of the last three -- nothing is transposed or copied.  dK / dV accumulate in TMEM over the query
// tiles and are written once per work item; the dQ partial of a sub-block is read back one sub-block later (x d^-1/2)
// and stored directly when the sequence is a single key block, else reduced into the (zeroed) dQ columns with
// red.global.add.bf16x2.  The softmax warps release the S / dP tiles as soon as they hold them in registers, so the
// next sub-block's first two products run under the exp / dS arithmetic of the current one.
// The per-row statistics of a sub-block (L2 = the forward's log-sum-exp, delta = rowsum(dO * O) from attn_delta_kernel)
// are staged by the two otherwise idle warps of warp group 0 into a two-deep shared-memory ring, one sub-block ahead:
// the softmax threads read them with one ld.shared each instead of carrying prefetched global loads across iterations.
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  int B, S, H, Dm, HG;
  int pack;          // two images per 128-row tile (S <= 64)
  int NT;            // 128-row query tiles == 128-key blocks per image
  int acc;           // NT == 2: a CTA takes both key blocks of a (image, head group) back to back and dQ accumulates in
                     // TMEM across them (no partial-dQ reduction in global memory)
  int bn_last;       // keys in the last key block (multiple of 64)
  int items;
  float c, scale;
  const float* lse;
  const float* delta;      // rowsum(dO * O) per (image, head, token), attn_delta_kernel
  __nv_bfloat16* dqkv;
};

// delta[(b*H + h)*S + s] = sum_j dO[b*S + s, h*DH + j] * O[b*S + s, h*DH + j].  A thread takes one 16-byte chunk of a
// (token, head) row piece, so a warp's loads are 512 contiguous bytes; the DH / 8 lanes of a piece combine with shuffles
// and the first of them writes.
template <int DH>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_out,
                                  float* __restrict__ delta, int B, int S, int H) {
  constexpr int CPR = DH / 8;                                   // 16-byte chunks per (token, head)
  pdl_wait();
  pdl_trigger();
  const long long total = static_cast<long long>(B) * S * H * CPR;      // a multiple of 4, the grid stride of 32
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t - (threadIdx.x & 31) < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    if (t < total) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(d_out) + t), bb = __ldg(reinterpret_cast<const uint4*>(o) + t);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(bw[e]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
#pragma unroll
    for (int o_ = 1; o_ < CPR; o_ <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o_);
    if (t < total && (t % CPR) == 0) {
      const long long idx = t / CPR;                           // (b * S + s) * H + h
      const int h = static_cast<int>(idx % H);
      const long long r = idx / H;
      const long long b = r / S, s_ = r % S;
      delta[(b * H + h) * S + s_] = acc;
    }
  }
}

__device__ __forceinline__ void red_add_bf16x2_v4(void* gptr, uint4 v) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// k-th work item of this CTA.  An item is (image or image pair, head group, key block), id = (.. * HG + hg) * NT + kb.
// Plain mode: items are dealt round-robin.  acc mode: units (image, head group) are dealt round-robin and a CTA walks
// the NT key blocks of its unit consecutively.
__device__ __forceinline__ int bwd_item(const BwdParams& p, int k) {
  if (!p.acc) return blockIdx.x + k * gridDim.x;
  return (blockIdx.x + (k / p.NT) * gridDim.x) * p.NT + k % p.NT;
}

constexpr int ATB_SM_WARPS = 16;
constexpr int ATB_THREADS = 32 * (4 + ATB_SM_WARPS);   // 640

template <int DH>
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const BwdParams p) {
  constexpr int NH = 64 / DH;
  constexpr int KS = DH / 16;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 256 + NH * DH, COL_DQ = 256 + 2 * NH * DH;
  static_assert(COL_DQ + 2 * NH * DH <= 512, "TMEM budget (acc mode: NT = 2 query tiles x NH heads x DH columns of dQ)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;                 // [2 stages][K 16 KB | V 16 KB]
  uint8_t* sQO = smem + 65536;         // [2 stages][Q 16 KB | dO 16 KB]
  uint8_t* sP = smem + 131072;         // [2 slabs of 64 keys][128 query rows][128 B]
  uint8_t* sdS = sP + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 32768);
  uint64_t* kv_full = bars;            // [2]
  uint64_t* kv_empty = bars + 2;       // [2]
  uint64_t* q_full = bars + 4;         // [2]
  uint64_t* q_empty = bars + 6;        // [2]
  uint64_t* sdp_full = bars + 8;       // S / dP products complete
  uint64_t* sdp_free = bars + 9;       // S / dP tiles are in registers
  uint64_t* pds_full = bars + 10;      // P / dS tiles written
  uint64_t* grads_done = bars + 11;    // dV / dK / dQ products of a sub-block complete
  uint64_t* dq_free = bars + 12;       // [2] dQ partial buffer read out
  uint64_t* dkv_free = bars + 14;      // dK / dV of an item read out
  uint64_t* st_full = bars + 15;       // [2] row statistics of a sub-block staged (warps 2 and 3)
  uint64_t* st_empty = bars + 17;      // [2] ... and read by the 16 softmax warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  float* sL = reinterpret_cast<float*>(bars + 20);   // [2][128] L2 = log-sum-exp (base 2) of the row; +inf = padded row
  float* sD = sL + 256;                              // [2][128] delta = rowsum(dO * O)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&dq_free[i], ATB_SM_WARPS);
      mbar_init(&st_full[i], 2);
      mbar_init(&st_empty[i], ATB_SM_WARPS);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, ATB_SM_WARPS);
    mbar_init(pds_full, ATB_SM_WARPS);
    mbar_init(grads_done, 1);
    mbar_init(dkv_free, ATB_SM_WARPS);
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
  const int NT = p.NT;

  if (warp < 4) {
    // The register pool setmaxnreg.inc draws from holds only what this CTA's own warps released: the CTA was
    // launched with 640 x 96 registers, so 4 x 32 x 64 + 16 x 32 x 104 = 61440 is the most that can be redistributed.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp >= 2) {
      // ------------------------------ row statistics (warps 2 and 3) ------------------------------
      // Same sub-block sequence as the other roles.  Warp 2 stages L2[row] (the forward's log-sum-exp, +inf for padded
      // rows), warp 3 delta[row] = rowsum(dO * O) (attn_delta_kernel), four rows per lane, one sub-block ahead of the
      // softmax warps.  (Forming delta here from the dO tile in shared memory and the O rows was built and measured:
      // the two warps then issue as much as a softmax warp on two of the four schedulers, 181 -> 241 us.)
      const float* src = warp == 2 ? p.lse : p.delta;
      float* dst = warp == 2 ? sL : sD;
      const float pad = warp == 2 ? INFINITY : 0.f;      // padded query row: p = exp2(s - inf) = 0
      uint32_t n = 0;
      for (int k = 0;; ++k) {
        const int item = bwd_item(p, k);
        if (item >= p.items) break;
        const int r_ = item / NT;
        const int hg = r_ % p.HG;
        const int bi = r_ / p.HG;
        const int nh = min(NH, p.H - hg * NH);
        for (int i = 0; i < NT; ++i) {
          for (int hh = 0; hh < nh; ++hh, ++n) {
            const uint32_t sb = n & 1;
            const int h = hg * NH + hh;
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = lane + 32 * k;
              const int b_img = p.pack ? 2 * bi + (r >> 6) : bi;
              const int tok = p.pack ? (r & 63) : i * 128 + r;
              v[k] = pad;
              if (tok < p.S && b_img < p.B) v[k] = __ldg(src + (static_cast<size_t>(b_img) * p.H + h) * p.S + tok);
            }
            mbar_wait_wd(&st_empty[sb], ((n >> 1) & 1) ^ 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[sb * 128 + lane + 32 * k] = v[k];
            __syncwarp();
            if (lane == 0) mbar_arrive(&st_full[sb]);
          }
        }
      }
    } else if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t ic = 0, qc = 0;
        for (;; ++ic) {
          const int item = bwd_item(p, static_cast<int>(ic));
          if (item >= p.items) break;
          const int kb = item % NT;
          const int r = item / NT;
          const int hg = r % p.HG;
          const int bi = r / p.HG;
          const int nbox = p.pack ? 2 : 1;
          {
            const uint32_t ks = ic & 1;
            mbar_wait_wd(&kv_empty[ks], ((ic >> 1) & 1) ^ 1);
            mbar_expect_tx(&kv_full[ks], 32768);
            uint8_t* k = sKV + ks * 32768;
            for (int t = 0; t < nbox; ++t) {
              const int row = p.pack ? (2 * bi + t) * p.S : bi * p.S + kb * 128;
              tma_load_2d(k + t * 8192, &tmap_qkv, &kv_full[ks], p.Dm + hg * 64, row);
              tma_load_2d(k + 16384 + t * 8192, &tmap_qkv, &kv_full[ks], 2 * p.Dm + hg * 64, row);
            }
          }
          for (int i = 0; i < NT; ++i, ++qc) {
            const uint32_t qs = qc & 1;
            mbar_wait_wd(&q_empty[qs], ((qc >> 1) & 1) ^ 1);
            mbar_expect_tx(&q_full[qs], 32768);
            uint8_t* q = sQO + qs * 32768;
            for (int t = 0; t < nbox; ++t) {
              const int row = p.pack ? (2 * bi + t) * p.S : bi * p.S + i * 128;
              tma_load_2d(q + t * 8192, &tmap_qkv, &q_full[qs], hg * 64, row);
              tma_load_2d(q + 16384 + t * 8192, &tmap_do, &q_full[qs], hg * 64, row);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer ------------------------------
      // The whole warp walks the schedule (warp-uniform control flow: descriptors and loop state live in uniform
      // registers); one elected lane issues.  Every descriptor is base_lo + a constant: base_lo is the 16-byte-unit
      // address of the shared-memory window, broadcast so that the compiler knows it is uniform.
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, 1, 1);     // dV / dK: A and B MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);     // dQ: A K-major, B MN-major
      constexpr uint32_t HI = umma_desc_hi_sw128(1024);
      constexpr uint32_t LBO_K = (16u >> 4) << 16, LBO_MN = (8192u >> 4) << 16, LBO_PT = (16384u >> 4) << 16;
      constexpr uint32_t OFF_QO = 65536u >> 4, OFF_P = 131072u >> 4, OFF_DS = (131072u + 32768u) >> 4;
      const uint32_t base_lo = __shfl_sync(0xffffffffu, (smem_u32(smem) & 0x3FFFFu) >> 4, 0);
      const uint32_t p_lo_mn = base_lo + OFF_P + LBO_PT, ds_lo_mn = base_lo + OFF_DS + LBO_PT;
      const uint32_t ds_lo_k = base_lo + OFF_DS + LBO_K;
      uint32_t n = 0, ic = 0, qc = 0;
      // pending sub-block: its gradient products are issued after the NEXT sub-block's S / dP products
      bool pend = false;
      uint32_t pd_n = 0, pd_ic = 0, pd_ks = 0, pd_qs = 0;
      int pd_hh = 0, pd_i = 0, pd_bn = 0, pd_kb = 0;
      bool pd_last_tile = false, pd_last_item = false;
#ifdef CSM_ATTN_TIMING
      long long acc_t[16] = {0};
#endif
      auto grads = [&]() {
        ATT_T(g0);
        mbar_wait_wd(pds_full, pd_n & 1);                                   // P / dS tiles are in shared memory
        ATT_T(g1);
        ATT_ACC(8, g1, g0);
        if (pd_i == 0 && pd_hh == 0 && pd_ic > 0) mbar_wait_wd(dkv_free, (pd_ic - 1) & 1);   // previous item's dK / dV read
        if (!p.acc) {
          if (pd_n >= 2) mbar_wait_wd(&dq_free[pd_n & 1], ((pd_n >> 1) - 1) & 1);            // dQ buffer read out
        } else if (pd_kb == 0 && pd_i == 0 && pd_hh == 0 && pd_ic >= static_cast<uint32_t>(NT)) {
          mbar_wait_wd(&dq_free[0], (pd_ic / NT - 1) & 1);                                   // previous unit's dQ read out
        }
        tc_fence_after();
        const uint32_t hoff = DH == 32 ? static_cast<uint32_t>(pd_hh) * 4u : 0u;            // 64 bytes per head
        const uint32_t q_lo = base_lo + OFF_QO + pd_qs * 2048u + hoff + LBO_MN;
        const uint32_t do_lo = q_lo + 1024u;                                                // dO tile: + 16 KB
        const uint32_t k_lo = base_lo + pd_ks * 2048u + hoff + LBO_MN;
        const uint32_t tm_dv = tmem_base + COL_DV + pd_hh * DH, tm_dk = tmem_base + COL_DK + pd_hh * DH;
        // plain: the partial of a sub-block goes to one of two buffers; acc: one accumulator per (query tile, head)
        const uint32_t tm_dq = tmem_base + COL_DQ + (p.acc ? (pd_i * NH + pd_hh) * DH : (pd_n & 1) * DH);
        const uint32_t accq = (p.acc && pd_kb > 0) ? 1u : 0u;
        const uint32_t acc0 = pd_i > 0 ? 1u : 0u;
        const int ksteps = pd_bn >> 4;             // 4 or 8
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)           // dV += P^T dO_i   (reduction over the 128 query rows)
            umma_f16_lh(tm_dv, p_lo_mn + kk * 128, HI, do_lo + kk * 128, HI, idesc_g, kk > 0 ? 1u : acc0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)           // dK += dS^T Q_i
            umma_f16_lh(tm_dk, ds_lo_mn + kk * 128, HI, q_lo + kk * 128, HI, idesc_g, kk > 0 ? 1u : acc0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {         // dQ_i (partial) = dS K   (reduction over the keys of the block)
            if (ks < ksteps)
              umma_f16_lh(tm_dq, ds_lo_k + (ks >> 2) * 1024 + (ks & 3) * 2, HI, k_lo + ks * 128, HI, idesc_q,
                          ks > 0 ? 1u : accq);
          }
          umma_commit(grads_done);
          if (pd_last_tile) umma_commit(&q_empty[pd_qs]);
          if (pd_last_item) umma_commit(&kv_empty[pd_ic & 1]);
        }
        __syncwarp();
        ATT_T(g2);
        ATT_ACC(9, g2, g1);
      };
      for (;; ++ic) {
        const int item = bwd_item(p, static_cast<int>(ic));
        if (item >= p.items) break;
        const int kb = item % NT;
        const int hg = (item / NT) % p.HG;
        const int nh = min(NH, p.H - hg * NH);
        const int bn = (kb == NT - 1) ? p.bn_last : 128;
        const uint32_t idesc_s = umma_idesc_bf16(128, bn, 0, 0);
        const uint32_t ks = ic & 1;
        mbar_wait_wd(&kv_full[ks], (ic >> 1) & 1);
        const uint32_t k_lo_k = base_lo + ks * 2048u + LBO_K, v_lo_k = k_lo_k + 1024u;
        for (int i = 0; i < NT; ++i, ++qc) {
          const uint32_t qs = qc & 1;
          mbar_wait_wd(&q_full[qs], (qc >> 1) & 1);
          const uint32_t q_lo_k = base_lo + OFF_QO + qs * 2048u + LBO_K, do_lo_k = q_lo_k + 1024u;
          for (int hh = 0; hh < nh; ++hh, ++n) {
            ATT_T(m0);
            if (n > 0) mbar_wait_wd(sdp_free, (n - 1) & 1);      // the previous S / dP tiles are in registers
            tc_fence_after();
            ATT_T(m1);
            ATT_ACC(10, m1, m0);
            const uint32_t ko = hh * KS * 2;                      // 32 bytes per k-step, in 16-byte units
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < KS; ++kk)
                umma_f16_lh(tmem_base + COL_S, q_lo_k + ko + kk * 2, HI, k_lo_k + ko + kk * 2, HI, idesc_s,
                            kk > 0 ? 1u : 0u);
#pragma unroll
              for (int kk = 0; kk < KS; ++kk)
                umma_f16_lh(tmem_base + COL_DP, do_lo_k + ko + kk * 2, HI, v_lo_k + ko + kk * 2, HI, idesc_s,
                            kk > 0 ? 1u : 0u);
              umma_commit(sdp_full);
            }
            __syncwarp();
            ATT_T(m2);
            ATT_ACC(11, m2, m1);
            if (pend) grads();
            pend = true;
            pd_n = n; pd_ic = ic; pd_ks = ks; pd_qs = qs; pd_hh = hh; pd_i = i; pd_bn = bn; pd_kb = kb;
            pd_last_tile = (hh == nh - 1);
            pd_last_item = pd_last_tile && (i == NT - 1);
          }
        }
      }
      if (pend) grads();
#ifdef CSM_ATTN_TIMING
      if (lane == 0)
        for (int i = 8; i < 12; ++i) atomicAdd(&g_attn_phase[i], static_cast<unsigned long long>(acc_t[i]));
#endif
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ------------------------------ softmax / gradient warps ------------------------------
    // 16 warps: thread = (query row, quarter of the key block) -- no reduction runs along a row in the backward, so
    // the split is free and four warps per scheduler hide each other's latencies
    const int cq = (warp - 4) >> 2;           // column quarter
    const int quarter = warp & 3;             // TMEM lane quarter
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.c;
    // pinned bases (csm::pin): under the 96-register cap of this 640-thread CTA the compiler otherwise re-derives the
    // shared-memory window address (~8 instructions) and this thread's TMEM lane offset (~6) at every barrier
    // operation, statistics read and tcgen05.ld of the loop
    const uint32_t sbar = pin(smem_u32(bars));                           // + 8 * index of a barrier
    const uint32_t s_stat = pin(smem_u32(sL) + row * 4);                 // L2 of this row; delta 1024 B further
    const uint32_t tlane = pin(tmem_base + lane_off);
    const uint32_t p_row = pin(smem_u32(sP) + row * 128);         // the dS tile is 32 KB further
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);
    const uint32_t ld3 = 3u * p.Dm;
    const bool reduce_dq = NT > 1 && !p.acc;
    constexpr uint32_t B_SDP_FULL = 8 * 8, B_SDP_FREE = 9 * 8, B_PDS_FULL = 10 * 8, B_GRADS = 11 * 8, B_DQ_FREE = 12 * 8,
                       B_DKV_FREE = 14 * 8, B_ST_FULL = 15 * 8, B_ST_EMPTY = 17 * 8;
    constexpr int OQ = DH / 4;                // gradient columns read back by one thread (8 or 16)

    auto pack8 = [&](const uint32_t* v, float mul) -> uint4 {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(v[0]) * mul, __uint_as_float(v[1]) * mul);
      o.y = pack_bf16x2(__uint_as_float(v[2]) * mul, __uint_as_float(v[3]) * mul);
      o.z = pack_bf16x2(__uint_as_float(v[4]) * mul, __uint_as_float(v[5]) * mul);
      o.w = pack_bf16x2(__uint_as_float(v[6]) * mul, __uint_as_float(v[7]) * mul);
      return o;
    };
    auto tmem_ld_oq = [&](uint32_t addr, uint32_t* v) {
      if (OQ == 16) tmem_ld_32x16(addr, v);
      else tmem_ld_32x8(addr, v);
    };

    uint32_t n = 0;
    // deferred read-backs of the previous sub-block
    uint32_t have_prev;
    asm volatile("mov.u32 %0, 0;" : "=r"(have_prev));
    bool prev_last_item = false, prev_last_kb = false;
    int prev_qrow = -1, prev_krow = -1;       // global token rows (or -1: padded)
    int prev_col = 0, prev_col0 = 0, prev_nh = 0, prev_img = 0;

    // dQ partial of sub-block n - 1 and, after the last sub-block of an item, its dK / dV
    auto read_back = [&]() {
      if (!p.acc) {
        uint32_t vq[OQ];
        tmem_ld_oq(tlane + COL_DQ + ((n - 1) & 1) * DH + cq * OQ, vq);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sbar + B_DQ_FREE + ((n - 1) & 1) * 8);
        if (prev_qrow >= 0) {
          __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(prev_qrow) * ld3 + prev_col + cq * OQ;
#pragma unroll
          for (int q8 = 0; q8 < OQ / 8; ++q8) {
            const uint4 pk = pack8(vq + q8 * 8, p.scale);
            if (reduce_dq) red_add_bf16x2_v4(dst + q8 * 8, pk);
            else *reinterpret_cast<uint4*>(dst + q8 * 8) = pk;
          }
        }
      } else if (prev_last_item && prev_last_kb) {
        // acc mode, end of a unit: the dQ accumulators of all its (query tile, head) pairs, summed over the key blocks
        // (bringing all accumulators into registers first and releasing TMEM before the stores was measured slower:
        //  160 -> 170 us on the decoder shape, the wider live range costs more than the earlier release returns)
#pragma unroll 1
        for (int t = 0; t < NT * prev_nh; ++t) {
          const int i = t / prev_nh, hh = t - i * prev_nh;
          uint32_t vq[OQ];
          tmem_ld_oq(tlane + COL_DQ + (i * NH + hh) * DH + cq * OQ, vq);
          tmem_ld_wait();
          const int tq = i * 128 + row;
          if (tq < p.S) {
            __nv_bfloat16* dst = p.dqkv + (static_cast<size_t>(prev_img) * p.S + tq) * ld3 + prev_col0 + hh * DH + cq * OQ;
#pragma unroll
            for (int q8 = 0; q8 < OQ / 8; ++q8) *reinterpret_cast<uint4*>(dst + q8 * 8) = pack8(vq + q8 * 8, p.scale);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sbar + B_DQ_FREE);
      }
      if (prev_last_item) {
#pragma unroll 1
        for (int hh = 0; hh < prev_nh; ++hh) {
          uint32_t vk[OQ], vv[OQ];
          tmem_ld_oq(tlane + COL_DK + hh * DH + cq * OQ, vk);
          tmem_ld_oq(tlane + COL_DV + hh * DH + cq * OQ, vv);
          tmem_ld_wait();
          if (prev_krow >= 0) {
            __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(prev_krow) * ld3 + prev_col0 + hh * DH + cq * OQ;
#pragma unroll
            for (int q8 = 0; q8 < OQ / 8; ++q8) {
              *reinterpret_cast<uint4*>(dst + p.Dm + q8 * 8) = pack8(vk + q8 * 8, p.scale);
              *reinterpret_cast<uint4*>(dst + 2 * p.Dm + q8 * 8) = pack8(vv + q8 * 8, 1.0f);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sbar + B_DKV_FREE);
      }
    };

#ifdef CSM_ATTN_TIMING
    long long acc_t[16] = {0};
#endif
#pragma unroll 1
    for (int k = 0;; ++k) {
      const int item = bwd_item(p, k);
      if (item >= p.items) break;
      // ---- per item: key block, head group, image(s); this thread's keys and its token rows ----
      const int kb = item % NT;
      const int r_ = item / NT;
      const int hg = r_ % p.HG;
      const int bi = r_ / p.HG;
      const int nh = min(NH, p.H - hg * NH);
      const int bn = (kb == NT - 1) ? p.bn_last : 128;
      const int bnq = bn >> 2;                  // keys of this thread's quarter: one or two 16-key groups
      const int colbase = cq * bnq;
      int kvalid = p.pack ? (((row >> 6) == (cq >> 1)) ? p.S - (cq & 1) * 32 : 0) : (p.S - kb * 128 - colbase);
      kvalid = max(0, min(kvalid, bnq));
      const int b_img = p.pack ? 2 * bi + (row >> 6) : bi;
      const bool img_ok = b_img < p.B;
      const int tk = p.pack ? (row & 63) : kb * 128 + row;
      const int krow = (tk < p.S && img_ok) ? b_img * p.S + tk : -1;
      const int col0 = hg * NH * DH;
      const uint32_t ts = tlane + COL_S + colbase;
      const uint32_t td = tlane + COL_DP + colbase;
      // shared-memory offsets of this thread's two 16-byte chunks per 16-key group in the 128B-swizzled [query][key]
      // tiles: fixed for the item, kept in registers (pin_hard) instead of ~20 instructions per group and sub-block
      const uint32_t gc = static_cast<uint32_t>(colbase >> 3);
      uint32_t st_o0[2], st_o1[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t g0 = gc + 2 * q;
        st_o0[q] = pin_hard(p_row + (g0 >> 3) * 16384 + (((g0 & 7) ^ sw7) << 4));
        st_o1[q] = pin_hard(p_row + ((g0 + 1) >> 3) * 16384 + ((((g0 + 1) & 7) ^ sw7) << 4));
      }
#pragma unroll 1
      for (int i = 0; i < NT; ++i) {
        const int tq = p.pack ? (row & 63) : i * 128 + row;
        const int qrow = (tq < p.S && img_ok) ? b_img * p.S + tq : -1;
#pragma unroll 1
        for (int hh = 0; hh < nh; ++hh, ++n) {
          ATT_T(t0);
          const uint32_t sb = n & 1;
          mbar_wait_bounded_a(sbar + B_ST_FULL + sb * 8, (n >> 1) & 1);
          const float Lc = lds_f32(s_stat + sb * 512), dc = lds_f32(s_stat + 1024 + sb * 512);
          ATT_T(t1);
          mbar_wait_bounded_a(sbar + B_SDP_FULL, n & 1);
          tc_fence_after();
          ATT_T(t2);
          uint32_t su[32], du[32];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q * 16 < kvalid) {
              tmem_ld_32x16(ts + q * 16, su + q * 16);
              tmem_ld_32x16(td + q * 16, du + q * 16);
            }
          }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(sbar + B_SDP_FREE);       // the next sub-block's S / dP products may overwrite the tiles
          ATT_T(t3);

          // P and dS of this thread's keys, packed to bf16 pairs (in place over the scores)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int lim = kvalid - q * 16;          // valid keys of the group (warp-uniform)
            if (lim > 0) {
              if (lim < 16) {
#pragma unroll
                for (int e = 1; e < 16; ++e)
                  if (e >= lim) su[q * 16 + e] = 0xff800000u;          // masked key: p = 0, dS = 0
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float p0 = ex2f(fmaf(__uint_as_float(su[q * 16 + 2 * e]), c, -Lc));
                const float p1 = ex2f(fmaf(__uint_as_float(su[q * 16 + 2 * e + 1]), c, -Lc));
                su[q * 16 + e] = pack_bf16x2(p0, p1);
                du[q * 16 + e] = pack_bf16x2(p0 * (__uint_as_float(du[q * 16 + 2 * e]) - dc),
                                             p1 * (__uint_as_float(du[q * 16 + 2 * e + 1]) - dc));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) su[q * 16 + e] = du[q * 16 + e] = 0u;
            }
          }
          // the previous sub-block's gradient products are done: the P / dS tiles may be rewritten
          ATT_T(t4);
          if (have_prev) {
            mbar_wait_bounded_a(sbar + B_GRADS, (n - 1) & 1);
            tc_fence_after();
          }
          ATT_T(t5);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q * 16 < bnq) {
              sts_v4(st_o0[q], make_uint4(su[q * 16 + 0], su[q * 16 + 1], su[q * 16 + 2], su[q * 16 + 3]));
              sts_v4(st_o1[q], make_uint4(su[q * 16 + 4], su[q * 16 + 5], su[q * 16 + 6], su[q * 16 + 7]));
              sts_v4(st_o0[q] + 32768, make_uint4(du[q * 16 + 0], du[q * 16 + 1], du[q * 16 + 2], du[q * 16 + 3]));
              sts_v4(st_o1[q] + 32768, make_uint4(du[q * 16 + 4], du[q * 16 + 5], du[q * 16 + 6], du[q * 16 + 7]));
            }
          }
          ATT_T(t6);
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_a(sbar + B_PDS_FULL);
            mbar_arrive_a(sbar + B_ST_EMPTY + sb * 8);   // every lane has consumed its statistics: the ring slot may be refilled
          }
          ATT_T(t7);
          // the read-back of the previous sub-block's dQ (dK / dV) runs under this sub-block's gradient products
          if (have_prev) read_back();
          ATT_T(t8);
          ATT_ACC(0, t1, t0);      // wait: row statistics
          ATT_ACC(1, t2, t1);      // wait: S / dP products
          ATT_ACC(2, t3, t2);      // tcgen05.ld of S / dP + release
          ATT_ACC(3, t4, t3);      // exp / dS arithmetic
          ATT_ACC(4, t5, t4);      // wait: previous gradient products
          ATT_ACC(5, t6, t5);      // st.shared of P / dS
          ATT_ACC(7, t7, t6);      // proxy fence + arrive
          ATT_ACC(6, t8, t7);      // read-back of dQ (dK / dV)

          have_prev = 1;
          prev_qrow = qrow;
          prev_krow = krow;
          prev_col0 = col0;
          prev_col = col0 + hh * DH;
          prev_nh = nh;
          prev_last_item = (hh == nh - 1) && (i == NT - 1);
          prev_last_kb = (kb == NT - 1);
          prev_img = b_img;
        }
      }
    }
    if (have_prev) {
      mbar_wait_bounded_a(sbar + B_GRADS, (n - 1) & 1);
      tc_fence_after();
      read_back();
    }
#ifdef CSM_ATTN_TIMING
    if (warp == 4 && lane == 0) {
      for (int i = 0; i < 8; ++i) atomicAdd(&g_attn_phase[i], static_cast<unsigned long long>(acc_t[i]));
      atomicAdd(&g_attn_phase[15], static_cast<unsigned long long>(n));
    }
#endif
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#ifdef CSM_ATTN_TIMING
}  // namespace
// development: reads and clears the phase counters
extern "C" int csm_attn_phase_read(unsigned long long* host16) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host16, g_attn_phase, sizeof(unsigned long long) * 16);
  unsigned long long zero[16] = {0};
  cudaMemcpyToSymbol(g_attn_phase, zero, sizeof(zero));
  return CSM_OK;
}
namespace {
#endif

template <int DH>
int attn_bwd_tc_launch(const void* qkv, const void* o, const void* d_out, const float* lse, float* delta, void* dqkv,
                       int B, int S, int H, cudaStream_t stream) {
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
    p.bn_last = ((S - (p.NT - 1) * 128) + 63) & ~63;
    p.items = B * p.HG * p.NT;
    p.acc = p.NT == 2 ? 1 : 0;      // NT * (64 / DH) * DH = 128 dQ accumulator columns fit beside S, dP, dK, dV
  }
  p.scale = 1.0f / sqrtf(static_cast<float>(DH));
  p.c = 1.4426950408889634f * p.scale;
  p.lse = lse;
  p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  {
    const long long total = static_cast<long long>(B) * S * H * (DH / 8);
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 32LL * device_sms()));
    cudaError_t de = csm_launch_pdl(attn_delta_kernel<DH>, dim3(blocks), dim3(256), 0, stream,
                                    reinterpret_cast<const __nv_bfloat16*>(o),
                                    reinterpret_cast<const __nv_bfloat16*>(d_out), delta, B, S, H);
    if (de != cudaSuccess) {
      csm_set_error("attention_bwd: delta launch failed: %s", cudaGetErrorString(de));
      return CSM_ERR_CUDA;
    }
  }
  if (p.NT > 1 && !p.acc) {
    // several key blocks reduce their dQ partials into the dQ columns of dqkv: start from zero
    cudaError_t me = cudaMemset2DAsync(dqkv, static_cast<size_t>(3) * Dm * 2, 0, static_cast<size_t>(Dm) * 2,
                                       static_cast<size_t>(B) * S, stream);
    if (me != cudaSuccess) {
      csm_set_error("attention_bwd: memset failed: %s", cudaGetErrorString(me));
      return CSM_ERR_CUDA;
    }
  }
  const size_t smem = 1024 + 4 * 32768 + 65536 + 256 + 2048;
  auto kern = attn_bwd_tc_kernel<DH>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      csm_set_error("attention_bwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return CSM_ERR_CUDA;
    }
    configured = true;
  }
  CUtensorMap tq, tdo;
  const uint32_t box_rows = p.pack ? 64 : 128;
  int rc = csm_tensor_map_2d(&tq, qkv, 3ull * Dm, static_cast<uint64_t>(B) * S, 3ull * Dm, 64, box_rows, 2, 128);
  if (rc) return rc;
  rc = csm_tensor_map_2d(&tdo, d_out, Dm, static_cast<uint64_t>(B) * S, Dm, 64, box_rows, 2, 128);
  if (rc) return rc;
  const int units = p.acc ? p.items / p.NT : p.items;
  const int grid = units < device_sms() ? units : device_sms();
  cudaError_t le = csm_launch_pdl(kern, dim3(grid), dim3(ATB_THREADS), smem, stream, tq, tdo, p);
  if (le != cudaSuccess) {
    csm_set_error("attention_bwd: launch failed: %s", cudaGetErrorString(le));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

}  // namespace

extern "C" int csm_colsum_bf16(const void* dy_bf16, float* db, int rows, int N, int skip_period, int num_sms,
                               cudaStream_t stream);

// variant: 0 = probabilities through TMEM (A operand of P.V from TMEM), 1 = through shared memory (diagnostic)
extern "C" int csm_attention_fwd_tc(const void* qkv_bf16, void* out_bf16, float* lse, int B, int S, int H, int head_dim,
                                    int variant, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && S > 0 && H > 0, "csm_attention_fwd: bad sizes B=%d S=%d H=%d", B, S, H);
  CSM_CHECK_ARG((H * head_dim) % 8 == 0, "csm_attention_fwd: H * head_dim must be a multiple of 8");
  const bool tm = variant == 0;
  if (head_dim == 32) {
    if (S <= 128) {
      return tm ? attn_fwd_tc_launch<32, 4, 4, true>(qkv_bf16, out_bf16, lse, B, S, H, stream)
                : attn_fwd_tc_launch<32, 4, 4, false>(qkv_bf16, out_bf16, lse, B, S, H, stream);
    }
    return tm ? attn_fwd_tc_launch<32, 6, 7, true>(qkv_bf16, out_bf16, lse, B, S, H, stream)
              : attn_fwd_tc_launch<32, 6, 7, false>(qkv_bf16, out_bf16, lse, B, S, H, stream);
  }
  if (head_dim == 64) {
    return tm ? attn_fwd_tc_launch<64, 4, 4, true>(qkv_bf16, out_bf16, lse, B, S, H, stream)
              : attn_fwd_tc_launch<64, 4, 4, false>(qkv_bf16, out_bf16, lse, B, S, H, stream);
  }
  csm_set_error("csm_attention_fwd: head_dim must be 32 or 64 (got %d)", head_dim);
  return CSM_ERR_ARG;
}

extern "C" int csm_attention_fwd(const void* qkv_bf16, void* out_bf16, float* lse, int B, int S, int H, int head_dim,
                                 cudaStream_t stream) {
  return csm_attention_fwd_tc(qkv_bf16, out_bf16, lse, B, S, H, head_dim, 0, stream);
}

extern "C" int csm_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse,
                                 float* delta_scratch, void* dqkv_bf16, float* dbias, int B, int S, int H,
                                 int head_dim, cudaStream_t stream) {
  CSM_CHECK_ARG(B > 0 && S > 0 && H > 0, "csm_attention_bwd: bad sizes B=%d S=%d H=%d", B, S, H);
  CSM_CHECK_ARG((H * head_dim) % 8 == 0, "csm_attention_bwd: H * head_dim must be a multiple of 8");
  CSM_CHECK_ARG(delta_scratch != nullptr, "csm_attention_bwd: needs the delta scratch buffer [B * H * S] f32");
  int rc;
  if (head_dim == 32)
    rc = attn_bwd_tc_launch<32>(qkv_bf16, out_bf16, d_out_bf16, lse, delta_scratch, dqkv_bf16, B, S, H, stream);
  else if (head_dim == 64)
    rc = attn_bwd_tc_launch<64>(qkv_bf16, out_bf16, d_out_bf16, lse, delta_scratch, dqkv_bf16, B, S, H, stream);
  else {
    csm_set_error("csm_attention_bwd: head_dim must be 32 or 64 (got %d)", head_dim);
    return CSM_ERR_ARG;
  }
  if (rc) return rc;
  // attn.qkv.bias gradient = column sums of dqkv over the tokens
  if (dbias != nullptr) return csm_colsum_bf16(dqkv_bf16, dbias, B * S, 3 * H * head_dim, 0, 0, stream);
  return CSM_OK;
}
