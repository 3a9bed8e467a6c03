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
  int items, pairs;
  float c;           // d^-1/2 * log2(e)
  __nv_bfloat16* out;
  float* lse;
};

// bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the device
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
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

template <int N>
struct IC {
  static constexpr int value = N;
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem, bf16 pairs packed along K] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Work of one CTA = "pairs" (stride gridDim.x); a pair is two UNITS that run concurrently on the two softmax warp
// groups: the two d = 32 heads of one (image, 64-column head group, query tile) item, which share the item's Q and
// K / V tiles, or -- for d = 64, one head per 64-column group -- two consecutive items.
struct FwdUnit {
  int valid, item, hg, b, row_q0, row_q1, row_k0, row_k1;
};
template <int NH>
__device__ __forceinline__ FwdUnit fwd_unit(const FwdParams& p, int pair, int w) {
  FwdUnit u;
  u.item = NH == 2 ? pair : 2 * pair + w;
  u.valid = u.item < p.items && (NH == 1 || w < min(NH, p.H - ((p.pack ? u.item % p.HG : (u.item / p.QT) % p.HG)) * NH));
  if (!p.pack) {
    const int qt = u.item % p.QT;
    const int r = u.item / p.QT;
    u.hg = r % p.HG;
    u.b = r / p.HG;
    u.row_q0 = u.b * p.S + qt * 128;
    u.row_k0 = u.b * p.S;
    u.row_q1 = u.row_k1 = 0;
  } else {
    u.hg = u.item % p.HG;
    u.b = 2 * (u.item / p.HG);
    u.row_q0 = u.row_k0 = u.b * p.S;
    u.row_q1 = u.row_k1 = (u.b + 1) * p.S;
  }
  return u;
}

// DH: head dim (32: two heads per 64-column group, 64: one).  BN: keys per block (multiple of 16; 208 covers the 197
// keys of the decoder in one block).  The probabilities go back to TMEM over the score columns they came from (bf16
// pairs) and feed the P.V product as its TMEM A operand.
template <int DH, int BN>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const FwdParams p) {
  constexpr int NH = 64 / DH;          // heads per 64-column group
  constexpr int KS = DH / 16;          // 16-wide k-steps of one head in the Q / K rows
  constexpr int NST = 4 / NH;          // shared-memory stages of the Q and of the K/V tiles
  constexpr int KV_BYTES = BN * 128;   // one K (or V) block
  constexpr uint32_t COL_O = 2 * BN;   // O of warp group g at COL_O + g * DH
  static_assert(2 * BN + 2 * DH <= 512, "TMEM budget");
  static_assert(BN % 16 == 0, "key block");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [NST][16 KB]
  uint8_t* sKV = smem + NST * AT_QBYTES;               // [NST][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NST * 2 * KV_BYTES);
  uint64_t* q_full = bars;                 // [NST]
  uint64_t* q_empty = bars + NST;          // [NST]
  uint64_t* kv_full = bars + 2 * NST;      // [NST]
  uint64_t* kv_empty = bars + 3 * NST;     // [NST]
  uint64_t* s_full = bars + 4 * NST;       // [2]  per warp group
  uint64_t* p_full = s_full + 2;           // [2]
  uint64_t* o_full = s_full + 4;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], NH);        // released by the last P.V of every unit that reads the tile
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], NH);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
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
  const int nb = p.nb;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t qc = 0, kvc = 0;
        auto load_q = [&](const FwdUnit& u) {
          const uint32_t qs = qc % NST;
          mbar_wait_wd(&q_empty[qs], ((qc / NST) & 1) ^ 1);
          mbar_expect_tx(&q_full[qs], AT_QBYTES);
          uint8_t* q = sQ + qs * AT_QBYTES;
          tma_load_2d(q, &tmap_q, &q_full[qs], u.hg * 64, u.row_q0);
          if (p.pack) tma_load_2d(q + 8192, &tmap_q, &q_full[qs], u.hg * 64, u.row_q1);
          ++qc;
        };
        auto load_kv = [&](const FwdUnit& u, int j) {
          const uint32_t ks = kvc % NST;
          mbar_wait_wd(&kv_empty[ks], ((kvc / NST) & 1) ^ 1);
          mbar_expect_tx(&kv_full[ks], 2u * KV_BYTES);
          uint8_t* k = sKV + ks * 2 * KV_BYTES;
          uint8_t* v = k + KV_BYTES;
          tma_load_2d(k, &tmap_kv, &kv_full[ks], p.Dm + u.hg * 64, u.row_k0 + j * BN);
          tma_load_2d(v, &tmap_kv, &kv_full[ks], 2 * p.Dm + u.hg * 64, u.row_k0 + j * BN);
          if (p.pack) {
            tma_load_2d(k + 8192, &tmap_kv, &kv_full[ks], p.Dm + u.hg * 64, u.row_k1);
            tma_load_2d(v + 8192, &tmap_kv, &kv_full[ks], 2 * p.Dm + u.hg * 64, u.row_k1);
          }
          ++kvc;
        };
        for (int pair = blockIdx.x; pair < p.pairs; pair += gridDim.x) {
          const FwdUnit u0 = fwd_unit<NH>(p, pair, 0), u1 = fwd_unit<NH>(p, pair, 1);
          if (NH == 2) {
            load_q(u0);
            for (int j = 0; j < nb; ++j) load_kv(u0, j);
          } else {
            load_q(u0);
            if (u1.valid) load_q(u1);
            for (int j = 0; j < nb; ++j) {
              load_kv(u0, j);
              if (u1.valid) load_kv(u1, j);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer (one thread) ------------------------------
      if (lane == 0) {
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
        // Sequence of (pair, key block, unit-of-pair) steps; unit w always runs on warp group w with score buffer w.
        // Per step: S = Q K^T is issued two steps ahead of the P.V that consumes the same TMEM buffer, so a warp
        // group finds its next score tile ready when it comes back from draining O.
        struct Ent {
          uint32_t q_addr, k_addr, v_addr, q_stage, kv_stage;
          int hh, bn, last_j, valid, releases;
        };
        Ent ent[2];
        ent[0].valid = ent[1].valid = 0;
        uint32_t qc = 0, kvc = 0, cnt[2] = {0, 0};
        uint32_t cur_q[2] = {0, 0}, cur_kv = 0;
        // iteration state of the "S issue" cursor
        int a_pair = blockIdx.x, a_j = 0, a_w = 0;
        auto issue_s = [&]() -> bool {           // issues S for the next step of the sequence into buffer a_w
          if (a_pair >= p.pairs) return false;
          const int w = a_w;
          const FwdUnit u = fwd_unit<NH>(p, a_pair, w);
          Ent& e = ent[w];
          e.valid = u.valid;
          if (u.valid) {
            if (a_j == 0 && (NH == 1 || w == 0)) {       // a new Q tile
              const uint32_t qs = qc % NST;
              mbar_wait_wd(&q_full[qs], (qc / NST) & 1);
              ++qc;
              cur_q[NH == 1 ? w : 0] = qs;
            }
            if (NH == 1 || w == 0) {                     // a new K / V block
              const uint32_t ks = kvc % NST;
              mbar_wait_wd(&kv_full[ks], (kvc / NST) & 1);
              ++kvc;
              cur_kv = ks;
            }
            e.q_stage = cur_q[NH == 1 ? w : 0];
            e.kv_stage = cur_kv;
            e.hh = NH == 2 ? w : 0;
            e.q_addr = smem_u32(sQ + e.q_stage * AT_QBYTES);
            e.k_addr = smem_u32(sKV + e.kv_stage * 2 * KV_BYTES);
            e.v_addr = e.k_addr + KV_BYTES;
            e.bn = p.pack ? 128 : min(BN, ((p.S - a_j * BN) + 15) & ~15);
            e.last_j = (a_j == nb - 1);
            // tiles shared by the two heads of a group are released by both units; a group with a single head
            // (odd H) releases twice from its only unit
            e.releases = (NH == 2 && w == 0 && !fwd_unit<NH>(p, a_pair, 1).valid) ? 2 : 1;
            tc_fence_after();
            const uint32_t idesc_s = umma_idesc_bf16(128, e.bn, 0, 0);
            const uint32_t tmem_s = tmem_base + w * BN;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
              const uint64_t da = umma_smem_desc_sw128(e.q_addr + (e.hh * KS + kk) * 32, 16, 1024);
              const uint64_t db = umma_smem_desc_sw128(e.k_addr + (e.hh * KS + kk) * 32, 16, 1024);
              umma_f16(tmem_s, da, db, idesc_s, kk > 0 ? 1u : 0u);
            }
            umma_commit(&s_full[w]);
          }
          if (++a_w == 2) {
            a_w = 0;
            if (++a_j == nb) {
              a_j = 0;
              a_pair += gridDim.x;
            }
          }
          return true;
        };
        auto issue_pv = [&](int w) {
          Ent& e = ent[w];
          if (!e.valid) return;
          mbar_wait_wd(&p_full[w], cnt[w] & 1);
          ++cnt[w];
          tc_fence_after();
          const uint32_t v_addr = e.v_addr + (DH == 32 ? e.hh * 64 : 0);
          const uint32_t tmem_p = tmem_base + w * BN;
          const uint32_t tmem_o = tmem_base + COL_O + w * DH;
          const int ksteps = e.bn >> 4;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t db = umma_smem_desc_sw128(v_addr + k * 2048, 8192, 1024);
            umma_f16_ts(tmem_o, tmem_p + k * 8, db, idesc_o, k > 0 ? 1u : 0u);
          }
          umma_commit(&o_full[w]);
          for (int r = 0; r < e.releases; ++r) {
            umma_commit(&kv_empty[e.kv_stage]);
            if (e.last_j) umma_commit(&q_empty[e.q_stage]);
          }
        };
        // prologue: the first step of both warp groups; then P.V of a step followed by S of the step after next
        bool more = issue_s();
        if (more) more = issue_s();
        int w = 0;
        // every issue_s() call fills ent[a_w] of the step it issued; P.V of that entry must be issued before the
        // entry is overwritten by the S two steps later -- the loop below keeps exactly that order
        while (ent[0].valid || ent[1].valid) {
          issue_pv(w);
          ent[w].valid = 0;
          if (more) more = issue_s();      // refills ent[w] (the cursor's a_w == w here by construction)
          w ^= 1;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------ softmax warp groups ------------------------------
    const int g = (warp - 4) >> 2;            // warp group = unit of the pair = score buffer
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.c;
    const uint32_t tmem_s = tmem_base + lane_off + g * BN;
    const uint32_t tmem_o = tmem_base + lane_off + COL_O + g * DH;
    uint32_t cnt = 0;

#pragma unroll 1
    for (int pair = blockIdx.x; pair < p.pairs; pair += gridDim.x) {
      const FwdUnit u = fwd_unit<NH>(p, pair, g);
      if (!u.valid) continue;
      const int h = u.hg * NH + (NH == 2 ? g : 0);
      int b, srow;
      if (!p.pack) {
        b = u.b;
        srow = (u.item % p.QT) * 128 + row;
      } else {
        b = u.b + (row >> 6);
        srow = row & 63;
      }
      const bool row_valid = (srow < p.S) && (b < p.B);
      // warps whose 32 rows are all padding skip the arithmetic (their P / O rows are garbage nobody reads)
      const bool warp_active = p.pack ? ((u.b + (quarter >> 1)) < p.B && (quarter & 1) * 32 < p.S)
                                      : ((u.item % p.QT) * 128 + quarter * 32 < p.S);
      float m_run = -INFINITY, l_run = 0.f;
      float o_acc[DH];
#pragma unroll
      for (int i = 0; i < DH; ++i) o_acc[i] = 0.f;
#pragma unroll 1
      for (int j = 0; j < nb; ++j, ++cnt) {
        // this row's keys inside the block: columns [c0, c0 + kvalid); everything else of the block's bn columns
        // receives zero probabilities
        const int bn = p.pack ? 128 : min(BN, ((p.S - j * BN) + 15) & ~15);
        const int c0 = p.pack ? (row >> 6) * 64 : 0;
        const int kvalid = p.pack ? p.S : min(BN, p.S - j * BN);
        mbar_wait_wd(&s_full[g], cnt & 1);
        tc_fence_after();
        float alpha = 1.f;
        if (warp_active) {
          // ---- pass 1: row maximum (the next chunk's TMEM load is in flight while a chunk is reduced)
          float mx0 = -INFINITY, mx1 = -INFINITY;
          {
            uint32_t a[32], bq[32];
            const int nch = (kvalid + 31) >> 5;
            tmem_ld_32x32(tmem_s + c0, a);
#pragma unroll 1
            for (int ch = 0; ch < nch; ch += 2) {
              tmem_ld_wait();
              if (ch + 1 < nch) tmem_ld_32x32(tmem_s + c0 + (ch + 1) * 32, bq);
              {
                const int lim = kvalid - ch * 32;
                if (lim >= 32) {
#pragma unroll
                  for (int e = 0; e < 32; e += 2) {
                    mx0 = fmaxf(mx0, __uint_as_float(a[e]));
                    mx1 = fmaxf(mx1, __uint_as_float(a[e + 1]));
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 32; ++e)
                    if (e < lim) mx0 = fmaxf(mx0, __uint_as_float(a[e]));
                }
              }
              if (ch + 1 < nch) {
                tmem_ld_wait();
                if (ch + 2 < nch) tmem_ld_32x32(tmem_s + c0 + (ch + 2) * 32, a);
                const int lim = kvalid - (ch + 1) * 32;
                if (lim >= 32) {
#pragma unroll
                  for (int e = 0; e < 32; e += 2) {
                    mx0 = fmaxf(mx0, __uint_as_float(bq[e]));
                    mx1 = fmaxf(mx1, __uint_as_float(bq[e + 1]));
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 32; ++e)
                    if (e < lim) mx0 = fmaxf(mx0, __uint_as_float(bq[e]));
                }
              }
            }
          }
          const float m_new = fmaxf(m_run, fmaxf(mx0, mx1));
          alpha = ex2f((m_run - m_new) * c);
          m_run = m_new;
          const float mc = m_new * c;
          // ---- pass 2: p = exp2(s*c - m*c) -> bf16 pairs, written over the score columns the P.V product reads as
          // its A operand (16-key groups, two per chunk); the next chunk's TMEM load is in flight meanwhile
          float ls0 = 0.f, ls1 = 0.f;
          const int ngroups = bn >> 4;                 // 16-key groups the P.V product reads
          const int g0 = c0 >> 4;                      // first group of this row's keys
          const int nch2 = (ngroups + 1) >> 1;
          auto chunk = [&](const uint32_t* s, int ch) {
            uint32_t pk[16];
            const int gi = 2 * ch;
            const int k0 = (gi - g0) * 16;             // key index (within this row's keys) of s[0]
            const bool two = gi + 1 < ngroups;
            const int lim = kvalid - k0;               // valid keys from s[0] on (may be <= 0 or >= 32)
            if (k0 >= 0 && (lim >= 32 || (!two && lim >= 16))) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float e0 = ex2f(fmaf(__uint_as_float(s[2 * e]), c, -mc));
                const float e1 = ex2f(fmaf(__uint_as_float(s[2 * e + 1]), c, -mc));
                ls0 += e0;
                ls1 += e1;
                pk[e] = pack_bf16x2(e0, e1);
              }
            } else if (k0 >= 0 && lim > 0) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float e0 = (2 * e < lim) ? ex2f(fmaf(__uint_as_float(s[2 * e]), c, -mc)) : 0.f;
                const float e1 = (2 * e + 1 < lim) ? ex2f(fmaf(__uint_as_float(s[2 * e + 1]), c, -mc)) : 0.f;
                ls0 += e0;
                ls1 += e1;
                pk[e] = pack_bf16x2(e0, e1);
              }
            } else {                                   // the other image's keys (packed tiles): zero probabilities
#pragma unroll
              for (int e = 0; e < 16; ++e) pk[e] = 0u;
            }
            if (two) tmem_st_32x16(tmem_s + gi * 8, pk);
            else tmem_st_32x8(tmem_s + gi * 8, pk);
          };
          {
            uint32_t sa[32], sb[32];
            tmem_ld_32x32(tmem_s, sa);
#pragma unroll 1
            for (int ch = 0; ch < nch2; ch += 2) {
              tmem_ld_wait();
              if (ch + 1 < nch2) tmem_ld_32x32(tmem_s + (ch + 1) * 32, sb);
              chunk(sa, ch);
              if (ch + 1 < nch2) {
                tmem_ld_wait();
                if (ch + 2 < nch2) tmem_ld_32x32(tmem_s + (ch + 2) * 32, sa);
                chunk(sb, ch + 1);
              }
            }
          }
          l_run = fmaf(l_run, alpha, ls0 + ls1);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        // ---- O of this block
        mbar_wait_wd(&o_full[g], cnt & 1);
        tc_fence_after();
        if (warp_active) {
          uint32_t o[DH];
#pragma unroll
          for (int q = 0; q < DH / 16; ++q) tmem_ld_32x16(tmem_o + q * 16, o + q * 16);
          tmem_ld_wait();
          if (nb == 1) {
#pragma unroll
            for (int i = 0; i < DH; ++i) o_acc[i] = __uint_as_float(o[i]);
          } else {
#pragma unroll
            for (int i = 0; i < DH; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(o[i]));
          }
        }
        tc_fence_before();
      }
      if (row_valid) {
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.S + srow) * p.Dm + h * DH;
#pragma unroll
        for (int q8 = 0; q8 < DH / 8; ++q8) {
          uint4 pk;
          pk.x = pack_bf16x2(o_acc[q8 * 8 + 0] * inv, o_acc[q8 * 8 + 1] * inv);
          pk.y = pack_bf16x2(o_acc[q8 * 8 + 2] * inv, o_acc[q8 * 8 + 3] * inv);
          pk.z = pack_bf16x2(o_acc[q8 * 8 + 4] * inv, o_acc[q8 * 8 + 5] * inv);
          pk.w = pack_bf16x2(o_acc[q8 * 8 + 6] * inv, o_acc[q8 * 8 + 7] * inv);
          *reinterpret_cast<uint4*>(dst + q8 * 8) = pk;
        }
        p.lse[(static_cast<size_t>(b) * p.H + h) * p.S + srow] = m_run * c + log2f(l_run);
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

int device_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

template <int DH, int BN>
int attn_fwd_tc_launch(const void* qkv, void* out, float* lse, int B, int S, int H, cudaStream_t stream) {
  constexpr int NH = 64 / DH;
  constexpr int NST = 4 / NH;
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
  p.pairs = NH == 2 ? p.items : (p.items + 1) / 2;
  p.c = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  const size_t smem = 1024 + NST * (AT_QBYTES + 2 * static_cast<size_t>(BN) * 128) + 256;
  auto kern = attn_fwd_tc_kernel<DH, BN>;
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
  const int grid = p.pairs < device_sms() ? p.pairs : device_sms();
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
// delta = rowsum(dO * O) comes from attn_delta_kernel.
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  int B, S, H, Dm, HG;
  int pack;          // two images per 128-row tile (S <= 64)
  int NT;            // 128-row query tiles == 128-key blocks per image
  int bn_last;       // keys in the last key block (multiple of 64)
  int items;
  float c, scale;
  const float* lse;
  const float* delta;      // rowsum(dO * O) per (image, head, token), attn_delta_kernel
  __nv_bfloat16* dqkv;
};

// delta[(b*H + h)*S + s] = sum_j dO[b*S + s, h*DH + j] * O[b*S + s, h*DH + j]: one thread per (token, head),
// consecutive threads read consecutive 2*DH-byte pieces of a row
template <int DH>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_out,
                                  float* __restrict__ delta, int B, int S, int H) {
  pdl_wait();
  pdl_trigger();
  const long long total = static_cast<long long>(B) * S * H;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(idx % H);
    const long long r = idx / H;                         // b * S + s
    const uint4* op = reinterpret_cast<const uint4*>(o + idx * DH);
    const uint4* dp = reinterpret_cast<const uint4*>(d_out + idx * DH);
    float acc = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < DH / 8; ++c8) {
      const uint4 a = __ldg(dp + c8), bb = __ldg(op + c8);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(bw[e]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
    const long long b = r / S, s_ = r % S;
    delta[(b * H + h) * S + s_] = acc;
  }
}

__device__ __forceinline__ void red_add_bf16x2_v4(void* gptr, uint4 v) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// the sub-block sequence of one CTA: items (stride gridDim.x) x query tiles x heads of the group
struct BwdIter {
  int item, i, hh, nh, hg, bi, kb;
  __device__ __forceinline__ void set_item(const BwdParams& p, int NH) {
    kb = item % p.NT;
    const int r = item / p.NT;
    hg = r % p.HG;
    bi = r / p.HG;
    nh = min(NH, p.H - hg * NH);
  }
  __device__ __forceinline__ void init(const BwdParams& p, int NH) {
    item = blockIdx.x;
    i = hh = 0;
    if (item < p.items) set_item(p, NH);
  }
  __device__ __forceinline__ bool valid(const BwdParams& p) const { return item < p.items; }
  __device__ __forceinline__ bool last_of_item(const BwdParams& p) const { return hh == nh - 1 && i == p.NT - 1; }
  __device__ __forceinline__ void next(const BwdParams& p, int NH) {
    if (++hh < nh) return;
    hh = 0;
    if (++i < p.NT) return;
    i = 0;
    item += gridDim.x;
    if (item < p.items) set_item(p, NH);
  }
};

constexpr int ATB_SM_WARPS = 16;
constexpr int ATB_THREADS = 32 * (4 + ATB_SM_WARPS);   // 640

template <int DH>
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const BwdParams p) {
  constexpr int NH = 64 / DH;
  constexpr int KS = DH / 16;
  constexpr int OC = DH / 2;           // gradient columns owned by one softmax thread
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 256 + NH * DH, COL_DQ = 256 + 2 * NH * DH;
  static_assert(COL_DQ + 2 * DH <= 512, "TMEM budget");
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t ic = 0, qc = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
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
      // ------------------------------ MMA issuer (one thread) ------------------------------
      if (lane == 0) {
        constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, 1, 1);     // dV / dK: A and B MN-major
        constexpr uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);     // dQ: A K-major, B MN-major
        const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
        uint32_t n = 0, ic = 0, qc = 0;
        // pending sub-block: its gradient products are issued after the NEXT sub-block's S / dP products
        bool pend = false;
        uint32_t pd_n = 0, pd_ic = 0, pd_k = 0, pd_q = 0, pd_qs = 0;
        int pd_hh = 0, pd_i = 0, pd_bn = 0;
        bool pd_last_tile = false, pd_last_item = false;
        auto grads = [&]() {
          mbar_wait_wd(pds_full, pd_n & 1);                                   // P / dS tiles are in shared memory
          if (pd_i == 0 && pd_hh == 0 && pd_ic > 0) mbar_wait_wd(dkv_free, (pd_ic - 1) & 1);   // previous item's dK / dV read
          if (pd_n >= 2) mbar_wait_wd(&dq_free[pd_n & 1], ((pd_n >> 1) - 1) & 1);              // dQ buffer read out
          tc_fence_after();
          const uint32_t hoff = DH == 32 ? pd_hh * 64 : 0;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {           // dV += P^T dO_i   (reduction over the 128 query rows)
            const uint64_t da = umma_smem_desc_sw128(p_addr + kk * 2048, 16384, 1024);
            const uint64_t db = umma_smem_desc_sw128(pd_q + 16384 + kk * 2048 + hoff, 8192, 1024);
            umma_f16(tmem_base + COL_DV + pd_hh * DH, da, db, idesc_g, (pd_i > 0 || kk > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {           // dK += dS^T Q_i
            const uint64_t da = umma_smem_desc_sw128(ds_addr + kk * 2048, 16384, 1024);
            const uint64_t db = umma_smem_desc_sw128(pd_q + kk * 2048 + hoff, 8192, 1024);
            umma_f16(tmem_base + COL_DK + pd_hh * DH, da, db, idesc_g, (pd_i > 0 || kk > 0) ? 1u : 0u);
          }
          const int ksteps = pd_bn >> 4;
          for (int ks = 0; ks < ksteps; ++ks) {      // dQ_i (partial) = dS K   (reduction over the keys of the block)
            const uint64_t da = umma_smem_desc_sw128(ds_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(pd_k + ks * 2048 + hoff, 8192, 1024);
            umma_f16(tmem_base + COL_DQ + (pd_n & 1) * DH, da, db, idesc_q, ks > 0 ? 1u : 0u);
          }
          umma_commit(grads_done);
          if (pd_last_tile) umma_commit(&q_empty[pd_qs]);
          if (pd_last_item) umma_commit(&kv_empty[pd_ic & 1]);
        };
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ic) {
          const int kb = item % NT;
          const int hg = (item / NT) % p.HG;
          const int nh = min(NH, p.H - hg * NH);
          const int bn = (kb == NT - 1) ? p.bn_last : 128;
          const uint32_t idesc_s = umma_idesc_bf16(128, bn, 0, 0);
          const uint32_t ks = ic & 1;
          mbar_wait_wd(&kv_full[ks], (ic >> 1) & 1);
          const uint32_t k_addr = smem_u32(sKV + ks * 32768), v_addr = k_addr + 16384;
          for (int i = 0; i < NT; ++i, ++qc) {
            const uint32_t qs = qc & 1;
            mbar_wait_wd(&q_full[qs], (qc >> 1) & 1);
            const uint32_t q_addr = smem_u32(sQO + qs * 32768), do_addr = q_addr + 16384;
            for (int hh = 0; hh < nh; ++hh, ++n) {
              if (n > 0) mbar_wait_wd(sdp_free, (n - 1) & 1);      // the previous S / dP tiles are in registers
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < KS; ++kk) {
                const uint32_t ko = (hh * KS + kk) * 32;
                umma_f16(tmem_base + COL_S, umma_smem_desc_sw128(q_addr + ko, 16, 1024),
                         umma_smem_desc_sw128(k_addr + ko, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
              }
#pragma unroll
              for (int kk = 0; kk < KS; ++kk) {
                const uint32_t ko = (hh * KS + kk) * 32;
                umma_f16(tmem_base + COL_DP, umma_smem_desc_sw128(do_addr + ko, 16, 1024),
                         umma_smem_desc_sw128(v_addr + ko, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
              }
              umma_commit(sdp_full);
              if (pend) grads();
              pend = true;
              pd_n = n; pd_ic = ic; pd_k = k_addr; pd_q = q_addr; pd_qs = qs; pd_hh = hh; pd_i = i; pd_bn = bn;
              pd_last_tile = (hh == nh - 1);
              pd_last_item = pd_last_tile && (i == NT - 1);
            }
          }
        }
        if (pend) grads();
      }
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
    const uint32_t p_row = smem_u32(sP) + row * 128;
    const uint32_t ds_row = smem_u32(sdS) + row * 128;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);
    const uint32_t ld3 = 3u * p.Dm;
    const bool reduce_dq = NT > 1;
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

    BwdIter cur, nxt;
    cur.init(p, NH);
    nxt = cur;
    // per-item token bookkeeping of this thread's TMEM lane (32-bit: B * H * S < 2^31)
    auto img_of = [&](const BwdIter& it) { return p.pack ? 2 * it.bi + (row >> 6) : it.bi; };
    auto tok_of = [&](int t) { return p.pack ? (row & 63) : t * 128 + row; };
    auto load_stats = [&](const BwdIter& it, float& L, float& d) {
      const int b_img = img_of(it), tok = tok_of(it.i);
      L = INFINITY;      // padded query row: p = exp2(-inf) = 0
      d = 0.f;
      if (tok < p.S && b_img < p.B) {
        const uint32_t o = (static_cast<uint32_t>(b_img) * p.H + it.hg * NH + it.hh) * p.S + tok;
        L = __ldg(p.lse + o);
        d = __ldg(p.delta + o);
      }
    };
    float Lc = INFINITY, dc = 0.f;
    if (cur.valid(p)) {
      load_stats(cur, Lc, dc);
      nxt.next(p, NH);
    }
    uint32_t n = 0;
    // deferred read-backs of the previous sub-block
    uint32_t have_prev;
    asm volatile("mov.u32 %0, 0;" : "=r"(have_prev));
    bool prev_last_item = false;
    int prev_qrow = -1, prev_krow = -1;       // global token rows (or -1: padded)
    int prev_col = 0, prev_col0 = 0, prev_nh = 0;

    // dQ partial of sub-block n - 1 and, after the last sub-block of an item, its dK / dV
    auto read_back = [&]() {
      uint32_t vq[OQ];
      tmem_ld_oq(tmem_base + lane_off + COL_DQ + ((n - 1) & 1) * DH + cq * OQ, vq);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_free[(n - 1) & 1]);
      if (prev_qrow >= 0) {
        __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(prev_qrow) * ld3 + prev_col + cq * OQ;
#pragma unroll
        for (int q8 = 0; q8 < OQ / 8; ++q8) {
          const uint4 pk = pack8(vq + q8 * 8, p.scale);
          if (reduce_dq) red_add_bf16x2_v4(dst + q8 * 8, pk);
          else *reinterpret_cast<uint4*>(dst + q8 * 8) = pk;
        }
      }
      if (prev_last_item) {
#pragma unroll 1
        for (int hh = 0; hh < prev_nh; ++hh) {
          uint32_t vk[OQ], vv[OQ];
          tmem_ld_oq(tmem_base + lane_off + COL_DK + hh * DH + cq * OQ, vk);
          tmem_ld_oq(tmem_base + lane_off + COL_DV + hh * DH + cq * OQ, vv);
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
        if (lane == 0) mbar_arrive(dkv_free);
      }
    };

#pragma unroll 1
    while (cur.valid(p)) {
      const int bn = (cur.kb == NT - 1) ? p.bn_last : 128;
      const int bnq = bn >> 2;                  // keys of this thread's quarter: one or two 16-key groups
      const int colbase = cq * bnq;
      int kvalid = p.pack ? (((row >> 6) == (cq >> 1)) ? p.S - (cq & 1) * 32 : 0) : (p.S - cur.kb * 128 - colbase);
      kvalid = max(0, min(kvalid, bnq));
      // statistics of the next sub-block's row: in flight during this sub-block
      float Ln = INFINITY, dn = 0.f;
      if (nxt.valid(p)) load_stats(nxt, Ln, dn);

      mbar_wait_wd(sdp_full, n & 1);
      tc_fence_after();
      uint32_t su[32], du[32];
      const uint32_t ts = tmem_base + lane_off + COL_S + colbase;
      const uint32_t td = tmem_base + lane_off + COL_DP + colbase;
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
      if (lane == 0) mbar_arrive(sdp_free);       // the next sub-block's S / dP products may overwrite the tiles

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
      // the previous sub-block's gradient products are done: the P / dS tiles may be rewritten and its dQ partial
      // (and dK / dV) read.  The stores go first so that they drain under the read-back, before the proxy fence.
      if (have_prev) {
        mbar_wait_wd(grads_done, (n - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q * 16 < bnq) {
          const uint32_t gc = static_cast<uint32_t>((colbase >> 3) + 2 * q);
          const uint32_t o0 = (gc >> 3) * 16384 + (((gc & 7) ^ sw7) << 4);
          const uint32_t o1 = ((gc + 1) >> 3) * 16384 + ((((gc + 1) & 7) ^ sw7) << 4);
          sts_v4(p_row + o0, make_uint4(su[q * 16 + 0], su[q * 16 + 1], su[q * 16 + 2], su[q * 16 + 3]));
          sts_v4(p_row + o1, make_uint4(su[q * 16 + 4], su[q * 16 + 5], su[q * 16 + 6], su[q * 16 + 7]));
          sts_v4(ds_row + o0, make_uint4(du[q * 16 + 0], du[q * 16 + 1], du[q * 16 + 2], du[q * 16 + 3]));
          sts_v4(ds_row + o1, make_uint4(du[q * 16 + 4], du[q * 16 + 5], du[q * 16 + 6], du[q * 16 + 7]));
        }
      }
      if (have_prev) read_back();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);

      have_prev = 1;
      {
        const int b_img = img_of(cur);
        const int tq = tok_of(cur.i), tk = tok_of(cur.kb);
        prev_qrow = (tq < p.S && b_img < p.B) ? b_img * p.S + tq : -1;
        prev_col0 = cur.hg * NH * DH;
        prev_col = prev_col0 + cur.hh * DH;
        prev_nh = cur.nh;
        prev_last_item = cur.last_of_item(p);
        if (prev_last_item) prev_krow = (tk < p.S && b_img < p.B) ? b_img * p.S + tk : -1;
      }
      cur = nxt;
      Lc = Ln;
      dc = dn;
      if (nxt.valid(p)) nxt.next(p, NH);
      ++n;
    }
    if (have_prev) {
      mbar_wait_wd(grads_done, (n - 1) & 1);
      tc_fence_after();
      read_back();
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
  }
  p.scale = 1.0f / sqrtf(static_cast<float>(DH));
  p.c = 1.4426950408889634f * p.scale;
  p.lse = lse;
  p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  {
    const long long total = static_cast<long long>(B) * S * H;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * device_sms()));
    cudaError_t de = csm_launch_pdl(attn_delta_kernel<DH>, dim3(blocks), dim3(256), 0, stream,
                                    reinterpret_cast<const __nv_bfloat16*>(o),
                                    reinterpret_cast<const __nv_bfloat16*>(d_out), delta, B, S, H);
    if (de != cudaSuccess) {
      csm_set_error("attention_bwd: delta launch failed: %s", cudaGetErrorString(de));
      return CSM_ERR_CUDA;
    }
  }
  if (p.NT > 1) {
    // several key blocks reduce their dQ partials into the dQ columns of dqkv: start from zero
    cudaError_t me = cudaMemset2DAsync(dqkv, static_cast<size_t>(3) * Dm * 2, 0, static_cast<size_t>(Dm) * 2,
                                       static_cast<size_t>(B) * S, stream);
    if (me != cudaSuccess) {
      csm_set_error("attention_bwd: memset failed: %s", cudaGetErrorString(me));
      return CSM_ERR_CUDA;
    }
  }
  const size_t smem = 1024 + 4 * 32768 + 65536 + 256;
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
  const int grid = p.items < device_sms() ? p.items : device_sms();
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

// (variant is ignored: the shared-memory P path of the first version is gone; kept for tools/attn_tc_check.py)
extern "C" int csm_attention_fwd_tc(const void* qkv_bf16, void* out_bf16, float* lse, int B, int S, int H, int head_dim,
                                    int variant, cudaStream_t stream) {
  (void)variant;
  CSM_CHECK_ARG(B > 0 && S > 0 && H > 0, "csm_attention_fwd: bad sizes B=%d S=%d H=%d", B, S, H);
  CSM_CHECK_ARG((H * head_dim) % 8 == 0, "csm_attention_fwd: H * head_dim must be a multiple of 8");
  if (head_dim == 32) {
    if (S <= 128) return attn_fwd_tc_launch<32, 128>(qkv_bf16, out_bf16, lse, B, S, H, stream);
    return attn_fwd_tc_launch<32, 208>(qkv_bf16, out_bf16, lse, B, S, H, stream);
  }
  if (head_dim == 64) return attn_fwd_tc_launch<64, 128>(qkv_bf16, out_bf16, lse, B, S, H, stream);
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
