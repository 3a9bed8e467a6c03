// tcgen05 GEMM for the Linear layers of the Cross-Scale MAE hot path (timm Block qkv/proj/fc1/fc2,
// decoder_embed, decoder_pred, patch-embed-as-GEMM, predictor MLP) and their dgrad / wgrad.
//
//   forward : Y[M,N]  = X[M,K] . W[N,K]^T          A K-major,  B K-major
//   dgrad   : dX[M,K] = dY[M,N] . W[N,K]           A K-major,  B MN-major (W is read as stored: no
//                                                  transposed weight copy exists anywhere)
//   wgrad   : dW[N,K] += sum_r dY[r,N] . X[r,K]    A MN-major, B MN-major, reduction over the token
//                                                  rows r split into work units, TMA reduce-add (f32)
//
// Design (B200: 148 SMs = 74 SM pairs, L2 -> SM feed ~43 B/clk/SM is the binding resource for bf16
// GEMM): a PERSISTENT kernel of 2-CTA clusters.  A cluster owns a 256 x BN output tile
// (BN <= 256, chosen per problem): each CTA TMA-loads its own 128 rows of A and its own half of the B
// tile (so a k-block costs 128 + BN/2 rows of traffic per SM instead of 128 + BN), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) into TMEM accumulators that live in both SMs.
//   warp 0      TMA producer (both CTAs; completion signalled on the LEADER's full barrier)
//   warp 1      TMEM allocator; in the leader CTA the single-thread MMA issuer
//   warps 2..17 epilogue (16 warps: the GELU / gelu' epilogues are FP32-issue bound, K/32 instructions per
//               output element is all the main loop hides): tcgen05.ld -> registers -> fused bias / GELU /
//               residual / dGELU -> 64B-swizzled 32 x 64 B staging tile in shared memory -> TMA store (or TMA
//               reduce-add); residual / gelu' inputs arrive by TMA load into the same staging tile
// The accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i overlaps
// the main loop of tile i+1; the smem ring is 5 stages of 32 KB.
//
// Rounding points follow the reference's CUDA-autocast graph (SURVEY.md 8a'): a Linear's output
// is rounded to bf16 before it is added to the fp32 residual stream; GELU is evaluated in fp32 on
// the bf16-rounded fc1 output and rounded again.  The fc1 epilogue stores gelu'(h) (bf16) instead of h:
// it is the only thing the backward needs from h, and the exp it shares with gelu(h) is already there,
// so the fc2-dgrad epilogue is a single multiply (one extra bf16 rounding on dh, documented in DESIGN.md).
#include "common.cuh"

#include <mutex>
#include <unordered_map>

namespace {

using namespace csm;

constexpr int BM = 128;            // rows per CTA; the cluster tile is 2 * BM = 256 rows
constexpr int BK = 64;             // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_BN = 256;
constexpr int STAGES = 5;
constexpr int EPI_WARPS = 16;                       // 4 per TMEM lane quarter = 4 per SM sub-partition
constexpr int GEMM_THREADS = 32 * (2 + EPI_WARPS);
constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
constexpr int B_BYTES_MAX = (MAX_BN / 2) * BK * 2;   // 16 KB (each CTA holds half of the B tile)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int STG_BYTES = 2048;                      // one staging tile: 32 rows x 64 B (32 bf16 or 16 f32 columns)
constexpr int STG_BUFS = 2;
constexpr int SMEM_RING = STAGES * STAGE_BYTES;
constexpr int SMEM_STG = EPI_WARPS * STG_BUFS * STG_BYTES;
constexpr int SMEM_BAR_OFFSET = SMEM_RING + SMEM_STG;
constexpr int SMEM_TOTAL = SMEM_BAR_OFFSET + 512 + 1024;   // + barriers + 1024 B alignment slack
constexpr int ACC_COLS = 256;                        // TMEM columns per accumulator buffer

enum Epi : int {
  EPI_BF16 = 0,        // out_bf16 = bf16(acc + bias)
  EPI_GELU = 1,        // h = bf16(acc + bias); out_bf16 = bf16(gelu'(h)); out2_bf16 = bf16(gelu(h))
  EPI_RESID = 2,       // out_f32 = aux_f32 + bf16(acc + bias)              (residual stream)
  EPI_DGELU = 3,       // out_bf16 = bf16(bf16(acc) * aux_bf16),  aux = the gelu'(h) stored by EPI_GELU
  EPI_F32_ATOMIC = 4,  // out_f32 += acc                                    (TMA reduce-add)
  EPI_F32 = 5          // out_f32 = acc + bias
};

struct GemmParams {
  int M, N;              // output rows / cols
  int num_kb;            // k-blocks of the whole reduction
  int kb_per_split;      // k-blocks handled by one work unit
  int splits;
  int bn;                // cluster tile width (multiple of 64, <= 256)
  int tiles_m, tiles_n;  // 256-row tiles, bn-col tiles
  const float* bias;     // [N] or nullptr
  int sk;                // stream-K schedule: every cluster gets an equal run of k-blocks (see Sched)
  unsigned* sk_flags;    // [clusters + 1] arrival counters of the partial tiles (zero between launches)
};

// ---------------------------------------------------------------------------------------------
// cluster / 2-CTA PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A (128 rows from each CTA) * B (bn/2 rows from each CTA)
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with the descriptors as {lo, hi} words: stepping through a tile is one 32-bit add on the low word
__device__ __forceinline__ void umma2_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once) on the mbarrier at this shared-memory offset in BOTH CTAs of the pair when all previously
// issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

enum Role : int { ROLE_FULL = 0, ROLE_TAIL = 1, ROLE_HEAD = 2 };
struct WorkUnit {
  int m_tile, n_tile, kb0, nkb;
  int role;
};
__device__ __forceinline__ WorkUnit decode_unit(const GemmParams& p, int u) {
  WorkUnit w;
  w.n_tile = u % p.tiles_n;
  const int r = u / p.tiles_n;
  w.m_tile = r % p.tiles_m;
  const int split = r / p.tiles_m;
  w.kb0 = split * p.kb_per_split;
  w.nkb = min(p.kb_per_split, p.num_kb - w.kb0);
  w.role = ROLE_FULL;
  return w;
}

// Work schedule of one cluster.  Data-parallel mode: whole units u = cluster, cluster + C, ...  Stream-K mode
// (p.sk): the units' k-blocks are laid end to end and cluster c takes the run [bound(c), bound(c+1)), so a
// cluster's FIRST piece may be the tail of a unit (its k-blocks kb0..end: the partial accumulator goes to the
// fp32 workspace slot of this cluster) and its LAST piece the head of a unit (k-blocks 0..n: it adds the partial
// that cluster c+1 wrote at the very start of its run -- long before -- and runs the fused epilogue).  Runs are
// at least one unit long (units >= clusters), so a unit is cut at most once; cuts closer than SK_SNAP k-blocks to
// a unit boundary are moved onto it.
constexpr int SK_SNAP = 3;
__device__ __forceinline__ int sk_bound(const GemmParams& p, int c, int num_clusters, int total_kb) {
  if (c >= num_clusters) return total_kb;
  int b = static_cast<int>(static_cast<long long>(c) * total_kb / num_clusters);
  const int off = b % p.num_kb;
  if (off < SK_SNAP) b -= off;
  else if (p.num_kb - off < SK_SNAP) b += p.num_kb - off;
  return b;
}
struct Sched {
  int sk, u, step, total_units, pos, end, nkb, tiles_n;
  __device__ __forceinline__ Sched(const GemmParams& p, int cluster_id, int num_clusters) {
    sk = p.sk;
    u = cluster_id;
    step = num_clusters;
    total_units = p.tiles_m * p.tiles_n * p.splits;
    nkb = p.num_kb;
    tiles_n = p.tiles_n;
    pos = end = 0;
    if (sk) {
      const int total_kb = total_units * nkb;
      pos = sk_bound(p, cluster_id, num_clusters, total_kb);
      end = sk_bound(p, cluster_id + 1, num_clusters, total_kb);
    }
  }
  __device__ __forceinline__ bool next(const GemmParams& p, WorkUnit& w) {
    if (!sk) {
      if (u >= total_units) return false;
      w = decode_unit(p, u);
      u += step;
      return true;
    }
    if (pos >= end) return false;
    const int unit = pos / nkb;
    w.kb0 = pos - unit * nkb;
    w.nkb = min(nkb - w.kb0, end - pos);
    w.n_tile = unit % tiles_n;
    w.m_tile = unit / tiles_n;
    w.role = w.kb0 > 0 ? ROLE_TAIL : (w.nkb < nkb ? ROLE_HEAD : ROLE_FULL);
    pos += w.nkb;
    return true;
  }
};
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
            const __grid_constant__ CUtensorMap tmap_ws, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint64_t* aux_bar = tmem_empty + 2;              // [EPI_WARPS][STG_BUFS]
  uint64_t* part_bar = aux_bar + EPI_WARPS * STG_BUFS;   // [EPI_WARPS] stream-K partial tile loads (used once)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int half_bn = p.bn >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (EPI == EPI_GELU || EPI == EPI_RESID || EPI == EPI_DGELU) tma_prefetch_desc(&tmap_aux);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * EPI_WARPS);
    }
    for (int i = 0; i < EPI_WARPS * STG_BUFS; ++i) mbar_init(&aux_bar[i], 1);
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&part_bar[i], 1);
    if (p.sk) tma_prefetch_desc(&tmap_ws);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 2 * ACC_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    // whole warp, warp-uniform control flow; one elected lane issues the copies
    {
      const uint32_t rank_u = __shfl_sync(0xffffffffu, rank, 0);
      const uint32_t stage_tx = 2u * static_cast<uint32_t>(A_BYTES + half_bn * BK * 2);
      uint32_t stage = 0, phase = 0;
      Sched sched(p, cluster_id, num_clusters);
      WorkUnit w;
      while (sched.next(p, w)) {
        const int m0 = w.m_tile * (2 * BM) + static_cast<int>(rank_u) * BM;
        const int n0 = w.n_tile * p.bn + static_cast<int>(rank_u) * half_bn;
        for (int i = 0; i < w.nkb; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          const int k = (w.kb0 + i) * BK;
          if (elect_one()) {
            if (rank_u == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
            // K-major operand : one box {64 k, rows}.
            // MN-major operand: boxes {64 mn, 64 k}; each 64-wide MN group is its own [64 k][128 B] slab.
            if (!A_MN) {
              tma2_load_2d(sa, &tmap_a, full_leader, k, m0);
            } else {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g) tma2_load_2d(sa + g * (BK * 128), &tmap_a, full_leader, m0 + g * 64, k);
            }
            if (!B_MN) {
              tma2_load_2d(sb, &tmap_b, full_leader, k, n0);
            } else {
              for (int g = 0; g < half_bn / 64; ++g)
                tma2_load_2d(sb + g * (BK * 128), &tmap_b, full_leader, n0 + g * 64, k);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA) --------------------------------
    // The whole warp walks the schedule with warp-uniform control flow (descriptors, stage and phase live in uniform
    // registers); one elected lane issues.  A descriptor is base_lo + stage * (STAGE_BYTES / 16) + a constant: one
    // 32-bit add per MMA on the low word instead of re-deriving the 64-bit value (see common.cuh, umma_desc_lo).
    if (__shfl_sync(0xffffffffu, rank, 0) == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM, p.bn, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t HI = umma_desc_hi_sw128(1024);
      // K-major, 128B swizzle : 8-row groups are 1024 B apart (SBO); a K step of 16 is 32 B inside the row.
      // MN-major, 128B swizzle: 64-wide MN groups are BK*128 B apart (LBO), 8-deep K groups 1024 B (SBO);
      //                         a K step of 16 is 16 rows = 2048 B.
      constexpr uint32_t LBO_A = ((A_MN ? BK * 128u : 16u) >> 4) << 16, LBO_B = ((B_MN ? BK * 128u : 16u) >> 4) << 16;
      constexpr uint32_t KSTEP_A = (A_MN ? UMMA_K * 128u : UMMA_K * 2u) >> 4;
      constexpr uint32_t KSTEP_B = (B_MN ? UMMA_K * 128u : UMMA_K * 2u) >> 4;
      const uint32_t base_lo = __shfl_sync(0xffffffffu, (smem_u32(smem) & 0x3FFFFu) >> 4, 0);
      uint32_t stage = 0, phase = 0, tile_it = 0;
      Sched sched(p, cluster_id, num_clusters);
      WorkUnit w;
      for (; sched.next(p, w); ++tile_it) {
        const uint32_t acc = tile_it & 1;
        const uint32_t acc_ph = (tile_it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);      // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
        for (int i = 0; i < w.nkb; ++i) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = base_lo + stage * (STAGE_BYTES >> 4) + LBO_A;
          const uint32_t b_lo = base_lo + stage * (STAGE_BYTES >> 4) + (A_BYTES >> 4) + LBO_B;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma2_f16_lh(tmem_d, a_lo + k * KSTEP_A, HI, b_lo + k * KSTEP_B, HI, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma2_commit_mc(&empty_bar[stage]);   // frees this smem slot in both CTAs once the MMAs have read it
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit_mc(&tmem_full[acc]);   // accumulator complete: wakes the epilogue warps of both CTAs
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, 16 warps) -------------------------
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int quad = ew >> 2;                      // which chunks of the tile this warp takes (c % 4 == quad)
    constexpr bool OUT_F32 = (EPI == EPI_RESID || EPI == EPI_F32 || EPI == EPI_F32_ATOMIC);
    constexpr bool HAS_AUX = (EPI == EPI_RESID || EPI == EPI_DGELU);
    constexpr int CW = OUT_F32 ? 16 : 32;          // columns per staging tile (64 bytes per row)
    uint8_t* stg = smem + SMEM_RING + ew * (STG_BUFS * STG_BYTES);
    uint64_t* my_aux_bar = aux_bar + ew * STG_BUFS;
    const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    const int nchunks = p.bn / CW;
    const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);   // 64B swizzle: 16-byte chunk ^= (row / 2) % 4
    uint32_t seq = 0;        // staging-buffer sequence number (continues across tiles)
    uint32_t aux_seq = 0;    // aux loads issued so far
    uint32_t tile_it = 0;

    auto issue_aux = [&](const WorkUnit& w, int c) {
      const uint32_t b = aux_seq & 1;
      mbar_expect_tx(&my_aux_bar[b], STG_BYTES);
      tma_load_2d(stg + b * STG_BYTES, &tmap_aux, &my_aux_bar[b], w.n_tile * p.bn + c * CW,
                  w.m_tile * (2 * BM) + static_cast<int>(rank) * BM + q * 32);
      ++aux_seq;
    };

    Sched sched(p, cluster_id, num_clusters);
    WorkUnit w;
    uint8_t* part = smem + ew * 8192;              // stream-K head piece: this warp's share of the partial tile
                                                   // (lands in the operand ring, idle by then)
    for (; sched.next(p, w); ++tile_it) {
      const uint32_t acc = tile_it & 1;
      const uint32_t acc_ph = (tile_it >> 1) & 1;
      const int row0 = w.m_tile * (2 * BM) + static_cast<int>(rank) * BM + q * 32;
      const int col_tile = w.n_tile * p.bn;
      const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_COLS;
      if (w.role == ROLE_TAIL) {
        // ---- stream-K tail piece (always a cluster's first piece): fp32 partial -> workspace slot of this
        //      cluster, then one arrival per warp on the slot's counter ----
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
        const int nch16 = p.bn >> 4;
        const int ws_row = cluster_id * (2 * BM) + static_cast<int>(rank) * BM + q * 32;
#pragma unroll 1
        for (int c = quad; c < nch16; c += 4) {
          uint32_t r[16];
          tmem_ld_32x16(tmem_row + c * 16, r);
          tmem_ld_wait();
          if (c + 4 >= nch16) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? tmem_empty_leader1 : tmem_empty_leader0);
          }
          const uint32_t b = seq & 1;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          const uint32_t sbase = smem_u32(stg + b * STG_BYTES) + lane * 64;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = r[g * 4 + 0]; o.y = r[g * 4 + 1]; o.z = r[g * 4 + 2]; o.w = r[g * 4 + 3];
            st_shared_v4(sbase + ((static_cast<uint32_t>(g) ^ sw) << 4), o);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_ws, stg + b * STG_BYTES, c * 16, ws_row);
            bulk_commit();
          }
          ++seq;
        }
        if (lane == 0) {
          bulk_wait_all();                          // the partial is written, not just read out of smem
          asm volatile("fence.proxy.async;" ::: "memory");
          __threadfence();
          atomicAdd(p.sk_flags + cluster_id, 1u);
        }
        __syncwarp();
        seq = 0;                                    // both staging buffers are idle again
        continue;
      }
      const bool head = (w.role == ROLE_HEAD);
      if (head && lane == 0) {
        // the cluster after this one wrote the rest of this unit's reduction at the very start of its run
        const unsigned* flag = p.sk_flags + cluster_id + 1;
        while (ld_acquire_gpu(flag) < 2u * EPI_WARPS) __nanosleep(64);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      if (HAS_AUX && lane == 0 && quad < nchunks) {
        // the staging buffer the load lands in must have been read out by its previous TMA store
        bulk_wait_read<0>();
        issue_aux(w, quad);
      }
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      if (quad >= nchunks) {     // narrow tile: this warp has no chunk, it only releases the accumulator
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? tmem_empty_leader1 : tmem_empty_leader0);
      }
      if (head && quad < nchunks) {
        // all MMAs of this cluster are done (the head piece is its last), so the operand ring is free: fetch this
        // warp's whole share of the partial tile (<= 4 groups of 32 rows x 16 f32 columns) in one go
        if (lane == 0) {
          int ngroups = 0;
          for (int c = quad; c < nchunks; c += 4) ngroups += CW / 16;
          mbar_expect_tx(&part_bar[ew], static_cast<uint32_t>(ngroups) * STG_BYTES);
          const int ws_row = (cluster_id + 1) * (2 * BM) + static_cast<int>(rank) * BM + q * 32;
          int gi = 0;
          for (int c = quad; c < nchunks; c += 4)
            for (int h = 0; h < CW / 16; ++h, ++gi)
              tma_load_2d(part + gi * STG_BYTES, &tmap_ws, &part_bar[ew], c * CW + h * 16, ws_row);
        }
        mbar_wait(&part_bar[ew], 0);
      }

#pragma unroll 1
      for (int c = quad; c < nchunks; c += 4) {
        const int col0 = col_tile + c * CW;
        const bool last = (c + 4 >= nchunks);
        // ---- accumulator chunk -> registers ----
        uint32_t r[CW];
        if (CW == 32) tmem_ld_32x32(tmem_row + c * CW, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        else tmem_ld_32x16(tmem_row + c * CW, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        tmem_ld_wait();
        if (last) {
          // all TMEM reads of this tile by this warp are done: hand the buffer back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc ? tmem_empty_leader1 : tmem_empty_leader0);
        }
        float v[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
        if (head) {
          const int gi0 = ((c - quad) >> 2) * (CW / 16);
#pragma unroll
          for (int h = 0; h < CW / 16; ++h) {
            const uint32_t pbase = smem_u32(part + (gi0 + h) * STG_BYTES) + lane * 64;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 x = ld_shared_v4(pbase + ((static_cast<uint32_t>(g) ^ sw) << 4));
              v[h * 16 + g * 4 + 0] += __uint_as_float(x.x);
              v[h * 16 + g * 4 + 1] += __uint_as_float(x.y);
              v[h * 16 + g * 4 + 2] += __uint_as_float(x.z);
              v[h * 16 + g * 4 + 3] += __uint_as_float(x.w);
            }
          }
        }
        if (EPI != EPI_DGELU && EPI != EPI_F32_ATOMIC && p.bias != nullptr) {
#pragma unroll
          for (int g = 0; g < CW / 4; ++g) {
            if (col0 + g * 4 < p.N) {       // N % 8 == 0, so a group of 4 is all-in or all-out
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + g * 4));
              v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
            }
          }
        }

        if (!HAS_AUX) {
          // ---- plain / GELU / f32 / reduce: registers -> staging tile -> TMA store ----
          const uint32_t b = seq & 1;
          if (lane == 0) bulk_wait_read<1>();     // the store that last used this buffer has read it out
          __syncwarp();
          const uint32_t sbase = smem_u32(stg + b * STG_BYTES) + lane * 64;
          if (OUT_F32) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o;
              o.x = __float_as_uint(v[g * 4 + 0]); o.y = __float_as_uint(v[g * 4 + 1]);
              o.z = __float_as_uint(v[g * 4 + 2]); o.w = __float_as_uint(v[g * 4 + 3]);
              st_shared_v4(sbase + ((static_cast<uint32_t>(g) ^ sw) << 4), o);
            }
          } else if (EPI == EPI_GELU) {
            // one pass: h rounded to bf16 (the reference's fc1 output); gelu'(h) and gelu(h) go to the two staging
            // tiles of this warp, one proxy fence, two TMA stores
            if (lane == 0) bulk_wait_read<0>();   // both buffers are rewritten
            __syncwarp();
            const uint32_t sbase2 = smem_u32(stg + (b ^ 1) * STG_BYTES) + lane * 64;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o, o2;
              uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
              uint32_t* ow2 = reinterpret_cast<uint32_t*>(&o2);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float ga, gpa, gb, gpb;
                bf16_round_pair(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
                gelu_and_grad2(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1], ga, gb, gpa, gpb);
                ow[j] = pack_bf16x2(gpa, gpb);
                ow2[j] = pack_bf16x2(ga, gb);
              }
              st_shared_v4(sbase + ((static_cast<uint32_t>(g) ^ sw) << 4), o);
              st_shared_v4(sbase2 + ((static_cast<uint32_t>(g) ^ sw) << 4), o2);
            }
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o;
              o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
              o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
              st_shared_v4(sbase + ((static_cast<uint32_t>(g) ^ sw) << 4), o);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (EPI == EPI_F32_ATOMIC) tma_reduce_add_2d(&tmap_out, stg + b * STG_BYTES, col0, row0);
            else tma_store_2d(&tmap_out, stg + b * STG_BYTES, col0, row0);
            bulk_commit();
          }
          ++seq;
          if (EPI == EPI_GELU) {
            if (lane == 0) {
              tma_store_2d(&tmap_aux, stg + (b ^ 1) * STG_BYTES, col0, row0);
              bulk_commit();
            }
            ++seq;
          }
        } else {
          // ---- residual / dGELU: the aux tile was TMA-loaded into the staging buffer; combine in place ----
          const uint32_t b = seq & 1;
          // prefetch the aux tile of this warp's next chunk into the other buffer
          if (lane == 0) {
            bulk_wait_read<0>();                    // the other buffer's last store has read it out
            if (!last) {
              issue_aux(w, c + 4);
            }
          }
          mbar_wait(&my_aux_bar[b], (seq >> 1) & 1);
          const uint32_t sbase = smem_u32(stg + b * STG_BYTES) + lane * 64;
          if (EPI == EPI_RESID) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t a = sbase + ((static_cast<uint32_t>(g) ^ sw) << 4);
              uint4 x = ld_shared_v4(a);
              bf16_round_pair(v[g * 4 + 0], v[g * 4 + 1]);
              bf16_round_pair(v[g * 4 + 2], v[g * 4 + 3]);
              x.x = __float_as_uint(__uint_as_float(x.x) + v[g * 4 + 0]);
              x.y = __float_as_uint(__uint_as_float(x.y) + v[g * 4 + 1]);
              x.z = __float_as_uint(__uint_as_float(x.z) + v[g * 4 + 2]);
              x.w = __float_as_uint(__uint_as_float(x.w) + v[g * 4 + 3]);
              st_shared_v4(a, x);
            }
          } else {   // EPI_DGELU
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t a = sbase + ((static_cast<uint32_t>(g) ^ sw) << 4);
              const uint4 hv = ld_shared_v4(a);
              const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
              uint4 o;
              uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 h2 = unpack_bf16x2(hw[j]);
                bf16_round_pair(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
                ow[j] = pack_bf16x2(v[g * 8 + 2 * j] * h2.x, v[g * 8 + 2 * j + 1] * h2.y);
              }
              st_shared_v4(a, o);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, stg + b * STG_BYTES, col0, row0);
            bulk_commit();
          }
          ++seq;
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * ACC_COLS);
  }
  if (p.sk && threadIdx.x == 0 && rank == 0) {
    // this cluster consumed the partial of cluster + 1 (if its run ended inside a unit): re-arm the counter
    const int total_kb = p.tiles_m * p.tiles_n * p.num_kb;
    const int e = sk_bound(p, cluster_id + 1, num_clusters, total_kb);
    if (e % p.num_kb != 0) p.sk_flags[cluster_id + 1] = 0u;
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
  uint32_t box_inner, box_outer, elem_bytes, swizzle_bytes;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer && elem_bytes == o.elem_bytes && swizzle_bytes == o.swizzle_bytes;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.inner * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.outer * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (k.ld * 31 + k.box_inner * 131 + k.box_outer * 7 + k.elem_bytes + k.swizzle_bytes * 3) + (h << 6) + (h >> 2);
    return h;
  }
};

// Row-major matrix [outer, inner] of bf16 (elem_bytes 2) or f32 (4) with leading dimension ld (elements);
// boxes {box_inner, box_outer}, 128B-swizzled (GEMM operands) or 64B-swizzled (epilogue staging tiles).
int get_tensor_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                   uint32_t box_inner, uint32_t box_outer, uint32_t elem_bytes = 2, uint32_t swizzle_bytes = 128) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer, elem_bytes, swizzle_bytes};
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
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * elem_bytes) % 16 != 0) {
    csm_set_error("tensor map: base pointer and row pitch must be 16-byte aligned (ptr=%p ld=%llu)", ptr,
                  (unsigned long long)ld);
    return CSM_ERR_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

// number of 2-CTA clusters of this kernel that can be co-resident on the device (<= SMs / 2)
template <typename Kern>
int max_clusters(Kern kern, int num_sms) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 74, 1, 1);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = num_sms > 1 ? num_sms / 2 : 1;
  }
  return n;
}

// Stream-K workspace (registered by the host, csm_gemm_set_workspace): [4096 B of arrival counters][clusters x
// 256 x 256 f32 partial tiles].  One stream at a time may run forward / dgrad GEMMs while it is registered.
struct SkWorkspace {
  unsigned* flags = nullptr;
  float* tiles = nullptr;
  int max_clusters = 0;
};
SkWorkspace g_sk_ws;
constexpr size_t SK_FLAG_BYTES = 4096;
constexpr size_t SK_TILE_BYTES = static_cast<size_t>(2 * BM) * MAX_BN * sizeof(float);

// BN for a [M, N] output: the k-block time of a 256 x bn cluster tile is bound by the L2 -> SM feed
// (~ 128 + bn/2 rows per SM), so the cost model is rounds(bn) * (256 + bn).
int choose_bn(int M, int N, int num_clusters, bool b_mn) {
  const int tiles_m = csm_cdiv(M, 2 * BM);
  int best_bn = 256;
  long long best_cost = -1;
  // any multiple of 32 is a legal width (UMMA N % 16 == 0, 8-row swizzle atoms per CTA half, 32-column epilogue chunks);
  // 224 and 160 are what the M = 6400 encoder shapes want: N = 2304 -> 275 units of 256 x 224 in 4 rounds instead of
  // 225 of 256 x 256 in 4, N = 768 -> 125 units of 256 x 160 in 2 rounds instead of 100 of 256 x 192 in 2
  const int cands_k[7] = {256, 224, 192, 160, 128, 96, 64};
  for (int i = 0; i < 7; ++i) {
    const int bn = cands_k[i];
    if (b_mn && (bn % 128) != 0) continue;     // MN-major B is loaded in 64-wide groups per CTA
    const long long tiles = static_cast<long long>(tiles_m) * csm_cdiv(N, bn);
    const long long rounds = (tiles + num_clusters - 1) / num_clusters;
    const long long cost = rounds * (256 + bn);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_bn = bn;
    }
  }
  return best_bn;
}

template <bool A_MN, bool B_MN, int EPI>
int launch_gemm(const void* a, uint64_t a_inner, uint64_t a_outer, const void* b, uint64_t b_inner, uint64_t b_outer,
                void* out, const void* aux, const float* bias, int M, int N, int red_len, int num_sms,
                cudaStream_t stream) {
  auto kern = gemm_kernel<A_MN, B_MN, EPI>;
  static int clusters = 0;
  if (clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) {
      csm_set_error("gemm: cudaFuncSetAttribute(smem=%d) failed: %s", SMEM_TOTAL, cudaGetErrorString(e));
      return CSM_ERR_CUDA;
    }
    clusters = max_clusters(kern, num_sms > 0 ? num_sms : 148);
  }
  constexpr bool OUT_F32 = (EPI == EPI_RESID || EPI == EPI_F32 || EPI == EPI_F32_ATOMIC);
  GemmParams p{};
  p.M = M; p.N = N;
  p.num_kb = csm_cdiv(red_len, BK);
  p.bias = bias;
  p.tiles_m = csm_cdiv(M, 2 * BM);
  if (EPI == EPI_F32_ATOMIC) {
    // wgrad: few output tiles, long reduction -> split the reduction so that every cluster has work
    p.bn = (N > 128 || B_MN) ? (N > 128 ? 256 : 128) : 64;
    if (B_MN && p.bn < 128) p.bn = 128;
    p.tiles_n = csm_cdiv(N, p.bn);
    const int tiles = p.tiles_m * p.tiles_n;
    // split the reduction so that the units fill whole rounds of `clusters`: best fill wins, fewer splits
    // (less reduce-add traffic) break ties; at least 4 k-blocks per unit
    int max_splits = p.num_kb / 4;
    if (max_splits < 1) max_splits = 1;
    int best_splits = 1;
    double best_fill = -1.0;
    for (int sp = 1; sp <= max_splits && tiles * sp <= 4 * clusters; ++sp) {
      const int kbs = csm_cdiv(p.num_kb, sp);
      const int real = csm_cdiv(p.num_kb, kbs);
      const long long units = static_cast<long long>(tiles) * real;
      const long long rounds = (units + clusters - 1) / clusters;
      // time ~ rounds * k-blocks per unit (+ a fixed cost per unit for prologue / epilogue ~ 6 k-blocks)
      const double t = static_cast<double>(rounds) * (kbs + 6);
      const double fill = 1.0 / t;
      if (fill > best_fill * 1.03) {
        best_fill = fill;
        best_splits = real;
      }
    }
    p.kb_per_split = csm_cdiv(p.num_kb, best_splits);
    p.splits = csm_cdiv(p.num_kb, p.kb_per_split);
  } else {
    p.bn = choose_bn(M, N, clusters, B_MN);
    p.tiles_n = csm_cdiv(N, p.bn);
    p.kb_per_split = p.num_kb;
    p.splits = 1;
    // stream-K: when the units do not fill whole rounds of clusters, cut the k-block stream evenly instead.  Cost
    // per cluster in k-block x (rows of operand traffic): data-parallel rounds * nkb * (256 + bn) against
    // (units * nkb / clusters + SK_OVERHEAD_KB) * (256 + bn').  The overhead (measured, ~6 us ~ 14 k-blocks) is the
    // partial tile's write -> flag -> read round trip at the end of the owner's run plus the slower k-block rate of
    // clusters that no longer read the same operand tiles at the same time; reductions of 12 k-blocks never pay.
    if (g_sk_ws.tiles != nullptr && g_sk_ws.max_clusters >= clusters && p.num_kb >= 2 * SK_SNAP) {
      const long long dp_units = static_cast<long long>(p.tiles_m) * p.tiles_n;
      const double dp_cost = static_cast<double>((dp_units + clusters - 1) / clusters) * p.num_kb * (256 + p.bn);
      double best = dp_cost * 0.94;
      const int cands[2] = {256, 128};
      for (int i = 0; i < 2; ++i) {
        const int bn = cands[i];
        const long long units = static_cast<long long>(p.tiles_m) * csm_cdiv(N, bn);
        if (units <= clusters || units % clusters == 0) continue;
        const double cost = (static_cast<double>(units) * p.num_kb / clusters + 14.0) * (256 + bn);
        if (cost < best) {
          best = cost;
          p.sk = 1;
          p.bn = bn;
        }
      }
      if (p.sk) {
        p.tiles_n = csm_cdiv(N, p.bn);
        p.sk_flags = g_sk_ws.flags;
      }
    }
  }
  CUtensorMap ta, tb, to, tx, tws;
  int rc;
  // A: K-major [M, red] box {64, 128};  MN-major [red, M] box {64, 64}
  rc = A_MN ? get_tensor_map(&ta, a, a_inner, a_outer, a_inner, 64, BK) : get_tensor_map(&ta, a, a_inner, a_outer, a_inner, BK, BM);
  if (rc) return rc;
  rc = B_MN ? get_tensor_map(&tb, b, b_inner, b_outer, b_inner, 64, BK)
            : get_tensor_map(&tb, b, b_inner, b_outer, b_inner, BK, p.bn / 2);
  if (rc) return rc;
  rc = get_tensor_map(&to, out, N, M, N, OUT_F32 ? 16 : 32, 32, OUT_F32 ? 4 : 2, 64);
  if (rc) return rc;
  tx = to;
  if (EPI == EPI_GELU || EPI == EPI_DGELU) {
    rc = get_tensor_map(&tx, aux, N, M, N, 32, 32, 2, 64);
    if (rc) return rc;
  } else if (EPI == EPI_RESID) {
    rc = get_tensor_map(&tx, aux, N, M, N, 16, 32, 4, 64);
    if (rc) return rc;
  }
  tws = to;
  if (p.sk) {
    rc = get_tensor_map(&tws, g_sk_ws.tiles, MAX_BN, static_cast<uint64_t>(g_sk_ws.max_clusters) * 2 * BM, MAX_BN, 16,
                        32, 4, 64);
    if (rc) return rc;
  }
  const int units = p.tiles_m * p.tiles_n * p.splits;
  const int grid = 2 * (units < clusters ? units : clusters);
  cudaError_t le = csm_launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), SMEM_TOTAL, stream, ta, tb, to, tx, tws, p);
  if (le != cudaSuccess) {
    csm_set_error("gemm_tcgen05: launch failed: %s", cudaGetErrorString(le));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

}  // namespace

// shared with the attention kernels (declared in common.cuh)
int csm_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer, uint32_t elem_bytes, uint32_t swizzle_bytes) {
  return get_tensor_map(out, ptr, inner, outer, ld, box_inner, box_outer, elem_bytes, swizzle_bytes);
}

// ---------------------------------------------------------------------------------------------
// C-ABI (declared in include/csmae_b200.h)
// ---------------------------------------------------------------------------------------------
extern "C" int csm_gemm_workspace_bytes(int num_sms) {
  if (num_sms <= 0) num_sms = 148;
  return static_cast<int>(SK_FLAG_BYTES + static_cast<size_t>(num_sms / 2) * SK_TILE_BYTES);
}

extern "C" int csm_gemm_set_workspace(void* ws, long long nbytes) {
  const size_t bytes = nbytes > 0 ? static_cast<size_t>(nbytes) : 0;
  if (ws == nullptr) {
    g_sk_ws = SkWorkspace{};
    return CSM_OK;
  }
  CSM_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "csm_gemm_set_workspace: pointer must be 256-byte aligned");
  CSM_CHECK_ARG(bytes >= SK_FLAG_BYTES + SK_TILE_BYTES, "csm_gemm_set_workspace: %zu bytes is too small", bytes);
  g_sk_ws.flags = reinterpret_cast<unsigned*>(ws);
  g_sk_ws.tiles = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + SK_FLAG_BYTES);
  g_sk_ws.max_clusters = static_cast<int>((bytes - SK_FLAG_BYTES) / SK_TILE_BYTES);
  if (static_cast<size_t>(g_sk_ws.max_clusters) + 1 > SK_FLAG_BYTES / sizeof(unsigned))
    g_sk_ws.max_clusters = static_cast<int>(SK_FLAG_BYTES / sizeof(unsigned)) - 1;
  return CSM_OK;
}

extern "C" int csm_linear_fwd(const void* x_bf16, const void* w_bf16, const float* bias, void* out, void* aux,
                              int M, int N, int K, int epilogue, cudaStream_t stream) {
  CSM_CHECK_ARG(M > 0 && N > 0 && K > 0, "csm_linear_fwd: empty problem M=%d N=%d K=%d", M, N, K);
  CSM_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "csm_linear_fwd: N and K must be multiples of 8 (N=%d K=%d)", N, K);
  switch (epilogue) {
    case EPI_BF16:
      return launch_gemm<false, false, EPI_BF16>(x_bf16, K, M, w_bf16, K, N, out, nullptr, bias, M, N, K, 0, stream);
    case EPI_GELU:
      CSM_CHECK_ARG(aux != nullptr, "csm_linear_fwd: GELU epilogue needs aux (activation output)");
      return launch_gemm<false, false, EPI_GELU>(x_bf16, K, M, w_bf16, K, N, out, aux, bias, M, N, K, 0, stream);
    case EPI_RESID:
      CSM_CHECK_ARG(aux != nullptr, "csm_linear_fwd: residual epilogue needs aux (f32 residual input)");
      return launch_gemm<false, false, EPI_RESID>(x_bf16, K, M, w_bf16, K, N, out, aux, bias, M, N, K, 0, stream);
    case EPI_F32:
      return launch_gemm<false, false, EPI_F32>(x_bf16, K, M, w_bf16, K, N, out, nullptr, bias, M, N, K, 0, stream);
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
  switch (epilogue) {
    case EPI_BF16:
      return launch_gemm<false, true, EPI_BF16>(dy_bf16, N, M, w_bf16, K, N, dx, nullptr, nullptr, M, K, N, 0, stream);
    case EPI_DGELU:
      CSM_CHECK_ARG(aux != nullptr, "csm_linear_dgrad: dGELU epilogue needs aux (bf16 pre-activation)");
      return launch_gemm<false, true, EPI_DGELU>(dy_bf16, N, M, w_bf16, K, N, dx, aux, nullptr, M, K, N, 0, stream);
    case EPI_F32:
      return launch_gemm<false, true, EPI_F32>(dy_bf16, N, M, w_bf16, K, N, dx, nullptr, nullptr, M, K, N, 0, stream);
    default:
      csm_set_error("csm_linear_dgrad: unsupported epilogue %d", epilogue);
      return CSM_ERR_ARG;
  }
}

extern "C" int csm_linear_wgrad(const void* dy_bf16, const void* x_bf16, float* dw, int rows, int N, int K,
                                int num_sms, cudaStream_t stream) {
  CSM_CHECK_ARG(rows > 0 && N > 0 && K > 0, "csm_linear_wgrad: empty problem rows=%d N=%d K=%d", rows, N, K);
  CSM_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "csm_linear_wgrad: N and K must be multiples of 8 (N=%d K=%d)", N, K);
  // dW[N,K] : GEMM "M" = N, "N" = K, reduction over the token rows; both operands MN-major
  return launch_gemm<true, true, EPI_F32_ATOMIC>(dy_bf16, N, rows, x_bf16, K, rows, dw, nullptr, nullptr, N, K, rows,
                                                 num_sms, stream);
}
