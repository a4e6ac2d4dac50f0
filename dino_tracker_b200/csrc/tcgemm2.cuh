// CTA-pair (cta_group::2) variant of the tcgen05 GEMM of tcgemm.cuh:  D[m][n] = sum_k A[m][k] * B[n][k].
//
// Two CTAs of a cluster (two SMs) compute one 256 x 256 output tile: each CTA stages its own 128 rows of A and HALF of
// the B tile (128 of the 256 rows), the leader CTA issues tcgen05.mma.cta_group::2 (UMMA M = 256), which reads both
// CTAs' shared memory and writes each CTA's 128 x 256 accumulator into its own TMEM.  Per SM that halves the B operand
// traffic (32 KB instead of 48 KB per K-block for the single-pass modes, which are L2-limited with single-CTA tiles) and
// deepens the smem ring (6 stages).
//
//   warp 0 (both CTAs) : TMA producer -- .cta_group::2 bulk-tensor loads completing on the LEADER's full barrier
//   warp 1 (leader)    : MMA issue + tcgen05.commit.multicast (frees the smem slot / publishes the accumulator in both CTAs)
//   warp 1 (both)      : TMEM allocation (cta_group::2, collective)
//   warps 2..5 (both)  : epilogue of the CTA's own 128 rows; "accumulator drained" arrives on the leader's barrier
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc05.cuh"
#include "tcgemm.cuh"

namespace dtk {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `local` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA loads: data lands in THIS CTA's shared memory, the transaction bytes complete on the barrier at the same
// offset in the pair's leader CTA (cluster rank 0; CUTLASS' SM100_TMA_2SM_LOAD clears the peer bit of the address)
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(mapa(smem_u32(bar), 0)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(mapa(smem_u32(bar), 0)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols));
}
template <bool kTF32>
__device__ __forceinline__ void mma2_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive (once all prior MMAs of this thread are done) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

}  // namespace tc

constexpr int TC2_BM = 256, TC2_BN = 256;   // pair tile; each CTA owns 128 rows of A / D and 128 rows of B

// EPI_WARPS / SCRATCH: epilogue warps of the kernel and whether it needs the per-warp transpose scratch (coalesced epilogues);
// 8 warps with scratch trade one ring stage for it.
template <TcMode MODE, int EPI_WARPS = 4, bool SCRATCH = true>
struct Tc2Cfg {
  using Base = TcCfg<MODE, TC2_BN>;
  static constexpr int kABytes = 128 * 128, kBBytes = (TC2_BN / 2) * 128;
  static constexpr int kStageBytes = Base::kOps * (kABytes + kBBytes);
  static constexpr int kStages = (Base::kOps == 2) ? 3 : ((EPI_WARPS == 8 && SCRATCH) ? 5 : 6);
  // no alignment slack: the kernel has no static shared memory, so the dynamic array starts 1024-byte aligned (trap otherwise)
  static constexpr int kSmem = kStages * kStageBytes + 256 + ((Base::kOps == 1 && SCRATCH) ? (EPI_WARPS / 4) * TC_EPI_SCRATCH : 0);
  static constexpr uint32_t kIdesc = tc::make_idesc(Base::kFmt, TC2_BM, TC2_BN);
};

// pb.tile_start: prefix of ceil(m / 256) per group.  Same Epi contract as tc_gemm_kernel.
// EPI_WARPS = 4: warps 2..5 own one TMEM lane quadrant each (all 256 columns of a tile); EPI_WARPS = 8: two warps per
// quadrant, each covering 128 columns -- a direct epilogue functor then sees "tiles" of 128 columns (tile_end's n_tile index
// = 2 * (n0 / 256) + half).  Launch with 64 + 32 * EPI_WARPS threads and Tc2Cfg<MODE, EPI_WARPS, scratch>::kSmem.
template <TcMode MODE, class Epi, int EPI_WARPS = 4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPI_WARPS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, TcProblem pb,
                Epi epi) {
  using Base = TcCfg<MODE, TC2_BN>;
  using Cfg = Tc2Cfg<MODE, EPI_WARPS, EPI_WARPS == 4 || EpiCoalesced<Epi>::value>;
  constexpr int BN = TC2_BN;
  extern __shared__ uint8_t smem_raw[];   // no static shared memory in this kernel: the dynamic window starts 1 KB-aligned
  uint8_t* smem = smem_raw;
  if (tc::smem_u32(smem) & 1023u) __trap();   // (checked: the 128B-swizzle atoms need 1024-byte aligned stage bases)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                       // [kStages]  (leader's copy is the one that is used)
  uint64_t* empty = bars + Cfg::kStages;       // [kStages]  per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * Cfg::kStages;   // [2]        per CTA (multicast commit)
  uint64_t* tempty = tfull + 2;                // [2]        leader's copy, 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_scratch = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);
  static_assert(!EpiCoalesced<Epi>::value || Base::kOps == 1, "coalesced epilogues need the scratch of the single-pass modes");
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "4 or 8 epilogue warps");
  constexpr int EPI_COLS = BN / (EPI_WARPS / 4);   // columns of a tile one epilogue warp covers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles_n = (pb.N + BN - 1) / BN;
  const int total_tiles = pb.tile_start[pb.n_groups] * n_tiles_n;
  const int KB = (pb.K + Base::kBK - 1) / Base::kBK;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi); tc::prefetch_tmap(&tmB_hi);
    if (Base::kOps == 2) { tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmB_lo); }
    for (int s = 0; s < Cfg::kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&tfull[b], 1); tc::mbar_init(&tempty[b], 2 * EPI_WARPS); }
    tc::mbar_fence_init();
  }
  tc::cluster_sync_all();                       // barrier inits visible cluster-wide before any remote use
  if (warp == 1) tc::tmem_alloc2(tmem_slot, 2 * BN);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& g, int& m0, int& n0) {
    int mt = tile / n_tiles_n;
    n0 = (tile - mt * n_tiles_n) * BN;
    int lo = 0, hi = pb.n_groups - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (pb.tile_start[mid] <= mt) lo = mid; else hi = mid - 1;
    }
    g = lo;
    m0 = (mt - pb.tile_start[g]) * TC2_BM;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (tc::elect_one()) {
      int stage = 0, phase = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        int g, m0, n0;
        decode(tile, g, m0, n0);
        const int arow = pb.grp_row0[g] + m0 + (int)rank * 128, batch = pb.grp_batch[g];
        const int brow = n0 + (int)rank * (BN / 2);
        for (int kb = 0; kb < KB; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          if (leader) tc::mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);   // both CTAs' bytes land on this barrier
          const int k0 = kb * Base::kBK;
          tc::tma2_load_2d(&tmA_hi, &full[stage], st, k0, arow);
          if (Base::kOps == 2) tc::tma2_load_2d(&tmA_lo, &full[stage], st + Cfg::kABytes, k0, arow);
          uint8_t* sb = st + Base::kOps * Cfg::kABytes;
          tc::tma2_load_3d(&tmB_hi, &full[stage], sb, k0, brow, batch);
          if (Base::kOps == 2) tc::tma2_load_3d(&tmB_lo, &full[stage], sb + Cfg::kBBytes, k0, brow, batch);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    int stage = 0, phase = 0, it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int buf = it & 1, aphase = (it >> 1) & 1;
      tc::mbar_wait(&tempty[buf], aphase ^ 1);
      tc::fence_after_sync();
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t sa = tc::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Base::kOps * Cfg::kABytes;
#pragma unroll
          for (int ks = 0; ks < Base::kBK / Base::kUmmaK; ++ks) {
            const uint32_t koff = ks * 32;
            const uint64_t a_hi = tc::smem_desc_sw128(sa + koff), b_hi = tc::smem_desc_sw128(sb + koff);
            const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
            if (Base::kOps == 2) {
              const uint64_t a_lo = tc::smem_desc_sw128(sa + Cfg::kABytes + koff);
              const uint64_t b_lo = tc::smem_desc_sw128(sb + Cfg::kBBytes + koff);
              tc::mma2_ss<Base::kTF32>(tmem_d, a_lo, b_hi, Cfg::kIdesc, first);
              tc::mma2_ss<Base::kTF32>(tmem_d, a_hi, b_lo, Cfg::kIdesc, 1u);
              tc::mma2_ss<Base::kTF32>(tmem_d, a_hi, b_hi, Cfg::kIdesc, 1u);
            } else {
              tc::mma2_ss<Base::kTF32>(tmem_d, a_hi, b_hi, Cfg::kIdesc, first);
            }
          }
          tc::mma2_commit_mc(&empty[stage]);
          if (kb == KB - 1) tc::mma2_commit_mc(&tfull[buf]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quad = warp & 3;
    const int chalf = ((warp - 2) >> 2) * EPI_COLS;   // first column of this warp's share of the tile
    int it = 0;
    const uint32_t tempty_leader0 = tc::mapa(tc::smem_u32(&tempty[0]), 0), tempty_leader1 = tc::mapa(tc::smem_u32(&tempty[1]), 0);
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int buf = it & 1, aphase = (it >> 1) & 1;
      int g, m0, n0;
      decode(tile, g, m0, n0);
      const int r = m0 + (int)rank * 128 + quad * 32 + lane;
      const bool row_ok = r < pb.grp_m[g];
      typename Epi::State est;
      epi.tile_begin(est);
      tc::mbar_wait(&tfull[buf], aphase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN + chalf;
      // column blocks of 32; the TMEM load of the next block is in flight while this one is consumed (the epilogue warps
      // have their scheduler to themselves, so nothing else hides that latency)
      auto consume = [&](const uint32_t (&v)[32], int cc) {
        const int c = chalf + cc;
        const int ncols = min(32, pb.N - (n0 + c));
        if constexpr (EpiCoalesced<Epi>::value) {
          if (!epi.direct(n0 + c)) {
            // transpose the warp's 32 x 32 block through shared memory so that global accesses run along rows:
            // lane (r4, c4) then owns 4 consecutive columns of rows it*4 + r4 -> 8 lanes cover 128 contiguous bytes
            float* sw = epi_scratch + (warp - 2) * (32 * 36);
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(sw + lane * 36 + i) =
                  make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            __syncwarp();
            const int c4 = (lane & 7) * 4, r4 = lane >> 3;
            const int row_base = m0 + (int)rank * 128 + quad * 32;
            if (c4 < ncols) {
              if constexpr (EpiPrefetch<Epi>::value) {
                float4 pre[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const int rr = it * 4 + r4;
                  pre[it] = row_base + rr < pb.grp_m[g] ? epi.fetch(g, row_base + rr, n0 + c + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const int rr = it * 4 + r4;
                  if (row_base + rr < pb.grp_m[g])
                    epi.vec4(g, row_base + rr, n0 + c + c4, *reinterpret_cast<const float4*>(sw + rr * 36 + c4), pre[it]);
                }
              } else {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const int rr = it * 4 + r4;
                  if (row_base + rr < pb.grp_m[g])
                    epi.vec4(g, row_base + rr, n0 + c + c4, *reinterpret_cast<const float4*>(sw + rr * 36 + c4));
                }
              }
            }
            __syncwarp();
            return;
          }
        }
        if (row_ok && ncols > 0) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          epi(est, g, r, n0 + c, f, ncols);
        }
      };
      {
        uint32_t va[32], vb[32];
        tc::tmem_ld32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < EPI_COLS; c += 64) {
          tc::tmem_ld_wait();
          tc::tmem_ld32(taddr + c + 32, vb);
          consume(va, c);
          tc::tmem_ld_wait();
          if (c + 64 < EPI_COLS) tc::tmem_ld32(taddr + c + 64, va);
          consume(vb, c + 32);
        }
      }
      if (row_ok) epi.tile_end(est, g, r, (n0 + chalf) / EPI_COLS);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(buf ? tempty_leader1 : tempty_leader0);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  tc::cluster_sync_all();                       // nobody may exit / free TMEM while the peer still uses remote state
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc2(tmem_base, 2 * BN);
  }
}

}  // namespace dtk
