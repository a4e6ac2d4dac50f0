// Fused attention for the ViT (sm_100a, tcgen05):  O = softmax(Q K^T) V  per (frame, head).  Q arrives pre-scaled by
// head_dim^-1/2 * log2(e), so the softmax is exp2(s - max) (one FADD + one MUFU.EX2 per score).
//
// One CTA per (frame*head, 128-query tile); 192 threads:
//   warp 0     : TMA producer (Q once; K_j [BKV keys][64] and V^T_j [64][BKV keys] through a 2-stage ring)
//   warp 1     : TMEM alloc + MMA issue:  S = Q K_j^T (kind::f16, M128 N=BKV K64) -> TMEM;  O_j = P_j V_j (M128 N64 K=BKV)
//                (BKV = 64: 66 KB of shared memory and 128 TMEM columns per CTA -> 3 CTAs per SM overlap each other)
//   warps 2..5 : online softmax, thread = query row: running max / sum in registers, P_j written as fp16 into a
//                128B-swizzled shared-memory tile (the A operand of the second MMA), O accumulated in registers
//                (O_j is read back from TMEM and added after the rescale), final normalise + fp32 store.
// The score matrix (8108 x 8108 per head) never leaves the SM.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc05.cuh"

namespace dtk {

#ifndef DTK_FA_BKV
#define DTK_FA_BKV 64
#endif
constexpr int FA_BQ = 128, FA_BKV = DTK_FA_BKV, FA_D = 64, FA_THREADS = 192;
constexpr int FA_NSUB = FA_BKV / 64;               // 64-key sub-tiles (one 128-byte swizzle row of fp16 each)
constexpr int FA_SQ = FA_BQ * 128;                 // Q tile bytes (128 rows x 64 fp16)
constexpr int FA_SK = FA_BKV * 128;                // K tile bytes
constexpr int FA_SV = FA_NSUB * FA_D * 128;        // V^T tile: sub-tiles [64 d][64 keys]
constexpr int FA_SP = FA_NSUB * FA_BQ * 128;       // P tile: sub-tiles [128 rows][64 keys]
constexpr int FA_TMEM = FA_BKV + 64 <= 128 ? 128 : 256;
constexpr int FA_STAGE = FA_SK + FA_SV;
constexpr int FA_SMEM = FA_SQ + 2 * FA_STAGE + FA_SP + 256;   // extern smem is declared 1024-byte aligned

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct FlashParams {
  int N1;          // tokens per frame (keys = queries)
  int D;           // model dim (output row pitch)
  int heads;
  void* out;       // [B*N1][D] fp32 or fp16; head h writes columns [h*64, h*64+64)
  int out_f16;
};

__global__ void __launch_bounds__(FA_THREADS, FA_BKV == 64 ? 2 : 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, FlashParams fp) {
  extern __shared__ __align__(1024) uint8_t fa_smem[];
  uint8_t* sQ = fa_smem;
  uint8_t* sKV = sQ + FA_SQ;                 // stage s: K at sKV + s*FA_STAGE, V at + FA_SK
  uint8_t* sP = sKV + 2 * FA_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FA_SP);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;      // 1
  uint64_t* p_ready = bars + 6;     // 1 (4 arrivals)
  uint64_t* o_full = bars + 7;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, q0 = blockIdx.x * FA_BQ;
  const int N1 = fp.N1;
  const int n_kv = (N1 + FA_BKV - 1) / FA_BKV;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    tc::mbar_init(s_full, 1); tc::mbar_init(p_ready, 4); tc::mbar_init(o_full, 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, FA_TMEM);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + FA_BKV;
  constexpr uint32_t kIdescS = tc::make_idesc(0, 128, FA_BKV);   // f16 inputs, fp32 accumulate
  constexpr uint32_t kIdescO = tc::make_idesc(0, 128, 64);

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, FA_SQ);
      tc::tma_load_2d(&tmQ, q_full, sQ, 0, bh * N1 + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        tc::mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        tc::mbar_expect_tx(&kv_full[s], FA_STAGE);
        uint8_t* st = sKV + s * FA_STAGE;
        tc::tma_load_3d(&tmK, &kv_full[s], st, 0, j * FA_BKV, bh);
#pragma unroll
        for (int kb = 0; kb < FA_NSUB; ++kb)
          tc::tma_load_3d(&tmV, &kv_full[s], st + FA_SK + kb * FA_D * 128, j * FA_BKV + kb * 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    auto mma_S = [&](int s) {   // S = Q K^T : 4 k-steps of 16 over d = 64
      const uint32_t a = tc::smem_u32(sQ), b = tc::smem_u32(sKV + s * FA_STAGE);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tc::mma_ss<false>(tmem_S, tc::smem_desc_sw128(a + ks * 32), tc::smem_desc_sw128(b + ks * 32), kIdescS, ks ? 1u : 0u);
    };
    auto mma_O = [&](int s) {   // O_j = P V : 2 sub-tiles x 4 k-steps over 128 keys
      const uint32_t a = tc::smem_u32(sP), b = tc::smem_u32(sKV + s * FA_STAGE + FA_SK);
#pragma unroll
      for (int kb = 0; kb < FA_NSUB; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc::mma_ss<false>(tmem_O, tc::smem_desc_sw128(a + kb * (FA_BQ * 128) + ks * 32),
                            tc::smem_desc_sw128(b + kb * (FA_D * 128) + ks * 32), kIdescO, (kb | ks) ? 1u : 0u);
    };
    tc::mbar_wait(q_full, 0);
    tc::mbar_wait(&kv_full[0], 0);
    tc::fence_after_sync();
    if (tc::elect_one()) { mma_S(0); tc::mma_commit(s_full); }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      if (j + 1 < n_kv) tc::mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
      tc::mbar_wait(p_ready, j & 1);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        mma_O(s);
        tc::mma_commit(o_full);
        tc::mma_commit(&kv_empty[s]);
        if (j + 1 < n_kv) { mma_S((j + 1) & 1); tc::mma_commit(s_full); }
      }
      __syncwarp();
    }
  } else {
    // ---------------- online softmax: thread = query row ----------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                 // row inside the Q tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o[FA_D];
#pragma unroll
    for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int kbase = j * FA_BKV;
      tc::mbar_wait(s_full, j & 1);
      tc::fence_after_sync();
      const bool tail = kbase + FA_BKV > N1;          // only the last key tile needs masking
      // pass 1: row maximum over the valid keys of this tile
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < FA_BKV; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_S + lane_addr + c, v);
        tc::tmem_ld_wait();
        if (!tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c + i < N1) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      // fold in the previous tile's P V (computed relative to the current running max)
      if (j > 0) {
        tc::mbar_wait(o_full, (j - 1) & 1);
        tc::fence_after_sync();
#pragma unroll
        for (int c = 0; c < FA_D; c += 32) {
          uint32_t v[32];
          tc::tmem_ld32(tmem_O + lane_addr + c, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2(m_run - m_new);      // m_run = -inf on the first tile -> 0
      l_run *= alpha;
#pragma unroll
      for (int i = 0; i < FA_D; ++i) o[i] *= alpha;
      m_run = m_new;
      // pass 2: p = exp(s - m), row sum, fp16 P tile in the 128B-swizzled K-major layout of the MMA A operand
#pragma unroll
      for (int c = 0; c < FA_BKV; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_S + lane_addr + c, v);
        tc::tmem_ld_wait();
        uint32_t packed[16];
        float lpart = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = fast_exp2(__uint_as_float(v[i]) - m_new), p1 = fast_exp2(__uint_as_float(v[i + 1]) - m_new);
          if (tail) { if (kbase + c + i >= N1) p0 = 0.f; if (kbase + c + i + 1 >= N1) p1 = 0.f; }
          __half2 h = __floats2half2_rn(p0, p1);
          float2 hf = __half22float2(h);
          lpart += hf.x + hf.y;
          packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        l_run += lpart;
        // 32 keys = 4 chunks of 16 bytes; key (c + 8*cc .. ) -> sub-tile kb = c / 64, chunk ((c % 64) / 8 + cc)
        uint8_t* base = sP + (c >> 6) * (FA_BQ * 128) + row * 128;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int chunk = ((c & 63) >> 3) + cc;
          *reinterpret_cast<uint4*>(base + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(packed[cc * 4], packed[cc * 4 + 1], packed[cc * 4 + 2], packed[cc * 4 + 3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the MMA
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_ready);
    }
    // last tile's P V, normalise, store
    tc::mbar_wait(o_full, (n_kv - 1) & 1);
    tc::fence_after_sync();
#pragma unroll
    for (int c = 0; c < FA_D; c += 32) {
      uint32_t v[32];
      tc::tmem_ld32(tmem_O + lane_addr + c, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
    }
    const int qrow = q0 + row;
    if (qrow < N1) {
      const float inv = 1.f / l_run;
      const int b = bh / fp.heads, hd = bh - b * fp.heads;
      const size_t off = ((size_t)b * N1 + qrow) * fp.D + hd * FA_D;
      if (fp.out_f16) {
        __half* dst = reinterpret_cast<__half*>(fp.out) + off;
#pragma unroll
        for (int i = 0; i < FA_D; i += 8) {
          __half2 h0 = __floats2half2_rn(o[i] * inv, o[i + 1] * inv), h1 = __floats2half2_rn(o[i + 2] * inv, o[i + 3] * inv);
          __half2 h2 = __floats2half2_rn(o[i + 4] * inv, o[i + 5] * inv), h3 = __floats2half2_rn(o[i + 6] * inv, o[i + 7] * inv);
          *reinterpret_cast<uint4*>(dst + i) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                          *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
      } else {
        float* dst = reinterpret_cast<float*>(fp.out) + off;
#pragma unroll
        for (int i = 0; i < FA_D; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, FA_TMEM);
  }
}

}  // namespace dtk
