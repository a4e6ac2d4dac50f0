// Block-cooperative trilinear descriptor sampling from the token-major feature video.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace dtk {

constexpr int SAMPLE_THREADS = 128;

// All SAMPLE_THREADS threads of the block call this with identical arguments.
// out[C] (optional) = sum over the (up to) 8 in-bounds corners, accumulated in ATen's order; norm_out (optional)
// receives |out|_2 (source_embeddings.norm(dim=1), models/tracker.py:164).  Corners whose weight is exactly 0 are not
// read (x + 0 * v = x for finite v: same value).  out_hi / out_lo (optional, [C] fp16 each) receive the split
// out = hi + lo that the tensor-core correlation GEMM consumes (same rounding as split_f16_kernel).
__device__ __forceinline__ void sample_point(const float* __restrict__ tpc, int C, int P, const TriCorners& c,
                                             int frame0, int frame1, float* __restrict__ out,
                                             float* __restrict__ norm_out, __half* __restrict__ out_hi = nullptr,
                                             __half* __restrict__ out_lo = nullptr) {
  const float4* rows[8];
  float wts[8];
#pragma unroll
  for (int z = 0; z < 2; ++z) {
    int f = z == 0 ? frame0 : frame1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bool ok = f >= 0 && c.tok[k] >= 0 && c.wxy[k][z] != 0.f;
      rows[z * 4 + k] = ok ? reinterpret_cast<const float4*>(tpc + ((size_t)f * P + c.tok[k]) * C) : nullptr;
      wts[z * 4 + k] = c.wxy[k][z];
    }
  }
  float sq = 0.f;
  for (int i = threadIdx.x; i < C / 4; i += SAMPLE_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (rows[k] != nullptr) {
        float4 v = __ldg(rows[k] + i);
        acc.x = fmaf(v.x, wts[k], acc.x); acc.y = fmaf(v.y, wts[k], acc.y);
        acc.z = fmaf(v.z, wts[k], acc.z); acc.w = fmaf(v.w, wts[k], acc.w);
      }
    }
    if (out != nullptr) reinterpret_cast<float4*>(out)[i] = acc;
    if (out_hi != nullptr) {
      __half h0 = __float2half_rn(acc.x), h1 = __float2half_rn(acc.y), h2 = __float2half_rn(acc.z), h3 = __float2half_rn(acc.w);
      __half l0 = __float2half_rn(acc.x - __half2float(h0)), l1 = __float2half_rn(acc.y - __half2float(h1));
      __half l2 = __float2half_rn(acc.z - __half2float(h2)), l3 = __float2half_rn(acc.w - __half2float(h3));
      __half2 a = __halves2half2(h0, h1), b = __halves2half2(h2, h3), cc = __halves2half2(l0, l1), d = __halves2half2(l2, l3);
      reinterpret_cast<uint2*>(out_hi)[i] = make_uint2(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b));
      reinterpret_cast<uint2*>(out_lo)[i] = make_uint2(*reinterpret_cast<unsigned*>(&cc), *reinterpret_cast<unsigned*>(&d));
    }
    sq = fmaf(acc.x, acc.x, sq); sq = fmaf(acc.y, acc.y, sq);
    sq = fmaf(acc.z, acc.z, sq); sq = fmaf(acc.w, acc.w, sq);
  }
  if (norm_out != nullptr) {
    __shared__ float red[SAMPLE_THREADS / 32];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < SAMPLE_THREADS / 32; ++k) s += red[k];
      *norm_out = sqrtf(s);
    }
  }
}

}  // namespace dtk
