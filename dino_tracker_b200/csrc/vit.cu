// DINOv2 ViT feature extractor: ImageNet normalisation, 14x14 / stride-7 patch embedding,
// cls + interpolated position embedding, pre-LN blocks (LayerNorm eps 1e-6, MHA scale 1/8, LayerScale, MLP 4x with
// exact GELU), tap = output of block `layer` before the final norm, cls dropped, written straight into the
// token-major feature video [T][P][C]   (models/extractor.py:41-85,137-150; utils.py:32-72; the block arithmetic is
// facebookresearch/dinov2's -- parity unpinned).
//
// Default path (gemm_f16 = 1, attn_materialized = 0): fp16 operands / fp32 accumulation everywhere.  The linear layers
// (patch embedding, qkv, proj, fc1, fc2) run on the tcgen05 GEMMs of tcgemm.cuh / tcgemm2.cuh (CTA pairs when gemm_pair = 1)
// with bias / position embedding / GELU / LayerScale + residual / head scatter as coalesced epilogues on the accumulator;
// the residual stream stays fp32.  Attention is the fused kernel of flash.cuh (scores never leave the SM).
// Validation path (attn_materialized = 1): TF32 GEMMs, attention scores materialised per (frame, row chunk) in a workspace
// (tensor-core GEMM -> row softmax -> tensor-core GEMM).
#include "common.cuh"
#include "corr.cuh"
#include "tcgemm.cuh"
#include "flash.cuh"
#include "tcgemm2.cuh"

namespace dtk {

constexpr int HD = 64;  // head dim of every DINOv2 ViT

// ---------------------------------------------------------------------------------------------- small kernels
// patches of ImageNet-normalised frames: out[(b*P + p)][c*196 + ky*14 + kx], row length Kp (zero padded)
template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }

template <typename OutT>
__global__ void vit_im2col_kernel(const float* __restrict__ frames, OutT* __restrict__ out, int B, int H, int W, int h,
                                  int w, int patch, int stride, int Kp) {
  const size_t row = blockIdx.x;  // b*P + p
  const int P = h * w;
  const int b = (int)(row / P), p = (int)(row - (size_t)b * P);
  const int py = (p / w) * stride, px = (p % w) * stride;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const int K = 3 * patch * patch;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float v = 0.f;
    if (k < K) {
      int c = k / (patch * patch), r = k - c * patch * patch;
      int ky = r / patch, kx = r - ky * patch;
      float x = frames[(((size_t)b * 3 + c) * H + py + ky) * W + px + kx];
      v = __fdiv_rn(__fsub_rn(x, mean[c]), stdv[c]);  // torchvision Normalize: (x - mean) / std
    }
    out[row * Kp + k] = cvt_out<OutT>(v);
  }
}

__global__ void vit_cls_kernel(float* __restrict__ x, const float* __restrict__ cls_pos, int N1, int D) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += blockDim.x) x[(size_t)b * N1 * D + i] = cls_pos[i];
}

// warp per row, D <= 2048
template <typename OutT>
__global__ void vit_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gw, const float* __restrict__ gb,
                                     OutT* __restrict__ y, size_t rows, int D) {
  const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[16];
  const int n4 = D >> 2;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int idx = lane + 32 * i;
    if (idx < n4) { v[i] = xr[idx]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
  }
  const float mu = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int idx = lane + 32 * i;
    if (idx < n4) {
      float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-6f);
  OutT* yr = y + row * D;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int idx = lane + 32 * i;
    if (idx < n4) {
      float4 w4 = __ldg(reinterpret_cast<const float4*>(gw) + idx), b4 = __ldg(reinterpret_cast<const float4*>(gb) + idx);
      float4 o;
      o.x = (v[i].x - mu) * rstd * w4.x + b4.x; o.y = (v[i].y - mu) * rstd * w4.y + b4.y;
      o.z = (v[i].z - mu) * rstd * w4.z + b4.z; o.w = (v[i].w - mu) * rstd * w4.w + b4.w;
      if constexpr (sizeof(OutT) == 4) {
        reinterpret_cast<float4*>(yr)[idx] = o;
      } else {
        __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
        reinterpret_cast<uint2*>(yr)[idx] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
    }
  }
}

// in-place row softmax over n columns (row stride ld); block per row
__global__ void vit_softmax_kernel(float* __restrict__ s, int n, int ld, size_t group_stride) {
  extern __shared__ float row[];
  __shared__ float red[32];
  float* p = s + (size_t)blockIdx.y * group_stride + (size_t)blockIdx.x * ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { float v = p[i]; row[i] = v; m = fmaxf(m, v); }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int k = 1; k < nw; ++k) m = fmaxf(m, red[k]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { float e = expf(row[i] - m); row[i] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int k = 0; k < nw; ++k) sum += red[k];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = row[i] * inv;
}

// x[b][1 + p][:] -> tpc[b][p][:]
__global__ void vit_tap_kernel(const float* __restrict__ x, float* __restrict__ tpc, int B, int P, int D) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  const int n4 = D >> 2;
  size_t total = (size_t)B * P * n4;
  if (i >= total) return;
  size_t row = i / n4;
  int c = (int)(i - row * n4);
  size_t b = row / P, p = row - b * P;
  reinterpret_cast<float4*>(tpc)[i] = reinterpret_cast<const float4*>(x)[((b * (P + 1) + 1 + p)) * n4 + c];
}

// group tables for the GEMMs: kind 0: one group of m rows; kind 1: `heads` groups (attention)
__global__ void vit_plan_kernel(int* batch, int* row0, int* m, int* tile_start, int n_groups, int rows, int row_stride,
                                int row_base, int batch_base, int tile_rows) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int g = 0; g < n_groups; ++g) {
      batch[g] = batch_base + g; row0[g] = row_base + g * row_stride; m[g] = rows; tile_start[g] = acc;
      acc += (rows + tile_rows - 1) / tile_rows;
    }
    tile_start[n_groups] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- epilogues
struct EpiBase {
  struct State {};
  __device__ __forceinline__ void tile_begin(State&) const {}
  __device__ __forceinline__ void tile_end(State&, int, int, int) const {}
  __device__ __forceinline__ bool direct(int) const { return false; }
};
// r / d for 0 <= r < 2^24 without the ~40-instruction integer division (the epilogues do it per 4 output values)
__device__ __forceinline__ int fast_div(int r, int d) {
  int q = __float2int_rd(__int2float_rn(r) * __frcp_rn(__int2float_rn(d)));
  q += ((q + 1) * d <= r) ? 1 : 0;
  q -= (q * d > r) ? 1 : 0;
  return q;
}
__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
  __half2 x = __floats2half2_rn(a, b), y = __floats2half2_rn(c, d);
  return make_uint2(*reinterpret_cast<uint32_t*>(&x), *reinterpret_cast<uint32_t*>(&y));
}

// tokens: x[b][1 + p][col] = acc + bias[col] + pos[p][col]   (row r = b*P + p)
struct EpiPatch : EpiBase {
  static constexpr bool kCoalesced = true;
  float* x; const float* bias; const float* pos; int P, D;
  __device__ __forceinline__ void vec4(int, int r, int col, float4 v) const {
    const int b = fast_div(r, P), p = r - b * P;
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col)), pp = __ldg(reinterpret_cast<const float4*>(pos + (size_t)p * D + col));
    *reinterpret_cast<float4*>(x + ((size_t)b * (P + 1) + 1 + p) * D + col) =
        make_float4(v.x + bb.x + pp.x, v.y + bb.y + pp.y, v.z + bb.z + pp.z, v.w + bb.w + pp.w);
  }
  __device__ __forceinline__ void operator()(State&, int, int r, int col0, const float (&f)[32], int ncols) const {
    const int b = r / P, p = r - b * P;
    float* o = x + ((size_t)b * (P + 1) + 1 + p) * D + col0;
    const float* ps = pos + (size_t)p * D + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      if (i < ncols) {   // ncols is a multiple of 4 (D % 64 == 0)
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0 + i)), pp = __ldg(reinterpret_cast<const float4*>(ps + i));
        *reinterpret_cast<float4*>(o + i) = make_float4(f[i] + bb.x + pp.x, f[i + 1] + bb.y + pp.y, f[i + 2] + bb.z + pp.z, f[i + 3] + bb.w + pp.w);
      }
  }
};

// qkv: scatter to q [b][hd][n][64] (scaled 1/8), k [b][hd][n][64], vT [b][hd][64][n]
struct EpiQKV : EpiBase {
  float* q; float* k; float* vT; const float* bias; int N1, D, heads, N1p;
  __device__ __forceinline__ void operator()(State&, int, int r, int col0, const float (&f)[32], int ncols) const {
    const int b = r / N1, n = r - b * N1;
    const int which = col0 / D, c = col0 - which * D, hd = c / HD, e0 = c - hd * HD;
    const size_t bh = (size_t)b * heads + hd;
    if (which < 2) {
      float* o = (which == 0 ? q : k) + (bh * N1 + n) * HD + e0;
      const float sc = which == 0 ? 0.125f : 1.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = (f[i] + __ldg(bias + col0 + i)) * sc;
    } else {
      float* o = vT + (bh * HD + e0) * N1p + n;
#pragma unroll
      for (int i = 0; i < 32; ++i) if (i < ncols) o[(size_t)i * N1p] = f[i] + __ldg(bias + col0 + i);
    }
  }
};

// qkv for the fused attention: fp16 q [b][hd][n][64] (scaled 1/8), k [b][hd][n][64], vT [b][hd][64][n] (pitch N1p)
struct EpiQKV16 : EpiBase {
  static constexpr bool kCoalesced = true;
  __half* q; __half* k; __half* vT; const float* bias; int N1, D, heads, N1p; float qscale;
  int direct_from = 1 << 30;   // columns >= direct_from take the thread-per-row call (host: 2 * D, or 0 = every column)
  // v goes out transposed ([d][n]): thread-per-row already writes consecutive n per lane -> keep the direct call there
  __device__ __forceinline__ bool direct(int col0) const { return col0 >= direct_from; }
  __device__ __forceinline__ void vec4(int, int r, int col, float4 v) const {
    const int b = fast_div(r, N1), n = r - b * N1;
    const int which = (col >= D) + (col >= 2 * D), c = col - which * D, hd = c / HD, e0 = c - hd * HD;
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col));
    const float sc = which == 0 ? qscale : 1.f;
    __half* o = (which == 0 ? q : k) + (((size_t)b * heads + hd) * N1 + n) * HD + e0;
    *reinterpret_cast<uint2*>(o) = pack_half4((v.x + bb.x) * sc, (v.y + bb.y) * sc, (v.z + bb.z) * sc, (v.w + bb.w) * sc);
  }
  __device__ __forceinline__ void operator()(State&, int, int r, int col0, const float (&f)[32], int ncols) const {
    const int b = r / N1, n = r - b * N1;
    const int which = col0 / D, c = col0 - which * D, hd = c / HD, e0 = c - hd * HD;
    const size_t bh = (size_t)b * heads + hd;
    if (which < 2) {
      __half* o = (which == 0 ? q : k) + (bh * N1 + n) * HD + e0;
      const float sc = which == 0 ? qscale : 1.f;
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        if (i < ncols) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + i)), b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + i + 4));
          __half2 h0 = __floats2half2_rn((f[i] + b0.x) * sc, (f[i + 1] + b0.y) * sc), h1 = __floats2half2_rn((f[i + 2] + b0.z) * sc, (f[i + 3] + b0.w) * sc);
          __half2 h2 = __floats2half2_rn((f[i + 4] + b1.x) * sc, (f[i + 5] + b1.y) * sc), h3 = __floats2half2_rn((f[i + 6] + b1.z) * sc, (f[i + 7] + b1.w) * sc);
          *reinterpret_cast<uint4*>(o + i) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                        *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
    } else {
      __half* o = vT + (bh * HD + e0) * N1p + n;
#pragma unroll
      for (int i = 0; i < 32; ++i) if (i < ncols) o[(size_t)i * N1p] = __float2half_rn(f[i] + __ldg(bias + col0 + i));
    }
  }
};

// plain store: out[(g * rows_per_group + r)][col] (attention scores)
struct EpiStore : EpiBase {
  float* out; int ld, rows_per_group;
  __device__ __forceinline__ void operator()(State&, int g, int r, int col0, const float (&f)[32], int ncols) const {
    float* o = out + ((size_t)g * rows_per_group + r) * ld + col0;
    if (ncols == 32) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = f[i];
    }
  }
};

// p.v: group g = head; out[(frame_row0 + r)][g*64 + col]
struct EpiPV : EpiBase {
  float* out; size_t frame_row0; int D;
  __device__ __forceinline__ void operator()(State&, int g, int r, int col0, const float (&f)[32], int ncols) const {
    float* o = out + (frame_row0 + r) * D + g * HD + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      if (i < ncols) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
  }
};

// x[r][col] += ls[col] * (acc + bias[col])
struct EpiResidual : EpiBase {
  static constexpr bool kCoalesced = true;
  static constexpr bool kPrefetch = true;
  float* x; const float* bias; const float* ls; int D;
  __device__ __forceinline__ float4 fetch(int, int r, int col) const { return *reinterpret_cast<const float4*>(x + (size_t)r * D + col); }
  __device__ __forceinline__ void vec4(int, int r, int col, float4 v, float4 xv) const {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col)), ll = __ldg(reinterpret_cast<const float4*>(ls + col));
    xv.x += (v.x + bb.x) * ll.x; xv.y += (v.y + bb.y) * ll.y; xv.z += (v.z + bb.z) * ll.z; xv.w += (v.w + bb.w) * ll.w;
    *reinterpret_cast<float4*>(x + (size_t)r * D + col) = xv;
  }
  __device__ __forceinline__ void vec4(int, int r, int col, float4 v) const {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col)), ll = __ldg(reinterpret_cast<const float4*>(ls + col));
    float4* o = reinterpret_cast<float4*>(x + (size_t)r * D + col);
    float4 xv = *o;
    xv.x += (v.x + bb.x) * ll.x; xv.y += (v.y + bb.y) * ll.y; xv.z += (v.z + bb.z) * ll.z; xv.w += (v.w + bb.w) * ll.w;
    *o = xv;
  }
  __device__ __forceinline__ void operator()(State&, int, int r, int col0, const float (&f)[32], int ncols) const {
    float* o = x + (size_t)r * D + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      if (i < ncols) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0 + i)), ll = __ldg(reinterpret_cast<const float4*>(ls + col0 + i));
        float4 xv = *reinterpret_cast<const float4*>(o + i);
        xv.x += (f[i] + bb.x) * ll.x; xv.y += (f[i + 1] + bb.y) * ll.y; xv.z += (f[i + 2] + bb.z) * ll.z; xv.w += (f[i + 3] + bb.w) * ll.w;
        *reinterpret_cast<float4*>(o + i) = xv;
      }
  }
};

// 0.5 x (1 + erf(x / sqrt 2)) with erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, two MUFU + 9 FMA-pipe
// instructions, branch-free): libdevice's erff is ~3x the instructions and made the fc1 epilogue, not its MMAs, pace that
// GEMM.  The result is rounded to fp16 (relative 4.9e-4) right after.
__device__ __forceinline__ float gelu_exact(float v) {
  const float z = fabsf(v) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));      // one MUFU (1 ulp: below the 1.5e-7 of the fit)
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = fast_exp2(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p * t, e, 1.f);          // erf(|v| / sqrt 2)
  return 0.5f * v + 0.5f * fabsf(v) * erf_abs;         // 0.5 v (1 + sign(v) erf(|v| / sqrt 2))
}

// h[r][col] = gelu(acc + bias[col])   (exact: 0.5 x (1 + erf(x / sqrt 2)))
template <typename OutT>
struct EpiGelu : EpiBase {
  static constexpr bool kCoalesced = true;
  OutT* h; const float* bias; int ld;
  int all_direct = 0;
  __device__ __forceinline__ bool direct(int) const { return all_direct != 0; }
  __device__ __forceinline__ void vec4(int, int r, int col, float4 v) const {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col));
    float t[4] = {v.x + bb.x, v.y + bb.y, v.z + bb.z, v.w + bb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = sizeof(OutT) == 4 ? 0.5f * t[j] * (1.f + erff(t[j] * 0.70710678118654752f)) : gelu_exact(t[j]);
    if constexpr (sizeof(OutT) == 4) *reinterpret_cast<float4*>(h + (size_t)r * ld + col) = make_float4(t[0], t[1], t[2], t[3]);
    else *reinterpret_cast<uint2*>(h + (size_t)r * ld + col) = pack_half4(t[0], t[1], t[2], t[3]);
  }
  __device__ __forceinline__ void operator()(State&, int, int r, int col0, const float (&f)[32], int ncols) const {
    OutT* o = h + (size_t)r * ld + col0;
    float gl[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0 + (i < ncols ? i : 0)));
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = f[i + j] + bv[j];
        gl[i + j] = sizeof(OutT) == 4 ? 0.5f * v * (1.f + erff(v * 0.70710678118654752f)) : gelu_exact(v);
      }
    }
    if constexpr (sizeof(OutT) == 4) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < ncols) *reinterpret_cast<float4*>(o + i) = make_float4(gl[i], gl[i + 1], gl[i + 2], gl[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        if (i < ncols) {
          __half2 h0 = __floats2half2_rn(gl[i], gl[i + 1]), h1 = __floats2half2_rn(gl[i + 2], gl[i + 3]);
          __half2 h2 = __floats2half2_rn(gl[i + 4], gl[i + 5]), h3 = __floats2half2_rn(gl[i + 6], gl[i + 7]);
          *reinterpret_cast<uint4*>(o + i) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                        *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
    }
  }
};

struct Plan { int* batch; int* row0; int* m; int* tile_start; };

template <class Epi, int BN, TcMode MODE = TcMode::TF32>
static int run_gemm(const void* A, uint64_t a_rows, const void* Bm, uint64_t b_batch, uint64_t b_rows, int K,
                    const Plan& pl, int n_groups, int max_tiles, const Epi& epi, int prof_cls, cudaStream_t st,
                    uint64_t ld = 0) {
  using Cfg = TcCfg<MODE, BN>;
  constexpr int kT = MODE == TcMode::F16 ? TMAP_F16 : TMAP_F32;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d(&tmA, A, a_rows, K, TC_BM, Cfg::kBK, kT, ld))) return rc;
  if ((rc = make_tmap_3d(&tmB, Bm, b_batch, b_rows, K, BN, Cfg::kBK, kT, ld))) return rc;
  auto kern = tc_gemm_kernel<MODE, Epi, BN>;
  static PerDev<bool> attr_dev;  // one static per template instantiation
  bool& attr = attr_dev.get();
  if (!attr) {
    DTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr = true;
  }
  TcProblem pb{pl.batch, pl.row0, pl.m, pl.tile_start, n_groups, (int)b_rows, K};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int tiles = max_tiles * cdiv((int)b_rows, BN);
  int grid = tiles < sms ? tiles : sms;
  ProfRange pr(prof_cls, st);
  kern<<<grid, TC_THREADS, Cfg::kSmem, st>>>(tmA, tmA, tmB, tmB, pb, epi);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

static int plan(const Plan& pl, int n_groups, int rows, int row_stride, int row_base, int batch_base, cudaStream_t st,
                int tile_rows = TC_BM) {
  ProfRange pr(PROF_VIT_MISC, st);
  vit_plan_kernel<<<1, 32, 0, st>>>(pl.batch, pl.row0, pl.m, pl.tile_start, n_groups, rows, row_stride, row_base, batch_base,
                                    tile_rows);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

// CTA-pair variant (cta_group::2, 256 x 256 tiles) for the single-pass fp16 linear layers; the plan must be in
// 256-row tiles.
template <class Epi>
static int run_gemm_pair(const void* A, uint64_t a_rows, const void* Bm, uint64_t b_rows, int K, const Plan& pl,
                         int max_tiles, const Epi& epi, int prof_cls, cudaStream_t st) {
  // 8 epilogue warps (two per TMEM lane quadrant): the fused epilogues (GELU, LayerScale + residual, head scatter) run on
  // warps that have their scheduler to themselves, so their latency chains, not the MMAs, paced these GEMMs with 4 warps
  constexpr int kEpiWarps = 8;
  using Base = TcCfg<TcMode::F16, TC2_BN>;
  using Cfg = Tc2Cfg<TcMode::F16, kEpiWarps, true>;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d(&tmA, A, a_rows, K, 128, Base::kBK, TMAP_F16))) return rc;
  if ((rc = make_tmap_3d(&tmB, Bm, 1, b_rows, K, TC2_BN / 2, Base::kBK, TMAP_F16))) return rc;
  auto kern = tc_gemm2_kernel<TcMode::F16, Epi, kEpiWarps>;
  static PerDev<bool> attr_dev;
  bool& attr = attr_dev.get();
  if (!attr) {
    DTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr = true;
  }
  TcProblem pb{pl.batch, pl.row0, pl.m, pl.tile_start, 1, (int)b_rows, K};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = max_tiles * cdiv((int)b_rows, TC2_BN);
  int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);
  ProfRange pr(prof_cls, st);
  kern<<<grid, 64 + 32 * kEpiWarps, Cfg::kSmem, st>>>(tmA, tmA, tmB, tmB, pb, epi);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

constexpr int VIT_ROW_CHUNK = 1024;  // query rows per attention-score chunk

// fused attention over fp16 q [B*heads][N1][64] (pre-scaled by 64^-1/2 * log2 e), k [B*heads][N1][64] and
// v^T [B*heads][64][N1p] (row pitch N1p, a multiple of 8): out[b*N1 + n][h*64 + e], row pitch D, fp32 or fp16
static int launch_flash(const __half* q16, const __half* k16, const __half* v16, int B, int heads, int N1, int N1p, int D,
                        void* out, bool out_f16, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, q16, (uint64_t)B * heads * N1, HD, FA_BQ, HD, TMAP_F16))) return rc;
  if ((rc = make_tmap_3d(&tmK, k16, (uint64_t)B * heads, N1, HD, FA_BKV, HD, TMAP_F16))) return rc;
  if ((rc = make_tmap_3d(&tmV, v16, (uint64_t)B * heads, HD, N1, HD, 64, TMAP_F16, (uint64_t)N1p))) return rc;
  // share of the exponentials evaluated on the FMA pipe (DTK_FA_POLY = 0 / 25 / 37 / 50 %, default 25)
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("DTK_FA_POLY");
    // measured on ViT-L (8108 tokens, 2 frames x 16 blocks): 0 % 11.3 ms, 25 % 10.2 ms, 37 % 10.3 ms, 50 % 10.9 ms -- since
    // P goes through tensor memory the softmax warps are MUFU-bound enough for the packed polynomial to pay
    poly = e ? atoi(e) : 25;
    DTK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    DTK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<0x88>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    DTK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<0xA8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    DTK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<0xAA>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
  }
  FlashParams fpar{N1, D, heads, out, out_f16 ? 1 : 0};
  ProfRange pr(PROF_VIT_ATTN, st);
  const dim3 fgrid(cdiv(N1, FA_BQ), B * heads);
  if (poly <= 0) flash_attn_kernel<0><<<fgrid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, fpar);
  else if (poly <= 25) flash_attn_kernel<0x88><<<fgrid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, fpar);
  else if (poly <= 37) flash_attn_kernel<0xA8><<<fgrid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, fpar);
  else flash_attn_kernel<0xAA><<<fgrid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, fpar);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

}  // namespace dtk

using namespace dtk;

extern "C" {

// row length of the im2col matrix / patch weight: 16-byte multiple of the operand type
static int vit_kp(const dinotrk_vit_config* c) {
  const bool f16 = c->gemm_f16 != 0 && c->attn_materialized == 0;
  return (int)align_up((size_t)3 * c->patch * c->patch, f16 ? 8 : 4);
}

size_t dinotrk_vit_workspace_bytes(const dinotrk_vit_config* c, const dinotrk_geom* g, int B) {
  if (!c || !g) return 0;
  const size_t P = (size_t)g->h * g->w, N1 = P + 1, D = c->dim;
  size_t b = 0;
  const size_t N1p = align_up(N1, 4);
  b += align_up(B * N1 * D * 4, 256) * 2;                                  // x, y
  b += align_up(B * N1 * D * 4, 256) * 2 + align_up(B * N1p * D * 4, 256); // q, k, vT
  size_t hid = B * N1 * 4 * D * 4, col = B * P * (size_t)vit_kp(c) * 4;
  b += align_up(hid > col ? hid : col, 256);                               // MLP hidden / im2col (aliased)
  b += align_up((size_t)c->heads * VIT_ROW_CHUNK * N1p * 4, 256);          // attention scores of one row chunk
  b += 4 * align_up((size_t)(c->heads + 2) * 4, 256) + 4096;               // plan tables
  return b;
}

int dinotrk_vit_attention(const void* q16, const void* k16, const void* vT16, int B, int heads, int N1, int N1p,
                          float* out, void* stream) {
  DTK_CHECK_ARG(q16 && k16 && vT16 && out, "vit_attention: null pointer");
  DTK_CHECK_ARG(B > 0 && heads > 0 && N1 > 0 && N1p >= N1 && N1p % 8 == 0, "vit_attention: bad sizes");
  return launch_flash(reinterpret_cast<const __half*>(q16), reinterpret_cast<const __half*>(k16),
                      reinterpret_cast<const __half*>(vT16), B, heads, N1, N1p, heads * HD, out, false, (cudaStream_t)stream);
}

int dinotrk_vit_forward(const float* frames, int B, const dinotrk_geom* g, const dinotrk_vit_config* c,
                        const dinotrk_vit_weights* wt, float* out_tpc, void* workspace, size_t workspace_bytes,
                        void* stream) {
  NvtxRange nvtx_range("dinotrk.vit_forward");
  DTK_CHECK_ARG(frames && g && c && wt && out_tpc && wt->blocks, "vit_forward: null pointer");
  const int D = c->dim, heads = c->heads, P = g->h * g->w, N1 = P + 1, Kp = vit_kp(c);
  DTK_CHECK_ARG(D == heads * HD && D % 64 == 0 && D <= 2048, "vit_forward: dim must be heads x 64 (<= 2048)");
  DTK_CHECK_ARG(c->tap_layer >= 0 && c->tap_layer < c->depth, "vit_forward: tap layer out of range");
  const int N1p = (int)align_up((size_t)N1, 4);   // row pitch of the score / v^T arrays (TMA strides are 16-byte multiples)
  DTK_CHECK_ARG(workspace && workspace_bytes >= dinotrk_vit_workspace_bytes(c, g, B), "vit_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes);
  const size_t rows = (size_t)B * N1;
  float* x = ar.take<float>(rows * D);
  float* y = ar.take<float>(rows * D);
  float* q = ar.take<float>(rows * D);
  float* k = ar.take<float>(rows * D);
  float* vT = ar.take<float>((size_t)B * N1p * D);
  size_t hid = rows * 4 * D, col = (size_t)B * P * Kp;
  float* hbuf = ar.take<float>(hid > col ? hid : col);
  float* S = ar.take<float>((size_t)heads * VIT_ROW_CHUNK * N1p);
  Plan pl{ar.take<int>(heads + 2), ar.take<int>(heads + 2), ar.take<int>(heads + 2), ar.take<int>(heads + 2)};
  DTK_CHECK_ARG(ar.ok(), "vit_forward: workspace arena overflow");
  int rc;

  // fp16 operand mode (default): LayerNorm / GELU / attention write fp16 activations, weights are fp16 (11-bit
  // significand like TF32, twice the tensor rate, half the operand traffic).  The validation path
  // (attn_materialized) keeps every operand fp32 / TF32.
  const bool f16 = c->gemm_f16 != 0 && c->attn_materialized == 0;
  const bool pairs = f16 && c->gemm_pair != 0;   // linear layers on CTA pairs (cta_group::2)
  const int pair_tiles = cdiv((int)((size_t)B * (g->h * g->w + 1)), TC2_BM);
  __half* y16 = reinterpret_cast<__half*>(y);
  __half* h16 = reinterpret_cast<__half*>(hbuf);

  // ---- patch embedding + cls + position embedding
  {
    ProfRange pr(PROF_VIT_MISC, st);
    if (f16) vit_im2col_kernel<__half><<<B * P, 128, 0, st>>>(frames, h16, B, g->H, g->W, g->h, g->w, c->patch, c->stride, Kp);
    else vit_im2col_kernel<float><<<B * P, 128, 0, st>>>(frames, hbuf, B, g->H, g->W, g->h, g->w, c->patch, c->stride, Kp);
    DTK_LAUNCHED();
  }
  if ((rc = plan(pl, 1, B * P, 0, 0, 0, st))) return rc;
  {
    EpiPatch ep{{}, x, wt->patch_b, wt->pos, P, D};
    rc = f16 ? run_gemm<EpiPatch, 256, TcMode::F16>(h16, (uint64_t)B * P, wt->patch_w, 1, D, Kp, pl, 1, cdiv(B * P, TC_BM), ep, PROF_VIT_GEMM, st)
             : run_gemm<EpiPatch, 256>(hbuf, (uint64_t)B * P, wt->patch_w, 1, D, Kp, pl, 1, cdiv(B * P, TC_BM), ep, PROF_VIT_GEMM, st);
    if (rc) return rc;
  }
  {
    ProfRange pr(PROF_VIT_MISC, st);
    vit_cls_kernel<<<B, 256, 0, st>>>(x, wt->cls_pos, N1, D);
    DTK_LAUNCHED();
  }

  const int all_tiles = cdiv((int)rows, TC_BM);
  auto layernorm = [&](const float* gw, const float* gb) -> int {
    ProfRange pr(PROF_VIT_MISC, st);
    if (f16) vit_layernorm_kernel<__half><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, gw, gb, y16, rows, D);
    else vit_layernorm_kernel<float><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, gw, gb, y, rows, D);
    DTK_LAUNCHED();
    return DINOTRK_OK;
  };
  for (int l = 0; l <= c->tap_layer; ++l) {
    const float* const* w = wt->blocks + (size_t)l * 14;
    // w: 0 norm1.w 1 norm1.b 2 qkv.w 3 qkv.b 4 proj.w 5 proj.b 6 ls1 7 norm2.w 8 norm2.b 9 fc1.w 10 fc1.b 11 fc2.w 12 fc2.b 13 ls2
    // (the four weight matrices are fp16 arrays in fp16 operand mode)
    if ((rc = layernorm(w[0], w[1]))) return rc;
    if ((rc = plan(pl, 1, (int)rows, 0, 0, 0, st, pairs ? TC2_BM : TC_BM))) return rc;
    if (c->attn_materialized == 0) {
      // fused attention: fp16 q / k / v^T, scores stay in TMEM / shared memory
      __half* q16 = reinterpret_cast<__half*>(q);
      __half* k16 = reinterpret_cast<__half*>(k);
      __half* v16 = reinterpret_cast<__half*>(vT);
      const int N1p8 = (int)align_up((size_t)N1, 8);
      EpiQKV16 eq{{}, q16, k16, v16, w[3], N1, D, heads, N1p8, 0.125f * 1.4426950408889634f};  // 1/sqrt(64) * log2(e)
      static const int epi_direct = getenv("DTK_EPI_DIRECT") ? atoi(getenv("DTK_EPI_DIRECT")) : 2;   // bit 0: q / k thread-per-row too (slower)
      eq.direct_from = (epi_direct & 1) ? 0 : 2 * D;
      rc = pairs ? run_gemm_pair<EpiQKV16>(y16, rows, w[2], 3 * D, D, pl, pair_tiles, eq, PROF_VIT_GEMM, st)
         : f16 ? run_gemm<EpiQKV16, 256, TcMode::F16>(y16, rows, w[2], 1, 3 * D, D, pl, 1, all_tiles, eq, PROF_VIT_GEMM, st)
               : run_gemm<EpiQKV16, 256>(y, rows, w[2], 1, 3 * D, D, pl, 1, all_tiles, eq, PROF_VIT_GEMM, st);
      if (rc) return rc;
      if ((rc = launch_flash(q16, k16, v16, B, heads, N1, N1p8, D, f16 ? (void*)y16 : (void*)y, f16, st))) return rc;
    } else {
      if ((rc = run_gemm<EpiQKV, 256>(y, rows, w[2], 1, 3 * D, D, pl, 1, all_tiles, EpiQKV{{}, q, k, vT, w[3], N1, D, heads, N1p},
                                      PROF_VIT_GEMM, st))) return rc;
      // attention, per frame and chunk of query rows: S = q k^T (all heads) -> softmax -> y = S v
      for (int b = 0; b < B; ++b) {
        for (int c0 = 0; c0 < N1; c0 += VIT_ROW_CHUNK) {
          const int rc_rows = N1 - c0 < VIT_ROW_CHUNK ? N1 - c0 : VIT_ROW_CHUNK;
          if ((rc = plan(pl, heads, rc_rows, N1, (b * heads) * N1 + c0, b * heads, st))) return rc;
          if ((rc = run_gemm<EpiStore, 256>(q, rows * heads, k, (uint64_t)B * heads, N1, HD, pl, heads,
                                            heads * cdiv(rc_rows, TC_BM), EpiStore{{}, S, N1p, VIT_ROW_CHUNK},
                                            PROF_VIT_ATTN, st))) return rc;
          {
            ProfRange pr(PROF_VIT_ATTN, st);  // rows of S live at (head * VIT_ROW_CHUNK + r)
            vit_softmax_kernel<<<dim3(rc_rows, heads), 256, (size_t)N1 * 4, st>>>(S, N1, N1p, (size_t)VIT_ROW_CHUNK * N1p);
            DTK_LAUNCHED();
          }
          if ((rc = plan(pl, heads, rc_rows, VIT_ROW_CHUNK, 0, b * heads, st))) return rc;
          if ((rc = run_gemm<EpiPV, 64>(S, (uint64_t)heads * VIT_ROW_CHUNK, vT, (uint64_t)B * heads, HD, N1, pl, heads,
                                        heads * cdiv(rc_rows, TC_BM), EpiPV{{}, y, (size_t)b * N1 + c0, D},
                                        PROF_VIT_ATTN, st, (uint64_t)N1p))) return rc;
        }
      }
    }
    if ((rc = plan(pl, 1, (int)rows, 0, 0, 0, st, pairs ? TC2_BM : TC_BM))) return rc;
    {
      EpiResidual er{{}, x, w[5], w[6], D};
      rc = pairs ? run_gemm_pair<EpiResidual>(y16, rows, w[4], D, D, pl, pair_tiles, er, PROF_VIT_GEMM, st)
         : f16 ? run_gemm<EpiResidual, 256, TcMode::F16>(y16, rows, w[4], 1, D, D, pl, 1, all_tiles, er, PROF_VIT_GEMM, st)
               : run_gemm<EpiResidual, 256>(y, rows, w[4], 1, D, D, pl, 1, all_tiles, er, PROF_VIT_GEMM, st);
      if (rc) return rc;
    }
    if ((rc = layernorm(w[7], w[8]))) return rc;
    if (pairs) {
      EpiGelu<__half> eg{{}, h16, w[10], 4 * D};
      static const int epi_direct2 = getenv("DTK_EPI_DIRECT") ? atoi(getenv("DTK_EPI_DIRECT")) : 2;  // bit 1: fp16 GELU rows written thread-per-row (64 B per thread, whole sectors; measured 6.46 -> 6.32 ms per 2 x 16 blocks)
      eg.all_direct = (epi_direct2 & 2) ? 1 : 0;
      if ((rc = run_gemm_pair<EpiGelu<__half>>(y16, rows, w[9], 4 * D, D, pl, pair_tiles, eg, PROF_VIT_GEMM, st))) return rc;
      if ((rc = run_gemm_pair<EpiResidual>(h16, rows, w[11], D, 4 * D, pl, pair_tiles, EpiResidual{{}, x, w[12], w[13], D},
                                           PROF_VIT_GEMM, st))) return rc;
    } else if (f16) {
      if ((rc = run_gemm<EpiGelu<__half>, 256, TcMode::F16>(y16, rows, w[9], 1, 4 * D, D, pl, 1, all_tiles,
                                                            EpiGelu<__half>{{}, h16, w[10], 4 * D}, PROF_VIT_GEMM, st))) return rc;
      if ((rc = run_gemm<EpiResidual, 256, TcMode::F16>(h16, rows, w[11], 1, D, 4 * D, pl, 1, all_tiles,
                                                        EpiResidual{{}, x, w[12], w[13], D}, PROF_VIT_GEMM, st))) return rc;
    } else {
      if ((rc = run_gemm<EpiGelu<float>, 256>(y, rows, w[9], 1, 4 * D, D, pl, 1, all_tiles,
                                              EpiGelu<float>{{}, hbuf, w[10], 4 * D}, PROF_VIT_GEMM, st))) return rc;
      if ((rc = run_gemm<EpiResidual, 256>(hbuf, rows, w[11], 1, D, 4 * D, pl, 1, all_tiles,
                                           EpiResidual{{}, x, w[12], w[13], D}, PROF_VIT_GEMM, st))) return rc;
    }
  }
  {
    ProfRange pr(PROF_VIT_MISC, st);
    size_t tot = (size_t)B * P * (D / 4);
    vit_tap_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, out_tpc, B, P, D);
    DTK_LAUNCHED();
  }
  return DINOTRK_OK;
}

}  // extern "C"
