// Tracker head: arg-max, 2-layer normalised-conv refiner, spatial softmax, disc-masked soft-argmax
// with the numerical-stability fallback (models/networks/tracker_head.py:107-121, :68-98, :100-105;
// conv_norm.py:34-46; data/dataset.py:21-53).
//
// One persistent CTA per SM; one warp per 4-row band of the map, one lane per 4x4-pixel tile
// (band = 32 tiles = 128 columns >= w).  Per hidden channel a lane computes its 16 hidden values from
// the 6x6 input window it keeps in registers, publishes its top/bottom rows to shared memory (the only
// cross-warp traffic), takes the side/corner halo from its lane neighbours by shuffle and
// accumulates the second convolution into 16 registers.  The next map is prefetched with cp.async
// while the current one is being refined.
#include "common.cuh"
#include "corr.cuh"

namespace dtk {

constexpr int HEAD_MAX_W = 128, HEAD_MAX_H = 128;

struct HeadParams {
  int h, w, P, map_stride;
  int stride_px, half_patch, radius2;  // pixel geometry: centre = half_patch + stride * index
  float normW, normH;                  // W - 1, H - 1
  int out_stride, out_mode;
};

__device__ __forceinline__ void cp_async16_head(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

template <typename T>
__device__ __forceinline__ T block_bcast_reduce(T v, T* red, int warp, int lane, int nwarps, T (*op)(T, T)) {
  // warp-level then cross-warp; every thread returns the result
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  T r = red[0];
  for (int k = 1; k < nwarps; ++k) r = op(r, red[k]);
  return r;
}
__device__ __forceinline__ float opmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float opadd(float a, float b) { return a + b; }
// arg-max key: larger value wins, ties -> smaller index (torch.argmax returns the first maximum).
// Values are >= 0 (ReLU'd), so the float bit pattern orders like the value.
__device__ __forceinline__ unsigned long long opkey(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// online-softmax pair (max, sum of exp relative to max)
struct MS { float m, s; };
__device__ __forceinline__ MS ms_merge(MS a, MS b) {
  float m = fmaxf(a.m, b.m);
  float sa = a.m == -INFINITY ? 0.f : a.s * __expf(a.m - m);
  float sb = b.m == -INFINITY ? 0.f : b.s * __expf(b.m - m);
  return MS{m, sa + sb};
}
struct F3 { float a, b, c; };

// kWrap: the band's last lane (tile 31) lies completely outside the map (w <= 124), so the side halo can be
// taken with rotating shuffles (lane 0 reads the all-zero tile 31) -- no edge selects in the hot loop.
template <int MAXT, bool kWrap>
__global__ void __launch_bounds__(MAXT, 1)
head_kernel(const float* __restrict__ maps, int n_maps, HeadParams hp, dinotrk_head_weights wts,
            const int* __restrict__ out_index, float* __restrict__ out, int* __restrict__ aux) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int h = hp.h, w = hp.w, P = hp.P;
  const int lin_elems = (hp.map_stride + 3) & ~3;
  float* sm_lin[2] = {smem, smem + lin_elems};
  float* sm_red = smem + 2 * lin_elems;                   // [6 * 32] floats
  unsigned long long* sm_red64 = reinterpret_cast<unsigned long long*>(sm_red + 192);  // [32]

  const int r0 = warp * 4, c0 = lane * 4;  // this lane's tile
  const int nchunks = hp.map_stride / 4;
  constexpr float kNeg = -1e30f;
  // additive masks: hidden activations of pixels outside the map must be exactly 0 (zero padding of the
  // second convolution); bias + kNeg makes the ReLU do that without per-pixel selects.
  float rmask6[6], cmask[4];
#pragma unroll
  for (int i = 0; i < 6; ++i) rmask6[i] = (r0 - 1 + i >= 0 && r0 - 1 + i < h) ? 0.f : kNeg;
#pragma unroll
  for (int i = 0; i < 4; ++i) cmask[i] = (c0 + i < w) ? 0.f : kNeg;
  const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;

  int map = blockIdx.x;
  if (map < n_maps) {
    const float4* src = reinterpret_cast<const float4*>(maps + (size_t)map * hp.map_stride);
    for (int i = threadIdx.x; i < nchunks; i += blockDim.x) cp_async16_head(sm_lin[0] + 4 * i, src + i);
  }
  asm volatile("cp.async.commit_group;\n" ::);

  for (int it = 0; map < n_maps; map += gridDim.x, ++it) {
    const float* lin = sm_lin[it & 1];
    {  // prefetch the next map into the other buffer
      int nmap = map + gridDim.x;
      if (nmap < n_maps) {
        const float4* src = reinterpret_cast<const float4*>(maps + (size_t)nmap * hp.map_stride);
        float* dst = sm_lin[(it + 1) & 1];
        for (int i = threadIdx.x; i < nchunks; i += blockDim.x) cp_async16_head(dst + 4 * i, src + i);
      }
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 1;\n" ::);
    }
    __syncthreads();

    // ---- arg-max of the (already ReLU'd) map: first maximal index ----------------------------
    unsigned long long key = 0ull;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      float v = lin[i] + 0.f;  // -0.0 -> +0.0 so that the bit pattern orders like the value
      unsigned long long k = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(0x7fffffff - i);
      key = k > key ? k : key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o); key = t > key ? t : key; }
    if (lane == 0) sm_red64[warp] = key;

    // ---- 8x6 input window: rows r0-2..r0+5, cols c0-1..c0+4 (zero outside the map = conv zero padding) --
    float m[8][6];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = r0 - 2 + i;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        int c = c0 - 1 + j;
        m[i][j] = (r >= 0 && r < h && c >= 0 && c < w) ? lin[r * w + c] : 0.f;
      }
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = wts.b2;

    __syncthreads();  // publishes sm_red64 (arg-max partials); no block-wide sync inside the channel loop
    // Each lane computes the hidden layer for its 4 columns on 6 rows (own 4 + the row above and below:
    // 1.5x recompute instead of a shared-memory exchange + barrier per channel), takes the two side columns
    // from its lane neighbours, and folds each hidden row into the (up to 3) output rows it touches.
#pragma unroll 1
    for (int o = 0; o < 16; ++o) {
      float w1[9], w2[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) { w1[k] = wts.w1[o][k]; w2[k] = wts.w2[o][k]; }
      const float b1 = wts.b1[o];
#pragma unroll
      for (int hr = 0; hr < 6; ++hr) {
        float hrow[6];
        const float br = b1 + rmask6[hr];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = br + cmask[j];
#pragma unroll
          for (int ki = 0; ki < 3; ++ki)
#pragma unroll
            for (int kj = 0; kj < 3; ++kj) a = fmaf(w1[ki * 3 + kj], m[hr + ki][j + kj], a);
          hrow[j + 1] = fmaxf(a, 0.f);
        }
        if (kWrap) {
          hrow[0] = __shfl_sync(0xffffffffu, hrow[4], lane_l);
          hrow[5] = __shfl_sync(0xffffffffu, hrow[1], lane_r);
        } else {
          float l = __shfl_up_sync(0xffffffffu, hrow[4], 1);
          float r = __shfl_down_sync(0xffffffffu, hrow[1], 1);
          hrow[0] = lane > 0 ? l : 0.f;
          hrow[5] = lane < 31 ? r : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ki = hr - i;  // hidden row hr is row (i - 1 + ki) of output row i's 3x3 window
          if (ki >= 0 && ki < 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a = acc[i][j];
#pragma unroll
              for (int kj = 0; kj < 3; ++kj) a = fmaf(w2[ki * 3 + kj], hrow[j + kj], a);
              acc[i][j] = a;
            }
          }
        }
      }
    }

    // ---- arg-max result (partials were published before the first barrier) ---------------------
    unsigned long long kbest = sm_red64[0];
    for (int k = 1; k < nwarps; ++k) { unsigned long long t = sm_red64[k]; kbest = t > kbest ? t : kbest; }
    const int amax = 0x7fffffff - (int)(kbest & 0xffffffffu);
    const int arow = amax / w, acol = amax - arow * w;

    // ---- softmax statistics over the whole map: one online (max, sum) reduction --------------------
    MS ms{-INFINITY, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + i < h && c0 + j < w) ms.m = fmaxf(ms.m, acc[i][j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + i < h && c0 + j < w) ms.s += __expf(acc[i][j] - ms.m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      MS t{__shfl_xor_sync(0xffffffffu, ms.m, o), __shfl_xor_sync(0xffffffffu, ms.s, o)};
      ms = ms_merge(ms, t);
    }
    if (lane == 0) { sm_red[warp] = ms.m; sm_red[32 + warp] = ms.s; }
    __syncthreads();
    float zmax = -INFINITY;
    for (int k = 0; k < nwarps; ++k) zmax = fmaxf(zmax, sm_red[k]);

    // ---- disc-masked soft-argmax (mask: |token centre - argmax centre| <= radius px) ----------
    // Only the few threads whose tile meets the disc do any work; exact expf / division here.
    float s = 0.f, sx = 0.f, sy = 0.f, gx = 0.f, gy = 0.f, cnt = 0.f, ssum_part = 0.f;
    {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int dr = (r0 + i - arow) * hp.stride_px, dc = (c0 + j - acol) * hp.stride_px;
          if (r0 + i < h && c0 + j < w && dr * dr + dc * dc <= hp.radius2) {
            float e = expf(acc[i][j] - zmax);
            float x = (float)(hp.half_patch + (c0 + j) * hp.stride_px), y = (float)(hp.half_patch + (r0 + i) * hp.stride_px);
            s += e; sx = fmaf(x, e, sx); sy = fmaf(y, e, sy);
            gx += x; gy += y; cnt += 1.f;
          }
        }
    }
    // global sum of exp(z - zmax) from the per-warp (max, sum) pairs
    for (int k = 0; k < nwarps; ++k) {
      float mk = sm_red[k];
      if (mk != -INFINITY) ssum_part += sm_red[32 + k] * expf(mk - zmax);
    }
    const float ssum = ssum_part;
    // reduce (s, sx, sy, gx, gy, cnt) in one pass
    float vals[6] = {s, sx, sy, gx, gy, cnt};
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vals[q] += __shfl_xor_sync(0xffffffffu, vals[q], o);
    __syncthreads();  // everyone has read sm_red (max/sum pairs) -> reuse it
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 6; ++q) sm_red[q * 32 + warp] = vals[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot[6];
      for (int q = 0; q < 6; ++q) { float t = 0.f; for (int k = 0; k < nwarps; ++k) t += sm_red[q * 32 + k]; tot[q] = t; }
      // p_i = e_i / S_all; s = sum p_i over the disc   (softmax then mask, tracker_head.py:84-86)
      float sp = __fdiv_rn(tot[0], ssum), spx = __fdiv_rn(tot[1], ssum), spy = __fdiv_rn(tot[2], ssum);
      const bool fallback = sp < 1e-8f;
      if (fallback) {  // heatmap <- (heatmap + 1/|mask|) * mask  (tracker_head.py:87-94)
        float u = __fdiv_rn(1.f, tot[5]);
        sp = fmaf(tot[5], u, sp); spx = fmaf(tot[3], u, spx); spy = fmaf(tot[4], u, spy);
      }
      float px = __fdiv_rn(spx, sp), py = __fdiv_rn(spy, sp);
      // RangeNormalizer((W, H)) dst=(-1,1): x / (W-1); * 2; + (-1)      (data/dataset.py:33-35)
      float nx = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(px, hp.normW)), -1.f);
      float ny = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(py, hp.normH)), -1.f);
      if (hp.out_mode == 0) {  // unnormalize(src=(-1,1)): (v - (-1)) / 2 * (W-1)   (data/dataset.py:50-52)
        nx = __fmul_rn(__fdiv_rn(__fadd_rn(nx, 1.f), 2.f), hp.normW);
        ny = __fmul_rn(__fdiv_rn(__fadd_rn(ny, 1.f), 2.f), hp.normH);
      }
      size_t oi = (size_t)(out_index ? out_index[map] : map) * hp.out_stride;
      out[oi] = nx; out[oi + 1] = ny;
      if (aux) { aux[2 * map] = amax; aux[2 * map + 1] = fallback ? 1 : 0; }
    }
    __syncthreads();  // lin / sm_hid / sm_red are reused by the next iteration
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
}

int launch_head(const float* maps, int n_maps, int map_stride, const dinotrk_geom& g,
                const dinotrk_head_weights& hw, const int* out_index, float* out, int out_stride, int out_mode,
                int* aux, cudaStream_t st) {
  if (n_maps <= 0) return DINOTRK_OK;
  DTK_CHECK_ARG(g.w <= HEAD_MAX_W && g.h <= HEAD_MAX_H, "head: token grid %dx%d exceeds the supported %dx%d",
                g.h, g.w, HEAD_MAX_H, HEAD_MAX_W);
  HeadParams hp;
  hp.h = g.h; hp.w = g.w; hp.P = g.h * g.w; hp.map_stride = map_stride;
  hp.stride_px = g.stride; hp.half_patch = g.patch / 2; hp.radius2 = g.radius * g.radius;
  hp.normW = (float)(g.W - 1); hp.normH = (float)(g.H - 1);
  hp.out_stride = out_stride; hp.out_mode = out_mode;
  const int nwarps = cdiv(g.h, 4);
  const int threads = nwarps * 32;
  const int lin_elems = (map_stride + 3) & ~3;
  size_t smem = (size_t)(2 * lin_elems + 6 * 32) * sizeof(float) + 32 * sizeof(unsigned long long);
  static size_t attr_smem[4] = {0, 0, 0, 0};
  const bool wrap = g.w <= 124;                 // tile 31 of every band lies outside the map
  const int variant = (threads <= 544 ? 0 : 2) + (wrap ? 0 : 1);  // <= 544 threads: 120 registers/thread; else 64
  auto kern = variant == 0 ? head_kernel<544, true> : variant == 1 ? head_kernel<544, false>
            : variant == 2 ? head_kernel<1024, true> : head_kernel<1024, false>;
  if (smem > attr_smem[variant]) {
    DTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[variant] = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = n_maps < sms ? n_maps : sms;
  ProfRange pr(PROF_HEAD, st);
  kern<<<grid, threads, smem, st>>>(maps, n_maps, hp, hw, out_index, out, aux);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

}  // namespace dtk

using namespace dtk;

extern "C" int dinotrk_head(const float* maps, int n_maps, const dinotrk_geom* g, const dinotrk_head_weights* hw,
                            const int* out_index, float* out, int out_stride, int out_mode, int* aux, void* stream) {
  DTK_CHECK_ARG(maps && g && hw && out, "head: null pointer");
  DTK_CHECK_ARG(out_stride >= 2 && (out_mode == 0 || out_mode == 1), "head: bad out_stride/out_mode");
  return launch_head(maps, n_maps, dinotrk_map_stride(g), *g, *hw, out_index, out, out_stride, out_mode, aux,
                     (cudaStream_t)stream);
}
