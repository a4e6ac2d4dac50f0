// Tracker head: arg-max, 2-layer normalised-conv refiner, spatial softmax, disc-masked soft-argmax
// with the numerical-stability fallback (models/networks/tracker_head.py:107-121, :68-98, :100-105;
// conv_norm.py:34-46; data/dataset.py:21-53).
//
// Two kernels:
//  * head_window_kernel (fast path, every map): arg-max, then the EXACT refiner only on the 11x11 box around
//    the arg-max (hidden layer on 13x13, input window 15x15).  The soft-argmax needs nothing else as long as
//    the numerical-stability branch (disc mass < 1e-8 of the whole softmax) does not fire; that is certified
//    with a rigorous, monotone upper bound on every logit outside the box (from the largest map value outside
//    the 7x7 core and the positive parts of the normalised weights).  Maps that cannot be certified are
//    queued for
//  * head_kernel (full map): the complete refiner + softmax, exactly as the reference evaluates it.
//
// head_kernel: one persistent CTA per SM; one warp per 4-row band of the map, one lane per 4x4-pixel tile
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
  float P1[16], P2[16];                // sums of the positive parts of the normalised 3x3 kernels (logit bound)
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
head_kernel(const float* __restrict__ maps, int n_maps_arg, const int* __restrict__ map_list,
            const int* __restrict__ list_count, HeadParams hp, dinotrk_head_weights wts,
            const int* __restrict__ out_index, float* __restrict__ out, int* __restrict__ aux) {
  extern __shared__ __align__(16) float smem[];
  // with a list: process map_list[0 .. *list_count) (maps the window kernel could not certify)
  const int n_maps = map_list ? *list_count : n_maps_arg;
  auto map_of = [&](int k) { return map_list ? map_list[k] : k; };
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int h = hp.h, w = hp.w, P = hp.P;
  const int lin_elems = (hp.map_stride + 3) & ~3;
  float* sm_lin[2] = {smem, smem + lin_elems};
  float* sm_hid = smem + 2 * lin_elems;                   // [2 buffers][nwarps][2 rows][128]
  float* sm_red = sm_hid + 2 * nwarps * 2 * 128;          // [3 * 32] floats
  unsigned long long* sm_red64 = reinterpret_cast<unsigned long long*>(sm_red + 96);  // [32]

  const int r0 = warp * 4, c0 = lane * 4;  // this lane's tile
  const int nchunks = hp.map_stride / 4;
  constexpr float kNeg = -1e30f;
  // additive masks: hidden activations of pixels outside the map must be exactly 0 (zero padding of the
  // second convolution); bias + kNeg makes the ReLU do that without per-pixel selects.
  float rmask[4], cmask[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { rmask[i] = (r0 + i < h) ? 0.f : kNeg; cmask[i] = (c0 + i < w) ? 0.f : kNeg; }
  const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;

  int mk = blockIdx.x;
  if (mk < n_maps) {
    const float4* src = reinterpret_cast<const float4*>(maps + (size_t)map_of(mk) * hp.map_stride);
    for (int i = threadIdx.x; i < nchunks; i += blockDim.x) cp_async16_head(sm_lin[0] + 4 * i, src + i);
  }
  asm volatile("cp.async.commit_group;\n" ::);

  for (int it = 0; mk < n_maps; mk += gridDim.x, ++it) {
    const int map = map_of(mk);
    const float* lin = sm_lin[it & 1];
    {  // prefetch the next map into the other buffer
      int nmk = mk + gridDim.x;
      if (nmk < n_maps) {
        const float4* src = reinterpret_cast<const float4*>(maps + (size_t)map_of(nmk) * hp.map_stride);
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

    // ---- 6x6 input window (zero outside the map = conv zero padding) -------------------------
    float m[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      int r = r0 - 1 + i;
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

    float hid[6][6];  // [1..4][1..4] own values, ring = halo
    auto conv1 = [&](int o) {
      float w1[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w1[k] = wts.w1[o][k];
      const float b1 = wts.b1[o];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float br = b1 + rmask[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = br + cmask[j];
#pragma unroll
          for (int ki = 0; ki < 3; ++ki)
#pragma unroll
            for (int kj = 0; kj < 3; ++kj) a = fmaf(w1[ki * 3 + kj], m[i + ki][j + kj], a);
          hid[i + 1][j + 1] = fmaxf(a, 0.f);
        }
      }
      float* hb = sm_hid + ((o & 1) * nwarps + warp) * 2 * 128;
      *reinterpret_cast<float4*>(hb + c0) = make_float4(hid[1][1], hid[1][2], hid[1][3], hid[1][4]);
      *reinterpret_cast<float4*>(hb + 128 + c0) = make_float4(hid[4][1], hid[4][2], hid[4][3], hid[4][4]);
    };

    conv1(0);
    __syncthreads();  // also publishes sm_red64 (arg-max partials)
#pragma unroll 1
    for (int o = 0; o < 16; ++o) {
      float w2[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w2[k] = wts.w2[o][k];
      // halo rows of channel o were published before the previous barrier
      float4 top = make_float4(0.f, 0.f, 0.f, 0.f), bot = top;
      if (warp > 0) top = *reinterpret_cast<const float4*>(sm_hid + ((o & 1) * nwarps + warp - 1) * 2 * 128 + 128 + c0);
      if (warp + 1 < nwarps) bot = *reinterpret_cast<const float4*>(sm_hid + ((o & 1) * nwarps + warp + 1) * 2 * 128 + c0);
      hid[0][1] = top.x; hid[0][2] = top.y; hid[0][3] = top.z; hid[0][4] = top.w;
      hid[5][1] = bot.x; hid[5][2] = bot.y; hid[5][3] = bot.z; hid[5][4] = bot.w;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (kWrap) {
          hid[i][0] = __shfl_sync(0xffffffffu, hid[i][4], lane_l);
          hid[i][5] = __shfl_sync(0xffffffffu, hid[i][1], lane_r);
        } else {
          float l = __shfl_up_sync(0xffffffffu, hid[i][4], 1);
          float r = __shfl_down_sync(0xffffffffu, hid[i][1], 1);
          hid[i][0] = lane > 0 ? l : 0.f;
          hid[i][5] = lane < 31 ? r : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = acc[i][j];
#pragma unroll
          for (int ki = 0; ki < 3; ++ki)
#pragma unroll
            for (int kj = 0; kj < 3; ++kj) a = fmaf(w2[ki * 3 + kj], hid[i + ki][j + kj], a);
          acc[i][j] = a;
        }
      if (o + 1 < 16) conv1(o + 1);   // next channel's hidden layer; published by the barrier below
      __syncthreads();
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


// ------------------------------------------------------------------------------------------------------
// Fast path: exact refiner on the 11x11 box around the arg-max + certified absence of the fallback branch.
constexpr int WIN_THREADS = 128;
constexpr int WB = 11, WH = 13, WM = 15;  // box, hidden window, input window (side lengths); disc radius <= 5 tokens
constexpr int WHC = 20;   // hidden window is stored [position][16 channels] with a 20-float pitch: float4 accesses of
                          // consecutive positions fall into distinct bank groups

// TM = true: the correlation GEMM already reduced every 256-token tile of the map to its maximum (tmax, corr.cuh), so the
// arg-max and the largest value outside the 7x7 core come from ~1.3 k tokens instead of two passes over all 8107, and the
// map is never staged in shared memory (12 KB instead of 44 KB per CTA: 16 CTAs per SM).
template <bool TM>
__global__ void __launch_bounds__(WIN_THREADS)
head_window_kernel(const float* __restrict__ maps, const float* __restrict__ tmax, int n_tiles, int n_maps, HeadParams hp,
                   dinotrk_head_weights wts, const int* __restrict__ out_index, float* __restrict__ out,
                   int* __restrict__ aux, int* __restrict__ slow_list, int* __restrict__ slow_count) {
  extern __shared__ __align__(16) float smem[];
  const int lin_elems = TM ? 0 : (hp.map_stride + 3) & ~3;
  float* lin = smem;                           // one map (several CTAs per SM hide the load latency); unused when TM
  float* sm_m = smem + lin_elems;              // [WM][WM] input window, zero outside the map
  float* sm_h = sm_m + WM * WM + 3;            // [WH * WH][WHC] hidden window (channel-innermost), zero outside the map
  float* sm_red = sm_h + WHC * WH * WH;        // partials
  unsigned long long* sm_key = reinterpret_cast<unsigned long long*>(sm_red + 32);  // [4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = hp.h, w = hp.w, P = hp.P;
  const int nchunks = hp.map_stride / 4;       // float4 chunks; the tail of the last chunk (>= P) is masked below

  for (int map = blockIdx.x; map < n_maps; map += gridDim.x) {
    const float* src = TM ? maps + (size_t)map * hp.map_stride : lin;   // where map values are read from below
    int amax;
    float mout_tiles = 0.f;   // TM: largest tile maximum among the tiles that do not touch the core rows
    int tok_lo = 0, tok_n = 0;   // TM: token range of the tiles that do
    if constexpr (TM) {
      const float* tm = tmax + (size_t)map * n_tiles;
      if (__ldg(tm) < 0.f) {   // thin group (streaming kernel): no tile maxima -> full-map kernel
        if (tid == 0) slow_list[atomicAdd(slow_count, 1)] = map;
        continue;
      }
      // every warp redundantly: (max, first tile holding it), then the first token of that tile equal to the max
      unsigned long long key = 0ull;
      for (int t = lane; t < n_tiles; t += 32) {
        unsigned long long k = ((unsigned long long)__float_as_uint(__ldg(tm + t) + 0.f) << 32) | (unsigned)(0x7fffffff - t);
        key = k > key ? k : key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o); key = t > key ? t : key; }
      const int wt = 0x7fffffff - (int)(key & 0xffffffffu);
      const float vmax = __uint_as_float((unsigned)(key >> 32));
      int cand = 0x7fffffff;
      for (int i = lane; i < CORR_TILE; i += 32) {
        const int p = wt * CORR_TILE + i;
        if (p < P && __ldg(src + p) + 0.f == vmax) cand = min(cand, p);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      amax = cand < P ? cand : wt * CORR_TILE;   // (always found: the tile maximum is one of the stored values)
      const int ar = amax / w, ac = amax - ar * w;
      const int t_lo = (max(ar - 3, 0) * w + max(ac - 3, 0)) / CORR_TILE;
      const int t_hi = (min(ar + 3, h - 1) * w + min(ac + 3, w - 1)) / CORR_TILE;
      for (int t = lane; t < n_tiles; t += 32)
        if (t < t_lo || t > t_hi) mout_tiles = fmaxf(mout_tiles, __ldg(tm + t));
      tok_lo = t_lo * CORR_TILE;
      tok_n = min((t_hi + 1) * CORR_TILE, P) - tok_lo;
    } else {
    {
      const float4* gsrc = reinterpret_cast<const float4*>(maps + (size_t)map * hp.map_stride);
      for (int i = tid; i < nchunks; i += WIN_THREADS) cp_async16_head(lin + 4 * i, gsrc + i);
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();

    // ---- arg-max (first maximal index): chunk-wise maxima, then the first chunk / element holding the max ----
    float best = -1.f;
    int bchunk = 0;
    for (int i = tid; i < nchunks; i += WIN_THREADS) {
      float4 v = *reinterpret_cast<const float4*>(lin + 4 * i);
      const int base = 4 * i;
      float m4 = v.x;                                          // element base always < P
      if (base + 1 < P) m4 = fmaxf(m4, v.y);
      if (base + 2 < P) m4 = fmaxf(m4, v.z);
      if (base + 3 < P) m4 = fmaxf(m4, v.w);
      if (m4 > best) { best = m4; bchunk = i; }                // strict: keeps the first chunk of this thread
    }
    // threads without a chunk (tiny maps) must not win: key 0 (values are >= 0, so real keys order like floats)
    unsigned long long key = best < 0.f ? 0ull
        : (((unsigned long long)__float_as_uint(best + 0.f) << 32) | (unsigned)(0x7fffffff - bchunk));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o); key = t > key ? t : key; }
    if (lane == 0) sm_key[warp] = key;
    __syncthreads();
    unsigned long long kb = sm_key[0];
#pragma unroll
    for (int k = 1; k < WIN_THREADS / 32; ++k) { unsigned long long t = sm_key[k]; kb = t > kb ? t : kb; }
    const int wchunk = 0x7fffffff - (int)(kb & 0xffffffffu);
    const float vmax = __uint_as_float((unsigned)(kb >> 32));
    amax = 4 * wchunk;
    {
      const float* q = lin + 4 * wchunk;
      amax += (q[0] + 0.f == vmax) ? 0 : (q[1] + 0.f == vmax) ? 1 : (q[2] + 0.f == vmax) ? 2 : 3;
    }
    }
    const int arow = amax / w, acol = amax - arow * w;

    // ---- input window (15 x 15, zero outside the map) ---------------------------------------------
    for (int i = tid; i < WM * WM; i += WIN_THREADS) {
      int y = i / WM, x = i - y * WM;
      int r = arow - 7 + y, c = acol - 7 + x;
      sm_m[i] = (r >= 0 && r < h && c >= 0 && c < w) ? src[r * w + c] : 0.f;
    }
    __syncthreads();
    // ---- largest map value outside the 7x7 core: blank the core in the private copy, then a plain max ----
    if (!TM && tid < 49) {
      int r = arow - 3 + tid / 7, c = acol - 3 + tid % 7;
      if (r >= 0 && r < h && c >= 0 && c < w) lin[r * w + c] = 0.f;
    }
    // ---- hidden layer on the 13 x 13 window (zero outside the map: padding of the second conv) -----
    for (int i = tid; i < WH * WH; i += WIN_THREADS) {
      int y = i / WH, x = i - y * WH;
      int r = arow - 6 + y, c = acol - 6 + x;
      const bool inside = r >= 0 && r < h && c >= 0 && c < w;
      float mw[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) mw[ky * 3 + kx] = sm_m[(y + ky) * WM + x + kx];
      float4* hrow = reinterpret_cast<float4*>(sm_h + i * WHC);
#pragma unroll
      for (int o4 = 0; o4 < 4; ++o4) {
        float a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = o4 * 4 + j;
          a[j] = wts.b1[o];
#pragma unroll
          for (int k = 0; k < 9; ++k) a[j] = fmaf(wts.w1[o][k], mw[k], a[j]);
          a[j] = inside ? fmaxf(a[j], 0.f) : 0.f;
        }
        hrow[o4] = make_float4(a[0], a[1], a[2], a[3]);
      }
    }
    __syncthreads();
    float mout = mout_tiles;
    if constexpr (TM) {
      // tokens of the tiles that touch the core rows, core excluded; row by row (no division per token)
      const int r_lo = tok_lo / w, r_hi = (tok_lo + tok_n - 1) / w;
      for (int r = r_lo; r <= r_hi; ++r) {
        const bool core_row = abs(r - arow) <= 3;
        for (int c = tid; c < w; c += WIN_THREADS) {
          const int p = r * w + c;
          if (p >= tok_lo && p < tok_lo + tok_n && !(core_row && abs(c - acol) <= 3)) mout = fmaxf(mout, __ldg(src + p));
        }
      }
    } else {
      for (int i = tid; i < nchunks; i += WIN_THREADS) {
        float4 v = *reinterpret_cast<const float4*>(lin + 4 * i);
        const int base = 4 * i;
        float m4 = v.x;
        if (base + 1 < P) m4 = fmaxf(m4, v.y);
        if (base + 2 < P) m4 = fmaxf(m4, v.z);
        if (base + 3 < P) m4 = fmaxf(m4, v.w);
        mout = fmaxf(mout, m4);
      }
    }
    mout = warp_max(mout);

    // ---- logits on the 11 x 11 box; thread = box pixel ----------------------------------------------
    float z = -INFINITY;
    bool valid = false, indisc = false;
    float px = 0.f, py = 0.f;
    if (tid < WB * WB) {
      int y = tid / WB, x = tid - y * WB;
      int r = arow - 5 + y, c = acol - 5 + x;
      valid = r >= 0 && r < h && c >= 0 && c < w;
      if (valid) {
        float a = wts.b2;
        // same accumulation order as before: channel-major, then the 3 x 3 taps
#pragma unroll
        for (int o4 = 0; o4 < 4; ++o4) {
          float4 hv[9];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              hv[ky * 3 + kx] = *reinterpret_cast<const float4*>(sm_h + ((y + ky) * WH + x + kx) * WHC + o4 * 4);
#pragma unroll
          for (int k = 0; k < 9; ++k) a = fmaf(wts.w2[o4 * 4 + 0][k], hv[k].x, a);
#pragma unroll
          for (int k = 0; k < 9; ++k) a = fmaf(wts.w2[o4 * 4 + 1][k], hv[k].y, a);
#pragma unroll
          for (int k = 0; k < 9; ++k) a = fmaf(wts.w2[o4 * 4 + 2][k], hv[k].z, a);
#pragma unroll
          for (int k = 0; k < 9; ++k) a = fmaf(wts.w2[o4 * 4 + 3][k], hv[k].w, a);
        }
        z = a;
        int dr = (r - arow) * hp.stride_px, dc = (c - acol) * hp.stride_px;
        indisc = dr * dr + dc * dc <= hp.radius2;
        px = (float)(hp.half_patch + c * hp.stride_px);
        py = (float)(hp.half_patch + r * hp.stride_px);
      }
    }
    float zmax = warp_max(z);
    if (lane == 0) { sm_red[warp] = mout; sm_red[8 + warp] = zmax; }
    __syncthreads();
    mout = fmaxf(fmaxf(sm_red[0], sm_red[1]), fmaxf(sm_red[2], sm_red[3]));
    zmax = fmaxf(fmaxf(sm_red[8], sm_red[9]), fmaxf(sm_red[10], sm_red[11]));
    const float e = valid ? expf(z - zmax) : 0.f;
    float v5[5] = {e, indisc ? e : 0.f, indisc ? px * e : 0.f, indisc ? py * e : 0.f, valid ? 1.f : 0.f};
#pragma unroll
    for (int q = 0; q < 5; ++q) v5[q] = warp_sum(v5[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 5; ++q) sm_red[12 + q * 4 + warp] = v5[q];
    }
    __syncthreads();
    if (tid == 0) {
      float tot[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) tot[q] = sm_red[12 + q * 4] + sm_red[13 + q * 4] + sm_red[14 + q * 4] + sm_red[15 + q * 4];
      // every logit outside the box:  z <= b2 + sum_o P2_o * relu(b1_o + P1_o * mout)   (all terms monotone in m >= 0)
      float F = wts.b2;
#pragma unroll
      for (int o = 0; o < 16; ++o) F = fmaf(hp.P2[o], fmaxf(fmaf(hp.P1[o], mout, wts.b1[o]), 0.f), F);
      const float rest = ((float)P - tot[4]) * expf(fminf(F - zmax, 80.f));
      // certified: disc mass >= 2e-8 of (an upper bound of) the whole softmax  ->  the reference does not take the
      // stability branch and its result is sum(x e) / sum(e) over the disc (the normaliser cancels)
      const bool certified = tot[1] >= 2e-8f * (tot[0] + rest) && tot[1] > 0.f && isfinite(rest);
      if (certified) {
        float px_ = __fdiv_rn(tot[2], tot[1]), py_ = __fdiv_rn(tot[3], tot[1]);
        float nx = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(px_, hp.normW)), -1.f);
        float ny = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(py_, hp.normH)), -1.f);
        if (hp.out_mode == 0) {
          nx = __fmul_rn(__fdiv_rn(__fadd_rn(nx, 1.f), 2.f), hp.normW);
          ny = __fmul_rn(__fdiv_rn(__fadd_rn(ny, 1.f), 2.f), hp.normH);
        }
        size_t oi = (size_t)(out_index ? out_index[map] : map) * hp.out_stride;
        out[oi] = nx; out[oi + 1] = ny;
        if (aux) { aux[2 * map] = amax; aux[2 * map + 1] = 0; }
      } else {
        slow_list[atomicAdd(slow_count, 1)] = map;
      }
    }
    __syncthreads();  // lin / windows are reused by the next iteration
  }
}

__global__ void zero_int_kernel(int* p) { *p = 0; }

int launch_head(const float* maps, int n_maps, int map_stride, const dinotrk_geom& g,
                const dinotrk_head_weights& hw, const int* out_index, float* out, int out_stride, int out_mode,
                int* aux, int* scratch, cudaStream_t st, const float* tmax, bool counter_zeroed) {
  if (n_maps <= 0) return DINOTRK_OK;
  DTK_CHECK_ARG(g.w <= HEAD_MAX_W && g.h <= HEAD_MAX_H, "head: token grid %dx%d exceeds the supported %dx%d",
                g.h, g.w, HEAD_MAX_H, HEAD_MAX_W);
  HeadParams hp;
  hp.h = g.h; hp.w = g.w; hp.P = g.h * g.w; hp.map_stride = map_stride;
  hp.stride_px = g.stride; hp.half_patch = g.patch / 2; hp.radius2 = g.radius * g.radius;
  hp.normW = (float)(g.W - 1); hp.normH = (float)(g.H - 1);
  hp.out_stride = out_stride; hp.out_mode = out_mode;
  for (int o = 0; o < 16; ++o) {
    float p1 = 0.f, p2 = 0.f;
    for (int k = 0; k < 9; ++k) { p1 += hw.w1[o][k] > 0.f ? hw.w1[o][k] : 0.f; p2 += hw.w2[o][k] > 0.f ? hw.w2[o][k] : 0.f; }
    hp.P1[o] = p1 * (1.f + 1e-6f); hp.P2[o] = p2 * (1.f + 1e-6f);   // rounded up: the bound must stay a bound
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int lin_elems = (map_stride + 3) & ~3;
  // the window fast path needs the disc inside the 11 x 11 box and a scratch list for the uncertified maps
  const bool window_ok = scratch != nullptr && g.radius <= 5 * g.stride && g.w <= HEAD_MAX_W;
  int* slow_count = scratch;
  int* slow_list = scratch ? scratch + 1 : nullptr;
  if (window_ok) {
    const bool tm = tmax != nullptr;
    size_t smem = (size_t)((tm ? 0 : lin_elems) + WM * WM + 3 + WHC * WH * WH + 32) * sizeof(float) + 4 * sizeof(unsigned long long);
    static size_t attr_w = 0;
    if (!tm && smem > attr_w) {
      DTK_CUDA(cudaFuncSetAttribute(head_window_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_w = smem;
    }
    if (!counter_zeroed) {
      ProfRange pr(PROF_MISC, st);
      zero_int_kernel<<<1, 1, 0, st>>>(slow_count);
      DTK_LAUNCHED();
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > (tm ? 16 : 8)) per_sm = tm ? 16 : 8;
    int grid = n_maps < sms * per_sm ? n_maps : sms * per_sm;
    ProfRange pr(PROF_HEAD, st);
    if (tm)
      head_window_kernel<true><<<grid, WIN_THREADS, smem, st>>>(maps, tmax, cdiv(hp.P, CORR_TILE), n_maps, hp, hw, out_index,
                                                                out, aux, slow_list, slow_count);
    else
      head_window_kernel<false><<<grid, WIN_THREADS, smem, st>>>(maps, nullptr, 0, n_maps, hp, hw, out_index, out, aux,
                                                                 slow_list, slow_count);
    DTK_LAUNCHED();
  }
  // full-map kernel: every map (no scratch) or only the maps the window kernel could not certify
  const int nwarps = cdiv(g.h, 4);
  const int threads = nwarps * 32;
  size_t smem = (size_t)(2 * lin_elems + 2 * nwarps * 2 * 128 + 6 * 32) * sizeof(float) + 32 * sizeof(unsigned long long);
  static size_t attr_smem[4] = {0, 0, 0, 0};
  const bool wrap = g.w <= 124;                 // tile 31 of every band lies outside the map
  const int variant = (threads <= 576 ? 0 : 2) + (wrap ? 0 : 1);  // <= 576 threads: 112 registers/thread; else 64
  auto kern = variant == 0 ? head_kernel<576, true> : variant == 1 ? head_kernel<576, false>
            : variant == 2 ? head_kernel<1024, true> : head_kernel<1024, false>;
  if (smem > attr_smem[variant]) {
    DTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[variant] = smem;
  }
  int grid = n_maps < sms ? n_maps : sms;
  ProfRange pr(PROF_HEAD_FULL, st);
  kern<<<grid, threads, smem, st>>>(maps, n_maps, window_ok ? slow_list : nullptr, window_ok ? slow_count : nullptr, hp,
                                    hw, out_index, out, aux);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

}  // namespace dtk

using namespace dtk;

extern "C" int dinotrk_head(const float* maps, int n_maps, const dinotrk_geom* g, const dinotrk_head_weights* hw,
                            const int* out_index, float* out, int out_stride, int out_mode, int* aux, int* scratch,
                            void* stream) {
  DTK_CHECK_ARG(maps && g && hw && out, "head: null pointer");
  DTK_CHECK_ARG(out_stride >= 2 && (out_mode == 0 || out_mode == 1), "head: bad out_stride/out_mode");
  return launch_head(maps, n_maps, dinotrk_map_stride(g), *g, *hw, out_index, out, out_stride, out_mode, aux, scratch,
                     (cudaStream_t)stream);
}
