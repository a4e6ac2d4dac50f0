// Shared helpers for libdinotrk (sm_100a).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler injects itself

#include "dinotrk.h"

namespace dtk {

void set_error(const char* fmt, ...);
extern unsigned long long g_launches;

#define DTK_CHECK_ARG(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      dtk::set_error(__VA_ARGS__);            \
      return DINOTRK_EINVAL;                  \
    }                                         \
  } while (0)

#define DTK_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      dtk::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return DINOTRK_ECUDA;                                                         \
    }                                                                               \
  } while (0)

// count + check a kernel launch
#define DTK_LAUNCHED()                 \
  do {                                 \
    ++dtk::g_launches;                 \
    DTK_CUDA(cudaPeekAtLastError());   \
  } while (0)

// ---- optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline) ----
enum ProfClass {
  PROF_SAMPLE = 0, PROF_CORR_GEMM, PROF_CORR_STREAM, PROF_HEAD, PROF_COS, PROF_ANCHOR_LIST, PROF_OCCLUSION,
  PROF_PACK, PROF_CONV, PROF_BLUR, PROF_ALIGN, PROF_MISC, PROF_BB, PROF_VIT_GEMM, PROF_VIT_ATTN, PROF_VIT_MISC, PROF_HEAD_FULL,
  PROF_XW_COARSE, PROF_XW_PLAN, PROF_XW_GEMM, PROF_XW_HEAD, PROF_TRAIN_BWD,
  PROF_COUNT
};
extern bool g_prof_on;
void prof_begin(int cls, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfRange {
  cudaStream_t st;
  bool on;
  ProfRange(int cls, cudaStream_t s) : st(s), on(g_prof_on) { if (on) prof_begin(cls, st); }
  ~ProfRange() { if (on) prof_end(st); }
};

// per-device cache slot for values that belong to a device's context (function attributes, occupancy queries,
// auxiliary streams): indexed by the CURRENT device, so a second GPU in the same process gets its own
constexpr int DTK_MAX_DEVICES = 64;
template <typename T>
struct PerDev {
  T v[DTK_MAX_DEVICES] = {};
  T& get() {
    int d = 0;
    cudaGetDevice(&d);
    return v[(d >= 0 && d < DTK_MAX_DEVICES) ? d : 0];
  }
};

// NVTX range over a host-side phase (nsys / ncu --nvtx): the four phases of dinotrk_infer, the ViT and delta-DINO stages
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t size, off;
  Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = (T*)(base + off);
    off += count * sizeof(T);
    return p;
  }
  bool ok() const { return off <= size; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- exact restatement of the reference's coordinate arithmetic ---------------------------
// models/tracker.py:84-93: a, b are computed in Python doubles and stored as fp32.
struct PointAffine {
  float aw, ah, bw, bh;
};
static inline PointAffine make_point_affine(const dinotrk_geom& g) {
  double p = g.patch, s = g.stride;
  double last_h = double((g.H - g.patch) / g.stride) * s + p / 2;
  double last_w = double((g.W - g.patch) / g.stride) * s + p / 2;
  PointAffine a;
  a.ah = (float)(2.0 / (last_h - p / 2));
  a.aw = (float)(2.0 / (last_w - p / 2));
  a.bh = (float)(1.0 - last_h * 2.0 / (last_h - p / 2));
  a.bw = (float)(1.0 - last_w * 2.0 / (last_w - p / 2));
  return a;
}

// ATen grid_sampler (align_corners=True, padding_mode=border): unnormalise then clip.
__device__ __forceinline__ float gs_unnorm_clip(float coord, int size) {
  float x = __fmul_rn(__fdiv_rn(__fadd_rn(coord, 1.f), 2.f), (float)(size - 1));
  return fminf(fmaxf(x, 0.f), (float)(size - 1));
}

// Trilinear sampling set-up for one point: 8 corners in ATen order
// (tnw, tne, tsw, tse, bnw, bne, bsw, bse) -> token index, set slot (z) and weight; weight 0 and
// token -1 for corners ATen skips as out of bounds.
struct TriCorners {
  int tok[4];     // (x0,y0) (x1,y0) (x0,y1) (x1,y1); -1 if out of bounds
  float wxy[4][2];  // [corner][z0|z1] full 3-factor weights in ATen's multiplication order
  int z0, z1;     // set slots; -1 if out of bounds
};
__device__ __forceinline__ TriCorners tri_setup(float xn, float yn, float set_idx, int N, int h, int w) {
  // utils.py:96-99: idx / (N-1) (skipped when N == 1), * 2 - 1
  float tn = set_idx;
  if (N > 1) tn = __fdiv_rn(tn, (float)(N - 1));
  tn = __fadd_rn(__fmul_rn(tn, 2.f), -1.f);
  float ix = gs_unnorm_clip(xn, w), iy = gs_unnorm_clip(yn, h), iz = gs_unnorm_clip(tn, N);
  float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
  float x1 = x0 + 1.f, y1 = y0 + 1.f, z1 = z0 + 1.f;
  float wx0 = __fsub_rn(x1, ix), wx1 = __fsub_rn(ix, x0);
  float wy0 = __fsub_rn(y1, iy), wy1 = __fsub_rn(iy, y0);
  float wz0 = __fsub_rn(z1, iz), wz1 = __fsub_rn(iz, z0);
  TriCorners c;
  int X0 = (int)x0, Y0 = (int)y0, X1 = X0 + 1, Y1 = Y0 + 1;
  bool okx1 = X1 <= w - 1, oky1 = Y1 <= h - 1;
  c.tok[0] = Y0 * w + X0;
  c.tok[1] = okx1 ? Y0 * w + X1 : -1;
  c.tok[2] = oky1 ? Y1 * w + X0 : -1;
  c.tok[3] = (okx1 && oky1) ? Y1 * w + X1 : -1;
  float wx[4] = {wx0, wx1, wx0, wx1};
  float wy[4] = {wy0, wy0, wy1, wy1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c.wxy[k][0] = __fmul_rn(__fmul_rn(wx[k], wy[k]), wz0);
    c.wxy[k][1] = __fmul_rn(__fmul_rn(wx[k], wy[k]), wz1);
  }
  c.z0 = (int)z0;
  c.z1 = ((int)z0 + 1 <= N - 1) ? (int)z0 + 1 : -1;
  return c;
}

}  // namespace dtk
