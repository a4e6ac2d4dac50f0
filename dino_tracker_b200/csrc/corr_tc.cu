// Tensor-core correlation GEMM (tcgen05, split-precision fp16 hi/lo) + operand splitting + TMA tensor-map helpers.
//
//   corr[j][p] = relu( <d_j, F[frame][p]> / max(|d_j| |F[frame][p]|, 1e-8) )     (models/tracker.py:158-173)
//
// The contraction runs as lo*hi + hi*lo + hi*hi on the kind::f16 tensor pipe with fp32 accumulation in TMEM
// (operands pre-split into fp16 hi + fp16 lo, x = hi + lo up to 2^-22 |x|), which keeps the products
// faithful to ~2^-21; the cosine normalisation and ReLU are the epilogue on the accumulator as it leaves TMEM.
#include <cuda_fp16.h>

#include "common.cuh"
#include "corr.cuh"
#include "tcgemm.cuh"
#include "tcgemm2.cuh"

namespace dtk {

// ---- driver entry point for cuTensorMapEncodeTiled (resolved once; no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, int elem) {
  EncodeTiledFn fn = get_encode();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return DINOTRK_ECUDA; }
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem == TMAP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                         : elem == TMAP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DINOTRK_ECUDA; }
  return DINOTRK_OK;
}

int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                 int elem, uint64_t ld) {
  const int elem_bytes = elem == TMAP_F32 ? 4 : 2;
  if (ld == 0) ld = cols;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(map, base, 2, dims, strides, box, elem);
}
int make_tmap_3d(CUtensorMap* map, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint32_t box_rows,
                 uint32_t box_cols, int elem, uint64_t ld) {
  const int elem_bytes = elem == TMAP_F32 ? 4 : 2;
  if (ld == 0) ld = cols;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * (uint64_t)elem_bytes, rows * ld * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  return encode(map, base, 3, dims, strides, box, elem);
}

// rank-4 map over a [d3][d2][d1][d0] tensor (d0 contiguous): used for {channels, w, h, T} boxes of the feature video
int make_tmap_4d(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4], int elem) {
  EncodeTiledFn fn = get_encode();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return DINOTRK_ECUDA; }
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapDataType dt = elem == TMAP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                         : elem == TMAP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 4, const_cast<void*>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (4-d) failed (%d)", (int)r); return DINOTRK_ECUDA; }
  return DINOTRK_OK;
}

// x = hi + lo (+ residual <= 2^-22 |x| in the fp16 normal range): hi = rn_fp16(x), lo = rn_fp16(x - hi)
__global__ void split_f16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 v = __ldg(x + i);
    __half h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y), h2 = __float2half_rn(v.z), h3 = __float2half_rn(v.w);
    __half l0 = __float2half_rn(v.x - __half2float(h0)), l1 = __float2half_rn(v.y - __half2float(h1));
    __half l2 = __float2half_rn(v.z - __half2float(h2)), l3 = __float2half_rn(v.w - __half2float(h3));
    __half2 a = __halves2half2(h0, h1), b2 = __halves2half2(h2, h3), c = __halves2half2(l0, l1), d = __halves2half2(l2, l3);
    hi[i] = make_uint2(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b2));
    lo[i] = make_uint2(*reinterpret_cast<unsigned*>(&c), *reinterpret_cast<unsigned*>(&d));
  }
}

int launch_split_f16(const float* x, void* hi, void* lo, size_t n, cudaStream_t st) {
  if (n == 0) return DINOTRK_OK;
  DTK_CHECK_ARG(n % 4 == 0, "split_fp16: length must be a multiple of 4");
  size_t n4 = n / 4;
  unsigned grid = (unsigned)((n4 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  ProfRange pr(PROF_MISC, st);
  split_f16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(hi),
                                         reinterpret_cast<uint2*>(lo), n4);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

struct CorrEpi {
  const float* norms;      // [T][P]
  const float* desc_norm;  // [rows]
  const int* grp_frame;
  const int* grp_row0;
  const int* grp_map0;
  float* maps;
  int map_stride, P;
  unsigned long long* tkeys;   // optional [maps][n_tiles]: (tile maximum, first token holding it) keys for the head (corr.cuh)
  int n_tiles;
  struct State { float mx; int tok; };
  __device__ __forceinline__ void tile_begin(State& s) const { s.mx = -1.f; s.tok = 0; }   // map values are >= 0 (ReLU)
  __device__ __forceinline__ void tile_end(State& s, int g, int r, int nt) const {
    if (tkeys)   // + 0.f: never the bit pattern of -0
      tkeys[(size_t)(grp_map0[g] + r) * n_tiles + nt] =
          ((unsigned long long)__float_as_uint(s.mx + 0.f) << 32) | (unsigned)(0x7fffffff - s.tok);
  }
  __device__ __forceinline__ void operator()(State& s, int g, int r, int col0, const float (&f)[32], int ncols) const {
    const float dn = desc_norm[grp_row0[g] + r];
    const float* fn = norms + (size_t)grp_frame[g] * P + col0;
    float* out = maps + (size_t)(grp_map0[g] + r) * map_stride + col0;
    if (ncols == 32) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        // norms rows are only 4-byte aligned (P is odd): scalar broadcast loads
        float4 n4 = make_float4(__ldg(fn + i), __ldg(fn + i + 1), __ldg(fn + i + 2), __ldg(fn + i + 3));
        float4 o;
        o.x = fmaxf(__fdiv_rn(f[i + 0], fmaxf(__fmul_rn(dn, n4.x), 1e-8f)), 0.f);
        o.y = fmaxf(__fdiv_rn(f[i + 1], fmaxf(__fmul_rn(dn, n4.y), 1e-8f)), 0.f);
        o.z = fmaxf(__fdiv_rn(f[i + 2], fmaxf(__fmul_rn(dn, n4.z), 1e-8f)), 0.f);
        o.w = fmaxf(__fdiv_rn(f[i + 3], fmaxf(__fmul_rn(dn, n4.w), 1e-8f)), 0.f);
        *reinterpret_cast<float4*>(out + i) = o;
        // strict >: the first token of the tile holding the maximum (columns are visited in increasing order)
        if (o.x > s.mx) { s.mx = o.x; s.tok = col0 + i; }
        if (o.y > s.mx) { s.mx = o.y; s.tok = col0 + i + 1; }
        if (o.z > s.mx) { s.mx = o.z; s.tok = col0 + i + 2; }
        if (o.w > s.mx) { s.mx = o.w; s.tok = col0 + i + 3; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < ncols) {
          const float o = fmaxf(__fdiv_rn(f[i], fmaxf(__fmul_rn(dn, fn[i]), 1e-8f)), 0.f);
          out[i] = o;
          if (o > s.mx) { s.mx = o; s.tok = col0 + i; }
        }
    }
  }
};

int corr_tc_tile_rows() {
  static int rows = 0;
  if (rows == 0) {
    const char* e = getenv("DTK_CORR_PAIRS");
    rows = (e && atoi(e) == 0) ? TC_BM : TC2_BM;
  }
  return rows;
}

size_t corr_tc_workspace_bytes(int total_rows, int C) { return 2 * align_up((size_t)total_rows * C * 2, 256); }

// wide groups on tensor cores; tile_start must already hold the plan (corr_plan_kernel).
int launch_corr_gemm_tc(const void* tpc_hi, const void* tpc_lo, const float* norms, int T, int C, int P,
                        const float* desc, int desc_rows, const float* desc_norm, const int* grp_frame,
                        const int* grp_row0, const int* grp_m, const int* grp_map0, const int* tile_start, int n_groups,
                        int max_tiles, float* maps, int map_stride, float* desc_split_ws, cudaStream_t st,
                        unsigned long long* tkeys, bool split_ready, int tile_rows) {
  using Cfg = TcCfg<TcMode::F16X3>;
  static_assert(TC_BN == CORR_TILE, "the tile maxima are per GEMM N tile");
  DTK_CHECK_ARG(C % 8 == 0, "corr (tensor path): C must be a multiple of 8");
  char* d_hi = reinterpret_cast<char*>(desc_split_ws);
  char* d_lo = d_hi + align_up((size_t)desc_rows * C * 2, 256);
  int rc = split_ready ? DINOTRK_OK : launch_split_f16(desc, d_hi, d_lo, (size_t)desc_rows * C, st);
  if (rc) return rc;
  const bool pairs = (tile_rows > 0 ? tile_rows : corr_tc_tile_rows()) == TC2_BM;   // tile_start was planned with this M tile
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  const uint32_t b_box = pairs ? TC2_BN / 2 : TC_BN;   // a CTA of a pair stages half of the B tile
  if ((rc = make_tmap_2d(&tmA_hi, d_hi, desc_rows, C, TC_BM, Cfg::kBK, TMAP_F16))) return rc;
  if ((rc = make_tmap_2d(&tmA_lo, d_lo, desc_rows, C, TC_BM, Cfg::kBK, TMAP_F16))) return rc;
  if ((rc = make_tmap_3d(&tmB_hi, tpc_hi, T, P, C, b_box, Cfg::kBK, TMAP_F16))) return rc;
  if ((rc = make_tmap_3d(&tmB_lo, tpc_lo, T, P, C, b_box, Cfg::kBK, TMAP_F16))) return rc;
  static PerDev<bool> attr_dev;
  bool& attr = attr_dev.get();
  if (!attr) {
    DTK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TcMode::F16X3, CorrEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmem));
    DTK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<TcMode::F16X3, CorrEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Tc2Cfg<TcMode::F16X3>::kSmem));
    attr = true;
  }
  TcProblem pb{grp_frame, grp_row0, grp_m, tile_start, n_groups, P, C};
  CorrEpi epi{norms, desc_norm, grp_frame, grp_row0, grp_map0, maps, map_stride, P, tkeys, cdiv(P, CORR_TILE)};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int tiles_bound = max_tiles * cdiv(P, TC_BN);
  ProfRange pr(PROF_CORR_GEMM, st);
  if (pairs) {
    int grid = 2 * (tiles_bound < sms / 2 ? tiles_bound : sms / 2);
    if (grid < 2) grid = 2;
    tc_gemm2_kernel<TcMode::F16X3, CorrEpi><<<grid, TC_THREADS, Tc2Cfg<TcMode::F16X3>::kSmem, st>>>(tmA_hi, tmA_lo, tmB_hi,
                                                                                                   tmB_lo, pb, epi);
    DTK_LAUNCHED();
    return DINOTRK_OK;
  }
  int grid = tiles_bound < sms ? tiles_bound : sms;
  if (grid < 1) grid = 1;
  tc_gemm_kernel<TcMode::F16X3, CorrEpi><<<grid, TC_THREADS, Cfg::kSmem, st>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, pb, epi);
  DTK_LAUNCHED();
  return DINOTRK_OK;
}

}  // namespace dtk

using namespace dtk;

extern "C" int dinotrk_split_fp16(const float* x, void* hi, void* lo, size_t n, void* stream) {
  DTK_CHECK_ARG(x && hi && lo, "split_fp16: null pointer");
  return launch_split_f16(x, hi, lo, n, (cudaStream_t)stream);
}
