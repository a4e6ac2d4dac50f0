// Internal interface of corr.cu / head.cu (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace dtk {

constexpr int STREAM_MAX_M = 8;  // groups with at most this many descriptors use the streaming kernel

struct FeatView {
  const float* tpc; const float* norms; const void* hi; const void* lo;
  int T, C, P;
  bool tensor() const { return hi != nullptr && lo != nullptr; }
};
static inline FeatView make_view(const dinotrk_features& f, const dinotrk_geom& g) {
  return FeatView{f.tpc, f.norms, f.hi, f.lo, f.T, f.C, g.h * g.w};
}

constexpr int CORR_TILE = 256;   // token tile of the tensor-core correlation GEMM (= TC_BN); unit of the tile maxima

// Optional by-products / shortcuts of one launch_corr_maps call (all members may stay zero):
//   tmax        [total_maps][cdiv(P, CORR_TILE)]: per-map maxima of every 256-token tile, written by the GEMM epilogue
//               (maps of thin groups, which the streaming kernel computes, get -1 in tile 0 = "no tile maxima");
//               only produced on the tensor path (fv.tensor()).
//   zero_word   an int the plan kernel sets to 0 (the head's counter of uncertified maps: saves a launch)
//   split_ready the fp16 hi/lo copies of `desc` are already in split_ws (written by the sampler): skip the split kernel
//   no_thin     the caller knows that no group has <= STREAM_MAX_M descriptors: skip the streaming kernel launch
struct CorrAssist {
  float* tmax = nullptr;
  int* zero_word = nullptr;
  bool split_ready = false;
  bool no_thin = false;
};

size_t corr_plan_bytes(int n_groups);
int corr_tc_tile_rows();   // 256: CTA-pair (cta_group::2) kernel, the default; 128: single-CTA kernel (DTK_CORR_PAIRS=0)
size_t corr_tc_workspace_bytes(int total_rows, int C);
// desc_rows = number of rows of the desc array (bounds of its tensor map); split_ws: corr_tc_workspace_bytes
// (only touched when fv.tensor()).
int launch_corr_maps(const FeatView& fv, const float* desc, int desc_rows, const float* desc_norm,
                     const int* grp_frame, const int* grp_row0, const int* grp_m, const int* grp_map0, int n_groups,
                     int total_maps, int max_group_m, float* maps, int map_stride, int* tile_start, float* split_ws,
                     cudaStream_t st, const CorrAssist& assist = CorrAssist());
int launch_corr_gemm_tc(const void* tpc_hi, const void* tpc_lo, const float* norms, int T, int C, int P,
                        const float* desc, int desc_rows, const float* desc_norm, const int* grp_frame,
                        const int* grp_row0, const int* grp_m, const int* grp_map0, const int* tile_start, int n_groups,
                        int max_tiles, float* maps, int map_stride, float* desc_split_ws, cudaStream_t st,
                        float* tmax = nullptr, bool split_ready = false);
int launch_split_f16(const float* x, void* hi, void* lo, size_t n, cudaStream_t st);

int launch_head(const float* maps, int n_maps, int map_stride, const dinotrk_geom& g,
                const dinotrk_head_weights& hw, const int* out_index, float* out, int out_stride, int out_mode,
                int* aux, int* scratch /* n_maps + 1 ints, or NULL: full-map kernel for every map */, cudaStream_t st,
                const float* tmax = nullptr /* tile maxima of launch_corr_maps */, bool counter_zeroed = false);

}  // namespace dtk
