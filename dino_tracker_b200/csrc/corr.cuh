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
//   tkeys       [total_maps][cdiv(P, CORR_TILE)]: per map and 256-token tile, (bits of the tile maximum) << 32 |
//               (0x7fffffff - first token holding it), written by the GEMM epilogue: the maximum over a map's keys is its
//               first arg-max.  Maps of thin groups (streaming kernel) get ~0 in tile 0 = "no keys".  Only produced on
//               the tensor path (fv.tensor()).
//   zero_word   an int the plan kernel sets to 0 (the head's counter of uncertified maps: saves a launch)
//   split_ready the fp16 hi/lo copies of `desc` are already in split_ws (written by the sampler): skip the split kernel
//   no_thin     the caller knows that no group has <= STREAM_MAX_M descriptors: skip the streaming kernel launch
struct CorrAssist {
  unsigned long long* tkeys = nullptr;
  int* zero_word = nullptr;
  bool split_ready = false;
  bool no_thin = false;
  bool all_wide = false;   // no streaming kernel: groups of any size > 0 are GEMM tiles
  bool small_tiles = false;   // tensor path: 128-row single-CTA tiles instead of 256-row CTA-pair tiles (many tiny groups)
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
                        unsigned long long* tkeys = nullptr, bool split_ready = false, int tile_rows = 0 /* 0: default */);
int launch_split_f16(const float* x, void* hi, void* lo, size_t n, cudaStream_t st);

int launch_head(const float* maps, int n_maps, int map_stride, const dinotrk_geom& g,
                const dinotrk_head_weights& hw, const int* out_index, float* out, int out_stride, int out_mode,
                int* aux, int* scratch /* n_maps + 1 ints, or NULL: full-map kernel for every map */, cudaStream_t st,
                const unsigned long long* tkeys = nullptr /* tile keys of launch_corr_maps */, bool counter_zeroed = false,
                int ctas_per_sm = 0 /* > 0: cap of the fast-path grid (co-residency with a GEMM on another stream) */,
                int parts = 3 /* bit 0: fast path over all maps, bit 1: full-map kernel over the uncertified ones */);

}  // namespace dtk
