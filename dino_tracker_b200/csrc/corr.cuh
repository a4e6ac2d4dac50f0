// Internal interface of corr.cu / head.cu (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace dtk {

constexpr int STREAM_MAX_M = 8;  // groups with at most this many descriptors use the streaming kernel

size_t corr_plan_bytes(int n_groups);
int launch_corr_maps(const float* tpc, const float* norms, int C, int P, const float* desc,
                     const float* desc_norm, const int* grp_frame, const int* grp_row0, const int* grp_m,
                     const int* grp_map0, int n_groups, int total_maps, int max_group_m, float* maps,
                     int map_stride, int* tile_start, cudaStream_t st);

int launch_head(const float* maps, int n_maps, int map_stride, const dinotrk_geom& g,
                const dinotrk_head_weights& hw, const int* out_index, float* out, int out_stride, int out_mode,
                int* aux, cudaStream_t st);

}  // namespace dtk
