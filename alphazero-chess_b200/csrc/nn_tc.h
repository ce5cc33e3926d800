// Host-side interface of the tcgen05 convolution (nn_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace azb {
// activations: NHWC bf16 [max_boards][8][8][channels], channels in {64, 128}
int tc_make_act_map(CUtensorMap* map, const void* base, int channels, int max_boards, int wide = 0);
// the same buffer as a matrix [max_boards * 64 pixel rows][128 channels] for the tcgen05 heads (nn_heads_tc.cu)
int tc_make_rows_map(CUtensorMap* map, const void* base, int max_boards);
// weights: bf16 [9 taps][128 out][cin], BatchNorm already folded
int tc_make_weight_map(CUtensorMap* map, const void* base, int cin);
// out = act( conv3x3(in) + bias (+ residual) ); the board count is read from n_boards_dev when non-null
int tc_conv3x3_launch(cudaStream_t stream, const CUtensorMap* in_map, const CUtensorMap* w_map, int cin, const float* bias,
                      const void* residual, void* out, const int* n_boards_dev, int n_boards_static, int relu, int grid, int dbg = 0);   // dbg bit 5 (32): hand accumulators back with a release arrive
// the 20-layer residual tower in one persistent launch; maps_dev = device array {act0, act1, act2, w[0..19]}
int tc_tower_launch(cudaStream_t stream, const CUtensorMap* maps_dev, const float* bias, void* const* act, const int* n_boards_dev,
                    int n_boards_static, int n_layers, int stem, int grid, int tile_lo = 0, int tile_hi = 0x7FFFFFFF, int range_tiles = 0,
                    int release_arrive = 0, int wide = 0, int l2_hint = 0);   // wide: one 16-file box per channel half (maps_dev[25..27])
}  // namespace azb
