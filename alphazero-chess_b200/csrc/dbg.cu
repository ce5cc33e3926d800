// Debug / unit-test entry points (device pointers in, no engine needed).  Declared in include/az_b200.h.
#include "nn_tc.h"
#include <cstdio>
#include <cstdlib>

extern "C" int az_dbg_conv3x3_tc(const void* in_bf16, int cin, const void* w_bf16, const float* bias, const void* residual,
                                 void* out_bf16, int n_boards, int relu, int iters, float* ms_out) {
    CUtensorMap in_map, w_map;
    const char* dbg_env = getenv("AZ_DBG_CONV");
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    int r = azb::tc_make_act_map(&in_map, in_bf16, cin, n_boards, cin == 64 ? 0 : (dbg & 256) ? 2 : (dbg & 64) && cin == 128 ? 1 : 0);
    if (r) return r;
    r = azb::tc_make_weight_map(&w_map, w_bf16, cin);
    if (r) return r;
    const char* grid_env = getenv("AZ_DBG_GRID");
    const int grid = grid_env ? atoi(grid_env) : 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    r = azb::tc_conv3x3_launch(0, &in_map, &w_map, cin, bias, residual, out_bf16, nullptr, n_boards, relu, grid, dbg);
    if (r) return r;
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "az_dbg_conv3x3_tc: %s\n", cudaGetErrorString(cudaGetLastError())); return -10; }
    if (iters > 0) {
        cudaEventRecord(e0, 0);
        for (int i = 0; i < iters; i++)
            azb::tc_conv3x3_launch(0, &in_map, &w_map, cin, bias, residual, out_bf16, nullptr, n_boards, relu, grid, dbg);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -11;
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms_out) *ms_out = ms / iters;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
