// Policy and value heads of agent.rs:124-141 for the bf16 path, fused into one kernel on warp-level tensor-core MMAs:
//   stage 1  [64 squares x 128 ch] x [128 x 40]   policy_conv_1 (32) and value_conv (8), BatchNorm folded, ReLU
//   stage 2  [64 co x 32] x [32 x 64 squares]     policy_conv_2, kept transposed so that logits land as [co][square]
//   softmax over the 4096 logits in registers (two warps share a board, 32 squares each), probabilities written as 32-byte sectors
//   value    [32 boards x 512] x [512 x 64] -> ReLU -> 64 -> 1 -> tanh, once per group of four rounds of a block
// The policy part of a board needs only its own pair of warps (named barrier of 64 threads), so the pairs of a block run
// through the rounds of a group without block-wide barriers; only the value head, which batches the boards of the whole group
// into two M = 16 MMA tiles, synchronises the block (ncu of the first version: "barrier" was the top stall reason).
// The heads are 0.3 % of the network's FLOPs; they use mma.sync (a warp reads its rows straight from global memory, no shared-memory staging of
// the activations) while the 99 % in the tower runs on tcgen05 (nn_tc.cu).
#include "nn.h"
#include "device_once.h"
#include <algorithm>

namespace azb {

constexpr int HW40_PITCH = 136;  // bf16 elements per row of W40 in shared memory (128 + 8: conflict-free fragment loads)
constexpr int HW2_PITCH = 40;    // 32 + 8
constexpr int HV1_PITCH = 520;   // 512 + 8
constexpr int HEXP_PITCH = 72;   // floats per output-channel row of the per-board softmax tile: 64 + 8, so that the float2 stores of
                                 // a half-warp (rows g = 0..3, columns 2 tig) hit 16 distinct bank pairs (68 gave two-way conflicts)
constexpr int HEADS_GROUP = 4;   // rounds of boards whose value heads are evaluated together (32 board rows = two M = 16 tiles)
constexpr int HEADS_DYN_SMEM = 8 * 64 * HEXP_PITCH * 4 + HEADS_GROUP * 8 * HV1_PITCH * 2;

__device__ __forceinline__ void mma_bf16_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

// Two warps share a board (32 squares each) so that 16 warps are resident per SM: the kernel is latency-bound (a
// handful of dependent MMA / shared-memory / global steps per board), and halving the per-warp accumulators (D2 is
// [64 co][32 squares] = 64 registers) is what buys the second set of warps.
__device__ __forceinline__ void pair_sync(int board_slot) { asm volatile("bar.sync %0, 64;" ::"r"(board_slot + 1) : "memory"); }

__global__ void __launch_bounds__(512, 1)
k_heads_mma(const __nv_bfloat16* __restrict__ tower, const __nv_bfloat16* __restrict__ w40, const float* __restrict__ b40,
            const __nv_bfloat16* __restrict__ wp2, const float* __restrict__ bp2, const __nv_bfloat16* __restrict__ wl1t,
            const float* __restrict__ bl1, const float* __restrict__ wl2, const float* __restrict__ bl2, float* __restrict__ policy_out,
            float* __restrict__ value_out, const int* __restrict__ n_dev, int n_static, HeadScatter sc, int bpi) {
    extern __shared__ float s_exp_all[];  // [8 boards][64 co][HEXP_PITCH]: exp(logit - max) of the board in flight, then s_v1
    __shared__ __align__(16) __nv_bfloat16 s_w40[40 * HW40_PITCH];
    __shared__ __align__(16) __nv_bfloat16 s_w2[64 * HW2_PITCH];
    __nv_bfloat16* s_v1 = reinterpret_cast<__nv_bfloat16*>(s_exp_all + 8 * 64 * HEXP_PITCH);   // [HEADS_GROUP * 8 board rows][HV1_PITCH]
    __shared__ float s_b40[40], s_b2[64], s_red[2][8][2], s_hid[2][HEADS_GROUP * 8][64], s_vsum[HEADS_GROUP][16];
    const int n = n_dev ? *n_dev : n_static;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, g = lane >> 2, tig = lane & 3;
    const int bs = warp >> 1, h = warp & 1;  // board slot in the block, square half
    // weights to shared memory with 16-byte copies (8 bf16): 40 rows x 16 chunks and 64 rows x 4 chunks
    for (int i = t; i < 40 * 16; i += 512)
        *reinterpret_cast<uint4*>(s_w40 + (i >> 4) * HW40_PITCH + (i & 15) * 8) = __ldg(reinterpret_cast<const uint4*>(w40) + i);
    if (t < 64 * 4)
        *reinterpret_cast<uint4*>(s_w2 + (t >> 2) * HW2_PITCH + (t & 3) * 8) = __ldg(reinterpret_cast<const uint4*>(wp2) + t);
    if (t < 40) s_b40[t] = b40[t];
    if (t < 64) s_b2[t] = bp2[t];
    __syncthreads();

    // bpi = boards per block and round (<= 8): the host picks it so that the last round is full (4096 boards on 148 blocks:
    // 4 rounds of 7 instead of 3.46 -> 4 rounds of 8)
    const int round_stride = gridDim.x * bpi;
    for (int base0 = blockIdx.x * bpi; base0 < n; base0 += round_stride * HEADS_GROUP) {
      for (int r = 0; r < HEADS_GROUP; r++) {
        const int base = base0 + r * round_stride;
        const int b = base + bs;
        const bool active = bs < bpi && b < n;
        __nv_bfloat16* vr = s_v1 + (r * 8 + bs) * HV1_PITCH;   // this board's row of the value head's input
        if (active) {
            // ---------------- stage 1: D1[square][40] for this warp's 32 squares
            float d1[2][5][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 5; nt++)
#pragma unroll
                    for (int i = 0; i < 4; i++) d1[mt][nt][i] = 0.0f;
            // A fragments with 16-byte loads: thread (g, tig) fetches channels blk*32 + tig*8 .. +7 of its two rows and
            // uses them for TWO k-steps.  The K order inside a 32-channel block is therefore permuted (k-step A takes
            // sub-channels 0-3 of every thread's group, k-step B 4-7); the weight fragments below use the same
            // permutation, which leaves the dot products unchanged.
            const uint4* X = reinterpret_cast<const uint4*>(tower + (size_t)b * 64 * 128) + h * 32 * 16;  // [square][16 x 16 B]
            uint4 lo[4][2], hi[4][2];
#pragma unroll
            for (int blk = 0; blk < 4; blk++)
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    lo[blk][mt] = __ldg(X + (mt * 16 + g) * 16 + blk * 4 + tig);
                    hi[blk][mt] = __ldg(X + (mt * 16 + g + 8) * 16 + blk * 4 + tig);
                }
#pragma unroll
            for (int blk = 0; blk < 4; blk++) {
                uint2 bfA[5], bfB[5];
#pragma unroll
                for (int nt = 0; nt < 5; nt++) {
                    const uint4 wv = *reinterpret_cast<const uint4*>(s_w40 + (nt * 8 + g) * HW40_PITCH + blk * 32 + tig * 8);
                    bfA[nt] = make_uint2(wv.x, wv.y);
                    bfB[nt] = make_uint2(wv.z, wv.w);
                }
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                    for (int nt = 0; nt < 5; nt++)
                        mma_bf16_16816(d1[mt][nt], lo[blk][mt].x, hi[blk][mt].x, lo[blk][mt].y, hi[blk][mt].y, bfA[nt].x, bfA[nt].y);
#pragma unroll
                    for (int nt = 0; nt < 5; nt++)
                        mma_bf16_16816(d1[mt][nt], lo[blk][mt].z, hi[blk][mt].z, lo[blk][mt].w, hi[blk][mt].w, bfB[nt].x, bfB[nt].y);
                }
            }
            // bias + ReLU; value hidden (columns 32..39) to shared memory as the flattened [c*64 + square] row
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                for (int nt = 0; nt < 5; nt++) {
                    const float bb0 = s_b40[nt * 8 + tig * 2], bb1 = s_b40[nt * 8 + tig * 2 + 1];
                    d1[mt][nt][0] = fmaxf(d1[mt][nt][0] + bb0, 0.0f);
                    d1[mt][nt][1] = fmaxf(d1[mt][nt][1] + bb1, 0.0f);
                    d1[mt][nt][2] = fmaxf(d1[mt][nt][2] + bb0, 0.0f);
                    d1[mt][nt][3] = fmaxf(d1[mt][nt][3] + bb1, 0.0f);
                }
                const int sq = h * 32 + mt * 16 + g, c = tig * 2;
                vr[c * 64 + sq] = __float2bfloat16_rn(d1[mt][4][0]);
                vr[(c + 1) * 64 + sq] = __float2bfloat16_rn(d1[mt][4][1]);
                vr[c * 64 + sq + 8] = __float2bfloat16_rn(d1[mt][4][2]);
                vr[(c + 1) * 64 + sq + 8] = __float2bfloat16_rn(d1[mt][4][3]);
            }
            // ---------------- stage 2 (transposed): D2[co][square] = W2[co][k] * P1[square][k]
            float d2[4][4][4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++)
#pragma unroll
                    for (int i = 0; i < 4; i++) d2[mt][nt][i] = 0.0f;
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                uint32_t bfr[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {  // squares 8nt..8nt+7 of this half come from D1 m-tile nt>>1, row half nt&1
                    const int h2 = (nt & 1) * 2;
                    bfr[nt][0] = pack_bf16(d1[nt >> 1][2 * ks][h2], d1[nt >> 1][2 * ks][h2 + 1]);
                    bfr[nt][1] = pack_bf16(d1[nt >> 1][2 * ks + 1][h2], d1[nt >> 1][2 * ks + 1][h2 + 1]);
                }
#pragma unroll
                for (int mt = 0; mt < 4; mt++) {
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_w2 + (mt * 16 + g) * HW2_PITCH + ks * 16 + tig * 2);
                    const uint32_t a0 = wp[0], a1 = wp[8 * HW2_PITCH / 2], a2 = wp[4], a3 = wp[8 * HW2_PITCH / 2 + 4];
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) mma_bf16_16816(d2[mt][nt], a0, a1, a2, a3, bfr[nt][0], bfr[nt][1]);
                }
            }
            // ---------------- softmax over the board's 4096 logits (agent.rs:130), two warps per board
            float mx = -INFINITY;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const float c0 = s_b2[mt * 16 + g], c1 = s_b2[mt * 16 + g + 8];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    d2[mt][nt][0] += c0; d2[mt][nt][1] += c0; d2[mt][nt][2] += c1; d2[mt][nt][3] += c1;
                    mx = fmaxf(mx, fmaxf(fmaxf(d2[mt][nt][0], d2[mt][nt][1]), fmaxf(d2[mt][nt][2], d2[mt][nt][3])));
                }
            }
            for (int d = 16; d; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            if (lane == 0) s_red[0][bs][h] = mx;
            pair_sync(bs);
            mx = fmaxf(s_red[0][bs][0], s_red[0][bs][1]);
            float sum = 0.0f;
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++)
#pragma unroll
                    for (int i = 0; i < 4; i++) { d2[mt][nt][i] = __expf(d2[mt][nt][i] - mx); sum += d2[mt][nt][i]; }
            for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            if (lane == 0) s_red[1][bs][h] = sum;
            if (sc.edge_P) {  // this half of the exp tile; the scatter below reads both halves
                float* se = s_exp_all + bs * 64 * HEXP_PITCH;
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) {
                        const int co = mt * 16 + g, sq = h * 32 + nt * 8 + tig * 2;
                        *reinterpret_cast<float2*>(se + co * HEXP_PITCH + sq) = make_float2(d2[mt][nt][0], d2[mt][nt][1]);
                        *reinterpret_cast<float2*>(se + (co + 8) * HEXP_PITCH + sq) = make_float2(d2[mt][nt][2], d2[mt][nt][3]);
                    }
            }
            pair_sync(bs);
            const float inv = __fdividef(1.0f, s_red[1][bs][0] + s_red[1][bs][1]);
            if (sc.edge_P) {  // priors of the legal moves only, written into the tree (tree.rs:84-104 reads nothing else)
                const float* se = s_exp_all + bs * 64 * HEXP_PITCH;
                const unsigned long long eo = sc.edge_off[b];
                const int L = sc.n_edges[b];
                for (int e2 = h * 32 + lane; e2 < L; e2 += 64) {
                    const uint32_t idx = sc.edge_mv[eo + e2] >> 16;
                    sc.edge_P[eo + e2] = se[(idx >> 6) * HEXP_PITCH + (idx & 63)] * inv;
                }
            }
            if (policy_out) {
                float* po = policy_out + (size_t)b * 4096;
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) {
                        const int co = mt * 16 + g, sq = h * 32 + nt * 8 + tig * 2;
                        *reinterpret_cast<float2*>(po + co * 64 + sq) = make_float2(d2[mt][nt][0] * inv, d2[mt][nt][1] * inv);
                        *reinterpret_cast<float2*>(po + (co + 8) * 64 + sq) = make_float2(d2[mt][nt][2] * inv, d2[mt][nt][3] * inv);
                    }
            }
        } else {
            for (int i = h * 32 + lane; i < 512; i += 64) vr[i] = __float2bfloat16_rn(0.0f);
        }
      }
        // ---------------- value head of the group's 32 board rows (row = round * 8 + slot): warp w sums K half (w >> 3) of
        // hidden[row][8(w&7) .. +7]; M tile m holds rows 16m .. 16m + 15 (rounds 2m and 2m + 1)
        {
            float acc[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
            const int hw = warp & 7, kh = warp >> 3;
            const uint32_t* W = reinterpret_cast<const uint32_t*>(wl1t + (size_t)(hw * 8 + g) * 512) + kh * 128;
            // the weight fragments do not depend on the boards: requested before the barrier (the policy accumulators are
            // dead by now), so their L2 latency hides behind the wait for the slowest pair
            uint32_t wf[16][2];
#pragma unroll
            for (int ks = 0; ks < 16; ks++) { wf[ks][0] = __ldg(W + ks * 8 + tig); wf[ks][1] = __ldg(W + ks * 8 + tig + 4); }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const uint32_t* V0 = reinterpret_cast<const uint32_t*>(s_v1 + (16 * m + g) * HV1_PITCH) + kh * 128;
                const uint32_t* V1 = reinterpret_cast<const uint32_t*>(s_v1 + (16 * m + g + 8) * HV1_PITCH) + kh * 128;
#pragma unroll
                for (int ks = 0; ks < 16; ks++)
                    mma_bf16_16816(acc[m], V0[ks * 8 + tig], V1[ks * 8 + tig], V0[ks * 8 + tig + 4], V1[ks * 8 + tig + 4], wf[ks][0], wf[ks][1]);
                const int hn = hw * 8 + tig * 2;   // acc[0], acc[1]: row 16m + g; acc[2], acc[3]: row 16m + g + 8; hidden units hn, hn + 1
                s_hid[kh][16 * m + g][hn] = acc[m][0];
                s_hid[kh][16 * m + g][hn + 1] = acc[m][1];
                s_hid[kh][16 * m + g + 8][hn] = acc[m][2];
                s_hid[kh][16 * m + g + 8][hn + 1] = acc[m][3];
            }
        }
        __syncthreads();
        {   // 32 rows x 64 hidden units on 512 threads: ReLU(sum of the two K halves + bias) * w2, reduced per row
            const int hn = t & 63;
            const float b1 = bl1[hn], w2 = wl2[hn];
#pragma unroll
            for (int r = 0; r < HEADS_GROUP; r++) {
                const int row = r * 8 + (t >> 6);
                float part = fmaxf(s_hid[0][row][hn] + s_hid[1][row][hn] + b1, 0.0f) * w2;
                for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                if (lane == 0) s_vsum[r][warp] = part;
            }
        }
        __syncthreads();
        if (t < HEADS_GROUP * 8) {
            const int r = t >> 3, slot = t & 7, board = base0 + r * round_stride + slot;
            if (slot < bpi && board < n) value_out[board] = tanhf(bl2[0] + s_vsum[r][2 * slot] + s_vsum[r][2 * slot + 1]);
        }
    }
}

int launch_heads_mma(az_engine* e, const __nv_bfloat16* tower, const int* n_dev, int n_static, float* policy_out, float* value_out,
                     const HeadScatter* scatter) {
    NetWeights* w = e->net;
    const int n_max = n_dev ? w->max_boards : n_static;
    if (n_max <= 0) return 0;
    int bpi = 8;
    if (n_max > 8 * e->sm_count) {   // several rounds per block: fewest board-slots (rounds x bpi) that cover the batch
        int best = 1 << 30;
        for (int c = 8; c >= 5; c--) {
            const int rounds = (n_max + e->sm_count * c - 1) / (e->sm_count * c);
            if (rounds * c < best) { best = rounds * c; bpi = c; }
        }
    }
    const int grid = std::min((n_max + bpi - 1) / bpi, e->sm_count);
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(k_heads_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, HEADS_DYN_SMEM);
    HeadScatter sc{nullptr, nullptr, nullptr, nullptr};
    if (scatter) sc = *scatter;
    k_heads_mma<<<grid, 512, HEADS_DYN_SMEM, e->stream>>>(tower, w->h_w40, w->f_b40, w->h_wp2, w->f_bp2, w->h_wl1t, w->f_bl1, w->f_wl2, w->f_bl2,
                                             policy_out, value_out, n_dev, n_static, sc, bpi);
    return check_cuda(e, cudaGetLastError(), "k_heads_mma");
}

}  // namespace azb
