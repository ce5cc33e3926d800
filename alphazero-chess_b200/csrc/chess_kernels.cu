// Batched chess.rs kernels: legal move generation, perft, play_move, the index codec and the f32 input planes.
// One thread per position for the rule kernels (branch-minimised bitboard code, move lists staged in shared memory and
// written back coalesced); one warp per position for the plane encoder.
#include "engine.h"
#include <climits>

namespace azb {

// ---------------------------------------------------------------------------------------------------- movegen
constexpr int MG_THREADS = 64;

__global__ void __launch_bounds__(MG_THREADS) k_movegen(const az_position* __restrict__ wire, int n, uint16_t* __restrict__ moves_out,
                                                        uint16_t* __restrict__ index_out, int32_t* __restrict__ count_out) {
    __shared__ __align__(16) uint16_t s_moves[MG_THREADS][AZ_MAX_MOVES];
    __shared__ int s_count[MG_THREADS];
    __shared__ int s_turn[MG_THREADS];
    const int i = blockIdx.x * MG_THREADS + threadIdx.x;
    int cnt = 0, turn = 0;
    if (i < n) {
        DPos p = dpos_from_wire(wire[i]);
        ListSink sink{s_moves[threadIdx.x], 0};
        gen_legal(p, sink);
        cnt = sink.n;
        turn = meta_turn(p.meta);
        count_out[i] = cnt;
    }
    s_count[threadIdx.x] = cnt;
    s_turn[threadIdx.x] = turn;
    __syncthreads();
    // coalesced write-back of the whole [64][256] tile
    const int base = blockIdx.x * MG_THREADS;
    for (int e = threadIdx.x; e < MG_THREADS * AZ_MAX_MOVES; e += MG_THREADS) {
        int row = e / AZ_MAX_MOVES, col = e % AZ_MAX_MOVES;
        if (base + row >= n) break;
        uint16_t m = col < s_count[row] ? s_moves[row][col] : (uint16_t)AZ_MOVE_NONE;
        size_t o = (size_t)(base + row) * AZ_MAX_MOVES + col;
        moves_out[o] = m;
        if (index_out) index_out[o] = col < s_count[row] ? (uint16_t)move_to_index(m, s_turn[row]) : (uint16_t)0xFFFF;
    }
}

void launch_movegen(cudaStream_t s, const az_position* wire, int n, uint16_t* moves, uint16_t* index, int32_t* count) {
    if (n <= 0) return;
    k_movegen<<<(n + MG_THREADS - 1) / MG_THREADS, MG_THREADS, 0, s>>>(wire, n, moves, index, count);
}

// one warp per position: the cooperative generator the search kernel uses (same output contract as k_movegen)
__global__ void __launch_bounds__(128) k_movegen_warp(const az_position* __restrict__ wire, int n, uint16_t* __restrict__ moves_out,
                                                      int32_t* __restrict__ count_out) {
    __shared__ uint16_t s_moves[4][AZ_MAX_MOVES];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + w;
    if (i >= n) return;
    DPos p = dpos_from_wire(wire[i]);
    int cnt = 0;
    warp_gen_legal(p, s_moves[w], lane, cnt);
    for (int k = lane; k < AZ_MAX_MOVES; k += 32) moves_out[(size_t)i * AZ_MAX_MOVES + k] = k < cnt ? s_moves[w][k] : (uint16_t)AZ_MOVE_NONE;
    if (lane == 0) count_out[i] = cnt;
}
void launch_movegen_warp(cudaStream_t s, const az_position* wire, int n, uint16_t* moves, int32_t* count) {
    if (n > 0) k_movegen_warp<<<(n + 3) / 4, 128, 0, s>>>(wire, n, moves, count);
}

// ---------------------------------------------------------------------------------------------------- play_move
// chess.rs:36-63 driven by a policy index (tree.rs:211-212)
__global__ void k_play_move(az_position* __restrict__ wire, const az_position* __restrict__ hist, const uint32_t* __restrict__ hist_off,
                            const uint16_t* __restrict__ action, int32_t* __restrict__ result, int n, RuleParams rp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DPos p = dpos_from_wire(wire[i]);
    uint16_t legal[AZ_MAX_MOVES];
    ListSink sink{legal, 0};
    gen_legal(p, sink);
    uint16_t mv = index_to_move(action[i], p, legal, sink.n);
    if (mv == 0xFFFF) { result[i] = AZ_RESULT_ILLEGAL; return; }
    DPos c = make_move(p, mv);
    CountSink cs{0};
    GenInfo gi = gen_legal(c, cs);
    wire[i] = dpos_to_wire(c);
    if (cs.n == 0) {
        result[i] = gi.checkers ? (meta_turn(c.meta) == 0 ? AZ_RESULT_BLACK_WINS : AZ_RESULT_WHITE_WINS) : AZ_RESULT_DRAW;
        return;
    }
    if (insufficient_material(c)) { result[i] = AZ_RESULT_DRAW; return; }
    set_key_bits(c, gi.has_legal_ep);
    int count = 1;
    if (hist) {
        for (uint32_t h = hist_off[i]; h < hist_off[i + 1]; h++) {
            DPos q = dpos_from_wire(hist[h]);
            if (q.pawn != c.pawn || q.white != c.white || q.knight != c.knight || q.bishop != c.bishop || q.rook != c.rook ||
                q.queen != c.queen || q.king != c.king || ((q.meta ^ c.meta) & 0x1F))
                continue;
            bool q_ep = false;
            if (meta_ep(q.meta) >= 0) { CountSink t{0}; q_ep = gen_legal(q, t).has_legal_ep; }
            set_key_bits(q, q_ep);
            if (same_position_key(q, c)) count++;
        }
    }
    const bool ongoing = count < rp.repetitions && meta_halfmoves(c.meta) < rp.num_halfmoves && meta_fullmoves(c.meta) < rp.num_fullmoves;
    result[i] = ongoing ? AZ_RESULT_ONGOING : AZ_RESULT_DRAW;
}

void launch_play_move(cudaStream_t s, az_position* wire, const az_position* hist, const uint32_t* hist_off, const uint16_t* action,
                      int32_t* result, int n, RuleParams rp) {
    if (n <= 0) return;
    k_play_move<<<(n + 63) / 64, 64, 0, s>>>(wire, hist, hist_off, action, result, n, rp);
}

// ---------------------------------------------------------------------------------------------------- codec
__global__ void k_move_to_index(const az_position* __restrict__ wire, const uint16_t* __restrict__ moves, uint16_t* __restrict__ index, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) index[i] = (uint16_t)move_to_index(moves[i], wire[i].turn & 1);
}
__global__ void k_index_to_move(const az_position* __restrict__ wire, const uint16_t* __restrict__ index, uint16_t* __restrict__ moves, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DPos p = dpos_from_wire(wire[i]);
    uint16_t legal[AZ_MAX_MOVES];
    ListSink sink{legal, 0};
    gen_legal(p, sink);
    moves[i] = index[i] < AZ_ACTION_SPACE ? index_to_move(index[i], p, legal, sink.n) : (uint16_t)0xFFFF;
}
void launch_move_to_index(cudaStream_t s, const az_position* wire, const uint16_t* moves, uint16_t* index, int n) {
    if (n > 0) k_move_to_index<<<(n + 127) / 128, 128, 0, s>>>(wire, moves, index, n);
}
void launch_index_to_move(cudaStream_t s, const az_position* wire, const uint16_t* index, uint16_t* moves, int n) {
    if (n > 0) k_index_to_move<<<(n + 63) / 64, 64, 0, s>>>(wire, index, moves, n);
}

// ---------------------------------------------------------------------------------------------------- planes (f32)
// to_tensor (chess.rs:191-245): one warp per position, 1216 coalesced floats
__global__ void k_encode_f32(const az_position* __restrict__ wire, float* __restrict__ planes, int n) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    DPos p = dpos_from_wire(wire[warp]);
    const u64 occ = occupied(p);
    const u64 ours = meta_turn(p.meta) == 0 ? p.white : occ ^ p.white;
    const int pep = pseudo_legal_ep(p);
    float* out = planes + (size_t)warp * (AZ_NUM_PLANES * 64);
    for (int e = lane; e < AZ_NUM_PLANES * 64; e += 32) out[e] = plane_value(p, e >> 6, e & 63, pep, ours, occ ^ ours);
}
void launch_encode_f32(cudaStream_t s, const az_position* wire, float* planes, int n) {
    if (n > 0) k_encode_f32<<<(n + 3) / 4, 128, 0, s>>>(wire, planes, n);
}

// ---------------------------------------------------------------------------------------------------- perft
__global__ void k_wire_to_dpos(const az_position* __restrict__ wire, DPos* __restrict__ out, uint32_t* __restrict__ root_ids, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = dpos_from_wire(wire[i]); if (root_ids) root_ids[i] = i; }
}
void launch_wire_to_dpos(cudaStream_t s, const az_position* wire, DPos* out, uint32_t* root_ids, int n) {
    if (n > 0) k_wire_to_dpos<<<(n + 127) / 128, 128, 0, s>>>(wire, out, root_ids, n);
}

// interior ply: every thread expands one position into its children; slots are reserved with one atomic per warp
__global__ void __launch_bounds__(128) k_perft_expand(const DPos* __restrict__ in, const uint32_t* __restrict__ in_root, int n_in,
                                                      DPos* __restrict__ out, uint32_t* __restrict__ out_root,
                                                      unsigned long long* __restrict__ out_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint16_t legal[AZ_MAX_MOVES];
    int cnt = 0;
    DPos p;
    uint32_t root = 0;
    if (i < n_in) {
        p = in[i];
        root = in_root[i];
        ListSink sink{legal, 0};
        gen_legal(p, sink);
        cnt = sink.n;
    }
    // warp exclusive scan of child counts
    int incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(out_count, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    unsigned long long o = base + (unsigned long long)(incl - cnt);
    for (int k = 0; k < cnt; k++) { out[o + k] = make_move(p, legal[k]); out_root[o + k] = root; }
}
void launch_perft_expand(cudaStream_t s, const DPos* in, const uint32_t* in_root, int n_in, DPos* out, uint32_t* out_root,
                         unsigned long long* out_count) {
    if (n_in > 0) k_perft_expand<<<(n_in + 127) / 128, 128, 0, s>>>(in, in_root, n_in, out, out_root, out_count);
}

// last ply: bulk count (popcounts, no move list), reduced per block when the block belongs to one root
__global__ void __launch_bounds__(256) k_perft_count(const DPos* __restrict__ in, const uint32_t* __restrict__ in_root, int n_in,
                                                     unsigned long long* __restrict__ nodes) {
    __shared__ unsigned long long s_sum;
    __shared__ uint32_t s_root;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    uint32_t root = 0xFFFFFFFFu;
    if (i < n_in) {
        DPos p = in[i];
        root = in_root[i];
        CountSink cs{0};
        gen_legal(p, cs);
        cnt = cs.n;
    }
    if (threadIdx.x == 0) { s_sum = 0; s_root = root; }
    __syncthreads();
    const bool uniform = __syncthreads_and(root == s_root || root == 0xFFFFFFFFu);
    if (uniform) {
        unsigned int w = __reduce_add_sync(0xffffffffu, (unsigned int)cnt);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_sum, (unsigned long long)w);
        __syncthreads();
        if (threadIdx.x == 0 && s_sum) atomicAdd(&nodes[s_root], s_sum);
    } else if (cnt) {
        atomicAdd(&nodes[root], (unsigned long long)cnt);
    }
}
void launch_perft_count(cudaStream_t s, const DPos* in, const uint32_t* in_root, int n_in, unsigned long long* nodes) {
    if (n_in > 0) k_perft_count<<<(n_in + 255) / 256, 256, 0, s>>>(in, in_root, n_in, nodes);
}

// ------------------------------------------------------------------------------------------------- minimax player
// chess.rs:247-318 (evaluate_material / negamax / get_best_move), full width and unpruned as in the reference, as a
// breadth-first sweep over level buffers: expand, score the horizon, then fold the children back with one integer
// atomicMax per child (order-independent, so the result is deterministic).
__device__ __forceinline__ int material_for_mover(const DPos& p) {  // chess.rs:254-264
    const u64 occ = occupied(p);
    const u64 ours = meta_turn(p.meta) == 0 ? p.white : occ ^ p.white;
    const u64 theirs = occ ^ ours;
    return 100 * (popc(p.pawn & ours) - popc(p.pawn & theirs)) + 320 * (popc(p.knight & ours) - popc(p.knight & theirs)) +
           330 * (popc(p.bishop & ours) - popc(p.bishop & theirs)) + 500 * (popc(p.rook & ours) - popc(p.rook & theirs)) +
           900 * (popc(p.queen & ours) - popc(p.queen & theirs));
}

// interior node with `depth` plies left (depth >= 1): a finished game scores itself (chess.rs:267-277), anything else is
// expanded and starts at INT_MIN.  The root (is_root) is never treated as finished: get_best_move walks its legal moves
// without asking (chess.rs:295-303).
__global__ void __launch_bounds__(128) k_mm_expand(const DPos* __restrict__ in, int n_in, int depth, int is_root, DPos* __restrict__ out,
                                                   uint32_t* __restrict__ out_parent, unsigned long long* __restrict__ out_count,
                                                   int* __restrict__ score, unsigned int* __restrict__ first_child,
                                                   int* __restrict__ n_child) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint16_t legal[AZ_MAX_MOVES];
    int cnt = 0;
    DPos p;
    if (i < n_in) {
        p = in[i];
        ListSink sink{legal, 0};
        const GenInfo gi = gen_legal(p, sink);
        cnt = sink.n;
        if (!is_root) {
            if (cnt == 0) score[i] = gi.checkers ? -20000 - depth : 0;
            else if (insufficient_material(p)) { score[i] = 0; cnt = 0; }
            else score[i] = INT_MIN;
        }
    }
    int incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(out_count, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    const unsigned long long o = base + (unsigned long long)(incl - cnt);
    if (i < n_in && first_child) { first_child[i] = (unsigned int)o; n_child[i] = cnt; }
    for (int k = 0; k < cnt; k++) { out[o + k] = make_move(p, legal[k]); out_parent[o + k] = (uint32_t)i; }
}

// horizon (depth 0): mate / stalemate / insufficient material, else material (chess.rs:267-280)
__global__ void __launch_bounds__(256) k_mm_leaf(const DPos* __restrict__ in, int n_in, int* __restrict__ score) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const DPos p = in[i];
    CountSink cs{0};
    const GenInfo gi = gen_legal(p, cs);
    int sc;
    if (cs.n == 0) sc = gi.checkers ? -20000 : 0;
    else if (insufficient_material(p)) sc = 0;
    else sc = material_for_mover(p);
    score[i] = sc;
}

__global__ void __launch_bounds__(256) k_mm_backup(const int* __restrict__ child_score, const uint32_t* __restrict__ child_parent,
                                                   int n_children, int* __restrict__ parent_score) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_children) atomicMax(&parent_score[child_parent[j]], -child_score[j]);
}

// scores of the root's legal moves, in legal-move order (the random choice among the best stays with the caller)
__global__ void __launch_bounds__(256) k_mm_root(const unsigned int* __restrict__ first_child, const int* __restrict__ n_child,
                                                 const int* __restrict__ child_score, int n_roots, int* __restrict__ scores_out,
                                                 int* __restrict__ count_out) {
    const int r = blockIdx.x, k = threadIdx.x;
    if (r >= n_roots) return;
    const int c = n_child[r];
    scores_out[(size_t)r * AZ_MAX_MOVES + k] = k < c ? -child_score[first_child[r] + k] : 0;
    if (k == 0) count_out[r] = c;
}

void launch_mm_expand(cudaStream_t s, const DPos* in, int n_in, int depth, int is_root, DPos* out, uint32_t* out_parent,
                      unsigned long long* out_count, int* score, unsigned int* first_child, int* n_child) {
    if (n_in > 0) k_mm_expand<<<(n_in + 127) / 128, 128, 0, s>>>(in, n_in, depth, is_root, out, out_parent, out_count, score, first_child, n_child);
}
void launch_mm_leaf(cudaStream_t s, const DPos* in, int n_in, int* score) {
    if (n_in > 0) k_mm_leaf<<<(n_in + 255) / 256, 256, 0, s>>>(in, n_in, score);
}
void launch_mm_backup(cudaStream_t s, const int* child_score, const uint32_t* child_parent, int n_children, int* parent_score) {
    if (n_children > 0) k_mm_backup<<<(n_children + 255) / 256, 256, 0, s>>>(child_score, child_parent, n_children, parent_score);
}
void launch_mm_root(cudaStream_t s, const unsigned int* first_child, const int* n_child, const int* child_score, int n_roots,
                    int* scores_out, int* count_out) {
    if (n_roots > 0) k_mm_root<<<n_roots, AZ_MAX_MOVES, 0, s>>>(first_child, n_child, child_score, n_roots, scores_out, count_out);
}

}  // namespace azb
