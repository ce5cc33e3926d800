// Device-resident PUCT search and self-play: one warp per game, struct-of-arrays tree pools in HBM.
//
// Restates tree.rs (MCTree::new / simulation / expand / traverse_new / apply_dirichlet_noise / max_subtree_depth)
// and training.rs run_episode with exactly the reference's arithmetic: one simulation in flight per game
// (tree.rs:170-172), PUCT = q + ((c*P)*sqrt(total))/(1+N) evaluated left to right in f32 with strict '>' and
// first-max-in-move-order ties (tree.rs:184-195), W += v ; N += 1 on the way back with a sign flip per level
// (tree.rs:197-206), terminal children never stored (tree.rs:233-234), no tree reuse between moves (tree.rs:239-256).
// This translation unit is compiled with -fmad=false and uses explicit _rn intrinsics so no FMA contraction can
// change a bit.  A "wave" lets every game run until it needs a network evaluation; the evaluations of all games form
// one batch (the reference's recv_many batch, training.rs:369).
#include "mcts.h"
#include "nn.h"
#include "det_math.cuh"
#include <cstdio>
#include <cstring>
#include <algorithm>

namespace azb {

constexpr int WARPS = 4;

struct WarpShared {
    double gam[AZ_MAX_MOVES];
    float fval[AZ_MAX_MOVES];
    uint16_t moves[AZ_MAX_MOVES];
    uint16_t sidx[AZ_MAX_MOVES];
    DPos child;
    DPos key;       // cache key of `child` (pseudo-legal ep, counters; no repetition bits)
    int n_moves;
    int term;       // 0 ongoing, 1 draw, 2 decisive (side to move in the child is mated)
    int has_ep;
    int pad;
};

// ------------------------------------------------------------------------------------------- deterministic RNG
// Counter-based generator + IEEE-exact log/exp, bit-identical to oracle/mcts.cpp (spec in DESIGN.md).
__device__ __forceinline__ u64 rng_u64(u64 seed, u64 game, u64 ply, u64 stream, u64 counter) {
    u64 h = splitmix64(seed);
    h = splitmix64(h ^ game);
    h = splitmix64(h ^ ply);
    h = splitmix64(h ^ stream);
    h = splitmix64(h ^ counter);
    return h;
}
__device__ __forceinline__ double rng_uniform(u64 seed, u64 game, u64 ply, u64 stream, u64 counter) {
    u64 h = rng_u64(seed, game, ply, stream, counter);
    return ((double)(h >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}
// Gamma(alpha, 1), Marsaglia-Tsang with the alpha < 1 boost (the scheme of rand_distr 0.4.3)
__device__ double gamma_sample(u64 seed, u64 game, u64 ply, u64 stream, double alpha) {
    u64 ctr = 0;
    double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double v, x;
    for (;;) {
        double u1, u2, s;
        do {
            u1 = 2.0 * rng_uniform(seed, game, ply, stream, ctr++) - 1.0;
            u2 = 2.0 * rng_uniform(seed, game, ply, stream, ctr++) - 1.0;
            s = u1 * u1 + u2 * u2;
        } while (s >= 1.0 || s == 0.0);
        x = u1 * sqrt(-2.0 * det_log(s) / s);
        v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = rng_uniform(seed, game, ply, stream, ctr++);
        if (det_log(u) < 0.5 * x * x + d - d * v + d * det_log(v)) break;
    }
    double g = d * v;
    if (alpha < 1.0) { double u = rng_uniform(seed, game, ply, stream, ctr++); g = g * det_exp(det_log(u) / alpha); }
    return g;
}

// ------------------------------------------------------------------------------------------- synthetic evaluator
__device__ __forceinline__ u64 stub_position_hash(u64 seed, const DPos& p) {
    const u64 occ = occupied(p);
    u64 h = splitmix64(seed);
    h = splitmix64(h ^ p.pawn); h = splitmix64(h ^ p.knight); h = splitmix64(h ^ p.bishop);
    h = splitmix64(h ^ p.rook); h = splitmix64(h ^ p.queen); h = splitmix64(h ^ p.king);
    h = splitmix64(h ^ p.white); h = splitmix64(h ^ (occ ^ p.white));
    const int ep = pseudo_legal_ep(p);
    u64 meta = (u64)meta_turn(p.meta) | ((u64)meta_castling(p.meta) << 8) | ((u64)(ep + 1) << 16) | ((u64)meta_halfmoves(p.meta) << 32) |
               ((u64)meta_fullmoves(p.meta) << 48);
    return splitmix64(h ^ meta);
}
// one warp per request: policy [4096] normalised integers, value in [-1, 1)
__global__ void k_stub_eval(const DPos* __restrict__ req_pos, const int* __restrict__ n_dev, int n_static, u64 seed,
                            float* __restrict__ policy, float* __restrict__ value) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n = n_dev ? *n_dev : n_static;
    if (warp >= n) return;
    const DPos p = req_pos[warp];
    const u64 h = stub_position_hash(seed, p);
    u64 total = 0;
    float* po = policy + (size_t)warp * AZ_ACTION_SPACE;
    for (int i = lane; i < AZ_ACTION_SPACE; i += 32) {
        u64 r = ((splitmix64(h + (u64)i) >> 52) << 12) + (u64)i + 1;
        total += r;
        po[i] = (float)r;
    }
    for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    const float tf = (float)total;
    __syncwarp();
    for (int i = lane; i < AZ_ACTION_SPACE; i += 32) po[i] = __fdiv_rn(po[i], tf);
    if (lane == 0) {
        u64 v = splitmix64(h ^ 0xA5A5A5A5A5A5A5A5ULL) >> 40;
        value[warp] = __fsub_rn(__fmul_rn((float)v, 1.0f / 8388608.0f), 1.0f);
    }
}

// ------------------------------------------------------------------------------------------- per-game helpers
struct Ctx {
    const SearchParams& prm;
    const SearchPtrs& ptr;
    WarpShared* sh;
    int g, lane;
    size_t nbase, ebase;   // first node / edge of this game
    GameCtl c;
    // statistics accumulated by lane 0
    unsigned int st_sims, st_pos, st_evals, st_term, st_games, st_depth, st_edges, st_hits, st_evict;
    unsigned int st_sdepth;        // sum of EpisodeStep::search_depth over the steps recorded (avg_search_depth, training.rs:91-97)
    unsigned long long new_link;   // edge_link value of the node create_node made last
};

// Expands `mv` from `parent` on lane 0: child position, its legal moves (into shared memory), mate / stalemate /
// insufficient material.  Repetition and move-count draws are decided afterwards by the whole warp.
__device__ __forceinline__ void lane0_make_child(Ctx& x, const DPos& parent, uint16_t mv) {
    // every lane builds the child (cheap, uniform), then the warp generates its legal moves cooperatively
    DPos child = make_move(parent, mv);
    int n = 0;
    const GenInfo gi = warp_gen_legal(child, x.sh->moves, x.lane, n);
    int term = 0;
    if (n == 0) term = gi.checkers ? 2 : 1;
    else if (insufficient_material(child)) term = 1;
    set_key_bits(child, gi.has_legal_ep);
    if (x.lane == 0) {
        x.sh->child = child;
        x.sh->n_moves = n;
        x.sh->term = term;
    }
    __syncwarp();
}

// chess.rs:52-60: pos_count[child] (this occurrence included) reaches REPETITIONS, or a move counter hits its limit.
// Candidates are the positions with the same side to move since the last irreversible move: they live on the
// search path (depth >= 1) or in the game history.  `depth_child` = edges from the root to the child.
__device__ __forceinline__ bool draw_by_rules(Ctx& x, const DPos& child, int depth_child) {
    const int hm = meta_halfmoves(child.meta);
    if (hm >= x.prm.num_halfmoves || meta_fullmoves(child.meta) >= x.prm.num_fullmoves) return true;
    const int root_v = (int)x.c.hist_len - 1;          // virtual index of the root
    const int v = root_v + depth_child;
    int matches = 0;
    const int n_cand = hm >> 1;
    for (int k0 = 0; k0 < n_cand; k0 += 32) {
        const int k = k0 + x.lane + 1;
        bool eq = false;
        const int j = v - 2 * k;
        if (k <= n_cand && j >= 0) {
            const DPos* q;
            if (j > root_v) q = &x.ptr.node_pos[x.nbase + (x.ptr.path[(size_t)x.g * x.prm.node_cap + (j - root_v)].x & 0xFFFF)];
            else q = &x.ptr.hist[(size_t)x.g * HIST_CAP + j];
            eq = same_position_key(*q, child);
        }
        matches += __popc(__ballot_sync(0xffffffffu, eq));
    }
    return matches + 1 >= x.prm.repetitions;
}

// tree.rs:272-289 on the edge list.  The reference draws one component per legal move INCLUDING the three
// under-promotion duplicates that share a policy index with the queen promotion; edges keep one entry per index,
// so a queen-promotion edge receives four consecutive components.
__device__ void apply_noise(Ctx& x, int node, u64 game_id, u64 ply) {
    const size_t off = x.ebase + x.ptr.node_edge_off[x.nbase + node];
    const int L = x.ptr.node_nedges[x.nbase + node], n = x.ptr.node_nmoves[x.nbase + node];
    if (n < 2) return;
    for (int i = x.lane; i < n; i += 32) x.sh->gam[i] = gamma_sample(x.prm.seed, game_id, ply, 1000 + (u64)i, (double)x.prm.alpha);
    __syncwarp();
    double sum = 0.0;
    if (x.lane == 0) for (int i = 0; i < n; i++) sum = sum + x.sh->gam[i];
    sum = __shfl_sync(0xffffffffu, sum, 0);
    for (int i = x.lane; i < n; i += 32) x.sh->fval[i] = (float)(x.sh->gam[i] / sum);
    // component offset of every edge
    if (x.lane == 0) {
        int comp = 0;
        for (int e = 0; e < L; e++) {
            x.sh->sidx[e] = (uint16_t)comp;
            comp += (((x.ptr.edge_mv[off + e] >> 12) & 7) == 4) ? 4 : 1;
        }
    }
    __syncwarp();
    const float keep = __fsub_rn(1.0f, x.prm.eps);
    for (int e = x.lane; e < L; e += 32) {
        float p = __fmul_rn(x.ptr.edge_P[off + e], keep);
        const int c0 = x.sh->sidx[e];
        const int reps = (((x.ptr.edge_mv[off + e] >> 12) & 7) == 4) ? 4 : 1;
        for (int r = 0; r < reps; r++) p = __fadd_rn(p, __fmul_rn(x.prm.eps, x.sh->fval[c0 + r]));
        x.ptr.edge_P[off + e] = p;
    }
    __syncwarp();
}

// Creates a tree node from sh->child / sh->moves (MCTree::new, tree.rs:84-104) without priors.
// Returns the node id or -1 on pool exhaustion.
__device__ __forceinline__ int create_node(Ctx& x, int depth) {
    const int n = x.sh->n_moves;
    int L = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + x.lane;
        bool keep = false;
        if (i < n) { int pr = (x.sh->moves[i] >> 12) & 7; keep = pr == 0 || pr == 4; }
        L += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if ((int)x.c.n_nodes >= x.prm.node_cap || (int)x.c.n_edges + L > x.prm.edge_cap) return -1;
    const int node = (int)x.c.n_nodes;
    const uint32_t eoff = x.c.n_edges;
    const int turn = meta_turn(x.sh->child.meta);
    int written = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + x.lane;
        bool keep = false;
        uint16_t mv = 0;
        if (i < n) { mv = x.sh->moves[i]; int pr = (mv >> 12) & 7; keep = pr == 0 || pr == 4; }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const size_t e = x.ebase + eoff + written + __popc(bal & ((1u << x.lane) - 1));
            x.ptr.edge_mv[e] = (uint32_t)mv | ((uint32_t)move_to_index(mv, turn) << 16);
            x.ptr.edge_N[e] = 0.0f;
            x.ptr.edge_W[e] = 0.0f;
            x.ptr.edge_P[e] = 0.0f;
            x.ptr.edge_link[e] = EDGE_NO_CHILD;
        }
        written += __popc(bal);
    }
    if (x.lane == 0) {
        const size_t ni = x.nbase + node;
        x.ptr.node_pos[ni] = x.sh->child;
        x.ptr.node_edge_off[ni] = eoff;
        x.ptr.node_nedges[ni] = (uint16_t)L;
        x.ptr.node_nmoves[ni] = (uint16_t)n;
        x.ptr.node_depth[ni] = (uint16_t)depth;
    }
    x.c.n_nodes++;
    x.c.n_edges += L;
    x.new_link = (unsigned long long)node | ((unsigned long long)L << 16) | ((unsigned long long)eoff << 24);
    __syncwarp();
    return node;
}

// Queues the position in sh->child (the node create_node made last: its edge range is in x.new_link) for evaluation;
// returns the batch slot.
__device__ __forceinline__ int submit_request(Ctx& x, int node) {
    int slot = 0;
    if (x.lane == 0) slot = atomicAdd(x.ptr.batch_count, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    const DPos child = x.sh->child;
    if (x.lane == 0) {
        x.ptr.req_pos[slot] = child;
        (void)node;
        x.ptr.req_edge_off[slot] = (unsigned long long)(x.ebase + (size_t)(x.new_link >> 24));
        x.ptr.req_nedges[slot] = (int)((x.new_link >> 16) & 0xFF);
    }
    if (x.prm.fp32_planes) {
        const u64 occ = occupied(child);
        const u64 ours = meta_turn(child.meta) == 0 ? child.white : occ ^ child.white;
        const int pep = pseudo_legal_ep(child);
        float* out = x.ptr.req_f32 + (size_t)slot * (AZ_NUM_PLANES * 64);
        for (int e = x.lane; e < AZ_NUM_PLANES * 64; e += 32) out[e] = plane_value(child, e >> 6, e & 63, pep, ours, occ ^ ours);
    } else {
        encode_bf16_warp(child, x.ptr.req_bf16 + (size_t)slot * 64 * x.prm.plane_ch, x.lane, x.prm.plane_ch);
    }
    if (x.lane == 0) x.st_evals++;
    return slot;
}

// W[a] += v ; N[a] += 1 along the stored path, leaf parent first (tree.rs:197-206).  `leaf_value` is what
// expand() returned (the child's point of view).  `spec` (nullable): this lane's path entry requested before the control block
// was known (levels 0..31), which takes one dependent load out of the chain ctl -> path -> edge.
__device__ __forceinline__ void backup(Ctx& x, float leaf_value, int path_len, const uint2* spec = nullptr) {
    for (int i0 = 0; i0 < path_len; i0 += 32) {
        const int i = i0 + x.lane;
        if (i < path_len) {
            const uint2 pe = (spec && i0 == 0) ? *spec : x.ptr.path[(size_t)x.g * x.prm.node_cap + i];
            const size_t e = x.ebase + pe.y;
            // levels above the leaf's parent: the sign flips once per level
            const int up = path_len - 1 - i;
            const float v = (up & 1) ? leaf_value : -leaf_value;
            x.ptr.edge_W[e] = __fadd_rn(x.ptr.edge_W[e], v);
            x.ptr.edge_N[e] = __fadd_rn(x.ptr.edge_N[e], 1.0f);
        }
    }
    __syncwarp();
}

// One PUCT descent from the root (tree.rs:180-202).  Fills the path and returns the leaf: its parent's node id, position and
// depth, and the chosen edge (pool index, wire move).
// The walk costs ONE dependent memory round trip per level: an edge carries its child's id together with the child's edge
// range (edge_link), the root's range is in the control block, and the chosen edge's move and every visited node's position
// (the leaf's parent needs it) are requested in the same batch of loads.
// total_visits (tree.rs:184: the sum of the node's visit counts) needs no loads at all: every simulation that passes through
// a node increments exactly one of its edges, and the simulations passing through a node are the visits of the edge leading
// into it minus the one that created it -- so total_visits + 1 is exactly that edge's N (at the root: simulations done + 1),
// the same integer in f32 the reference obtains by summing.
__device__ __forceinline__ void select_leaf(Ctx& x, int& leaf_node, size_t& leaf_pe, uint32_t& leaf_mv, int& depth, DPos& parent) {
    int node = 0;
    depth = 0;
    uint32_t eoff = 0;
    int L = (int)((x.c.flags >> 8) & 0xFF);
    float total = __fadd_rn((float)x.c.sims_done, 1.0f);
    for (;;) {
        const size_t off = x.ebase + eoff;
        u64 here = 0;   // lanes 0..7: the eight words of this node's position
        if (x.lane < 8) here = reinterpret_cast<const u64*>(&x.ptr.node_pos[x.nbase + node])[x.lane];
        const float sq = __fsqrt_rn(total);
        float best = -INFINITY;
        int best_e = 0x7FFFFFFF;
        // this lane's candidate: link, move and visit count of its best edge (lane 0 starts with edge 0 for the NaN fallback)
        u64 my_k = EDGE_NO_CHILD;
        uint32_t my_m = 0;
        float my_n = 0.0f;
        for (int e = x.lane; e < L; e += 32) {
            const float P = x.ptr.edge_P[off + e], N = x.ptr.edge_N[off + e], W = x.ptr.edge_W[off + e];
            const u64 K = x.ptr.edge_link[off + e];
            const uint32_t M = x.ptr.edge_mv[off + e];
            const float u = __fdiv_rn(__fmul_rn(__fmul_rn(x.prm.c_puct, P), sq), __fadd_rn(1.0f, N));
            const float q = N > 0.0f ? __fdiv_rn(W, N) : 0.0f;
            const float v = __fadd_rn(q, u);
            if (v > best) { best = v; best_e = e; my_k = K; my_m = M; my_n = N; }
            else if (e == x.lane) { my_k = K; my_m = M; my_n = N; }   // seed: lane 0 must hold edge 0 for the NaN fallback
        }
        // warp argmax over (score, edge): larger value wins, ties go to the earlier move (strict '>' in list order)
        for (int d = 16; d; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, d);
            const int oe = __shfl_xor_sync(0xffffffffu, best_e, d);
            if (ov > best || (ov == best && oe < best_e)) { best = ov; best_e = oe; }
        }
        if (best_e == 0x7FFFFFFF) best_e = 0;  // every score was NaN / -inf: the reference keeps max_index = 0 (first move)
        // the winning edge belongs to lane best_e % 32, whose candidate it is
        const int owner = best_e & 31;
        const u64 best_k = shfl_u64(my_k, owner);
        const uint32_t best_m = __shfl_sync(0xffffffffu, my_m, owner);
        const float best_n = __shfl_sync(0xffffffffu, my_n, owner);
        if (x.lane == 0) {
            x.ptr.path[(size_t)x.g * x.prm.node_cap + depth] =
                make_uint2((uint32_t)node | ((uint32_t)best_e << 16), eoff + (uint32_t)best_e);
            x.st_edges += L;
        }
        if (best_k == EDGE_NO_CHILD) {
            leaf_node = node; leaf_pe = off + best_e; leaf_mv = best_m;
            parent.pawn = shfl_u64(here, 0); parent.knight = shfl_u64(here, 1); parent.bishop = shfl_u64(here, 2);
            parent.rook = shfl_u64(here, 3); parent.queen = shfl_u64(here, 4); parent.king = shfl_u64(here, 5);
            parent.white = shfl_u64(here, 6); parent.meta = shfl_u64(here, 7);
            __syncwarp();
            return;
        }
        node = (int)(best_k & 0xFFFF);
        L = (int)((best_k >> 16) & 0xFF);
        eoff = (uint32_t)(best_k >> 24);
        total = best_n;   // = (visits through this child) + 1
        depth++;
    }
}


// ------------------------------------------------------------------------------------------- evaluation cache
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// entries are reused after eviction, so their contents are read from L2 (ld.cg), never from a possibly stale L1 line
__device__ __forceinline__ bool cache_key_equal(const CacheEntry* e, const DPos* key, int lane) {
    bool eq = true;
    if (lane < 8) eq = __ldcg(reinterpret_cast<const u64*>(&e->key) + lane) == reinterpret_cast<const u64*>(key)[lane];
    return __all_sync(0xffffffffu, eq);
}
constexpr int CACHE_PROBES = 8;
__device__ __forceinline__ uint32_t cache_tag(u64 h) { return (uint32_t)(h >> 46); }   // 18 bits
// whole warp; the key is in sh->key.  Returns the slot holding it (and the state word observed) or -1.
__device__ __forceinline__ int cache_lookup(Ctx& x, u64 h, uint32_t& seen) {
    const uint32_t tag = cache_tag(h);
    for (int probe = 0; probe < CACHE_PROBES; probe++) {
        const uint32_t slot = (uint32_t)(h + probe) & x.prm.cache_mask;
        const uint32_t s = ld_acquire_u32(&x.ptr.cache_state[slot]);
        if (s == 0) return -1;   // slots are replaced, never emptied: an empty slot ends every probe sequence
        if ((s & 3) == 2 && (s >> 14) == tag && cache_key_equal(&x.ptr.cache_entry[slot], &x.sh->key, x.lane)) { seen = s; return (int)slot; }
    }
    return -1;
}
// whole warp, after the entry's contents were copied: true if the slot still holds what cache_lookup saw
__device__ __forceinline__ bool cache_validate(Ctx& x, int slot, uint32_t seen) {
    __threadfence();   // the copies above are ordered before the second look at the state word
    __syncwarp();
    const uint32_t s2 = *reinterpret_cast<volatile uint32_t*>(&x.ptr.cache_state[slot]);
    const bool ok = __all_sync(0xffffffffu, ((s2 ^ seen) & ~CACHE_EPOCH_MASK) == 0);
    if (ok && x.lane == 0 && ((seen >> 2) & 63) != x.prm.cache_epoch)   // a hit keeps the entry young (failure is harmless)
        atomicCAS(&x.ptr.cache_state[slot], seen, (seen & ~CACHE_EPOCH_MASK) | (x.prm.cache_epoch << 2));
    return ok;
}
// whole warp: publishes (sh->key -> priors of `node`, value) unless it is already there; a full neighbourhood gives up
// its stalest entry (not inserted or hit for CACHE_MIN_AGE epochs or more)
__device__ __forceinline__ void cache_insert(Ctx& x, u64 h, int node, float value) {
    const uint32_t tag = cache_tag(h);
    const size_t off = x.ebase + x.ptr.node_edge_off[x.nbase + node];
    const int L = x.ptr.node_nedges[x.nbase + node];
    int target = -1, victim = -1, victim_age = CACHE_MIN_AGE - 1;
    uint32_t target_seq = 0, victim_s = 0;
    for (int probe = 0; probe < CACHE_PROBES && target < 0; probe++) {
        const uint32_t slot = (uint32_t)(h + probe) & x.prm.cache_mask;
        const uint32_t s = ld_acquire_u32(&x.ptr.cache_state[slot]);
        if ((s & 3) == 2 && (s >> 14) == tag && cache_key_equal(&x.ptr.cache_entry[slot], &x.sh->key, x.lane)) return;
        if (s == 0) {
            uint32_t old = 1;
            if (x.lane == 0) old = atomicCAS(&x.ptr.cache_state[slot], 0u, 1u);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == 0) target = (int)slot;
        } else if ((s & 3) == 2) {
            const int age = (int)((x.prm.cache_epoch - (s >> 2)) & 63);
            if (age > victim_age) { victim = (int)slot; victim_age = age; victim_s = s; }
        }
    }
    if (target < 0) {
        if (victim < 0) return;
        uint32_t old = ~victim_s;
        if (x.lane == 0) old = atomicCAS(&x.ptr.cache_state[victim], victim_s, (victim_s & ~3u) | 1u);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != victim_s) return;   // somebody else hit or replaced it meanwhile
        target = victim;
        target_seq = ((victim_s >> 8) + 1) & 63;
        if (x.lane == 0) x.st_evict++;
    }
    CacheEntry* e = &x.ptr.cache_entry[target];
    if (x.lane < 4) reinterpret_cast<uint4*>(&e->key)[x.lane] = reinterpret_cast<const uint4*>(&x.sh->key)[x.lane];
    if (x.lane == 4) { e->value = value; e->n_priors = (uint32_t)L; }
    for (int i = x.lane; i < L; i += 32) e->prior[i] = x.ptr.edge_P[off + i];
    __threadfence();
    __syncwarp();
    if (x.lane == 0) atomicExch(&x.ptr.cache_state[target], (tag << 14) | (target_seq << 8) | (x.prm.cache_epoch << 2) | 2u);
}

__device__ __forceinline__ void setup_root_from_shared(Ctx& x) {
    // fresh tree whose root is sh->child / sh->moves (priors are written by the caller)
    x.c.n_nodes = 0; x.c.n_edges = 0; x.c.sims_done = 0; x.c.max_depth = 0; x.c.pending_node = -1;
    create_node(x, 0);
    x.c.flags = (x.c.flags & 0xFFFF00FFu) | ((uint32_t)((x.new_link >> 16) & 0xFF) << 8);   // the root's edge count (its offset is 0)
}

// Start position, shared start-position priors, fresh noise (training.rs:352-361) -- or, when the generation's budget of
// games is used up (az_selfplay_begin_n), the slot goes idle.
__device__ void start_new_game(Ctx& x) {
    unsigned long long gid = 0;
    if (x.lane == 0) gid = atomicAdd(&x.ptr.counters->next_game_id, 1ULL);
    gid = __shfl_sync(0xffffffffu, gid, 0);
    if (x.prm.last_game_id != 0 && gid >= x.prm.last_game_id) { x.c.status = 1; x.c.n_samples = 0; return; }
    x.c.game_id = gid; x.c.ply = 0; x.c.n_samples = 0; x.c.hist_len = 1;
    if (x.lane == 0) {
        az_position sp;
        sp.roles[0] = 0x00FF00000000FF00ULL; sp.roles[1] = 0x4200000000000042ULL; sp.roles[2] = 0x2400000000000024ULL;
        sp.roles[3] = 0x8100000000000081ULL; sp.roles[4] = 0x0800000000000008ULL; sp.roles[5] = 0x1000000000000010ULL;
        sp.colors[0] = 0xFFFFULL; sp.colors[1] = 0xFFFF000000000000ULL;
        sp.turn = 0; sp.castling = 15; sp.ep_square = -1; sp.reserved = 0; sp.halfmoves = 0; sp.fullmoves = 1;
        DPos s = dpos_from_wire(sp);
        ListSink sink{x.sh->moves, 0};
        GenInfo gi = gen_legal(s, sink);
        set_key_bits(s, gi.has_legal_ep);
        x.sh->child = s; x.sh->n_moves = sink.n; x.sh->term = 0;
        x.ptr.hist[(size_t)x.g * HIST_CAP] = s;
    }
    __syncwarp();
    setup_root_from_shared(x);
    const int L0 = x.ptr.node_nedges[x.nbase];
    for (int e = x.lane; e < L0; e += 32) x.ptr.edge_P[x.ebase + e] = x.ptr.start_prior[e];
    __syncwarp();
    apply_noise(x, 0, x.c.game_id, 0);
}

// The game in this slot is over and `scale` = result x decay (training.rs:332-335).  Its staged EpisodeSteps move to the
// sample queue only if ALL of them fit (the counter never covers unwritten slots); otherwise the game parks (status 4)
// with its samples staged and tries again in the next wave, after the host has drained.  Nothing is ever dropped.
__device__ void finish_game(Ctx& x, float scale) {
    const int ns = (int)x.c.n_samples;
    unsigned long long base = 0;
    int ok = 0;
    if (x.lane == 0) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&x.ptr.counters->samples_out);
        for (;;) {
            if (cur + (unsigned long long)ns > (unsigned long long)x.prm.sample_cap) break;
            const unsigned long long prev = atomicCAS(&x.ptr.counters->samples_out, cur, cur + (unsigned long long)ns);
            if (prev == cur) { base = cur; ok = 1; break; }
            cur = prev;
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!ok) {
        x.c.status = 4;
        x.c.park_scale = __float_as_uint(scale);
        return;
    }
    x.c.status = 0;
    az_sample* src = x.ptr.game_samples + (size_t)x.g * MAX_SAMPLE_PLIES;
    for (int i = x.lane; i < ns; i += 32) src[i].final_value = __fmul_rn(src[i].final_value, scale);
    __syncwarp();
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(x.ptr.out_samples + base);
    const int n16 = ns * (int)(sizeof(az_sample) / 16);
    for (int i = x.lane; i < n16; i += 32) d4[i] = s4[i];
    if (x.lane == 0) x.st_games++;
    start_new_game(x);
}

// training.rs:303-335 for one finished search: record the EpisodeStep, pick the move, advance or finish the game.
__device__ void move_step(Ctx& x) {
    const size_t ni = x.nbase;  // root
    const size_t off = x.ebase + x.ptr.node_edge_off[ni];
    const int L = x.ptr.node_nedges[ni];
    const DPos root = x.ptr.node_pos[ni];
    const int turn = meta_turn(root.meta);
    // ---- sort the visited edges by policy index (the reference works on the dense 4096-vector)
    int nv = 0;
    for (int e0 = 0; e0 < L; e0 += 32) {
        const int e = e0 + x.lane;
        int rank = -1;
        if (e < L && x.ptr.edge_N[off + e] > 0.0f) {
            const uint32_t my = x.ptr.edge_mv[off + e] >> 16;
            rank = 0;
            for (int o = 0; o < L; o++)
                if (x.ptr.edge_N[off + o] > 0.0f && (x.ptr.edge_mv[off + o] >> 16) < my) rank++;
            x.sh->sidx[rank] = (uint16_t)e;
        }
        nv += __popc(__ballot_sync(0xffffffffu, rank >= 0));
    }
    __syncwarp();
    // ---- EpisodeStep
    const uint32_t sidx_slot = x.c.n_samples;
    az_sample* smp = nullptr;
    if (sidx_slot < MAX_SAMPLE_PLIES) {
        smp = x.ptr.game_samples + (size_t)x.g * MAX_SAMPLE_PLIES + sidx_slot;
        for (int i = x.lane; i < AZ_MAX_MOVES; i += 32) {  // unused tail entries are zero (deterministic records)
            uint16_t si = 0, sc = 0;
            if (i < nv) {
                const size_t e = off + x.sh->sidx[i];
                si = (uint16_t)(x.ptr.edge_mv[e] >> 16);
                sc = (uint16_t)x.ptr.edge_N[e];
            }
            smp->index[i] = si;
            smp->count[i] = sc;
        }
    }
    // ---- improved policy weights (tree.rs:173-175): w = visits^(1/T) per visited index, summed in index order (the zeros of
    // the reference's dense vector add nothing); T = 1 gives w = visits and the sum = S exactly
    for (int i = x.lane; i < nv; i += 32) x.sh->fval[i] = pow_inv_temperature(x.ptr.edge_N[off + x.sh->sidx[i]], x.prm.inv_temperature);
    __syncwarp();
    // ---- action (training.rs:310-321)
    int action_edge = 0;
    if (x.lane == 0) {
        float S = 0.0f;  // weights_sum
        for (int i = 0; i < nv; i++) S = __fadd_rn(S, x.sh->fval[i]);
        if ((uint32_t)meta_fullmoves(root.meta) >= x.prm.anneal) {
            float best = -1.0f;  // Iterator::max_by keeps the LAST maximum in index order
            for (int i = 0; i < nv; i++) {
                const float w = __fdiv_rn(x.sh->fval[i], S);
                if (w >= best) { best = w; action_edge = x.sh->sidx[i]; }
            }
        } else {
            float total = 0.0f;  // WeightedIndex: cumulative f32 weights in index order
            for (int i = 0; i < nv; i++) total = __fadd_rn(total, __fdiv_rn(x.sh->fval[i], S));
            const u64 h = rng_u64(x.prm.seed, x.c.game_id, x.c.ply, 1, 0);
            const float u = __fmul_rn((float)(h >> 40), 1.0f / 16777216.0f);
            const float chosen = __fmul_rn(u, total);
            float cum = 0.0f;
            action_edge = nv > 0 ? x.sh->sidx[nv - 1] : 0;
            for (int i = 0; i < nv; i++) {
                cum = __fadd_rn(cum, __fdiv_rn(x.sh->fval[i], S));
                if (cum > chosen) { action_edge = x.sh->sidx[i]; break; }
            }
        }
    }
    action_edge = __shfl_sync(0xffffffffu, action_edge, 0);
    const uint32_t amv = x.ptr.edge_mv[off + action_edge];
    if (smp && x.lane == 0) {
        smp->position = dpos_to_wire(root);
        smp->final_value = turn == 0 ? 1.0f : -1.0f;
        smp->search_depth = (int32_t)x.c.max_depth;
        smp->game_id = x.c.game_id;
        smp->ply = x.c.ply;
        smp->action = (uint16_t)(amv >> 16);
        smp->n_visits = (uint16_t)nv;
        x.st_pos++;
        x.st_sdepth += x.c.max_depth;
    }
    x.c.n_samples = min(x.c.n_samples + 1, (uint32_t)MAX_SAMPLE_PLIES);
    // ---- play it (chess.rs:36-63)
    const unsigned long long child_link = x.ptr.edge_link[off + action_edge];
    const int child_node = child_link == EDGE_NO_CHILD ? -1 : (int)(child_link & 0xFFFF);
    lane0_make_child(x, root, (uint16_t)(amv & 0xFFFF));
    int term = x.sh->term;
    if (term == 0 && draw_by_rules(x, x.sh->child, 1)) term = 1;
    if (term == 0) {
        // traverse_new (tree.rs:239-256): keep the child's priors, drop everything else, fresh noise
        float keepP[7];
        int Lc = 0;
        if (child_node >= 0) {
            const size_t coff = x.ebase + x.ptr.node_edge_off[x.nbase + child_node];
            Lc = x.ptr.node_nedges[x.nbase + child_node];
#pragma unroll
            for (int k = 0; k < 7; k++) { int e = x.lane + 32 * k; keepP[k] = e < Lc ? x.ptr.edge_P[coff + e] : 0.0f; }
        } else {
            x.c.status = 3;  // a move with visits but no stored child must be terminal
        }
        __syncwarp();
        if (x.lane == 0 && x.c.hist_len < HIST_CAP) x.ptr.hist[(size_t)x.g * HIST_CAP + x.c.hist_len] = x.sh->child;
        x.c.hist_len = min(x.c.hist_len + 1, (uint32_t)HIST_CAP);
        x.c.ply++;
        setup_root_from_shared(x);
#pragma unroll
        for (int k = 0; k < 7; k++) { int e = x.lane + 32 * k; if (e < Lc) x.ptr.edge_P[x.ebase + e] = keepP[k]; }
        __syncwarp();
        apply_noise(x, 0, x.c.game_id, x.c.ply);
        return;
    }
    // ---- game over: back-fill final_value (training.rs:332-335), publish the samples, start a new game
    const float mover = turn == 0 ? 1.0f : -1.0f;
    const float result = term == 1 ? 0.0f : mover;
    const float decay = __fsub_rn(1.0f, __fdiv_rn((float)meta_fullmoves(x.sh->child.meta), __fmul_rn(2.0f, (float)x.prm.num_fullmoves)));
    finish_game(x, __fmul_rn(result, decay));
}

// ------------------------------------------------------------------------------------------- cold paths of the wave kernel
// A finished search (once per S waves and game), a parked game and root noise are rare; compiled inline they sit in the middle
// of the hot instruction stream of a 20,000-instruction kernel whose warps stall on instruction fetch (ncu: "no instructions"
// is the second stall reason, profiles/r2_advance_v1_full.md).  They are real calls instead.  The game's control block goes in
// and out BY VALUE so that the hot path's copy stays in registers (a reference would pin it in local memory).
struct ColdOut {
    GameCtl c;
    unsigned int st_pos, st_games, st_sdepth;
};
__device__ __forceinline__ Ctx cold_ctx(const SearchParams* prm, const SearchPtrs* ptr, WarpShared* sh, int g, const GameCtl& c) {
    return Ctx{*prm, *ptr, sh, g, (int)(threadIdx.x & 31), (size_t)g * prm->node_cap, (size_t)g * prm->edge_cap, c, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
}
__device__ __noinline__ ColdOut move_step_cold(const SearchParams* prm, const SearchPtrs* ptr, WarpShared* sh, int g, GameCtl c) {
    Ctx x = cold_ctx(prm, ptr, sh, g, c);
    move_step(x);
    return ColdOut{x.c, x.st_pos, x.st_games, x.st_sdepth};
}
__device__ __noinline__ ColdOut finish_game_cold(const SearchParams* prm, const SearchPtrs* ptr, WarpShared* sh, int g, GameCtl c) {
    Ctx x = cold_ctx(prm, ptr, sh, g, c);
    finish_game(x, __uint_as_float(c.park_scale));
    return ColdOut{x.c, x.st_pos, x.st_games, x.st_sdepth};
}
__device__ __noinline__ void root_noise_cold(const SearchParams* prm, const SearchPtrs* ptr, WarpShared* sh, int g, GameCtl c) {
    Ctx x = cold_ctx(prm, ptr, sh, g, c);
    apply_noise(x, 0, c.game_id, c.noise_ply);
}

// ------------------------------------------------------------------------------------------- the wave kernel
#ifdef AZ_ADV_TIMING
#define ADV_T0() long long t_prev = clock64(), t_start = t_prev; unsigned long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define ADV_T(k) do { const long long t_now = clock64(); t_acc[k] += (unsigned long long)(t_now - t_prev); t_prev = t_now; } while (0)
#else
#define ADV_T0() do { } while (0)
#define ADV_T(k) do { } while (0)
#endif
// MINB = resident groups of four warps per SM the register allocation is bounded for: the kernel is a chain of dependent global loads
// per game, so resident warps (latency hiding) are worth more than registers (AZ_ADV_MINB selects, see launch below)
template <int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB * 4 / WARPS) k_advance(const __grid_constant__ SearchParams prm, const __grid_constant__ SearchPtrs ptr) {
    __shared__ WarpShared shared[WARPS];
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= prm.n_games) return;
    // the path of the pending simulation is requested together with the control block (its address needs only g)
    const uint2 path_spec = ptr.path[(size_t)g * prm.node_cap + min((int)(threadIdx.x & 31), prm.node_cap - 1)];
    Ctx x{prm, ptr, &shared[threadIdx.x >> 5], g, (int)(threadIdx.x & 31), (size_t)g * prm.node_cap, (size_t)g * prm.edge_cap,
          ptr.ctl[g], 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (g == 0 && x.lane == 0 && ptr.batch_zero) *ptr.batch_zero = 0;   // the other wave parity's request counter (see run_wave)
    if (x.c.status != 0 && x.c.status != 4) return;
    if (!prm.consume && x.c.pending_node >= 0) return;   // extra pass: this game already waits for the network

    ADV_T0();
    // ---- 0. game over, samples staged: publish now if the host has drained the queue, else stay parked
    if (x.c.status == 4) {
        const ColdOut o = finish_game_cold(&prm, &ptr, x.sh, g, x.c);
        x.c = o.c; x.st_pos += o.st_pos; x.st_games += o.st_games; x.st_sdepth += o.st_sdepth;
    }
    const bool runnable = x.c.status == 0;
    // ---- 1. the evaluation requested in the previous wave has arrived
    if (runnable && x.c.pending_node >= 0) {
        const int node = x.c.pending_node, slot = x.c.pending_slot;
        const size_t off = x.ebase + ptr.node_edge_off[x.nbase + node];
        const int L = ptr.node_nedges[x.nbase + node];
        if (!prm.priors_scattered) {
            const float* pol = ptr.res_policy + (size_t)slot * AZ_ACTION_SPACE;
            for (int e = x.lane; e < L; e += 32) ptr.edge_P[off + e] = pol[ptr.edge_mv[off + e] >> 16];
        }
        __syncwarp();
        if (prm.cache_mask && prm.mode == 1) {  // cache.insert (training.rs:413)
            if (x.lane == 0) x.sh->key = fen_key_of(ptr.node_pos[x.nbase + node]);
            __syncwarp();
            cache_insert(x, fen_key_hash(x.sh->key), node, ptr.res_value[slot]);
        }
        if (node == 0) {  // MCTree::init (tree.rs:37-64): root priors, optional noise, no backup
            if (x.c.flags & 1) root_noise_cold(&prm, &ptr, x.sh, g, x.c);
        } else {
            backup(x, ptr.res_value[slot], (int)x.c.path_len, &path_spec);
            x.c.sims_done++;
            if (x.lane == 0) x.st_sims++;
        }
        x.c.pending_node = -1;
    }

    ADV_T(0);
    // ---- 2. run until the network is needed again
    for (int iter = 0; runnable && iter < prm.max_iters; iter++) {
        if ((int)x.c.sims_done >= prm.S) {
            if (prm.mode == 0) { x.c.status = 1; break; }
            const ColdOut o = move_step_cold(&prm, &ptr, x.sh, g, x.c);
            x.c = o.c; x.st_pos += o.st_pos; x.st_games += o.st_games; x.st_sdepth += o.st_sdepth;
            ADV_T(6);
            if (x.c.status != 0) break;   // idle (generation complete), parked (sample queue full) or an error
            continue;
        }
        int node, depth;
        size_t pe;
        uint32_t leaf_mv;
        DPos parent;
        select_leaf(x, node, pe, leaf_mv, depth, parent);
        ADV_T(1);
        lane0_make_child(x, parent, (uint16_t)(leaf_mv & 0xFFFF));
        int term = x.sh->term;
        ADV_T(2);
        if (term == 0 && draw_by_rules(x, x.sh->child, depth + 1)) term = 1;
        ADV_T(3);
        if (x.lane == 0) x.st_depth += depth + 1;
        if (term != 0) {
            // Draw -> 0.0, decisive -> -1.0 from the child's side (tree.rs:233-234); nothing is stored
            backup(x, term == 1 ? 0.0f : -1.0f, depth + 1);
            x.c.sims_done++;
            if (x.lane == 0) { x.st_sims++; x.st_term++; }
            continue;
        }
        const int child = create_node(x, depth + 1);
        ADV_T(4);
        if (child < 0) { x.c.status = 2; if (x.lane == 0) atomicAdd(&ptr.counters->errors, 1ULL); break; }
        if (x.lane == 0) ptr.edge_link[pe] = x.new_link;
        x.c.max_depth = max(x.c.max_depth, (uint32_t)(depth + 1));
        if (prm.cache_mask && prm.mode == 1) {  // cache.get (tree.rs:214-218): a hit needs no network evaluation
            if (x.lane == 0) x.sh->key = fen_key_of(x.sh->child);
            __syncwarp();
            uint32_t seen = 0;
            const int slot = cache_lookup(x, fen_key_hash(x.sh->key), seen);
            if (slot >= 0) {
                const CacheEntry* ce = &ptr.cache_entry[slot];
                const size_t coff = x.ebase + ptr.node_edge_off[x.nbase + child];
                const int Lc = ptr.node_nedges[x.nbase + child];
                for (int e = x.lane; e < Lc; e += 32) ptr.edge_P[coff + e] = __ldcg(&ce->prior[e]);
                const float cv = __ldcg(&ce->value);
                if (cache_validate(x, slot, seen)) {   // else: the entry was being replaced; evaluate as a miss (edge_P is overwritten)
                    backup(x, cv, depth + 1);
                    x.c.sims_done++;
                    if (x.lane == 0) { x.st_sims++; x.st_hits++; }
                    continue;
                }
            }
        }
        x.c.pending_node = child;
        x.c.pending_slot = submit_request(x, child);
        x.c.path_len = depth + 1;
        ADV_T(5);
        break;
    }

    if (x.lane == 0) {
        ptr.ctl[g] = x.c;
        // statistics: same field order as Counters; 64 stripes keep 4096 warps from queueing on eight addresses
        unsigned long long* ct = ptr.stats + (size_t)(blockIdx.x & (STAT_STRIPES - 1)) * STAT_WIDTH;
        if (x.st_sims) atomicAdd(&ct[0], (unsigned long long)x.st_sims);
        if (x.st_pos) atomicAdd(&ct[1], (unsigned long long)x.st_pos);
        if (x.st_evals) atomicAdd(&ct[2], (unsigned long long)x.st_evals);
        if (x.st_hits) atomicAdd(&ct[3], (unsigned long long)x.st_hits);
        if (x.st_term) atomicAdd(&ct[4], (unsigned long long)x.st_term);
        if (x.st_games) atomicAdd(&ct[5], (unsigned long long)x.st_games);
        if (x.st_depth) atomicAdd(&ct[6], (unsigned long long)x.st_depth);
        if (x.st_edges) atomicAdd(&ct[7], (unsigned long long)x.st_edges);
        if (x.st_evict) atomicAdd(&ct[16], (unsigned long long)x.st_evict);
        if (x.st_sdepth) atomicAdd(&ct[17], (unsigned long long)x.st_sdepth);
#ifdef AZ_ADV_TIMING
        t_acc[7] = (unsigned long long)(clock64() - t_start);
        for (int k = 0; k < 8; k++) atomicAdd(&ct[8 + k], t_acc[k]);
        atomicAdd(&ct[24 + min(15, (int)(t_acc[7] >> 12))], 1ULL);
#endif
    }
}

// Root set-up for az_search (MCTree::init): node 0 from the given position, evaluation pending.
__global__ void __launch_bounds__(WARPS * 32) k_init_search(SearchParams prm, SearchPtrs ptr, const az_position* __restrict__ roots,
                                                            const az_position* __restrict__ hist, const uint32_t* __restrict__ hist_off,
                                                            const unsigned long long* __restrict__ noise_ids,
                                                            const uint32_t* __restrict__ noise_plies) {
    __shared__ WarpShared shared[WARPS];
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= prm.n_games) return;
    GameCtl c;
    memset(&c, 0, sizeof c);
    Ctx x{prm, ptr, &shared[threadIdx.x >> 5], g, (int)(threadIdx.x & 31), (size_t)g * prm.node_cap, (size_t)g * prm.edge_cap, c,
          0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    x.c.game_id = noise_ids ? noise_ids[g] : 0;
    x.c.noise_ply = noise_plies ? noise_plies[g] : 0;
    x.c.flags = noise_ids ? 1 : 0;
    // history (positions already counted, current root last); key bits need the legal ep square of each entry
    int hl = 1;
    if (hist) {
        const uint32_t h0 = hist_off[g], h1 = hist_off[g + 1];
        int n = (int)(h1 - h0);
        const uint32_t skip = n > HIST_CAP ? n - HIST_CAP : 0;
        n -= skip;
        for (int i = x.lane; i < n; i += 32) {
            DPos q = dpos_from_wire(hist[h0 + skip + i]);
            bool q_ep = false;
            if (meta_ep(q.meta) >= 0) { CountSink t{0}; q_ep = gen_legal(q, t).has_legal_ep; }
            set_key_bits(q, q_ep);
            ptr.hist[(size_t)g * HIST_CAP + i] = q;
        }
        hl = max(n, 1);
    }
    __syncwarp();
    if (x.lane == 0) {
        DPos r = dpos_from_wire(roots[g]);
        ListSink sink{x.sh->moves, 0};
        GenInfo gi = gen_legal(r, sink);
        set_key_bits(r, gi.has_legal_ep);
        x.sh->child = r; x.sh->n_moves = sink.n; x.sh->term = 0;
        if (!hist || hist_off[g + 1] == hist_off[g]) ptr.hist[(size_t)g * HIST_CAP] = r;
    }
    __syncwarp();
    x.c.hist_len = hl;
    setup_root_from_shared(x);
    x.c.pending_node = 0;
    x.c.pending_slot = submit_request(x, 0);
    if (x.sh->n_moves == 0) x.c.status = 1;  // nothing to search in a terminal position
    if (x.lane == 0) ptr.ctl[g] = x.c;
}

// Game set-up for self-play (training.rs:352-361): start position, shared priors, Dirichlet noise.
__global__ void __launch_bounds__(WARPS * 32) k_init_selfplay(SearchParams prm, SearchPtrs ptr, unsigned long long first_game_id) {
    __shared__ WarpShared shared[WARPS];
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= prm.n_games) return;
    GameCtl c;
    memset(&c, 0, sizeof c);
    Ctx x{prm, ptr, &shared[threadIdx.x >> 5], g, (int)(threadIdx.x & 31), (size_t)g * prm.node_cap, (size_t)g * prm.edge_cap, c,
          0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    x.c.game_id = first_game_id + g;
    x.c.hist_len = 1;
    if (x.lane == 0) {
        az_position sp;
        sp.roles[0] = 0x00FF00000000FF00ULL; sp.roles[1] = 0x4200000000000042ULL; sp.roles[2] = 0x2400000000000024ULL;
        sp.roles[3] = 0x8100000000000081ULL; sp.roles[4] = 0x0800000000000008ULL; sp.roles[5] = 0x1000000000000010ULL;
        sp.colors[0] = 0xFFFFULL; sp.colors[1] = 0xFFFF000000000000ULL;
        sp.turn = 0; sp.castling = 15; sp.ep_square = -1; sp.reserved = 0; sp.halfmoves = 0; sp.fullmoves = 1;
        DPos s = dpos_from_wire(sp);
        ListSink sink{x.sh->moves, 0};
        GenInfo gi = gen_legal(s, sink);
        set_key_bits(s, gi.has_legal_ep);
        x.sh->child = s; x.sh->n_moves = sink.n; x.sh->term = 0;
        ptr.hist[(size_t)g * HIST_CAP] = s;
    }
    __syncwarp();
    setup_root_from_shared(x);
    const int L0 = ptr.node_nedges[x.nbase];
    for (int e = x.lane; e < L0; e += 32) ptr.edge_P[x.ebase + e] = ptr.start_prior[e];
    __syncwarp();
    apply_noise(x, 0, x.c.game_id, 0);
    if (x.lane == 0) ptr.ctl[g] = x.c;
}

// start-position priors from a policy row (the single shared forward of training.rs:344-350)
__global__ void k_start_prior(const float* __restrict__ policy_row, float* __restrict__ start_prior) {
    if (threadIdx.x == 0) {
        az_position sp;
        sp.roles[0] = 0x00FF00000000FF00ULL; sp.roles[1] = 0x4200000000000042ULL; sp.roles[2] = 0x2400000000000024ULL;
        sp.roles[3] = 0x8100000000000081ULL; sp.roles[4] = 0x0800000000000008ULL; sp.roles[5] = 0x1000000000000010ULL;
        sp.colors[0] = 0xFFFFULL; sp.colors[1] = 0xFFFF000000000000ULL;
        sp.turn = 0; sp.castling = 15; sp.ep_square = -1; sp.reserved = 0; sp.halfmoves = 0; sp.fullmoves = 1;
        DPos s = dpos_from_wire(sp);
        uint16_t mv[AZ_MAX_MOVES];
        ListSink sink{mv, 0};
        gen_legal(s, sink);
        for (int i = 0; i < sink.n && i < 32; i++) start_prior[i] = policy_row[move_to_index(mv[i], 0)];
    }
}

// dense export of the root statistics (Box<[f32; 4096]> visits / scores)
__global__ void k_export_root(SearchParams prm, SearchPtrs ptr, float* __restrict__ visits, float* __restrict__ scores,
                              int32_t* __restrict__ depth) {
    const int g = blockIdx.x;
    float* vo = visits + (size_t)g * AZ_ACTION_SPACE;
    float* so = scores ? scores + (size_t)g * AZ_ACTION_SPACE : nullptr;
    for (int i = threadIdx.x; i < AZ_ACTION_SPACE; i += blockDim.x) { vo[i] = 0.0f; if (so) so[i] = 0.0f; }
    __syncthreads();
    const size_t nb = (size_t)g * prm.node_cap, eb = (size_t)g * prm.edge_cap + ptr.node_edge_off[nb];
    const int L = ptr.node_nedges[nb];
    for (int e = threadIdx.x; e < L; e += blockDim.x) {
        const int idx = ptr.edge_mv[eb + e] >> 16;
        vo[idx] = ptr.edge_N[eb + e];
        if (so) so[idx] = ptr.edge_W[eb + e];
    }
    if (threadIdx.x == 0 && depth) depth[g] = (int32_t)ptr.ctl[g].max_depth;
}

__global__ void k_count_active(SearchParams prm, SearchPtrs ptr, int* out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < prm.n_games) {
        const uint32_t s = ptr.ctl[g].status;
        if (s == 0) atomicAdd(&out[0], 1);
        if (s == 2 || s == 3) atomicAdd(&out[1], 1);
        if (s == 4) atomicAdd(&out[2], 1);
    }
}

// ------------------------------------------------------------------------------------------- host side
template <class T>
static int salloc(az_engine* e, SearchState* st, T** p, size_t n) {
    cudaError_t r = cudaMalloc(p, n * sizeof(T));
    if (r != cudaSuccess) return check_cuda(e, r, "cudaMalloc(search)");
    st->allocs.push_back(*p);
    return 0;
}

int search_create(az_engine* e) {
    SearchState* st = new SearchState;
    e->search = st;
    const az_config& c = e->cfg;
    const int G = c.max_games;
    st->G = G;
    SearchParams& p = st->prm;
    p.n_games = G; p.S = c.num_simulations; p.c_puct = c.c_puct; p.alpha = c.dirichlet_alpha; p.eps = c.dirichlet_epsilon;
    p.anneal = c.temperature_annealing; p.num_halfmoves = (int)c.num_halfmoves; p.num_fullmoves = (int)c.num_fullmoves;
    p.repetitions = (int)c.repetitions; p.seed = c.seed;
    p.node_cap = c.num_simulations + 2;
    if (p.node_cap > 65535) return set_err(e, AZ_ERR_INVALID_ARGUMENT, "num_simulations must be < 65534");
    // a game ends by the fullmove limit, so its history and its sample staging hold < 2 * num_fullmoves positions
    if (2 * (int)c.num_fullmoves > std::min(HIST_CAP, MAX_SAMPLE_PLIES) || c.num_fullmoves == 0 || c.num_halfmoves == 0 || c.repetitions == 0)
        return set_err(e, AZ_ERR_INVALID_ARGUMENT, "num_fullmoves must be in 1..256, num_halfmoves and repetitions positive");
    if (!(c.c_puct > 0.0f) || !(c.dirichlet_alpha > 0.0f) || c.dirichlet_epsilon < 0.0f || c.dirichlet_epsilon > 1.0f)
        return set_err(e, AZ_ERR_INVALID_ARGUMENT, "c_puct and dirichlet_alpha must be positive, dirichlet_epsilon in [0, 1]");
    const int per_node = c.edge_capacity_per_node > 0 ? c.edge_capacity_per_node : 96;
    p.edge_cap = std::max(p.node_cap * std::min(per_node, 218), 256);
    // A game whose leaf was terminal could go on selecting within the same wave, but those few warps (about 8 of 4096) then
    // run two or three times longer than the rest and set the kernel's duration; they continue in the next wave instead
    // (measured per wave: 8 -> 73.6 us, 2 -> 66.4 us, 1 -> 64.2 us; results do not depend on the schedule).
    // With the evaluation cache a hit completes a simulation without the network, so games may go further per wave.
    p.mode = 0; p.max_iters = c.cache_log2 > 0 ? 4 : 2;
    if (const char* v = getenv("AZ_ADV_MAX_ITERS")) p.max_iters = std::max(1, atoi(v));
    p.last_game_id = 0; p.cache_epoch = 0; p.consume = 1;
    if (const char* v = getenv("AZ_ADV_PASSES")) st->adv_passes = std::max(1, std::min(4, atoi(v)));
    if (!(c.temperature > 0.0f)) return set_err(e, AZ_ERR_INVALID_ARGUMENT, "temperature must be positive");
    p.inv_temperature = 1.0f / c.temperature;
    p.fp32_planes = c.precision == 1 ? 1 : 0;
    p.plane_ch = e->net->in_ch;
    // finished games' samples wait here for az_selfplay_drain / az_replay_add_pending; a game that does not fit parks until
    // the host has drained (finish_game), so the only hard requirement is room for one whole game
    p.sample_cap = std::max(G * 128, 1 << 16);
    if (const char* v = getenv("AZ_SAMPLE_CAP")) p.sample_cap = std::max(MAX_SAMPLE_PLIES, atoi(v));   // tests: force parking
    SearchPtrs& q = st->ptr;
    const size_t NN = (size_t)G * p.node_cap, NE = (size_t)G * p.edge_cap;
    int r = 0;
    r |= salloc(e, st, &q.node_pos, NN); r |= salloc(e, st, &q.node_edge_off, NN); r |= salloc(e, st, &q.node_nedges, NN);
    r |= salloc(e, st, &q.node_nmoves, NN); r |= salloc(e, st, &q.node_depth, NN);
    r |= salloc(e, st, &q.edge_P, NE); r |= salloc(e, st, &q.edge_N, NE); r |= salloc(e, st, &q.edge_W, NE);
    r |= salloc(e, st, &q.edge_link, NE); r |= salloc(e, st, &q.edge_mv, NE);
    r |= salloc(e, st, &q.ctl, (size_t)G); r |= salloc(e, st, &q.path, NN); r |= salloc(e, st, &q.hist, (size_t)G * HIST_CAP);
    r |= salloc(e, st, &q.batch_count, 8); r |= salloc(e, st, &q.req_pos, (size_t)e->max_batch);
    r |= salloc(e, st, &q.req_f32, (size_t)e->max_batch * AZ_NUM_PLANES * 64);
    r |= salloc(e, st, &q.req_edge_off, (size_t)e->max_batch); r |= salloc(e, st, &q.req_nedges, (size_t)e->max_batch);
    r |= salloc(e, st, &q.start_prior, 32); r |= salloc(e, st, &q.counters, 1);
    r |= salloc(e, st, &q.stats, (size_t)STAT_STRIPES * STAT_WIDTH);
    r |= salloc(e, st, &st->d_noise_ids, (size_t)G); r |= salloc(e, st, &st->d_noise_plies, (size_t)G);
    if (r) return AZ_ERR_OUT_OF_MEMORY;
    p.cache_mask = 0; q.cache_state = nullptr; q.cache_entry = nullptr;
    if (c.cache_log2 > 0) {
        if (c.cache_log2 > 26) return set_err(e, AZ_ERR_INVALID_ARGUMENT, "cache_log2 must be <= 26");
        const size_t slots = (size_t)1 << c.cache_log2;
        if (salloc(e, st, &q.cache_state, slots) || salloc(e, st, &q.cache_entry, slots)) return AZ_ERR_OUT_OF_MEMORY;
        cudaMemset(q.cache_state, 0, slots * sizeof(uint32_t));
        p.cache_mask = (uint32_t)(slots - 1);
    }
    q.req_bf16 = e->net->a_in;
    q.res_policy = e->d_policy;
    q.res_value = e->d_value;
    q.game_samples = nullptr;
    q.out_samples = nullptr;
    cudaMemset(q.counters, 0, sizeof(Counters));
    cudaMemset(q.stats, 0, (size_t)STAT_STRIPES * STAT_WIDTH * sizeof(unsigned long long));
    cudaMemset(q.batch_count, 0, 32);
    st->batch_base = q.batch_count;
    q.batch_zero = nullptr;
    return 0;
}

void search_destroy(az_engine* e) {
    SearchState* st = e->search;
    if (!st) return;
    for (void* p : st->allocs) cudaFree(p);
    delete st;
    e->search = nullptr;
}

// evaluates the queued requests: the network (bf16 or fp32) or the synthetic evaluator
static int evaluate_batch(az_engine* e, SearchState* st) {
    const SearchPtrs& q = st->ptr;
    if (e->stub_kind == 1) {
        const int warps = e->max_batch;
        e->n_launches++;
        k_stub_eval<<<(warps + 3) / 4, 128, 0, e->stream>>>(q.req_pos, q.batch_count, 0, e->stub_seed, e->d_policy, e->d_value);
        return check_cuda(e, cudaGetLastError(), "k_stub_eval");
    }
    if (e->cfg.precision == 1) return net_forward_fp32(e, q.req_f32, q.batch_count, 0, e->d_policy, e->d_value);
    HeadScatter sc{q.req_edge_off, q.req_nedges, q.edge_mv, q.edge_P};
    return net_forward_bf16(e, q.batch_count, 0, nullptr, e->d_value, &sc);
}

int search_pending_samples(az_engine* e, const az_sample** d_samples, int* n) {
    SearchState* st = e->search;
    if (!st->selfplay_active) return set_err(e, AZ_ERR_STATE, "az_selfplay_begin has not been called");
    Counters c;
    AZ_CUDA(e, cudaMemcpyAsync(&c, st->ptr.counters, sizeof c, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    *d_samples = st->ptr.out_samples;
    *n = (int)std::min<unsigned long long>(c.samples_out, st->prm.sample_cap);
    return 0;
}
int search_clear_pending(az_engine* e) {
    SearchState* st = e->search;
    AZ_CUDA(e, cudaMemsetAsync(&st->ptr.counters->samples_out, 0, sizeof(unsigned long long), e->stream));
    return 0;
}

static int run_wave(az_engine* e, SearchState* st) {
    const int blocks = (st->prm.n_games + WARPS - 1) / WARPS;
    st->prm.priors_scattered = (e->stub_kind == 0 && e->cfg.precision != 1) ? 1 : 0;
    if (st->prm.mode == 1) st->prm.cache_epoch = (uint32_t)((st->wave_counter++ / (unsigned long long)std::max(st->prm.S, 1)) & 63);
    // Two request counters alternate between waves: this wave's k_advance counts into one and clears the other, which the
    // previous wave's network kernels (complete by stream order) were the last to read -- no memset launch per wave.
    st->batch_parity ^= 1;
    st->ptr.batch_count = st->batch_base + st->batch_parity;
    st->ptr.batch_zero = st->batch_base + (st->batch_parity ^ 1);
    if (e->prof_every > 0 && e->stub_kind == 0 && e->cfg.precision != 1 && (e->prof_counter % (uint64_t)e->prof_every) == 0 &&
        !e->prof_adv_event) {
        cudaEventCreate(&e->prof_adv_event);
        cudaEventRecord(e->prof_adv_event, e->stream);
    }
    // AZ_ADV_PASSES > 1 (experiment, default 1): extra k_advance passes per wave in which only the games that completed their
    // max_iters network-free simulations without needing the network go on.  Measured at 14 % / 45 % / 63 % cache hits
    // (profiles/r2_passes_raw.jsonl): always slower -- the fuller batch costs tower time in proportion, and more simulations
    // per wave mean more games evaluating the same position in the same wave before anybody's result reaches the cache.
    const int passes = st->prm.mode == 1 && st->prm.cache_mask ? st->adv_passes : 1;
    for (int pass = 0; pass < passes; pass++) {
    st->prm.consume = pass == 0 ? 1 : 0;
    e->n_launches++;
    const int minb = e->knobs.adv_minb;
    // groups of four warps per SM (register bound): 3 -> 162 registers, 4 -> 128, 5 -> 96, 6 -> 80, 7 -> 72 with 216 bytes of
    // spills.  With 7, all 4096 games of a wave are resident at once (148 SMs x 28 warps) and the kernel is one round of
    // latency chains instead of two: measured per wave 4 -> 67.6 us, 5 -> 68.5, 6 -> 69.6, 7 -> 54.7.
    switch (minb) {
        case 4: k_advance<4><<<blocks, WARPS * 32, 0, e->stream>>>(st->prm, st->ptr); break;
        case 5: k_advance<5><<<blocks, WARPS * 32, 0, e->stream>>>(st->prm, st->ptr); break;
        case 6: k_advance<6><<<blocks, WARPS * 32, 0, e->stream>>>(st->prm, st->ptr); break;
        case 7: k_advance<7><<<blocks, WARPS * 32, 0, e->stream>>>(st->prm, st->ptr); break;
        default: k_advance<3><<<blocks, WARPS * 32, 0, e->stream>>>(st->prm, st->ptr); break;
    }
    }
    st->prm.consume = 1;
    AZ_CUDA(e, cudaGetLastError());
    return evaluate_batch(e, st);
}

}  // namespace azb

using namespace azb;

extern "C" {

int az_set_evaluator_stub(az_engine* e, int kind, uint64_t seed) {
    if (!e || kind < 0 || kind > 1) return AZ_ERR_INVALID_ARGUMENT;
    e->stub_kind = kind;
    e->stub_seed = seed;
    return AZ_OK;
}

int az_search(az_engine* e, int n, const az_position* roots, const az_position* history, const uint32_t* hist_offsets,
              int num_simulations, const uint64_t* noise_game_ids, const uint32_t* noise_plies, float* visits_out, float* scores_out,
              int32_t* depth_out) {
    if (!e) return AZ_ERR_INVALID_ARGUMENT;
    SearchState* st = e->search;
    if (n < 0 || n > st->G) return set_err(e, AZ_ERR_CAPACITY, "more roots than az_config.max_games");
    if (n == 0) return AZ_OK;
    if (!roots || !visits_out) return set_err(e, AZ_ERR_INVALID_ARGUMENT, "null buffer");
    if (num_simulations <= 0 || num_simulations + 2 > st->prm.node_cap)
        return set_err(e, AZ_ERR_CAPACITY, "num_simulations exceeds az_config.num_simulations (pool size)");
    if (e->stub_kind == 0 && !e->net->loaded) return set_err(e, AZ_ERR_NO_WEIGHTS, "az_load_weights has not been called");
    cudaSetDevice(e->cfg.device);
    st->selfplay_active = false;
    SearchParams prm = st->prm;
    prm.n_games = n; prm.S = num_simulations; prm.mode = 0;
    SearchState run = *st;
    run.prm = prm;
    // stage inputs
    AZ_CUDA(e, cudaMemcpyAsync(e->d_wire, roots, (size_t)n * sizeof(az_position), cudaMemcpyHostToDevice, e->stream));
    const az_position* d_hist = nullptr;
    if (history && hist_offsets) {
        // GameState::pos_count already holds the current position (chess.rs:23,52-53): the last history entry IS the root
        for (int g = 0; g < n; g++) {
            if (hist_offsets[g + 1] < hist_offsets[g]) return set_err(e, AZ_ERR_INVALID_ARGUMENT, "hist_offsets must be non-decreasing");
            if (hist_offsets[g + 1] == hist_offsets[g]) continue;
            const az_position& last = history[hist_offsets[g + 1] - 1];
            const az_position& r0 = roots[g];
            if (std::memcmp(last.roles, r0.roles, sizeof last.roles) || std::memcmp(last.colors, r0.colors, sizeof last.colors) ||
                last.turn != r0.turn || last.castling != r0.castling || last.ep_square != r0.ep_square)
                return set_err(e, AZ_ERR_INVALID_ARGUMENT, "the last history entry of a game must be its root position");
        }
        size_t total = hist_offsets[n];
        if (total > e->hist_cap) {
            cudaFree(e->d_hist); e->d_hist = nullptr; e->hist_cap = 0;
            AZ_CUDA(e, cudaMalloc(&e->d_hist, std::max<size_t>(total, 1024) * sizeof(az_position)));
            e->hist_cap = std::max<size_t>(total, 1024);
        }
        if (total) AZ_CUDA(e, cudaMemcpyAsync(e->d_hist, history, total * sizeof(az_position), cudaMemcpyHostToDevice, e->stream));
        AZ_CUDA(e, cudaMemcpyAsync(e->d_hist_off, hist_offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, e->stream));
        d_hist = e->d_hist ? e->d_hist : (const az_position*)e->d_wire;
    }
    unsigned long long* d_ids = nullptr;
    uint32_t* d_plies = nullptr;
    if (noise_game_ids) {
        d_ids = st->d_noise_ids;   // persistent staging: no allocation (= no implicit device synchronisation) per call
        d_plies = st->d_noise_plies;
        AZ_CUDA(e, cudaMemcpyAsync(d_ids, noise_game_ids, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream));
        if (noise_plies) AZ_CUDA(e, cudaMemcpyAsync(d_plies, noise_plies, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
        else AZ_CUDA(e, cudaMemsetAsync(d_plies, 0, (size_t)n * 4, e->stream));
    }
    const int blocks = (n + WARPS - 1) / WARPS;
    run.batch_parity = 0;
    run.ptr.batch_count = run.batch_base;
    run.ptr.batch_zero = nullptr;
    AZ_CUDA(e, cudaMemsetAsync(run.batch_base, 0, 16, e->stream));
    e->n_launches++;
    k_init_search<<<blocks, WARPS * 32, 0, e->stream>>>(run.prm, run.ptr, e->d_wire, d_hist, e->d_hist_off, d_ids, d_plies);
    AZ_CUDA(e, cudaGetLastError());
    int r = evaluate_batch(e, &run);
    if (r) return r;
    // every wave completes at least one simulation per active game
    int* d_flags = run.batch_base + 4;
    int rc = AZ_OK;
    for (int wave = 0;; wave++) {
        r = run_wave(e, &run);
        if (r) { rc = r; break; }
        if (wave + 1 >= num_simulations && (wave + 1 - num_simulations) % 4 == 0) {
            int flags[4] = {0, 0, 0, 0};
            AZ_CUDA(e, cudaMemsetAsync(d_flags, 0, 16, e->stream));
            k_count_active<<<(n + 127) / 128, 128, 0, e->stream>>>(run.prm, run.ptr, d_flags);
            AZ_CUDA(e, cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, e->stream));
            AZ_CUDA(e, cudaStreamSynchronize(e->stream));
            if (flags[1]) { rc = set_err(e, AZ_ERR_CAPACITY, "a per-game node/edge pool overflowed (raise edge_capacity_per_node)"); break; }
            if (flags[0] == 0) break;
            if (wave > 4 * num_simulations + 16) { rc = set_err(e, AZ_ERR_STATE, "search did not converge"); break; }
        }
    }
    if (rc == AZ_OK) {
        float* d_vis = e->d_policy;  // reuse the result buffers for the dense export
        float* d_sc = nullptr;
        if (scores_out) {
            if (!e->d_scores) AZ_CUDA(e, cudaMalloc(&e->d_scores, (size_t)e->max_batch * AZ_ACTION_SPACE * 4));
            d_sc = e->d_scores;
        }
        k_export_root<<<n, 256, 0, e->stream>>>(run.prm, run.ptr, d_vis, d_sc, e->d_count);
        AZ_CUDA(e, cudaGetLastError());
        AZ_CUDA(e, cudaMemcpyAsync(visits_out, d_vis, (size_t)n * AZ_ACTION_SPACE * 4, cudaMemcpyDeviceToHost, e->stream));
        if (scores_out) AZ_CUDA(e, cudaMemcpyAsync(scores_out, d_sc, (size_t)n * AZ_ACTION_SPACE * 4, cudaMemcpyDeviceToHost, e->stream));
        if (depth_out) AZ_CUDA(e, cudaMemcpyAsync(depth_out, e->d_count, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
        AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    }
    return rc;
}

int az_selfplay_begin(az_engine* e, int n_games, uint64_t first_game_id) { return az_selfplay_begin_n(e, n_games, first_game_id, 0); }

int az_selfplay_begin_n(az_engine* e, int n_games, uint64_t first_game_id, uint64_t total_games) {
    if (!e) return AZ_ERR_INVALID_ARGUMENT;
    SearchState* st = e->search;
    if (total_games != 0 && (uint64_t)n_games > total_games) n_games = (int)total_games;   // never more slots than games to play
    if (n_games <= 0 || n_games > st->G) return set_err(e, AZ_ERR_CAPACITY, "more games than az_config.max_games");
    if (e->stub_kind == 0 && !e->net->loaded) return set_err(e, AZ_ERR_NO_WEIGHTS, "az_load_weights has not been called");
    cudaSetDevice(e->cfg.device);
    SearchPtrs& q = st->ptr;
    if (!q.game_samples) {
        int r = salloc(e, st, &q.game_samples, (size_t)st->G * MAX_SAMPLE_PLIES);
        r |= salloc(e, st, &q.out_samples, (size_t)st->prm.sample_cap);
        if (r) return AZ_ERR_OUT_OF_MEMORY;
    }
    st->prm.n_games = n_games; st->prm.mode = 1; st->prm.S = e->cfg.num_simulations;
    st->prm.last_game_id = total_games ? first_game_id + total_games : 0;
    st->wave_counter = 0; st->prm.cache_epoch = 0;
    if (st->prm.cache_mask)  // a new cache per generation (training.rs:342)
        AZ_CUDA(e, cudaMemsetAsync(q.cache_state, 0, ((size_t)st->prm.cache_mask + 1) * sizeof(uint32_t), e->stream));
    Counters zero;
    std::memset(&zero, 0, sizeof zero);
    zero.next_game_id = first_game_id + n_games;
    AZ_CUDA(e, cudaMemcpyAsync(q.counters, &zero, sizeof zero, cudaMemcpyHostToDevice, e->stream));
    AZ_CUDA(e, cudaMemsetAsync(q.stats, 0, (size_t)STAT_STRIPES * STAT_WIDTH * sizeof(unsigned long long), e->stream));
    // the one shared forward of the start position (training.rs:344-350)
    az_position sp;
    az_position_start(&sp);
    AZ_CUDA(e, cudaMemcpyAsync(e->d_wire, &sp, sizeof sp, cudaMemcpyHostToDevice, e->stream));
    int r;
    if (e->stub_kind == 1) {
        DPos d = dpos_from_wire(sp);
        AZ_CUDA(e, cudaMemcpyAsync(q.req_pos, &d, sizeof d, cudaMemcpyHostToDevice, e->stream));
        k_stub_eval<<<1, 128, 0, e->stream>>>(q.req_pos, nullptr, 1, e->stub_seed, e->d_policy, e->d_value);
        r = check_cuda(e, cudaGetLastError(), "k_stub_eval");
    } else if (e->cfg.precision == 1) {
        launch_encode_f32(e->stream, e->d_wire, e->d_planes, 1);
        r = net_forward_fp32(e, e->d_planes, nullptr, 1, e->d_policy, e->d_value);
    } else {
        launch_encode_bf16_wire(e->stream, e->d_wire, e->net->a_in, 1, e->net->in_ch);
        r = net_forward_bf16(e, nullptr, 1, e->d_policy, e->d_value);
    }
    if (r) return r;
    st->batch_parity = 0;
    q.batch_count = st->batch_base;
    q.batch_zero = nullptr;
    AZ_CUDA(e, cudaMemsetAsync(st->batch_base, 0, 8, e->stream));   // both request counters (run_wave alternates them)
    k_start_prior<<<1, 32, 0, e->stream>>>(e->d_policy, q.start_prior);
    const int blocks = (n_games + WARPS - 1) / WARPS;
    e->n_launches += 2;
    k_init_selfplay<<<blocks, WARPS * 32, 0, e->stream>>>(st->prm, st->ptr, first_game_id);
    AZ_CUDA(e, cudaGetLastError());
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    st->selfplay_active = true;
    return AZ_OK;
}

int az_selfplay_step(az_engine* e, int waves, az_selfplay_stats* out) {
    if (!e) return AZ_ERR_INVALID_ARGUMENT;
    SearchState* st = e->search;
    if (!st->selfplay_active) return set_err(e, AZ_ERR_STATE, "az_selfplay_begin has not been called");
    cudaSetDevice(e->cfg.device);
    for (int w = 0; w < waves; w++) {
        int r = run_wave(e, st);
        if (r) return r;
    }
    Counters c;
    unsigned long long stripes[STAT_STRIPES * STAT_WIDTH];
    int flags[4] = {0, 0, 0, 0};
    int* d_flags = st->batch_base + 4;
    AZ_CUDA(e, cudaMemsetAsync(d_flags, 0, 16, e->stream));
    k_count_active<<<(st->prm.n_games + 127) / 128, 128, 0, e->stream>>>(st->prm, st->ptr, d_flags);
    AZ_CUDA(e, cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(&c, st->ptr.counters, sizeof c, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(stripes, st->ptr.stats, sizeof stripes, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    {
        unsigned long long sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned long long evictions = 0, sdepth = 0;
        for (int i = 0; i < STAT_STRIPES * STAT_WIDTH; i++) {
            const int f = i % STAT_WIDTH;
            if (f < 8) sum[f] += stripes[i];
            if (f == 16) evictions += stripes[i];
            if (f == 17) sdepth += stripes[i];
        }
        st->cache_evictions = evictions;
        st->sum_search_depth = sdepth;
#ifdef AZ_ADV_TIMING
        unsigned long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < STAT_STRIPES * STAT_WIDTH; i++) if (i % STAT_WIDTH >= 8 && i % STAT_WIDTH < 16) tc[i % STAT_WIDTH - 8] += stripes[i];
        unsigned long long hist[16] = {0};
        for (int i = 0; i < STAT_STRIPES * STAT_WIDTH; i++) if (i % STAT_WIDTH >= 24) hist[i % STAT_WIDTH - 24] += stripes[i];
        {   // the device counters are cumulative since az_selfplay_begin: report this call's share
            static unsigned long long prev_tc[8], prev_hist[16];
            for (int k = 0; k < 8; k++) { const unsigned long long c = tc[k]; tc[k] = c >= prev_tc[k] ? c - prev_tc[k] : c; prev_tc[k] = c; }
            for (int k = 0; k < 16; k++) { const unsigned long long c = hist[k]; hist[k] = c >= prev_hist[k] ? c - prev_hist[k] : c; prev_hist[k] = c; }
        }
        if (waves > 0) {
        fprintf(stderr, "azb: k_advance warp-time histogram (buckets of 4096 clocks):");
        for (int k = 0; k < 16; k++) fprintf(stderr, " %llu", hist[k]);
        fprintf(stderr, "\n");
        fprintf(stderr, "azb: k_advance phase clocks per warp-wave: consume %.0f select %.0f child %.0f rules %.0f create %.0f submit %.0f move %.0f total %.0f\n",
                (double)tc[0] / ((double)st->prm.n_games * waves), (double)tc[1] / ((double)st->prm.n_games * waves), (double)tc[2] / ((double)st->prm.n_games * waves),
                (double)tc[3] / ((double)st->prm.n_games * waves), (double)tc[4] / ((double)st->prm.n_games * waves), (double)tc[5] / ((double)st->prm.n_games * waves),
                (double)tc[6] / ((double)st->prm.n_games * waves), (double)tc[7] / ((double)st->prm.n_games * waves));
        }
#endif
        c.simulations = sum[0]; c.positions = sum[1]; c.evaluations = sum[2]; c.cache_hits = sum[3];
        c.terminal_leaves = sum[4]; c.games_finished = sum[5]; c.sum_leaf_depth = sum[6]; c.sum_edges = sum[7];
    }
    if (out) {
        out->simulations = c.simulations; out->positions = c.positions; out->evaluations = c.evaluations; out->cache_hits = c.cache_hits;
        out->terminal_leaves = c.terminal_leaves; out->games_finished = c.games_finished; out->sum_leaf_depth = c.sum_leaf_depth;
        out->sum_edges = c.sum_edges; out->waves = (uint64_t)waves; out->pending_samples = std::min<unsigned long long>(c.samples_out, st->prm.sample_cap);
        out->active_games = (uint64_t)(flags[0] + flags[2]); out->parked_games = (uint64_t)flags[2]; out->cache_evictions = st->cache_evictions;
        out->sum_search_depth = st->sum_search_depth;
    }
    if (c.errors || flags[1]) return set_err(e, AZ_ERR_CAPACITY, "a per-game node/edge pool overflowed during self-play (raise edge_capacity_per_node)");
    return AZ_OK;
}

int az_selfplay_drain(az_engine* e, az_sample* out, int max_samples, int* n_out) {
    if (!e || !n_out) return AZ_ERR_INVALID_ARGUMENT;
    SearchState* st = e->search;
    if (!st->selfplay_active) return set_err(e, AZ_ERR_STATE, "az_selfplay_begin has not been called");
    cudaSetDevice(e->cfg.device);
    Counters c;
    AZ_CUDA(e, cudaMemcpy(&c, st->ptr.counters, sizeof c, cudaMemcpyDeviceToHost));
    unsigned long long avail = std::min<unsigned long long>(c.samples_out, st->prm.sample_cap);
    if (avail > (unsigned long long)max_samples) return set_err(e, AZ_ERR_CAPACITY, "drain buffer smaller than the pending sample count");
    if (avail && out) AZ_CUDA(e, cudaMemcpy(out, st->ptr.out_samples, (size_t)avail * sizeof(az_sample), cudaMemcpyDeviceToHost));
    unsigned long long zero = 0;
    AZ_CUDA(e, cudaMemcpy(&st->ptr.counters->samples_out, &zero, sizeof zero, cudaMemcpyHostToDevice));
    *n_out = (int)avail;
    return AZ_OK;
}

int az_selfplay_drain_dev(az_engine* e, az_sample* out_dev, int max_samples, int* n_out) {
    if (!e || !n_out) return AZ_ERR_INVALID_ARGUMENT;
    SearchState* st = e->search;
    if (!st->selfplay_active) return set_err(e, AZ_ERR_STATE, "az_selfplay_begin has not been called");
    cudaSetDevice(e->cfg.device);
    Counters c;
    AZ_CUDA(e, cudaMemcpyAsync(&c, st->ptr.counters, sizeof c, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    const unsigned long long avail = std::min<unsigned long long>(c.samples_out, st->prm.sample_cap);
    if (avail > (unsigned long long)max_samples) return set_err(e, AZ_ERR_CAPACITY, "drain buffer smaller than the pending sample count");
    if (avail && out_dev)
        AZ_CUDA(e, cudaMemcpyAsync(out_dev, st->ptr.out_samples, (size_t)avail * sizeof(az_sample), cudaMemcpyDeviceToDevice, e->stream));
    AZ_CUDA(e, cudaMemsetAsync(&st->ptr.counters->samples_out, 0, sizeof(unsigned long long), e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    *n_out = (int)avail;
    return AZ_OK;
}

/* test hook: the EpisodeSteps the game in slot `slot` has staged so far (they are published when the game ends) */
int az_dbg_selfplay_staged(az_engine* e, int slot, az_sample* out, int max_samples, int* n_out) {
    if (!e || !out || !n_out) return AZ_ERR_INVALID_ARGUMENT;
    SearchState* st = e->search;
    if (!st->selfplay_active || !st->ptr.game_samples) return set_err(e, AZ_ERR_STATE, "az_selfplay_begin has not been called");
    if (slot < 0 || slot >= st->prm.n_games) return set_err(e, AZ_ERR_INVALID_ARGUMENT, "slot out of range");
    cudaSetDevice(e->cfg.device);
    GameCtl c;
    AZ_CUDA(e, cudaMemcpy(&c, st->ptr.ctl + slot, sizeof c, cudaMemcpyDeviceToHost));
    const int n = std::min<int>((int)c.n_samples, max_samples);
    if (n) AZ_CUDA(e, cudaMemcpy(out, st->ptr.game_samples + (size_t)slot * MAX_SAMPLE_PLIES, (size_t)n * sizeof(az_sample), cudaMemcpyDeviceToHost));
    *n_out = n;
    return AZ_OK;
}

}  // extern "C"
