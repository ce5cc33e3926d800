// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  C entry points for ctypes.
//
// run_episode restates training.rs:294-338; the caching evaluator restates the FEN-keyed
// moka cache of training.rs:342,413 / tree.rs:214-218 (EnPassantMode::PseudoLegal keys).
#include "oracle.hpp"
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <unordered_map>
#include <string>
#include <thread>
#include <atomic>
#include <mutex>

using namespace orc;

namespace orc {
struct Net;
Net* net_create_random(u64 seed);
Net* net_create_from_arrays(const float* const* arrays, int n_arrays);
void net_destroy(Net*);
void net_forward(const Net*, const float* planes /*19*64*/, float* policy /*4096*/, float* value);
int net_num_arrays();
size_t net_array_size(int i);
const float* net_array(const Net*, int i);
}

namespace {

struct SharedCache {  // moka Cache<Fen, CacheEntry> (training.rs:342), shared by all games of a generation
    std::unordered_map<std::string, std::pair<std::vector<float>, float>> map;
    std::mutex mu;
    size_t capacity = 500000;  // parameters.rs:4
};

struct CacheCtx {
    eval_fn inner; void* inner_ctx;
    SharedCache* cache;  // nullable: no caching, only counting
    long hits = 0, misses = 0;
};

std::string cache_key(const Pos& p) {
    struct K { u64 b[8]; uint8_t turn, castling; int8_t ep; uint8_t pad; uint16_t hm, fm; } k;
    std::memset(&k, 0, sizeof k);
    std::memcpy(k.b, p.role, 48); std::memcpy(k.b + 6, p.color, 16);
    k.turn = p.turn; k.castling = p.castling; k.ep = (int8_t)pseudo_legal_ep_square(p); k.hm = p.halfmoves; k.fm = p.fullmoves;
    return std::string((const char*)&k, sizeof k);
}

void cached_eval(void* vctx, const Pos* pos, float* policy, float* value) {
    CacheCtx* c = (CacheCtx*)vctx;
    if (!c->cache) { c->misses++; c->inner(c->inner_ctx, pos, policy, value); return; }
    std::string key = cache_key(*pos);
    {
        std::lock_guard<std::mutex> g(c->cache->mu);
        auto it = c->cache->map.find(key);
        if (it != c->cache->map.end()) {
            c->hits++;
            std::memcpy(policy, it->second.first.data(), sizeof(float) * ACTION_SPACE);
            *value = it->second.second;
            return;
        }
    }
    c->misses++;
    c->inner(c->inner_ctx, pos, policy, value);
    std::lock_guard<std::mutex> g(c->cache->mu);
    if (c->cache->map.size() < c->cache->capacity)
        c->cache->map.emplace(key, std::make_pair(std::vector<float>(policy, policy + ACTION_SPACE), *value));
}

void net_eval(void* ctx, const Pos* pos, float* policy, float* value) {
    float planes[19 * 64];
    to_tensor(*pos, planes);
    net_forward((const Net*)ctx, planes, policy, value);
}

GameState state_from_history(const Pos& pos, const Pos* history, int n_hist) {
    // history = every position counted so far (including the current one); empty => GameState::new semantics
    GameState st(pos);
    if (n_hist > 0) {
        st.pos_count.clear();
        for (int i = 0; i < n_hist; i++) st.pos_count[make_key(history[i])] += 1;
    }
    return st;
}

}  // namespace

extern "C" {

struct OrcSearchParams {
    int32_t num_simulations;
    float c_puct, dirichlet_alpha, dirichlet_eps;
    uint32_t temperature_annealing;
    uint64_t seed;
    float temperature;   // TEMPERATURE (parameters.rs:33); 0 is read as 1.0
};

static SearchParams to_sp(const OrcSearchParams* p) {
    SearchParams sp;
    sp.num_simulations = p->num_simulations; sp.c_puct = p->c_puct; sp.dirichlet_alpha = p->dirichlet_alpha;
    sp.dirichlet_eps = p->dirichlet_eps; sp.temperature_annealing = p->temperature_annealing; sp.seed = p->seed;
    sp.temperature = p->temperature > 0.0f ? p->temperature : 1.0f;
    return sp;
}

void orc_init() { init_tables(); }
int orc_pos_from_fen(const char* fen, Pos* out) { return pos_from_fen(fen, out) ? 0 : -1; }
void orc_startpos(Pos* out) { *out = startpos(); }

int orc_legal_moves(const Pos* p, uint16_t* moves_out, uint16_t* index_out) {
    MoveList ml; legal_moves(*p, ml);
    for (int i = 0; i < ml.n; i++) {
        if (moves_out) moves_out[i] = encode_move(ml.m[i]);
        if (index_out) index_out[i] = (uint16_t)move_to_index(ml.m[i], p->turn);
    }
    return ml.n;
}

uint64_t orc_perft(const Pos* p, int depth) { return perft(*p, depth); }
int orc_minimax_scores(const Pos* p, int depth, int32_t* scores) { return minimax_scores(*p, depth, scores); }
int orc_outcome(const Pos* p) { return outcome(*p); }
int orc_legal_ep_square(const Pos* p) { return legal_ep_square(*p); }
int orc_pseudo_legal_ep_square(const Pos* p) { return pseudo_legal_ep_square(*p); }
void orc_to_tensor(const Pos* p, float* out) { to_tensor(*p, out); }

// play a wire-encoded move without any rule glue (for building test positions)
int orc_play_encoded(Pos* p, uint16_t mv) {
    MoveList ml; legal_moves(*p, ml);
    for (int i = 0; i < ml.n; i++) if (encode_move(ml.m[i]) == mv) { play_unchecked(*p, ml.m[i]); return 0; }
    return -1;
}

int orc_move_to_index(const Pos* p, uint16_t mv) {
    MoveList ml; legal_moves(*p, ml);
    for (int i = 0; i < ml.n; i++) if (encode_move(ml.m[i]) == mv) return move_to_index(ml.m[i], p->turn);
    return -1;
}

// returns 1 and the encoded move, or 0 (None)
int orc_index_to_move(const Pos* p, int index, uint16_t* mv_out) {
    Move m;
    if (!index_to_move(index, *p, &m)) return 0;
    *mv_out = encode_move(m);
    return 1;
}

// chess.rs play_move through an action index (tree.rs:211-212). pos is updated in place when legal.
int orc_play_move(Pos* pos, const Pos* history, int n_hist, int action_index) {
    GameState st = state_from_history(*pos, history, n_hist);
    Move m;
    if (!index_to_move(action_index, st.position, &m)) return ILLEGAL;
    int r = play_move(st, m);
    if (r != ILLEGAL) *pos = st.position;
    return r;
}

// ---- batched forms for the 65,536-position differential of BASELINE config 2 (one ctypes call instead of 65,536) ----------
// Seeded random playouts: position i starts from roots[i % n_roots] and plays a uniformly random legal
// move for a random number of plies in [0, max_plies] (stopping early at a finished game).  hist_out (nullable) receives
// every position on the way INCLUDING the final one, hist_off[n + 1] delimits games; returns the total history length.
int64_t orc_playout_corpus(uint64_t seed, int n, int max_plies, const Pos* roots, int n_roots, Pos* pos_out, Pos* hist_out,
                           int64_t hist_cap, uint32_t* hist_off) {
    int64_t total = 0;
    for (int i = 0; i < n; i++) {
        Pos p = roots[i % n_roots];
        const int plies = (int)(rng_u64(seed, (uint64_t)i, 0, 7, 0) % (uint64_t)(max_plies + 1));
        if (hist_off) hist_off[i] = (uint32_t)total;
        if (hist_out && total < hist_cap) hist_out[total] = p;
        total++;
        for (int k = 0; k < plies; k++) {
            MoveList ml; legal_moves(p, ml);
            if (ml.n == 0 || outcome(p) != 0) break;
            play_unchecked(p, ml.m[rng_u64(seed, (uint64_t)i, (uint64_t)k + 1, 7, 1) % (uint64_t)ml.n]);
            if (hist_out && total < hist_cap) hist_out[total] = p;
            total++;
        }
        pos_out[i] = p;
    }
    if (hist_off) hist_off[n] = (uint32_t)total;
    return total;
}

// legal_moves + move_to_index for n positions: moves_out / index_out [n][256] (unused entries 0xFFFF / untouched), count_out [n]
void orc_legal_moves_batch(const Pos* pos, int n, uint16_t* moves_out, uint16_t* index_out, int32_t* count_out) {
    for (int i = 0; i < n; i++) {
        MoveList ml; legal_moves(pos[i], ml);
        count_out[i] = ml.n;
        for (int k = 0; k < 256; k++) {
            moves_out[(size_t)i * 256 + k] = k < ml.n ? encode_move(ml.m[k]) : (uint16_t)0xFFFF;
            if (k < ml.n) index_out[(size_t)i * 256 + k] = (uint16_t)move_to_index(ml.m[k], pos[i].turn);
        }
    }
}

void orc_to_tensor_batch(const Pos* pos, int n, float* out) {
    for (int i = 0; i < n; i++) to_tensor(pos[i], out + (size_t)i * 19 * 64);
}

// play_move (chess.rs:36-63) for n games: pos updated in place when legal, result_out [n]
void orc_play_move_batch(Pos* pos, int n, const Pos* history, const uint32_t* hist_off, const uint16_t* action_index, int32_t* result_out) {
    for (int i = 0; i < n; i++) {
        const Pos* h = history ? history + hist_off[i] : nullptr;
        const int nh = history ? (int)(hist_off[i + 1] - hist_off[i]) : 0;
        result_out[i] = orc_play_move(&pos[i], h, nh, (int)action_index[i]);
    }
}

// index_to_move for n (position, index) pairs: 0xFFFF where the reference returns None
void orc_index_to_move_batch(const Pos* pos, int n, const uint16_t* index, uint16_t* moves_out) {
    for (int i = 0; i < n; i++) {
        Move m;
        moves_out[i] = index_to_move((int)index[i], pos[i], &m) ? encode_move(m) : (uint16_t)0xFFFF;
    }
}

void orc_stub_eval(uint64_t seed, const Pos* p, float* policy, float* value) { stub_evaluator(&seed, p, policy, value); }

uint64_t orc_rng_u64(uint64_t seed, uint64_t game, uint64_t ply, uint64_t stream, uint64_t counter) {
    return rng_u64(seed, game, ply, stream, counter);
}
void orc_improved_policy(const float* visits, float temperature, float* out) { improved_policy(visits, temperature, out); }
double orc_det_log(double x) { return det_log(x); }
double orc_det_exp(double x) { return det_exp(x); }
void orc_dirichlet(uint64_t seed, uint64_t game, uint64_t ply, float alpha, int n, float* out) {
    dirichlet_noise(seed, game, ply, alpha, n, out);
}

// ---- network -----------------------------------------------------------------
void* orc_net_create_random(uint64_t seed) { return net_create_random(seed); }
void* orc_net_create(const float* const* arrays, int n) { return net_create_from_arrays(arrays, n); }
void orc_net_destroy(void* n) { net_destroy((Net*)n); }
int orc_net_num_arrays() { return net_num_arrays(); }
uint64_t orc_net_array_size(int i) { return net_array_size(i); }
const float* orc_net_array(void* n, int i) { return net_array((const Net*)n, i); }
void orc_net_forward_planes(void* n, const float* planes, int count, float* policy, float* value) {
    for (int i = 0; i < count; i++) net_forward((const Net*)n, planes + (size_t)i * 19 * 64, policy + (size_t)i * ACTION_SPACE, value + i);
}
void orc_net_forward_pos(void* n, const Pos* pos, int count, float* policy, float* value) {
    for (int i = 0; i < count; i++) net_eval(n, pos + i, policy + (size_t)i * ACTION_SPACE, value + i);
}

// ---- search ------------------------------------------------------------------
// evaluator selection: kind 0 = stub(seed = stub_seed), 1 = oracle fp32 net (ctx = Net*), 2 = callback
struct OrcEvaluator {
    int32_t kind;
    uint64_t stub_seed;
    void* net;
    eval_fn callback;
    void* callback_ctx;
};

static void pick_eval(const OrcEvaluator* e, eval_fn* fn, void** ctx, uint64_t* seed_store) {
    if (e->kind == 0) { *seed_store = e->stub_seed; *fn = stub_evaluator; *ctx = seed_store; }
    else if (e->kind == 1) { *fn = net_eval; *ctx = e->net; }
    else { *fn = e->callback; *ctx = e->callback_ctx; }
}

// MCTree::init + monte_carlo_tree_search (tree.rs:37-64,106-115). noise_game < 0 disables root noise.
// visits_out[4096] receives raw visit counts (improved_policy = visits / num_simulations when T = 1).
int orc_search(const Pos* root, const Pos* history, int n_hist, const OrcSearchParams* params, const OrcEvaluator* evr,
               int64_t noise_game, int64_t noise_ply, float* visits_out, float* scores_out, int* depth_out, long* evals_out) {
    SearchParams sp = to_sp(params);
    eval_fn ev; void* ctx; uint64_t seed_store;
    pick_eval(evr, &ev, &ctx, &seed_store);
    GameState st = state_from_history(*root, history, n_hist);
    std::array<float, ACTION_SPACE> pol; float v;
    ev(ctx, &st.position, pol.data(), &v);
    long evals = 1;
    std::vector<float> noise;
    const float* noise_ptr = nullptr;
    if (noise_game >= 0) {
        MoveList ml; legal_moves(st.position, ml);
        noise.resize(ml.n > 0 ? ml.n : 1);
        if (ml.n >= 2) dirichlet_noise(sp.seed, (u64)noise_game, (u64)noise_ply, sp.dirichlet_alpha, ml.n, noise.data());
        noise_ptr = noise.data();
    }
    MCTree tree(pol.data(), st, noise_ptr, sp.dirichlet_eps);
    for (int i = 0; i < sp.num_simulations; i++) tree.simulation(sp, ev, ctx, &evals);
    std::memcpy(visits_out, tree.visits->data(), sizeof(float) * ACTION_SPACE);
    if (scores_out) std::memcpy(scores_out, tree.scores->data(), sizeof(float) * ACTION_SPACE);
    if (depth_out) *depth_out = tree.max_subtree_depth();
    if (evals_out) *evals_out = evals;
    return 0;
}

// ---- self-play episode (training.rs:294-338) ----------------------------------------
struct OrcEpisodeStats {
    int32_t n_steps;
    int32_t result;          // GameResult of the final move
    int64_t simulations;
    int64_t evals;           // evaluator calls that reached the network (cache misses)
    int64_t cache_hits;
    double seconds;
};

void* orc_cache_create() { return new SharedCache(); }
void orc_cache_destroy(void* c) { delete (SharedCache*)c; }
uint64_t orc_cache_size(void* c) { return ((SharedCache*)c)->map.size(); }

// positions_out[max_steps], visits_out[max_steps][4096] (nullable), final_value_out[max_steps], depth_out[max_steps],
// action_out[max_steps].  cache: handle from orc_cache_create (the FEN-keyed NN cache) or NULL.
// max_steps bounds the number of plies played (a truncated game reports result ONGOING and zero final values).
int orc_selfplay_episode(const OrcSearchParams* params, const OrcEvaluator* evr, uint64_t game_id, void* cache_handle,
                         int max_steps, Pos* positions_out, float* visits_out, float* final_value_out,
                         int32_t* depth_out, int32_t* action_out, OrcEpisodeStats* stats) {
    SearchParams sp = to_sp(params);
    eval_fn ev; void* ctx; uint64_t seed_store;
    pick_eval(evr, &ev, &ctx, &seed_store);
    CacheCtx cc{ev, ctx, (SharedCache*)cache_handle};
    auto t0 = std::chrono::steady_clock::now();

    GameState state;
    std::array<float, ACTION_SPACE> root_pol; float root_v;
    ev(ctx, &state.position, root_pol.data(), &root_v);  // training.rs:344-350 (one forward of the start position)
    std::vector<float> noise(256);
    auto make_noise = [&](const Pos& p, u64 ply) -> const float* {
        MoveList ml; legal_moves(p, ml);
        if (ml.n >= 2) dirichlet_noise(sp.seed, game_id, ply, sp.dirichlet_alpha, ml.n, noise.data());
        return noise.data();
    };
    std::unique_ptr<MCTree> tree(new MCTree(root_pol.data(), state, make_noise(state.position, 0), sp.dirichlet_eps));

    int n = 0; int result = ONGOING; float outcome_value = 0.0f; long sims = 0;
    std::vector<float> turn_sign;
    for (u64 ply = 0; (int)ply < max_steps; ply++) {
        for (int i = 0; i < sp.num_simulations; i++) tree->simulation(sp, cached_eval, &cc, nullptr);
        sims += sp.num_simulations;
        // improved policy = visits^(1/T)/sum (tree.rs:173-177; T = 1.0 by default, then powf is the identity)
        std::array<float, ACTION_SPACE> improved;
        improved_policy(tree->visits->data(), sp.temperature, improved.data());
        float turn = state.position.turn == WHITE ? 1.0f : -1.0f;
        positions_out[n] = state.position;
        if (visits_out) std::memcpy(visits_out + (size_t)n * ACTION_SPACE, tree->visits->data(), sizeof(float) * ACTION_SPACE);
        depth_out[n] = tree->max_subtree_depth();
        turn_sign.push_back(turn);
        int action_index;
        if (state.position.fullmoves >= sp.temperature_annealing) {
            // Iterator::max_by keeps the LAST maximum (training.rs:311-316)
            action_index = 0; float best = improved[0];
            for (int i = 1; i < ACTION_SPACE; i++) if (improved[i] >= best) { best = improved[i]; action_index = i; }
        } else {
            // WeightedIndex restated: cumulative f32 weights, uniform draw in [0, total), first cum > draw
            float total = 0.0f;
            for (int i = 0; i < ACTION_SPACE; i++) total += improved[i];
            u64 h = rng_u64(sp.seed, game_id, ply, 1, 0);
            float u = (float)(h >> 40) * (1.0f / 16777216.0f);
            float chosen = u * total;
            float cum = 0.0f; action_index = -1; int last_pos = 0;
            for (int i = 0; i < ACTION_SPACE; i++) {
                if (improved[i] > 0.0f) last_pos = i;
                cum += improved[i];
                if (cum > chosen && improved[i] > 0.0f) { action_index = i; break; }
            }
            if (action_index < 0) action_index = last_pos;
        }
        action_out[n] = action_index;
        n++;
        Move action;
        if (!index_to_move(action_index, state.position, &action)) abort();
        int r = play_move(state, action);
        if (r != ONGOING) {
            result = r;
            outcome_value = (r == DRAW) ? 0.0f : turn;
            break;
        }
        // traverse_new (tree.rs:239-256): keep the child's policy/moves/state, fresh noise, zero statistics
        auto it = tree->nodes.find(action_index);
        if (it == tree->nodes.end()) abort();
        std::unique_ptr<MCTree> child = std::move(it->second);
        const float* nz = make_noise(child->state.position, ply + 1);
        tree.reset(new MCTree(child->policy->data(), child->state, nz, sp.dirichlet_eps));
    }
    float decay = 1.0f - ((float)state.position.fullmoves / (2.0f * (float)NUM_FULLMOVES));
    for (int i = 0; i < n; i++) final_value_out[i] = turn_sign[i] * (outcome_value * decay);
    auto t1 = std::chrono::steady_clock::now();
    stats->n_steps = n; stats->result = result; stats->simulations = sims;
    stats->evals = cc.misses + 1;
    stats->cache_hits = cc.hits;
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
}

// Perft over many roots on `threads` host threads (CPU baseline for BASELINE config 2).
void orc_perft_batch(const Pos* roots, int n, int depth, int threads, uint64_t* out) {
    std::atomic<int> next(0);
    auto work = [&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; out[i] = perft(roots[i], depth); } };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(work);
    for (auto& t : th) t.join();
}

}  // extern "C"
