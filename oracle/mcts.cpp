// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// Restates tree.rs:25-289 (MCTree: new / simulation / expand / traverse_new /
// max_subtree_depth / apply_dirichlet_noise) and training.rs:294-338 (run_episode),
// keeping the reference's data layout (dense 4096-wide policy/visits/scores per node,
// boxed children in a map, GameState cloned into every node).
//
// Randomness: the reference draws from unseeded thread_rng (tree.rs:277, training.rs:320),
// so only distributional parity exists.  This project defines a counter-based generator
// (rng_u64) and IEEE-exact log/exp so that the oracle and the CUDA path draw bit-identical
// noise and move samples; the spec is repeated in DESIGN.md.
#include "oracle.hpp"
#include <cstring>
#include <cmath>
#include <unordered_map>
#include <string>

namespace orc {

// ---------------------------------------------------------------------------
// counter-based RNG + deterministic transcendental helpers
// ---------------------------------------------------------------------------
static inline u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

u64 rng_u64(u64 seed, u64 game, u64 ply, u64 stream, u64 counter) {
    u64 h = splitmix64(seed);
    h = splitmix64(h ^ game);
    h = splitmix64(h ^ ply);
    h = splitmix64(h ^ stream);
    h = splitmix64(h ^ counter);
    return h;
}

double rng_uniform(u64 seed, u64 game, u64 ply, u64 stream, u64 counter) {
    u64 h = rng_u64(seed, game, ply, stream, counter);
    return ((double)(h >> 12) + 0.5) * (1.0 / 4503599627370496.0);  // (k + 0.5) * 2^-52, exact
}

// log(x) for x > 0 using only + - * / and exponent extraction (bit-identical on CPU and GPU).
double det_log(double x) {
    u64 bits; std::memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;
    double m; std::memcpy(&m, &bits, 8);            // m in [1,2)
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double t = 1.0 / 23.0;
    t = t * s2 + 1.0 / 21.0; t = t * s2 + 1.0 / 19.0; t = t * s2 + 1.0 / 17.0; t = t * s2 + 1.0 / 15.0;
    t = t * s2 + 1.0 / 13.0; t = t * s2 + 1.0 / 11.0; t = t * s2 + 1.0 / 9.0; t = t * s2 + 1.0 / 7.0;
    t = t * s2 + 1.0 / 5.0; t = t * s2 + 1.0 / 3.0; t = t * s2 + 1.0;
    return (double)e * 0.6931471805599453 + 2.0 * s * t;
}

// exp(x) for |x| < 700
double det_exp(double x) {
    double kf = x * 1.4426950408889634;
    long long k = (long long)(kf < 0 ? kf - 0.5 : kf + 0.5);
    double r = x - (double)k * 0.6931471805599453;
    double t = 1.0 / 6227020800.0;  // 1/13!
    t = t * r + 1.0 / 479001600.0; t = t * r + 1.0 / 39916800.0; t = t * r + 1.0 / 3628800.0;
    t = t * r + 1.0 / 362880.0; t = t * r + 1.0 / 40320.0; t = t * r + 1.0 / 5040.0; t = t * r + 1.0 / 720.0;
    t = t * r + 1.0 / 120.0; t = t * r + 1.0 / 24.0; t = t * r + 1.0 / 6.0; t = t * r + 0.5; t = t * r + 1.0;
    t = t * r + 1.0;
    u64 bits = (u64)(k + 1023) << 52;
    double sc; std::memcpy(&sc, &bits, 8);
    return t * sc;
}

// Gamma(alpha, 1) by Marsaglia-Tsang (boosted for alpha < 1), the same scheme rand_distr 0.4.3 uses.
static double gamma_sample(u64 seed, u64 game, u64 ply, u64 stream, double alpha) {
    u64 ctr = 0;
    auto U = [&]() { return rng_uniform(seed, game, ply, stream, ctr++); };
    double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / std::sqrt(9.0 * d);
    double v, x;
    for (;;) {
        double u1, u2, s;
        do { u1 = 2.0 * U() - 1.0; u2 = 2.0 * U() - 1.0; s = u1 * u1 + u2 * u2; } while (s >= 1.0 || s == 0.0);
        x = u1 * std::sqrt(-2.0 * det_log(s) / s);
        v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = U();
        if (det_log(u) < 0.5 * x * x + d - d * v + d * det_log(v)) break;
    }
    double g = d * v;
    if (alpha < 1.0) { double u = U(); g = g * det_exp(det_log(u) / alpha); }
    return g;
}

void dirichlet_noise(u64 seed, u64 game, u64 ply, float alpha, int n, float* out) {
    std::vector<double> g(n);
    double sum = 0.0;
    for (int i = 0; i < n; i++) { g[i] = gamma_sample(seed, game, ply, 1000 + (u64)i, (double)alpha); }
    for (int i = 0; i < n; i++) sum = sum + g[i];
    for (int i = 0; i < n; i++) out[i] = (float)(g[i] / sum);
}

// ---------------------------------------------------------------------------
// synthetic evaluator (test hook shared with the CUDA path: az_set_evaluator_stub)
// ---------------------------------------------------------------------------
static u64 position_hash(u64 seed, const Pos& p) {
    u64 h = splitmix64(seed);
    for (int i = 0; i < 6; i++) h = splitmix64(h ^ p.role[i]);
    for (int i = 0; i < 2; i++) h = splitmix64(h ^ p.color[i]);
    int ep = pseudo_legal_ep_square(p);
    u64 meta = (u64)p.turn | ((u64)p.castling << 8) | ((u64)(ep + 1) << 16) | ((u64)p.halfmoves << 32) | ((u64)p.fullmoves << 48);
    return splitmix64(h ^ meta);
}

void stub_evaluator(void* ctx, const Pos* pos, float* policy, float* value) {
    u64 seed = ctx ? *(u64*)ctx : 0;
    u64 h = position_hash(seed, *pos);
    u64 total = 0;
    for (int i = 0; i < ACTION_SPACE; i++) {
        u64 r = ((splitmix64(h + (u64)i) >> 52) << 12) + (u64)i + 1;  // <= 2^24, distinct per index
        policy[i] = (float)r; total += r;
    }
    float tf = (float)total;
    for (int i = 0; i < ACTION_SPACE; i++) policy[i] = policy[i] / tf;
    u64 v = splitmix64(h ^ 0xA5A5A5A5A5A5A5A5ULL) >> 40;  // 24 bits
    *value = (float)v * (1.0f / 8388608.0f) - 1.0f;
}

// ---------------------------------------------------------------------------
// tree.rs
// ---------------------------------------------------------------------------
void apply_dirichlet_noise(float* policy, const std::vector<int>& legal, const float* noise, float eps) {  // tree.rs:272-289
    if (legal.size() < 2) return;
    for (int i = 0; i < ACTION_SPACE; i++) policy[i] *= 1.0f - eps;
    for (size_t i = 0; i < legal.size(); i++) policy[legal[i]] += eps * noise[i];
}

static std::vector<int> legal_indices(const Pos& p) {
    MoveList ml; legal_moves(p, ml);
    std::vector<int> v(ml.n);
    for (int i = 0; i < ml.n; i++) v[i] = move_to_index(ml.m[i], p.turn);
    return v;
}

MCTree::MCTree(const float* policy_in, const GameState& st, const float* noise, float eps) : state(st) {  // tree.rs:84-104
    policy.reset(new std::array<float, ACTION_SPACE>);
    visits.reset(new std::array<float, ACTION_SPACE>);
    scores.reset(new std::array<float, ACTION_SPACE>);
    std::memcpy(policy->data(), policy_in, sizeof(float) * ACTION_SPACE);
    visits->fill(0.0f); scores->fill(0.0f);
    moves = legal_indices(state.position);
    if (noise) apply_dirichlet_noise(policy->data(), moves, noise, eps);
}

// tree.rs:117-144 / 180-207
float MCTree::simulation(const SearchParams& sp, eval_fn ev, void* ctx, long* evals) {
    float max_value = -INFINITY; int max_index = 0;
    float total_visits = 0.0f;
    for (int i = 0; i < ACTION_SPACE; i++) total_visits += (*visits)[i];
    total_visits += 1.0f;
    for (int i : moves) {
        float u_value = sp.c_puct * (*policy)[i] * std::sqrt(total_visits) / (1.0f + (*visits)[i]);
        float q_value = (*visits)[i] > 0.0f ? (*scores)[i] / (*visits)[i] : 0.0f;
        float value = q_value + u_value;
        if (value > max_value) { max_value = value; max_index = i; }
    }
    float value;
    auto it = nodes.find(max_index);
    if (it != nodes.end()) value = -it->second->simulation(sp, ev, ctx, evals);
    else value = -expand(sp, ev, ctx, max_index, evals);
    (*scores)[max_index] += value;
    (*visits)[max_index] += 1.0f;
    return value;
}

// tree.rs:146-167 / 209-237
float MCTree::expand(const SearchParams& sp, eval_fn ev, void* ctx, int max_index, long* evals) {
    GameState leaf = state;
    Move action;
    if (!index_to_move(max_index, leaf.position, &action)) abort();  // expect("Illegal move!")
    int res = play_move(leaf, action);
    if (res == ONGOING) {
        std::array<float, ACTION_SPACE> pol; float value;
        ev(ctx, &leaf.position, pol.data(), &value);
        if (evals) (*evals)++;
        nodes[max_index].reset(new MCTree(pol.data(), leaf, nullptr, 0.0f));
        return value;
    }
    if (res == DRAW) return 0.0f;
    if (res == ILLEGAL) abort();
    return -1.0f;
}

float pow_inv_temperature(float visits, float inv_t) {
    if (inv_t == 1.0f || visits == 0.0f) return visits;
    return (float)det_exp(det_log((double)visits) * (double)inv_t);
}

void improved_policy(const float* visits, float temperature, float* out) {  // tree.rs:173-177
    const float inv_t = 1.0f / temperature;
    float wsum = 0.0f;
    for (int i = 0; i < ACTION_SPACE; i++) { out[i] = pow_inv_temperature(visits[i], inv_t); wsum += out[i]; }
    for (int i = 0; i < ACTION_SPACE; i++) out[i] = out[i] / wsum;
}

int MCTree::max_subtree_depth() const {  // tree.rs:258-269
    int best = -1;
    for (auto& kv : nodes) { int d = kv.second->max_subtree_depth(); if (d > best) best = d; }
    return best < 0 ? 0 : 1 + best;
}

}  // namespace orc
