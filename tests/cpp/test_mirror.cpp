// Exercises the C++ mirror (include/az_b200.hpp) the way the reference's callers use the Rust items, and prints the
// results as JSON lines for tests/test_cpp_mirror_gpu.py to compare with the oracle.
#include <cstdio>
#include <cstdlib>
#include "az_b200.hpp"

using namespace az;

static std::size_t index_of(Engine& e, const Position& p, int from, int to) {
    std::vector<uint16_t> idx;
    auto mv = legal_moves(e, p, &idx);
    for (std::size_t i = 0; i < mv.size(); i++)
        if ((mv[i] & 63) == from && ((mv[i] >> 6) & 63) == to) return idx[i];
    std::fprintf(stderr, "move %d-%d not legal\n", from, to);
    std::exit(2);
}

int main(int argc, char** argv) {
    const uint64_t stub_seed = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 5;
    az_config cfg = Engine::default_config();
    cfg.max_games = 8; cfg.num_simulations = 48; cfg.seed = 42;
    Engine eng(cfg);
    if (az_set_evaluator_stub(eng.handle(), 1, stub_seed) != AZ_OK) return 3;
    AlphaZero model(eng);

    // chess.rs: knights out and back twice = threefold repetition (chess.rs:52-60)
    GameState st;
    const int shuffle[8][2] = {{6, 21}, {62, 45}, {21, 6}, {45, 62}, {6, 21}, {62, 45}, {21, 6}, {45, 62}};
    GameResult last = GameResult::Ongoing;
    int plies = 0;
    for (auto& m : shuffle) {
        last = play_move(eng, st, index_of(eng, st.position, m[0], m[1]));
        plies++;
        if (last != GameResult::Ongoing) break;
    }
    std::printf("{\"test\": \"repetition\", \"plies\": %d, \"result\": %d, \"counted\": %zu}\n", plies, (int)last, st.pos_count.size());
    GameState fresh;
    std::printf("{\"test\": \"illegal\", \"result\": %d, \"none\": %d, \"e2e4\": %zu}\n", (int)play_move(eng, fresh, 0),
                (int)!index_to_move(eng, 0, fresh.position).has_value(), index_of(eng, fresh.position, 12, 28));
    auto planes = to_tensor(eng, fresh.position);
    float psum = 0; for (float v : planes) psum += v;
    std::printf("{\"test\": \"planes\", \"sum\": %.6f}\n", psum);

    // tree.rs: MCTree::init + monte_carlo_tree_search, then traverse_new and search again
    MCTree tree = MCTree::init(model, GameState(), false);
    Policy pol = tree.monte_carlo_tree_search(48);
    std::printf("{\"test\": \"search\", \"depth\": %zu, \"visits\": [", tree.max_subtree_depth());
    bool first = true;
    for (std::size_t i = 0; i < ACTION_SPACE; i++)
        if (pol[i] > 0) { std::printf("%s[%zu, %.1f]", first ? "" : ", ", i, pol[i] * 48.0f); first = false; }
    std::printf("]}\n");
    std::size_t best = 0;
    for (std::size_t i = 0; i < ACTION_SPACE; i++) if (pol[i] >= pol[best]) best = i;
    MCTree next = std::move(tree).traverse_new(best, false);
    Policy pol2 = next.monte_carlo_tree_search(48);
    std::printf("{\"test\": \"search2\", \"action\": %zu, \"depth\": %zu, \"visits\": [", best, next.max_subtree_depth());
    first = true;
    for (std::size_t i = 0; i < ACTION_SPACE; i++)
        if (pol2[i] > 0) { std::printf("%s[%zu, %.1f]", first ? "" : ", ", i, pol2[i] * 48.0f); first = false; }
    std::printf("]}\n");

    // training.rs: run_all_episodes
    auto [avg_batch, steps] = run_all_episodes(model, 4, 0, 16);
    double vsum = 0; std::size_t dsum = 0;
    for (auto& s : steps) { vsum += s.final_value; dsum += s.search_depth; }
    std::printf("{\"test\": \"episodes\", \"steps\": %zu, \"value_sum\": %.6f, \"depth_sum\": %zu, \"avg_batch\": %.3f}\n", steps.size(), vsum,
                dsum, avg_batch);

    // memory.rs: the finished games' steps go from device memory into the replay buffer; sample a batch
    {
        az_config c2 = Engine::default_config();
        c2.max_games = 8; c2.num_simulations = 24; c2.seed = 7;
        Engine e2(c2);
        if (az_set_evaluator_stub(e2.handle(), 1, stub_seed) != AZ_OK) return 3;
        ReplayBuffer rb(e2, 1000, 64);
        if (az_selfplay_begin(e2.handle(), 8, 0) != AZ_OK) return 4;
        az_selfplay_stats st2{};
        std::size_t added = 0, unique = 0;
        while (st2.games_finished < 8) {
            e2.check(az_selfplay_step(e2.handle(), 32, &st2), "az_selfplay_step");
            if (st2.pending_samples) { auto [n, nu] = rb.add_pending(); added += n; unique += nu; }
        }
        TrainingBatch b = rb.sample(64, 3);
        double psum = 0; for (float v : b.policy) psum += v;
        std::printf("{\"test\": \"replay\", \"added\": %zu, \"unique\": %zu, \"len\": %zu, \"batch\": %zu, \"policy_sum\": %.4f}\n", added, unique,
                    rb.len(), b.n, psum);
    }

    // chess.rs:295-318: the minimax bot finds the mate in one and reports None without legal moves
    {
        Position mate{}, mated{};
        az_position_from_fen("6k1/5ppp/8/8/8/8/8/R6K w - - 0 1", &mate);
        az_position_from_fen("R5k1/5ppp/8/8/8/8/8/7K b - - 1 1", &mated);
        auto mv = get_best_move(eng, mate, 2);
        std::printf("{\"test\": \"minimax\", \"from\": %d, \"to\": %d, \"none\": %d}\n", mv ? (int)(*mv & 63) : -1, mv ? (int)((*mv >> 6) & 63) : -1,
                    (int)!get_best_move(eng, mated, 2).has_value());
    }
    return 0;
}
