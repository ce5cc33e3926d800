// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the self-play hot path of AlexandreGac/alphazero-chess
// (reference mounted read-only at /root/reference; nothing is copied from it).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (alphazero-chess_b200/)
// never links, imports or calls anything in this directory.
//
// Parity status: the reference has no tests and its chess rules live in the
// un-vendored crate shakmaty 0.29.0 (Cargo.lock:4573), so the rules here
// restate shakmaty's published algorithm and are pinned by the public perft
// tables (tests/golden/perft.json) and hand-derived codec vectors
// (SURVEY.md §8c).  Move ORDER inside legal_moves() follows shakmaty's
// generator structure from memory and is "parity unpinned" (it only matters
// on exact f32 ties and for which Dirichlet component lands on which move).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <map>
#include <array>
#include <memory>

namespace orc {

using u64 = uint64_t;

enum Role : uint8_t { PAWN = 0, KNIGHT = 1, BISHOP = 2, ROOK = 3, QUEEN = 4, KING = 5, NO_ROLE = 6 };
enum Color : uint8_t { WHITE = 0, BLACK = 1 };

// Wire format shared with include/az_b200.h (az_position, 72 bytes).
struct Pos {
    u64 role[6];       // pawn, knight, bishop, rook, queen, king
    u64 color[2];      // white, black
    uint8_t turn;      // 0 white, 1 black
    uint8_t castling;  // bit0 white K-side (h1), bit1 white Q-side (a1), bit2 black K (h8), bit3 black Q (a8)
    int8_t ep;         // -1, or the square behind the last double push (shakmaty keeps it after EVERY double push)
    uint8_t reserved;
    uint16_t halfmoves;
    uint16_t fullmoves;
};
static_assert(sizeof(Pos) == 72, "wire format");

enum MoveKind : uint8_t { NORMAL = 0, EN_PASSANT = 1, CASTLE = 2 };

// shakmaty::Move restated: Castle is king-takes-rook (from = king, to = rook).
struct Move {
    uint8_t kind;
    uint8_t role;
    uint8_t from;
    uint8_t to;
    uint8_t capture;    // NO_ROLE if none
    uint8_t promotion;  // NO_ROLE if none
};

// u16 wire encoding: from | to<<6 | promo<<12 (0 none,1 N,2 B,3 R,4 Q) | special<<15 (castle or en passant)
uint16_t encode_move(const Move& m);

struct MoveList {
    Move m[256];
    int n = 0;
    void push(const Move& mv) { m[n++] = mv; }
};

void init_tables();
Pos startpos();
bool pos_from_fen(const char* fen, Pos* out);

u64 occupied(const Pos& p);
u64 attacks_to(const Pos& p, int sq, int attacker_color, u64 occ);
u64 checkers(const Pos& p);
void legal_moves(const Pos& p, MoveList& out);
void play_unchecked(Pos& p, const Move& m);
bool is_legal(const Pos& p, const Move& m);
int legal_ep_square(const Pos& p);         // -1 if none
int pseudo_legal_ep_square(const Pos& p);  // -1 if none
bool is_insufficient_material(const Pos& p);
// 0 unknown, 1 draw, 2 white wins, 3 black wins   (shakmaty Position::outcome)
int outcome(const Pos& p);
u64 perft(const Pos& p, int depth);
int evaluate_material(const Pos& p);                         // chess.rs:247-264
int negamax(const Pos& p, int depth);                        // chess.rs:266-292
int minimax_scores(const Pos& p, int depth, int32_t* scores); // chess.rs:295-318 (scores of the legal moves, in order)

// ---- chess.rs restatement -------------------------------------------------
constexpr int ACTION_SPACE = 4096;
constexpr uint32_t NUM_HALFMOVES = 100;  // chess.rs:9
constexpr uint32_t NUM_FULLMOVES = 200;  // chess.rs:10
constexpr int REPETITIONS = 3;           // chess.rs:11

enum GameResult : int { ONGOING = 0, DRAW = 1, WHITE_WINS = 2, BLACK_WINS = 3, ILLEGAL = -1 };

// Key with shakmaty's Chess Eq semantics: board, turn, castling rights, LEGAL ep square.
struct PosKey {
    u64 role[6];
    u64 color[2];
    uint8_t turn, castling;
    int8_t legal_ep;
    bool operator<(const PosKey& o) const;
};
PosKey make_key(const Pos& p);

struct GameState {            // chess.rs:13-27
    Pos position;
    std::map<PosKey, int> pos_count;
    GameState();
    explicit GameState(const Pos& p);
};

int play_move(GameState& st, const Move& m);                 // chess.rs:36-63
int move_to_index(const Move& m, int turn);                  // chess.rs:73-116
bool index_to_move(int index, const Pos& p, Move* out);      // chess.rs:118-171
void to_tensor(const Pos& p, float* out /*19*64*/);          // chess.rs:191-245

// ---- evaluator ------------------------------------------------------------
// fills policy[4096] (softmax probabilities over ALL indices) and *value
typedef void (*eval_fn)(void* ctx, const Pos* pos, float* policy, float* value);
void stub_evaluator(void* ctx /* u64* seed */, const Pos* pos, float* policy, float* value);

// ---- deterministic RNG (project-defined; the reference uses unseeded thread_rng) ----
u64 rng_u64(u64 seed, u64 game, u64 ply, u64 stream, u64 counter);
double rng_uniform(u64 seed, u64 game, u64 ply, u64 stream, u64 counter);  // (0,1)
double det_log(double x);
double det_exp(double x);
void dirichlet_noise(u64 seed, u64 game, u64 ply, float alpha, int n, float* out);

// ---- tree.rs restatement --------------------------------------------------
struct SearchParams {
    int num_simulations = 256;    // parameters.rs:32
    float c_puct = 3.0f;          // parameters.rs:34
    float dirichlet_alpha = 0.3f; // parameters.rs:28
    float dirichlet_eps = 0.25f;  // parameters.rs:29
    uint32_t temperature_annealing = 15;  // parameters.rs:31
    u64 seed = 42;
    float temperature = 1.0f;     // parameters.rs:33
};
// tree.rs:173-177: weights[i] = visits[i].powf(1.0 / T); out[i] = weights[i] / sum(weights).  powf is restated as
// (float)exp(log((double)x) * (double)(1/T)) with the project's det_log / det_exp (identity when T == 1), shared with the
// CUDA engine; against libm's powf this can differ in the last bit for T != 1.
float pow_inv_temperature(float visits, float inv_t);
void improved_policy(const float* visits, float temperature, float* out);

struct MCTree {  // tree.rs:25-34 (dense 4096-wide arrays, as the reference)
    std::map<int, std::unique_ptr<MCTree>> nodes;
    std::vector<int> moves;
    GameState state;
    std::unique_ptr<std::array<float, ACTION_SPACE>> policy, visits, scores;

    MCTree(const float* policy_in, const GameState& st, const float* noise /*nullable, one per move*/,
           float eps);
    float simulation(const SearchParams& sp, eval_fn ev, void* ctx, long* evals);
    float expand(const SearchParams& sp, eval_fn ev, void* ctx, int max_index, long* evals);
    int max_subtree_depth() const;
};

void apply_dirichlet_noise(float* policy, const std::vector<int>& legal, const float* noise, float eps);

}  // namespace orc
