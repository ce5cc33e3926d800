#include <cstdint>
// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// Chess rules: restates shakmaty 0.29.0 (pinned in /root/reference/Cargo.lock:4573-4576,
// NOT vendored under /root/reference) at the call sites the reference uses:
//   legal_moves()      tree.rs:39,86
//   is_legal / play_unchecked / outcome   chess.rs:38,42,43
//   Hash/Eq of Chess   chess.rs:16,52
//   pseudo_legal_ep_square, castles().has  chess.rs:218-232
// and the reference's own glue in chess.rs:36-245.
#include "oracle.hpp"
#include <cstring>
#include <cstdlib>
#include <cctype>
#include <tuple>

namespace orc {

static u64 KNIGHT_ATT[64], KING_ATT[64], PAWN_ATT[2][64];
static u64 RAY[8][64];  // N, NE, E, SE, S, SW, W, NW
static u64 BETWEEN[64][64], LINE[64][64];
static bool g_init = false;

static inline int lsb(u64 b) { return __builtin_ctzll(b); }
static inline int msb(u64 b) { return 63 - __builtin_clzll(b); }
static inline int popcnt(u64 b) { return __builtin_popcountll(b); }
static inline u64 bit(int s) { return 1ULL << s; }
static const int DF[8] = {0, 1, 1, 1, 0, -1, -1, -1};
static const int DR[8] = {1, 1, 0, -1, -1, -1, 0, 1};

static u64 ray_attacks(int sq, u64 occ, int dir) {
    u64 r = RAY[dir][sq];
    u64 blockers = r & occ;
    if (blockers) {
        // directions with increasing square index: N(0), NE(1), E(2), NW(7)
        bool positive = (dir == 0 || dir == 1 || dir == 2 || dir == 7);
        int b = positive ? lsb(blockers) : msb(blockers);
        r ^= RAY[dir][b];
    }
    return r;
}
static u64 rook_attacks(int sq, u64 occ) {
    return ray_attacks(sq, occ, 0) | ray_attacks(sq, occ, 2) | ray_attacks(sq, occ, 4) | ray_attacks(sq, occ, 6);
}
static u64 bishop_attacks(int sq, u64 occ) {
    return ray_attacks(sq, occ, 1) | ray_attacks(sq, occ, 3) | ray_attacks(sq, occ, 5) | ray_attacks(sq, occ, 7);
}

void init_tables() {
    if (g_init) return;
    for (int s = 0; s < 64; s++) {
        int f = s & 7, r = s >> 3;
        static const int kf[8] = {1, 2, 2, 1, -1, -2, -2, -1}, kr[8] = {2, 1, -1, -2, -2, -1, 1, 2};
        u64 n = 0, k = 0;
        for (int i = 0; i < 8; i++) {
            int nf = f + kf[i], nr = r + kr[i];
            if (nf >= 0 && nf < 8 && nr >= 0 && nr < 8) n |= bit(nr * 8 + nf);
            nf = f + DF[i]; nr = r + DR[i];
            if (nf >= 0 && nf < 8 && nr >= 0 && nr < 8) k |= bit(nr * 8 + nf);
        }
        KNIGHT_ATT[s] = n; KING_ATT[s] = k;
        u64 w = 0, b = 0;
        if (r < 7) { if (f > 0) w |= bit(s + 7); if (f < 7) w |= bit(s + 9); }
        if (r > 0) { if (f > 0) b |= bit(s - 9); if (f < 7) b |= bit(s - 7); }
        PAWN_ATT[WHITE][s] = w; PAWN_ATT[BLACK][s] = b;
        for (int d = 0; d < 8; d++) {
            u64 ray = 0; int nf = f + DF[d], nr = r + DR[d];
            while (nf >= 0 && nf < 8 && nr >= 0 && nr < 8) { ray |= bit(nr * 8 + nf); nf += DF[d]; nr += DR[d]; }
            RAY[d][s] = ray;
        }
    }
    for (int a = 0; a < 64; a++) for (int b = 0; b < 64; b++) {
        BETWEEN[a][b] = 0; LINE[a][b] = 0;
        for (int d = 0; d < 8; d++) if (RAY[d][a] & bit(b)) {
            BETWEEN[a][b] = RAY[d][a] & RAY[(d + 4) & 7][b];
            LINE[a][b] = RAY[d][a] | RAY[(d + 4) & 7][a] | bit(a);
        }
    }
    g_init = true;
}

uint16_t encode_move(const Move& m) {
    static const int promo_code[7] = {0, 1, 2, 3, 4, 0, 0};
    uint16_t v = (uint16_t)(m.from | (m.to << 6) | (promo_code[m.promotion] << 12));
    if (m.kind != NORMAL) v |= 0x8000;
    return v;
}

Pos startpos() {
    Pos p; std::memset(&p, 0, sizeof p);
    pos_from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", &p);
    return p;
}

bool pos_from_fen(const char* fen, Pos* out) {
    init_tables();
    Pos p; std::memset(&p, 0, sizeof p); p.ep = -1; p.halfmoves = 0; p.fullmoves = 1;
    int r = 7, f = 0; const char* c = fen;
    for (; *c && *c != ' '; c++) {
        if (*c == '/') { r--; f = 0; continue; }
        if (isdigit((unsigned char)*c)) { f += *c - '0'; continue; }
        int col = isupper((unsigned char)*c) ? WHITE : BLACK; int role;
        switch (tolower(*c)) {
            case 'p': role = PAWN; break; case 'n': role = KNIGHT; break; case 'b': role = BISHOP; break;
            case 'r': role = ROOK; break; case 'q': role = QUEEN; break; case 'k': role = KING; break;
            default: return false;
        }
        if (r < 0 || f > 7) return false;
        p.role[role] |= bit(r * 8 + f); p.color[col] |= bit(r * 8 + f); f++;
    }
    if (*c != ' ') return false;
    c++;
    p.turn = (*c == 'b') ? BLACK : WHITE; c++;
    if (*c == ' ') c++;
    for (; *c && *c != ' '; c++) {
        if (*c == 'K') p.castling |= 1; else if (*c == 'Q') p.castling |= 2;
        else if (*c == 'k') p.castling |= 4; else if (*c == 'q') p.castling |= 8;
    }
    if (*c == ' ') c++;
    if (*c && *c != '-') { int ef = c[0] - 'a', er = c[1] - '1'; p.ep = (int8_t)(er * 8 + ef); c += 2; } else if (*c) c++;
    if (*c == ' ') { c++; p.halfmoves = (uint16_t)strtol(c, (char**)&c, 10); }
    if (*c == ' ') { c++; p.fullmoves = (uint16_t)strtol(c, (char**)&c, 10); }
    *out = p; return true;
}

u64 occupied(const Pos& p) { return p.color[0] | p.color[1]; }

// shakmaty Board::attacks_to(sq, attacker, occupied)
u64 attacks_to(const Pos& p, int sq, int attacker, u64 occ) {
    return p.color[attacker] &
           ((rook_attacks(sq, occ) & (p.role[ROOK] | p.role[QUEEN])) |
            (bishop_attacks(sq, occ) & (p.role[BISHOP] | p.role[QUEEN])) |
            (KNIGHT_ATT[sq] & p.role[KNIGHT]) | (KING_ATT[sq] & p.role[KING]) |
            (PAWN_ATT[attacker ^ 1][sq] & p.role[PAWN]));
}

static int king_of(const Pos& p, int c) { return lsb(p.role[KING] & p.color[c]); }

u64 checkers(const Pos& p) { return attacks_to(p, king_of(p, p.turn), p.turn ^ 1, occupied(p)); }

static int role_at(const Pos& p, int sq) {
    u64 b = bit(sq);
    for (int r = 0; r < 6; r++) if (p.role[r] & b) return r;
    return NO_ROLE;
}

static void push_promotions(MoveList& ml, int from, int to, int capture) {
    static const uint8_t order[4] = {QUEEN, ROOK, BISHOP, KNIGHT};
    for (int i = 0; i < 4; i++) ml.push(Move{NORMAL, PAWN, (uint8_t)from, (uint8_t)to, (uint8_t)capture, order[i]});
}

static u64 relative_rank(int turn, int rank) { return 0xFFULL << (8 * (turn == WHITE ? rank : 7 - rank)); }

// shakmaty gen_pawn_moves: captures, promotion captures, single pushes, promotion pushes, double pushes
static void gen_pawn_moves(const Pos& p, u64 target, MoveList& ml) {
    int us = p.turn; u64 ours = p.color[us], theirs = p.color[us ^ 1], occ = ours | theirs;
    u64 pawns = p.role[PAWN] & ours;
    u64 seventh = pawns & relative_rank(us, 6);
    for (u64 b = pawns & ~seventh; b; b &= b - 1) {
        int from = lsb(b);
        for (u64 t = PAWN_ATT[us][from] & theirs & target; t; t &= t - 1) {
            int to = lsb(t);
            ml.push(Move{NORMAL, PAWN, (uint8_t)from, (uint8_t)to, (uint8_t)role_at(p, to), NO_ROLE});
        }
    }
    for (u64 b = seventh; b; b &= b - 1) {
        int from = lsb(b);
        for (u64 t = PAWN_ATT[us][from] & theirs & target; t; t &= t - 1) {
            int to = lsb(t);
            push_promotions(ml, from, to, role_at(p, to));
        }
    }
    u64 single = (us == WHITE ? pawns << 8 : pawns >> 8) & ~occ;
    u64 dbl = (us == WHITE ? single << 8 : single >> 8) & relative_rank(us, 3) & ~occ;
    const u64 backranks = 0xFF000000000000FFULL;
    int back = us == WHITE ? -8 : 8;
    for (u64 t = single & target & ~backranks; t; t &= t - 1) {
        int to = lsb(t);
        ml.push(Move{NORMAL, PAWN, (uint8_t)(to + back), (uint8_t)to, NO_ROLE, NO_ROLE});
    }
    for (u64 t = single & target & backranks; t; t &= t - 1) {
        int to = lsb(t);
        push_promotions(ml, to + back, to, NO_ROLE);
    }
    for (u64 t = dbl & target; t; t &= t - 1) {
        int to = lsb(t);
        ml.push(Move{NORMAL, PAWN, (uint8_t)(to + 2 * back), (uint8_t)to, NO_ROLE, NO_ROLE});
    }
}

static void gen_piece(const Pos& p, int role, u64 target, MoveList& ml) {
    u64 occ = occupied(p);
    for (u64 b = p.role[role] & p.color[p.turn]; b; b &= b - 1) {
        int from = lsb(b); u64 att;
        switch (role) {
            case KNIGHT: att = KNIGHT_ATT[from]; break;
            case BISHOP: att = bishop_attacks(from, occ); break;
            case ROOK: att = rook_attacks(from, occ); break;
            default: att = rook_attacks(from, occ) | bishop_attacks(from, occ); break;
        }
        for (u64 t = att & target; t; t &= t - 1) {
            int to = lsb(t);
            ml.push(Move{NORMAL, (uint8_t)role, (uint8_t)from, (uint8_t)to, (uint8_t)role_at(p, to), NO_ROLE});
        }
    }
}

static void gen_non_king(const Pos& p, u64 target, MoveList& ml) {
    gen_pawn_moves(p, target, ml);
    gen_piece(p, KNIGHT, target, ml);
    gen_piece(p, BISHOP, target, ml);
    gen_piece(p, ROOK, target, ml);
    gen_piece(p, QUEEN, target, ml);
}

static void gen_safe_king(const Pos& p, int king, u64 target, MoveList& ml) {
    for (u64 t = KING_ATT[king] & target; t; t &= t - 1) {
        int to = lsb(t);
        if (!attacks_to(p, to, p.turn ^ 1, occupied(p)))
            ml.push(Move{NORMAL, KING, (uint8_t)king, (uint8_t)to, (uint8_t)role_at(p, to), NO_ROLE});
    }
}

static void gen_castling(const Pos& p, int king, int side /*0 king-side, 1 queen-side*/, MoveList& ml) {
    int us = p.turn;
    int right = (us == WHITE ? 0 : 2) + side;
    if (!(p.castling & (1 << right))) return;
    int base = us == WHITE ? 0 : 56;
    int rook = base + (side == 0 ? 7 : 0);
    int king_to = base + (side == 0 ? 6 : 2);
    int rook_to = base + (side == 0 ? 5 : 3);
    u64 occ = occupied(p);
    // path between king and rook must be empty (standard chess: equals shakmaty Castles::path)
    if (BETWEEN[king][rook] & occ) return;
    u64 king_path = BETWEEN[king][king_to] | bit(king);
    for (u64 t = king_path; t; t &= t - 1)
        if (attacks_to(p, lsb(t), us ^ 1, occ ^ bit(king))) return;
    if (attacks_to(p, king_to, us ^ 1, occ ^ bit(king) ^ bit(rook) ^ bit(rook_to))) return;
    ml.push(Move{CASTLE, KING, (uint8_t)king, (uint8_t)rook, NO_ROLE, NO_ROLE});
}

static bool gen_en_passant(const Pos& p, MoveList& ml) {
    if (p.ep < 0) return false;
    bool found = false;
    for (u64 b = p.role[PAWN] & p.color[p.turn] & PAWN_ATT[p.turn ^ 1][p.ep]; b; b &= b - 1) {
        ml.push(Move{EN_PASSANT, PAWN, (uint8_t)lsb(b), (uint8_t)p.ep, PAWN, NO_ROLE});
        found = true;
    }
    return found;
}

static u64 slider_blockers(const Pos& p, u64 enemy, int king) {
    u64 snipers = (rook_attacks(king, 0) & (p.role[ROOK] | p.role[QUEEN])) |
                  (bishop_attacks(king, 0) & (p.role[BISHOP] | p.role[QUEEN]));
    u64 blockers = 0, occ = occupied(p);
    for (u64 s = snipers & enemy; s; s &= s - 1) {
        u64 b = BETWEEN[king][lsb(s)] & occ;
        if (popcnt(b) <= 1) blockers |= b;
    }
    return blockers;
}

static bool is_safe(const Pos& p, int king, const Move& m, u64 blockers) {
    if (m.kind == NORMAL) {
        return !(blockers & bit(m.from)) || (LINE[m.from][m.to] & bit(king));
    } else if (m.kind == EN_PASSANT) {
        int cap = (m.from & 56) | (m.to & 7);
        u64 occ = (occupied(p) ^ bit(m.from) ^ bit(cap)) | bit(m.to);
        return (attacks_to(p, king, p.turn ^ 1, occ) & ~bit(cap)) == 0;
    }
    return true;
}

static void evasions(const Pos& p, int king, u64 chk, MoveList& ml) {
    u64 sliders = chk & (p.role[BISHOP] | p.role[ROOK] | p.role[QUEEN]);
    u64 attacked = 0;
    for (u64 s = sliders; s; s &= s - 1) {
        int c = lsb(s);
        attacked |= LINE[c][king] ^ bit(c);
    }
    gen_safe_king(p, king, ~p.color[p.turn] & ~attacked, ml);
    if (popcnt(chk) == 1) {
        int c = lsb(chk);
        gen_non_king(p, BETWEEN[king][c] | bit(c), ml);
    }
}

// shakmaty <Chess as Position>::legal_moves
void legal_moves(const Pos& p, MoveList& ml) {
    init_tables();
    ml.n = 0;
    int king = king_of(p, p.turn);
    bool has_ep = gen_en_passant(p, ml);
    u64 chk = checkers(p);
    if (!chk) {
        u64 target = ~p.color[p.turn];
        gen_non_king(p, target, ml);
        gen_safe_king(p, king, target, ml);
        gen_castling(p, king, 0, ml);
        gen_castling(p, king, 1, ml);
    } else {
        evasions(p, king, chk, ml);
    }
    u64 blockers = slider_blockers(p, p.color[p.turn ^ 1], king);
    if (blockers || has_ep) {
        int w = 0;
        for (int i = 0; i < ml.n; i++) if (is_safe(p, king, ml.m[i], blockers)) ml.m[w++] = ml.m[i];
        ml.n = w;
    }
}

bool is_legal(const Pos& p, const Move& m) {
    MoveList ml; legal_moves(p, ml);
    for (int i = 0; i < ml.n; i++)
        if (ml.m[i].kind == m.kind && ml.m[i].from == m.from && ml.m[i].to == m.to && ml.m[i].promotion == m.promotion &&
            ml.m[i].role == m.role)
            return true;
    return false;
}

static void discard_piece(Pos& p, int sq) {
    u64 m = ~bit(sq);
    for (int r = 0; r < 6; r++) p.role[r] &= m;
    p.color[0] &= m; p.color[1] &= m;
}
static void set_piece(Pos& p, int sq, int role, int color) {
    discard_piece(p, sq);
    p.role[role] |= bit(sq); p.color[color] |= bit(sq);
}
static void discard_rook_right(Pos& p, int sq) {
    if (sq == 7) p.castling &= ~1; else if (sq == 0) p.castling &= ~2;
    else if (sq == 63) p.castling &= ~4; else if (sq == 56) p.castling &= ~8;
}

// shakmaty do_move
void play_unchecked(Pos& p, const Move& m) {
    int color = p.turn;
    p.ep = -1;
    bool zeroing = (m.role == PAWN && m.kind != CASTLE) || (m.kind == NORMAL && m.capture != NO_ROLE) || m.kind == EN_PASSANT;
    p.halfmoves = zeroing ? 0 : (uint16_t)(p.halfmoves == 0xFFFF ? 0xFFFF : p.halfmoves + 1);
    if (m.kind == NORMAL) {
        if (m.role == PAWN && (m.to - m.from == 16 || m.from - m.to == 16)) p.ep = (int8_t)(m.from + (color == WHITE ? 8 : -8));
        if (m.role == KING) p.castling &= (color == WHITE ? ~3 : ~12);
        else if (m.role == ROOK) discard_rook_right(p, m.from);
        if (m.capture == ROOK) discard_rook_right(p, m.to);
        discard_piece(p, m.from);
        set_piece(p, m.to, m.promotion != NO_ROLE ? m.promotion : m.role, color);
    } else if (m.kind == CASTLE) {
        int king = m.from, rook = m.to; bool qs = rook < king;
        int base = color == WHITE ? 0 : 56;
        discard_piece(p, king); discard_piece(p, rook);
        set_piece(p, base + (qs ? 3 : 5), ROOK, color);
        set_piece(p, base + (qs ? 2 : 6), KING, color);
        p.castling &= (color == WHITE ? ~3 : ~12);
    } else {
        discard_piece(p, (m.from & 56) | (m.to & 7));
        discard_piece(p, m.from);
        set_piece(p, m.to, PAWN, color);
    }
    if (color == BLACK) p.fullmoves = (uint16_t)(p.fullmoves == 0xFFFF ? 0xFFFF : p.fullmoves + 1);
    p.turn = (uint8_t)(color ^ 1);
}

int legal_ep_square(const Pos& p) {
    if (p.ep < 0) return -1;
    MoveList ml; legal_moves(p, ml);
    for (int i = 0; i < ml.n; i++) if (ml.m[i].kind == EN_PASSANT) return p.ep;
    return -1;
}

int pseudo_legal_ep_square(const Pos& p) {
    if (p.ep < 0) return -1;
    return (PAWN_ATT[p.turn ^ 1][p.ep] & p.role[PAWN] & p.color[p.turn]) ? p.ep : -1;
}

static bool has_insufficient_material(const Pos& p, int c) {
    const u64 DARK = 0xAA55AA55AA55AA55ULL, LIGHT = 0x55AA55AA55AA55AAULL;
    if (p.color[c] & (p.role[PAWN] | p.role[ROOK] | p.role[QUEEN])) return false;
    if (p.color[c] & p.role[KNIGHT])
        return popcnt(p.color[c]) <= 2 && (p.color[c ^ 1] & ~p.role[KING] & ~p.role[QUEEN]) == 0;
    if (p.color[c] & p.role[BISHOP]) {
        bool same = (p.role[BISHOP] & DARK) == 0 || (p.role[BISHOP] & LIGHT) == 0;
        return same && p.role[KNIGHT] == 0 && p.role[PAWN] == 0;
    }
    return true;
}
bool is_insufficient_material(const Pos& p) { return has_insufficient_material(p, WHITE) && has_insufficient_material(p, BLACK); }

int outcome(const Pos& p) {
    MoveList ml; legal_moves(p, ml);
    if (ml.n == 0) {
        if (checkers(p)) return p.turn == WHITE ? BLACK_WINS : WHITE_WINS;
        return DRAW;
    }
    if (is_insufficient_material(p)) return DRAW;
    return 0;
}

// chess.rs:247-264 evaluate_material: pawn 100, knight 320, bishop 330, rook 500, queen 900, from the mover's side
int evaluate_material(const Pos& p) {
    static const int value[5] = {100, 320, 330, 500, 900};
    const int us = p.turn, them = us ^ 1;
    int score = 0;
    for (int r = 0; r < 5; r++) {
        score += __builtin_popcountll(p.role[r] & p.color[us]) * value[r];
        score -= __builtin_popcountll(p.role[r] & p.color[them]) * value[r];
    }
    return score;
}

// chess.rs:266-292 negamax: full width, no pruning; a decided game scores -(20000 + remaining depth) for the side to move
// that is mated, a drawn one (stalemate, insufficient material) 0; the horizon scores material.
int negamax(const Pos& p, int depth) {
    MoveList ml; legal_moves(p, ml);
    const bool over = ml.n == 0 || is_insufficient_material(p);
    if (depth == 0 || over) {
        if (ml.n == 0) return checkers(p) ? -20000 - depth : 0;
        if (over) return 0;
        return evaluate_material(p);
    }
    int best = INT32_MIN;
    for (int i = 0; i < ml.n; i++) {
        Pos c = p; play_unchecked(c, ml.m[i]);
        const int sc = -negamax(c, depth - 1);
        if (sc > best) best = sc;
    }
    return best;
}

// chess.rs:295-318 get_best_move, up to the random choice among the best: the score of every legal move, in move order
int minimax_scores(const Pos& p, int depth, int32_t* scores) {
    MoveList ml; legal_moves(p, ml);
    for (int i = 0; i < ml.n; i++) {
        Pos c = p; play_unchecked(c, ml.m[i]);
        scores[i] = -negamax(c, depth - 1);
    }
    return ml.n;
}

u64 perft(const Pos& p, int depth) {
    if (depth == 0) return 1;
    MoveList ml; legal_moves(p, ml);
    if (depth == 1) return (u64)ml.n;
    u64 n = 0;
    for (int i = 0; i < ml.n; i++) { Pos c = p; play_unchecked(c, ml.m[i]); n += perft(c, depth - 1); }
    return n;
}

// ---------------------------------------------------------------------------
// chess.rs glue
// ---------------------------------------------------------------------------
bool PosKey::operator<(const PosKey& o) const {
    int c = std::memcmp(role, o.role, sizeof role); if (c) return c < 0;
    c = std::memcmp(color, o.color, sizeof color); if (c) return c < 0;
    return std::tie(turn, castling, legal_ep) < std::tie(o.turn, o.castling, o.legal_ep);
}
PosKey make_key(const Pos& p) {
    PosKey k; std::memcpy(k.role, p.role, sizeof k.role); std::memcpy(k.color, p.color, sizeof k.color);
    k.turn = p.turn; k.castling = p.castling; k.legal_ep = (int8_t)legal_ep_square(p);
    return k;
}

GameState::GameState() : GameState(startpos()) {}
GameState::GameState(const Pos& p) : position(p) { pos_count[make_key(p)] = 1; }  // chess.rs:20-26

// chess.rs:36-63
int play_move(GameState& st, const Move& m) {
    if (!is_legal(st.position, m)) return ILLEGAL;
    play_unchecked(st.position, m);
    int oc = outcome(st.position);
    if (oc != 0) return oc;
    int& count = st.pos_count[make_key(st.position)];
    count += 1;
    if (count < REPETITIONS && st.position.halfmoves < NUM_HALFMOVES && st.position.fullmoves < NUM_FULLMOVES) return ONGOING;
    return DRAW;
}

// chess.rs:73-116
int move_to_index(const Move& m, int turn) {
    int file = m.from & 7, rank = turn == BLACK ? 7 - (m.from >> 3) : (m.from >> 3);
    int dfile = m.to & 7, drank = turn == BLACK ? 7 - (m.to >> 3) : (m.to >> 3);
    int df = dfile - file, dr = drank - rank, plane;
    if (df == 1 && dr == 2) plane = 0; else if (df == 2 && dr == 1) plane = 1;
    else if (df == 2 && dr == -1) plane = 2; else if (df == 1 && dr == -2) plane = 3;
    else if (df == -1 && dr == -2) plane = 4; else if (df == -2 && dr == -1) plane = 5;
    else if (df == -2 && dr == 1) plane = 6; else if (df == -1 && dr == 2) plane = 7;
    else if (df == 0 && dr > 0) plane = 7 + dr;
    else if (df > 0 && dr > 0) plane = 14 + dr;
    else if (df > 0 && dr == 0) plane = 21 + df;
    else if (df > 0 && dr < 0) plane = 28 + df;
    else if (df == 0 && dr < 0) plane = 35 - dr;
    else if (df < 0 && dr < 0) plane = 42 - dr;
    else if (df < 0 && dr == 0) plane = 49 - df;
    else plane = 56 - df;
    return plane * 64 + rank * 8 + file;
}

// chess.rs:118-171 (the UCI string round trip is restated as shakmaty's UciMove::to_move rules)
bool index_to_move(int index, const Pos& p, Move* out) {
    int plane = index / 64, sqi = index % 64, from_file = sqi % 8, crank = sqi / 8;
    int from_rank = p.turn == BLACK ? 7 - crank : crank;
    int df, dr;
    static const int kdf[8] = {1, 2, 2, 1, -1, -2, -2, -1}, kdr[8] = {2, 1, -1, -2, -2, -1, 1, 2};
    if (plane < 8) { df = kdf[plane]; dr = kdr[plane]; }
    else if (plane < 15) { df = 0; dr = plane - 7; }
    else if (plane < 22) { df = plane - 14; dr = plane - 14; }
    else if (plane < 29) { df = plane - 21; dr = 0; }
    else if (plane < 36) { df = plane - 28; dr = 28 - plane; }
    else if (plane < 43) { df = 0; dr = 35 - plane; }
    else if (plane < 50) { df = 42 - plane; dr = 42 - plane; }
    else if (plane < 57) { df = 49 - plane; dr = 0; }
    else { df = 56 - plane; dr = plane - 56; }
    if (p.turn == BLACK) dr = -dr;
    int dest_file = from_file + df, dest_rank = from_rank + dr;
    if (dest_file < 0 || dest_file > 7 || dest_rank < 0 || dest_rank > 7) return false;
    int from = from_rank * 8 + from_file, to = dest_rank * 8 + dest_file;
    int role = role_at(p, from);
    if (role == NO_ROLE) return false;  // `role_at(from_square)?`
    int promotion = (role == PAWN && (dest_rank == 0 || dest_rank == 7)) ? QUEEN : NO_ROLE;
    // UciMove::to_move
    Move m;
    // Deviation (documented in DESIGN.md): shakmaty tests `castling_rights().contains(to)` over BOTH colours'
    // rooks [recalled]; a king capturing an enemy rook that still has its right would then fail to parse and
    // the reference would panic (tree.rs:211).  We only treat our own castling rooks as castling targets.
    u64 rights = 0;
    if (p.turn == WHITE) { if (p.castling & 1) rights |= bit(7); if (p.castling & 2) rights |= bit(0); }
    else { if (p.castling & 4) rights |= bit(63); if (p.castling & 8) rights |= bit(56); }
    if (role == PAWN && to == legal_ep_square(p)) {
        m = Move{EN_PASSANT, PAWN, (uint8_t)from, (uint8_t)to, PAWN, NO_ROLE};
    } else if (role == KING && (rights & bit(to))) {
        m = Move{CASTLE, KING, (uint8_t)from, (uint8_t)to, NO_ROLE, NO_ROLE};
    } else if (role == KING && from == (p.turn == WHITE ? 4 : 60) && (to >> 3) == (p.turn == WHITE ? 0 : 7) &&
               std::abs((to & 7) - (from & 7)) == 2) {
        int rook = (from & 7) < (to & 7) ? (p.turn == WHITE ? 7 : 63) : (p.turn == WHITE ? 0 : 56);
        m = Move{CASTLE, KING, (uint8_t)from, (uint8_t)rook, NO_ROLE, NO_ROLE};
    } else {
        m = Move{NORMAL, (uint8_t)role, (uint8_t)from, (uint8_t)to, (uint8_t)role_at(p, to), (uint8_t)promotion};
    }
    if (!is_legal(p, m)) return false;
    *out = m; return true;
}

// chess.rs:191-245
void to_tensor(const Pos& p, float* out) {
    std::memset(out, 0, sizeof(float) * 19 * 64);
    int us = p.turn, them = us ^ 1;
    for (int sq = 0; sq < 64; sq++) {
        int role = role_at(p, sq);
        if (role == NO_ROLE) continue;
        int offset = (p.color[us] & bit(sq)) ? 0 : 6;
        int file = sq & 7, rank = us == BLACK ? 7 - (sq >> 3) : (sq >> 3);
        out[(role + offset) * 64 + rank * 8 + file] = 1.0f;
    }
    auto has = [&](int color, int side) { return (p.castling >> ((color == WHITE ? 0 : 2) + side)) & 1; };
    auto fill = [&](int plane, float v) { for (int i = 0; i < 64; i++) out[plane * 64 + i] = v; };
    if (has(us, 0)) fill(12, 1.0f);
    if (has(us, 1)) fill(13, 1.0f);
    if (has(them, 0)) fill(14, 1.0f);
    if (has(them, 1)) fill(15, 1.0f);
    int ep = pseudo_legal_ep_square(p);
    if (ep >= 0) {
        int file = ep & 7, rank = us == BLACK ? 7 - (ep >> 3) : (ep >> 3);
        out[16 * 64 + rank * 8 + file] = 1.0f;
    }
    fill(17, (float)p.halfmoves / (float)NUM_HALFMOVES);
    fill(18, (float)p.fullmoves / (float)NUM_FULLMOVES);
}

}  // namespace orc
