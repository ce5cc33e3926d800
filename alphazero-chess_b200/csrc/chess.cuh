// Device-side chess rules for the self-play hot path: bitboard legal move generation, make-move, game-end rules,
// the 4096-way move<->index codec and the 19x8x8 input planes.
//
// Behaviour mirrors what the reference gets from shakmaty 0.29.0 (Cargo.lock:4573) at its call sites
// (tree.rs:39,86; chess.rs:38-55,73-171,191-245).  The move ORDER of legal moves follows shakmaty's generator
// (en passant, pawn captures, promotion captures, pushes, promotion pushes, double pushes, N, B, R, Q, K, O-O, O-O-O;
// in check: king first) because tree.rs:191 breaks PUCT ties by list order and tree.rs:286 assigns Dirichlet
// components by list position.  Everything is table-free: sliding attacks use the hyperbola-quintessence identity
// with BREV, leapers use shifts, so a thread needs registers only.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/az_b200.h"

namespace azb {

typedef unsigned long long u64;

// 64-byte position: six role boards, white occupancy (black = occupied ^ white) and one packed word.
struct __align__(16) DPos {
    u64 pawn, knight, bishop, rook, queen, king, white, meta;
};

// meta: bit0 turn | bits1-4 castling (WK,WQ,BK,BQ) | bits8-14 ep+1 (0 = none; raw: kept after EVERY double push)
//       bits16-31 halfmoves | bits32-47 fullmoves | bits48-54 LEGAL ep+1 (valid when bit55 set; used for repetition keys)
constexpr u64 META_KEY_MASK = 0x1FULL | (0xFFULL << 48);  // turn, castling, legal ep (+valid bit)

__device__ __forceinline__ int meta_turn(u64 m) { return (int)(m & 1); }
__device__ __forceinline__ int meta_castling(u64 m) { return (int)((m >> 1) & 15); }
__device__ __forceinline__ int meta_ep(u64 m) { return (int)((m >> 8) & 127) - 1; }
__device__ __forceinline__ int meta_halfmoves(u64 m) { return (int)((m >> 16) & 0xFFFF); }
__device__ __forceinline__ int meta_fullmoves(u64 m) { return (int)((m >> 32) & 0xFFFF); }

__host__ __device__ inline DPos dpos_from_wire(const az_position& w) {
    DPos p;
    p.pawn = w.roles[0]; p.knight = w.roles[1]; p.bishop = w.roles[2]; p.rook = w.roles[3]; p.queen = w.roles[4]; p.king = w.roles[5];
    p.white = w.colors[0];
    p.meta = (u64)(w.turn & 1) | ((u64)(w.castling & 15) << 1) | ((u64)((w.ep_square + 1) & 127) << 8) |
             ((u64)w.halfmoves << 16) | ((u64)w.fullmoves << 32);
    return p;
}
__host__ __device__ inline az_position dpos_to_wire(const DPos& p) {
    az_position w;
    w.roles[0] = p.pawn; w.roles[1] = p.knight; w.roles[2] = p.bishop; w.roles[3] = p.rook; w.roles[4] = p.queen; w.roles[5] = p.king;
    u64 occ = p.pawn | p.knight | p.bishop | p.rook | p.queen | p.king;
    w.colors[0] = p.white; w.colors[1] = occ ^ p.white;
    w.turn = (uint8_t)(p.meta & 1); w.castling = (uint8_t)((p.meta >> 1) & 15);
    w.ep_square = (int8_t)((int)((p.meta >> 8) & 127) - 1); w.reserved = 0;
    w.halfmoves = (uint16_t)((p.meta >> 16) & 0xFFFF); w.fullmoves = (uint16_t)((p.meta >> 32) & 0xFFFF);
    return w;
}

#ifdef __CUDACC__

constexpr u64 FILE_A = 0x0101010101010101ULL, FILE_H = 0x8080808080808080ULL;
constexpr u64 NOT_A = ~FILE_A, NOT_H = ~FILE_H;
constexpr u64 NOT_AB = 0xFCFCFCFCFCFCFCFCULL, NOT_GH = 0x3F3F3F3F3F3F3F3FULL;
constexpr u64 RANK_1 = 0xFFULL, BACKRANKS = 0xFF000000000000FFULL;
constexpr u64 DIAG_A1H8 = 0x8040201008040201ULL, ANTI_H1A8 = 0x0102040810204080ULL;

__device__ __forceinline__ u64 bit(int s) { return 1ULL << s; }
__device__ __forceinline__ int lsb(u64 b) { return __ffsll((long long)b) - 1; }
__device__ __forceinline__ int popc(u64 b) { return __popcll(b); }

__device__ __forceinline__ u64 rank_mask(int sq) { return RANK_1 << (sq & 56); }
__device__ __forceinline__ u64 file_mask(int sq) { return FILE_A << (sq & 7); }
__device__ __forceinline__ u64 diag_mask(int sq) {
    int d = (sq & 7) - (sq >> 3);
    return d >= 0 ? DIAG_A1H8 >> (8 * d) : DIAG_A1H8 << (-8 * d);
}
__device__ __forceinline__ u64 anti_mask(int sq) {
    int d = 7 - (sq & 7) - (sq >> 3);
    return d >= 0 ? ANTI_H1A8 >> (8 * d) : ANTI_H1A8 << (-8 * d);
}
// attacks of a slider on `sq` along one line (mask includes sq): o^(o-2s) forwards, bit-reversed for the other ray
__device__ __forceinline__ u64 line_attacks(u64 occ, int sq, u64 mask) {
    u64 o = occ & mask, s = bit(sq);
    u64 fwd = o - 2 * s;
    u64 rev = __brevll(__brevll(o) - 2 * __brevll(s));
    return (fwd ^ rev) & mask;
}
__device__ __forceinline__ u64 rook_attacks(int sq, u64 occ) {
    return line_attacks(occ, sq, rank_mask(sq)) | line_attacks(occ, sq, file_mask(sq));
}
__device__ __forceinline__ u64 bishop_attacks(int sq, u64 occ) {
    return line_attacks(occ, sq, diag_mask(sq)) | line_attacks(occ, sq, anti_mask(sq));
}
__device__ __forceinline__ u64 knight_attacks_bb(u64 b) {
    u64 h1 = ((b >> 1) & ~FILE_H) | ((b << 1) & ~FILE_A);
    u64 h2 = ((b >> 2) & NOT_GH) | ((b << 2) & NOT_AB);
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
__device__ __forceinline__ u64 king_attacks_bb(u64 b) {
    u64 a = ((b << 1) & NOT_A) | ((b >> 1) & NOT_H);
    u64 c = b | a;
    return a | (c << 8) | (c >> 8);
}
// squares attacked by pawns of `color` standing on bitboard b
__device__ __forceinline__ u64 pawn_attacks_bb(int color, u64 b) {
    return color == 0 ? (((b << 7) & NOT_H) | ((b << 9) & NOT_A)) : (((b >> 9) & NOT_H) | ((b >> 7) & NOT_A));
}
// the full line through a and b (0 when not aligned)
__device__ __forceinline__ u64 line_through(int a, int b) {
    u64 bb = bit(b), m;
    m = rank_mask(a); if (m & bb) return m;
    m = file_mask(a); if (m & bb) return m;
    m = diag_mask(a); if (m & bb) return m;
    m = anti_mask(a); if (m & bb) return m;
    return 0;
}
__device__ __forceinline__ u64 between_bb(int a, int b) {
    u64 l = line_through(a, b);
    int lo = min(a, b), hi = max(a, b);
    return l & (~0ULL << lo) & ~bit(lo) & (bit(hi) - 1);
}

__device__ __forceinline__ u64 occupied(const DPos& p) { return p.pawn | p.knight | p.bishop | p.rook | p.queen | p.king; }

// pieces in `attackers` that attack sq given occupancy occ
__device__ __forceinline__ u64 attackers_to(const DPos& p, int sq, u64 attackers, int attacker_color, u64 occ) {
    u64 s = bit(sq);
    return attackers & ((rook_attacks(sq, occ) & (p.rook | p.queen)) | (bishop_attacks(sq, occ) & (p.bishop | p.queen)) |
                        (knight_attacks_bb(s) & p.knight) | (king_attacks_bb(s) & p.king) |
                        (pawn_attacks_bb(attacker_color ^ 1, s) & p.pawn));
}

// wire move: from | to<<6 | promo<<12 (0 none, 1 N, 2 B, 3 R, 4 Q) | special<<15 (castle = king takes rook, or en passant)
__device__ __forceinline__ uint16_t mk_move(int from, int to, int promo, int special) {
    return (uint16_t)(from | (to << 6) | (promo << 12) | (special << 15));
}

struct ListSink {
    uint16_t* mv;
    int n;
    __device__ __forceinline__ void add(int from, int to, int promo, int special) { mv[n++] = mk_move(from, to, promo, special); }
    __device__ __forceinline__ void targets(int from, u64 t) { for (; t; t &= t - 1) add(from, lsb(t), 0, 0); }
    __device__ __forceinline__ void promo_targets(int from, u64 t) {
        for (; t; t &= t - 1) { int to = lsb(t); add(from, to, 4, 0); add(from, to, 3, 0); add(from, to, 2, 0); add(from, to, 1, 0); }
    }
    __device__ __forceinline__ void pushes(u64 t, int delta) { for (; t; t &= t - 1) { int to = lsb(t); add(to - delta, to, 0, 0); } }
    __device__ __forceinline__ void promo_pushes(u64 t, int delta) {
        for (; t; t &= t - 1) { int to = lsb(t), f = to - delta; add(f, to, 4, 0); add(f, to, 3, 0); add(f, to, 2, 0); add(f, to, 1, 0); }
    }
};
struct CountSink {
    int n;
    __device__ __forceinline__ void add(int, int, int, int) { n++; }
    __device__ __forceinline__ void targets(int, u64 t) { n += popc(t); }
    __device__ __forceinline__ void promo_targets(int, u64 t) { n += 4 * popc(t); }
    __device__ __forceinline__ void pushes(u64 t, int) { n += popc(t); }
    __device__ __forceinline__ void promo_pushes(u64 t, int) { n += 4 * popc(t); }
};

struct GenInfo {
    u64 checkers;
    bool has_legal_ep;
};

// shakmaty gen_non_king with the pin filter folded in (same order as generate-then-retain)
template <class Sink>
__device__ __forceinline__ void gen_non_king(const DPos& p, Sink& s, int us, u64 ours, u64 theirs, u64 occ, int ksq, u64 blockers,
                                             u64 target) {
    const u64 pawns = p.pawn & ours;
    const u64 seventh = pawns & (us == 0 ? 0x00FF000000000000ULL : 0x000000000000FF00ULL);
    for (u64 b = pawns & ~seventh; b; b &= b - 1) {
        int from = lsb(b);
        u64 t = pawn_attacks_bb(us, bit(from)) & theirs & target;
        if (blockers & bit(from)) t &= line_through(ksq, from);
        s.targets(from, t);
    }
    for (u64 b = seventh; b; b &= b - 1) {
        int from = lsb(b);
        u64 t = pawn_attacks_bb(us, bit(from)) & theirs & target;
        if (blockers & bit(from)) t &= line_through(ksq, from);
        s.promo_targets(from, t);
    }
    // a pinned pawn may only push along the king's file
    const u64 pushers = pawns & ~(blockers & ~file_mask(ksq));
    const u64 single = (us == 0 ? pushers << 8 : pushers >> 8) & ~occ;
    const u64 dbl = (us == 0 ? single << 8 : single >> 8) & (us == 0 ? 0x00000000FF000000ULL : 0x000000FF00000000ULL) & ~occ;
    const int delta = us == 0 ? 8 : -8;
    s.pushes(single & target & ~BACKRANKS, delta);
    s.promo_pushes(single & target & BACKRANKS, delta);
    s.pushes(dbl & target, 2 * delta);
    for (u64 b = p.knight & ours & ~blockers; b; b &= b - 1) {  // a pinned knight never moves
        int from = lsb(b);
        s.targets(from, knight_attacks_bb(bit(from)) & target);
    }
    for (u64 b = p.bishop & ours; b; b &= b - 1) {
        int from = lsb(b);
        u64 t = bishop_attacks(from, occ) & target;
        if (blockers & bit(from)) t &= line_through(ksq, from);
        s.targets(from, t);
    }
    for (u64 b = p.rook & ours; b; b &= b - 1) {
        int from = lsb(b);
        u64 t = rook_attacks(from, occ) & target;
        if (blockers & bit(from)) t &= line_through(ksq, from);
        s.targets(from, t);
    }
    for (u64 b = p.queen & ours; b; b &= b - 1) {
        int from = lsb(b);
        u64 t = (rook_attacks(from, occ) | bishop_attacks(from, occ)) & target;
        if (blockers & bit(from)) t &= line_through(ksq, from);
        s.targets(from, t);
    }
}

// <Chess as Position>::legal_moves
template <class Sink>
__device__ __forceinline__ GenInfo gen_legal(const DPos& p, Sink& s) {
    const int us = meta_turn(p.meta);
    const u64 occ = occupied(p);
    const u64 ours = us == 0 ? p.white : occ ^ p.white;
    const u64 theirs = occ ^ ours;
    const int ksq = lsb(p.king & ours);
    const u64 kbit = bit(ksq);
    GenInfo gi;
    gi.has_legal_ep = false;

    // en passant first (gen_en_passant + is_safe)
    const int ep = meta_ep(p.meta);
    if (ep >= 0) {
        for (u64 b = p.pawn & ours & pawn_attacks_bb(us ^ 1, bit(ep)); b; b &= b - 1) {
            int from = lsb(b);
            int cap = (from & 56) | (ep & 7);
            u64 o2 = (occ ^ bit(from) ^ bit(cap)) | bit(ep);
            if ((attackers_to(p, ksq, theirs, us ^ 1, o2) & ~bit(cap)) == 0) { s.add(from, ep, 0, 1); gi.has_legal_ep = true; }
        }
    }

    const u64 checkers = attackers_to(p, ksq, theirs, us ^ 1, occ);
    gi.checkers = checkers;

    // slider_blockers: pieces standing alone between an enemy slider and our king
    u64 blockers = 0;
    {
        u64 snipers = theirs & ((rook_attacks(ksq, 0) & (p.rook | p.queen)) | (bishop_attacks(ksq, 0) & (p.bishop | p.queen)));
        for (; snipers; snipers &= snipers - 1) {
            u64 b = between_bb(ksq, lsb(snipers)) & occ;
            if ((b & (b - 1)) == 0) blockers |= b;
        }
    }

    const u64 occ_nok = occ ^ kbit;
    if (!checkers) {
        gen_non_king(p, s, us, ours, theirs, occ, ksq, blockers, ~ours);
        for (u64 t = king_attacks_bb(kbit) & ~ours; t; t &= t - 1) {
            int to = lsb(t);
            if (!attackers_to(p, to, theirs, us ^ 1, occ_nok)) s.add(ksq, to, 0, 0);
        }
        // gen_castling_moves (standard chess): king-side then queen-side, encoded king-takes-rook
        const int rights = meta_castling(p.meta) >> (us * 2);
        const int base = us * 56;
        if ((rights & 1) && ksq == base + 4 && !(occ & (0x60ULL << base))) {
            if (!attackers_to(p, base + 5, theirs, us ^ 1, occ_nok) &&
                !attackers_to(p, base + 6, theirs, us ^ 1, occ_nok ^ bit(base + 7) ^ bit(base + 5)))
                s.add(ksq, base + 7, 0, 1);
        }
        if ((rights & 2) && ksq == base + 4 && !(occ & (0x0EULL << base))) {
            if (!attackers_to(p, base + 3, theirs, us ^ 1, occ_nok) && !attackers_to(p, base + 2, theirs, us ^ 1, occ_nok) &&
                !attackers_to(p, base + 2, theirs, us ^ 1, occ_nok ^ bit(base) ^ bit(base + 3)))
                s.add(ksq, base, 0, 1);
        }
    } else {
        for (u64 t = king_attacks_bb(kbit) & ~ours; t; t &= t - 1) {
            int to = lsb(t);
            if (!attackers_to(p, to, theirs, us ^ 1, occ_nok)) s.add(ksq, to, 0, 0);
        }
        if ((checkers & (checkers - 1)) == 0) {
            int c = lsb(checkers);
            gen_non_king(p, s, us, ours, theirs, occ, ksq, blockers, between_bb(ksq, c) | checkers);
        }
    }
    return gi;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-cooperative legal move generation (same list, same order as gen_legal<ListSink>).
// Every lane owns one "query square" and computes the same five attack sets from it (rook-, bishop-, knight-, king- and
// pawn-pattern) without divergence; what the sets mean depends on the lane's job:
//   lanes  0..15  our i-th non-king piece (ascending square): its move targets
//   lanes 16..23  the king's eight neighbour squares: is the square attacked (king removed from the occupancy)?
//   lane  24      the king square: checkers              lane 25  snipers -> slider blockers (pins)
//   lanes 26,27   castling transit / destination squares (king side), lanes 28,29,30 (queen side)
//   lane  31      en passant: the king square with the capture played
// Items are then ordered exactly like shakmaty's generator (category, from) by a prefix sum over lanes.
__device__ __forceinline__ int nth_set_bit(u64 bb, int n) {  // position of the n-th (0-based) set bit, bb has more than n bits
    const unsigned lo = (unsigned)bb, hi = (unsigned)(bb >> 32);
    const int cl = __popc(lo);
    return n < cl ? (int)__fns(lo, 0, n + 1) : 32 + (int)__fns(hi, 0, n - cl + 1);
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int mask) {
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, mask), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), mask);
    return ((u64)hi << 32) | lo;
}

__device__ __forceinline__ GenInfo warp_gen_legal(const DPos& p, uint16_t* __restrict__ moves, int lane, int& n_out) {
    const int us = meta_turn(p.meta);
    const u64 occ = occupied(p);
    const u64 ours = us == 0 ? p.white : occ ^ p.white;
    const u64 theirs = occ ^ ours;
    const int ksq = lsb(p.king & ours);
    const u64 kbit = bit(ksq);
    const u64 occ_nok = occ ^ kbit;
    const int base = us * 56;
    const int ep = meta_ep(p.meta);
    const u64 pieces = ours ^ kbit;  // our non-king pieces
    const int n_pieces = popc(pieces);

    // ---- phase A: query square + occupancy of every lane
    int sq = ksq;
    u64 o = occ;
    bool job = false;
    int ep_from = -1, ep_from2 = -1;
    if (lane < 16) {
        job = lane < n_pieces;
        if (job) sq = nth_set_bit(pieces, lane);
    } else if (lane < 24) {
        const int d = lane - 16;
        const int df = (d == 0 || d == 3 || d == 5) ? -1 : (d == 1 || d == 6) ? 0 : 1;   // a1-relative 3x3 ring
        const int dr = d < 3 ? -1 : d < 5 ? 0 : 1;
        const int f = (ksq & 7) + df, r = (ksq >> 3) + dr;
        job = f >= 0 && f < 8 && r >= 0 && r < 8 && !(ours & bit((r & 7) * 8 + (f & 7)));
        if (job) sq = r * 8 + f;
        o = occ_nok;
    } else if (lane == 24) {
        job = true;
    } else if (lane == 25) {
        job = true; o = 0;
    } else if (lane <= 30) {
        // castling squares: 26: f-file, 27: g-file (rook hopped), 28: d-file, 29: c-file, 30: c-file (rook hopped)
        const int file = lane == 26 ? 5 : lane == 27 ? 6 : lane == 28 ? 3 : 2;
        sq = base + file;
        o = occ_nok;
        if (lane == 27) o ^= bit(base + 7) ^ bit(base + 5);
        if (lane == 30) o ^= bit(base) ^ bit(base + 3);
        job = true;
    } else {
        // en passant: up to two capturing pawns; the second is checked by re-using this lane after the first
        if (ep >= 0) {
            u64 c = p.pawn & ours & pawn_attacks_bb(us ^ 1, bit(ep));
            if (c) { ep_from = lsb(c); c &= c - 1; if (c) ep_from2 = lsb(c); }
        }
        job = ep_from >= 0;
        if (job) o = (occ ^ bit(ep_from) ^ bit((ep_from & 56) | (ep & 7))) | bit(ep);
    }
    // ---- the five attack patterns from (sq, o): identical instruction stream on every lane
    const u64 sb = bit(sq);
    u64 R = rook_attacks(sq, o);
    u64 B = bishop_attacks(sq, o);
    const u64 N = knight_attacks_bb(sb);
    const u64 K = king_attacks_bb(sb);
    const u64 P = pawn_attacks_bb(us, sb);
    const u64 attackers = theirs & ((R & (p.rook | p.queen)) | (B & (p.bishop | p.queen)) | (N & p.knight) | (K & p.king) | (P & p.pawn));

    // second en-passant candidate (rare): same test with the other pawn
    bool ep_ok1 = false, ep_ok2 = false;
    if (lane == 31 && job) {
        const int cap = (ep_from & 56) | (ep & 7);
        ep_ok1 = (attackers & ~bit(cap)) == 0;
        if (ep_from2 >= 0) {
            const u64 o2 = (occ ^ bit(ep_from2) ^ bit(cap)) | bit(ep);
            ep_ok2 = (attackers_to(p, ksq, theirs, us ^ 1, o2) & ~bit(cap)) == 0;
        }
    }
    // ---- phase B: checkers and blockers for everyone
    const u64 checkers = shfl_u64(attackers, 24);
    u64 blockers = 0;
    if (lane == 25) {
        u64 snipers = theirs & ((R & (p.rook | p.queen)) | (B & (p.bishop | p.queen)));
        for (; snipers; snipers &= snipers - 1) {
            const u64 b = between_bb(ksq, lsb(snipers)) & occ;
            if ((b & (b - 1)) == 0) blockers |= b;
        }
    }
    blockers = shfl_u64(blockers, 25);
    const bool in_check = checkers != 0;
    const bool single = in_check && (checkers & (checkers - 1)) == 0;
    const u64 target = !in_check ? ~ours : single ? (between_bb(ksq, lsb(checkers)) | checkers) : 0ULL;
    // king targets: neighbour lanes whose square is not attacked
    const unsigned safe_mask = __ballot_sync(0xffffffffu, lane >= 16 && lane < 24 && job && attackers == 0);
    u64 king_targets = 0;
    {
        u64 mine = (lane >= 16 && lane < 24 && job && attackers == 0) ? sb : 0ULL;
        for (int d = 4; d; d >>= 1) mine |= shfl_u64(mine, (lane ^ d));   // OR over the 8 lanes 16..23 (xor stays inside)
        king_targets = shfl_u64(mine, 16);
        (void)safe_mask;
    }
    // castling: gathered from lanes 26..30
    const unsigned att_mask = __ballot_sync(0xffffffffu, attackers != 0);
    const int rights = meta_castling(p.meta) >> (us * 2);
    const bool castle_k = (rights & 1) && ksq == base + 4 && !in_check && !(occ & (0x60ULL << base)) && !(att_mask & (3u << 26));
    const bool castle_q = (rights & 2) && ksq == base + 4 && !in_check && !(occ & (0x0EULL << base)) && !(att_mask & (7u << 28));
    const bool e1 = __shfl_sync(0xffffffffu, (int)ep_ok1, 31) != 0, e2 = __shfl_sync(0xffffffffu, (int)ep_ok2, 31) != 0;
    const int epf1 = __shfl_sync(0xffffffffu, ep_from, 31), epf2 = __shfl_sync(0xffffffffu, ep_from2, 31);

    // ---- phase C: one item per lane = (category, from, targets, kind)
    // categories: 0 ep, 1 king-in-check, 2 pawn captures, 3 promotion captures, 4 pushes, 5 promotion pushes, 6 double pushes,
    //             7 N, 8 B, 9 R, 10 Q, 11 king (not in check), 12 O-O, 13 O-O-O
    int cat = 15, from = 0, count = 0, kind = 0;   // kind: 0 targets, 1 promo targets, 2 pushes, 3 promo pushes, 4 single special
    u64 tg = 0;
    int delta = 0, special_to = 0;
    if (lane < 16) {
        if (job) {
            from = sq;
            const u64 pin = (blockers & sb) ? line_through(ksq, sq) : ~0ULL;
            if (p.pawn & sb) {
                tg = P & theirs & target & pin;
                const bool seventh = (sq >> 3) == (us == 0 ? 6 : 1);
                cat = seventh ? 3 : 2; kind = seventh ? 1 : 0;
            } else if (p.knight & sb) { tg = (blockers & sb) ? 0ULL : (N & target); cat = 7; }
            else if (p.bishop & sb) { tg = B & target & pin; cat = 8; }
            else if (p.rook & sb) { tg = R & target & pin; cat = 9; }
            else { tg = (R | B) & target & pin; cat = 10; }
            count = popc(tg) * (kind == 1 ? 4 : 1);
        }
    } else if (lane == 16 || lane == 17 || lane == 18) {
        // pawn pushes as sets (ordered by destination)
        const u64 pawns = p.pawn & ours;
        const u64 pushers = pawns & ~(blockers & ~file_mask(ksq));
        const u64 single_p = (us == 0 ? pushers << 8 : pushers >> 8) & ~occ;
        delta = us == 0 ? 8 : -8;
        if (lane == 16) { tg = single_p & target & ~BACKRANKS; cat = 4; kind = 2; count = popc(tg); }
        else if (lane == 17) { tg = single_p & target & BACKRANKS; cat = 5; kind = 3; count = 4 * popc(tg); }
        else {
            tg = (us == 0 ? single_p << 8 : single_p >> 8) & (us == 0 ? 0x00000000FF000000ULL : 0x000000FF00000000ULL) & ~occ & target;
            cat = 6; kind = 2; delta *= 2; count = popc(tg);
        }
    } else if (lane == 19) {
        tg = king_targets; from = ksq; cat = in_check ? 1 : 11; count = popc(tg);
    } else if (lane == 20) {
        if (castle_k) { cat = 12; kind = 4; from = ksq; special_to = base + 7; count = 1; }
    } else if (lane == 21) {
        if (castle_q) { cat = 13; kind = 4; from = ksq; special_to = base; count = 1; }
    } else if (lane == 22) {
        if (e1) { cat = 0; kind = 4; from = epf1; special_to = ep; count = 1; }
    } else if (lane == 23) {
        if (e2) { cat = 0; kind = 4; from = epf2; special_to = ep; count = 1; }
    }
    // ---- phase D: offsets in (category, from, lane) order
    const int key = (cat << 12) | (from << 5) | lane;
    int offset = 0, total = 0;
#pragma unroll
    for (int j = 0; j < 24; j++) {
        const int kj = __shfl_sync(0xffffffffu, key, j), cj = __shfl_sync(0xffffffffu, count, j);
        if (kj < key) offset += cj;
        total += cj;
    }
    // ---- phase E: emit
    if (count) {
        uint16_t* out = moves + offset;
        if (kind == 0) { for (u64 t = tg; t; t &= t - 1) *out++ = mk_move(from, lsb(t), 0, 0); }
        else if (kind == 1) {
            for (u64 t = tg; t; t &= t - 1) { const int to = lsb(t); *out++ = mk_move(from, to, 4, 0); *out++ = mk_move(from, to, 3, 0); *out++ = mk_move(from, to, 2, 0); *out++ = mk_move(from, to, 1, 0); }
        } else if (kind == 2) { for (u64 t = tg; t; t &= t - 1) { const int to = lsb(t); *out++ = mk_move(to - delta, to, 0, 0); } }
        else if (kind == 3) {
            for (u64 t = tg; t; t &= t - 1) { const int to = lsb(t), f = to - delta; *out++ = mk_move(f, to, 4, 0); *out++ = mk_move(f, to, 3, 0); *out++ = mk_move(f, to, 2, 0); *out++ = mk_move(f, to, 1, 0); }
        } else { *out = mk_move(from, special_to, 0, 1); }
    }
    __syncwarp();
    n_out = total;
    GenInfo gi;
    gi.checkers = checkers;
    gi.has_legal_ep = e1 || e2;
    return gi;
}

// shakmaty play_unchecked for a wire move (assumed legal)
__device__ __forceinline__ DPos make_move(const DPos& p, uint16_t mv) {
    const int from = mv & 63, to = (mv >> 6) & 63, promo = (mv >> 12) & 7, special = mv >> 15;
    const int us = meta_turn(p.meta);
    const u64 fb = bit(from), tb = bit(to);
    DPos n = p;
    int castling = meta_castling(p.meta);
    int halfmoves = meta_halfmoves(p.meta), fullmoves = meta_fullmoves(p.meta);
    int new_ep = -1;
    bool zeroing = false;
    if (special && (p.king & fb)) {
        const int base = us * 56;
        const bool qs = to < from;
        const u64 kto = bit(base + (qs ? 2 : 6)), rto = bit(base + (qs ? 3 : 5));
        n.king = (n.king ^ fb) | kto;
        n.rook = (n.rook ^ tb) | rto;
        if (us == 0) n.white = (n.white & ~(fb | tb)) | kto | rto;
        castling &= us == 0 ? ~3 : ~12;
    } else if (special) {
        const u64 cb = bit((from & 56) | (to & 7));
        n.pawn = (n.pawn ^ fb ^ cb) | tb;
        if (us == 0) n.white = (n.white ^ fb) | tb; else n.white &= ~cb;
        zeroing = true;
    } else {
        const u64 occ = occupied(p);
        if (occ & tb) {
            zeroing = true;
            n.pawn &= ~tb; n.knight &= ~tb; n.bishop &= ~tb; n.rook &= ~tb; n.queen &= ~tb; n.king &= ~tb;
            n.white &= ~tb;
        }
        if (p.pawn & fb) {
            zeroing = true;
            n.pawn ^= fb;
            if (promo == 0) n.pawn |= tb;
            else if (promo == 1) n.knight |= tb;
            else if (promo == 2) n.bishop |= tb;
            else if (promo == 3) n.rook |= tb;
            else n.queen |= tb;
            if ((to ^ from) == 16) new_ep = (from + to) >> 1;
        } else if (p.knight & fb) n.knight ^= fb | tb;
        else if (p.bishop & fb) n.bishop ^= fb | tb;
        else if (p.rook & fb) n.rook ^= fb | tb;
        else if (p.queen & fb) n.queen ^= fb | tb;
        else { n.king ^= fb | tb; castling &= us == 0 ? ~3 : ~12; }
        if (us == 0) n.white = (n.white ^ fb) | tb;
        // a right disappears when its rook leaves or is captured on the corner
        const u64 ft = fb | tb;
        if (ft & bit(7)) castling &= ~1;
        if (ft & bit(0)) castling &= ~2;
        if (ft & bit(63)) castling &= ~4;
        if (ft & bit(56)) castling &= ~8;
    }
    halfmoves = zeroing ? 0 : min(halfmoves + 1, 0xFFFF);
    if (us == 1) fullmoves = min(fullmoves + 1, 0xFFFF);
    n.meta = (u64)(us ^ 1) | ((u64)castling << 1) | ((u64)(new_ep + 1) << 8) | ((u64)halfmoves << 16) | ((u64)fullmoves << 32);
    return n;
}

// shakmaty is_insufficient_material (both sides)
__device__ __forceinline__ bool side_insufficient(const DPos& p, u64 ours, u64 theirs) {
    if (ours & (p.pawn | p.rook | p.queen)) return false;
    if (ours & p.knight) return popc(ours) <= 2 && (theirs & ~p.king & ~p.queen) == 0;
    if (ours & p.bishop) {
        const u64 DARK = 0xAA55AA55AA55AA55ULL;
        bool same = (p.bishop & DARK) == 0 || (p.bishop & ~DARK) == 0;
        return same && p.knight == 0 && p.pawn == 0;
    }
    return true;
}
__device__ __forceinline__ bool insufficient_material(const DPos& p) {
    u64 occ = occupied(p), black = occ ^ p.white;
    return side_insufficient(p, p.white, black) && side_insufficient(p, black, p.white);
}

__device__ __forceinline__ int pseudo_legal_ep(const DPos& p) {
    int ep = meta_ep(p.meta);
    if (ep < 0) return -1;
    int us = meta_turn(p.meta);
    u64 occ = occupied(p);
    u64 ours = us == 0 ? p.white : occ ^ p.white;
    return (pawn_attacks_bb(us ^ 1, bit(ep)) & p.pawn & ours) ? ep : -1;
}

// ---- the identity of a position as Fen::from_position(.., EnPassantMode::PseudoLegal) sees it (tree.rs:214, memory.rs:42):
// board, turn, castling rights, pseudo-legal ep square, halfmove clock, fullmove number.  Keys of the evaluation cache and
// of the replay buffer.
__device__ __forceinline__ u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ DPos fen_key_of(const DPos& p) {
    DPos k = p;
    k.meta = (p.meta & 0x1FULL) | ((u64)(pseudo_legal_ep(p) + 1) << 8) | (p.meta & (0xFFFFFFFFULL << 16));
    return k;
}
__device__ __forceinline__ u64 fen_key_hash(const DPos& k) {
    u64 h = splitmix64(k.pawn);
    h = splitmix64(h ^ k.knight); h = splitmix64(h ^ k.bishop); h = splitmix64(h ^ k.rook); h = splitmix64(h ^ k.queen);
    h = splitmix64(h ^ k.king); h = splitmix64(h ^ k.white); h = splitmix64(h ^ k.meta);
    return h;
}
__device__ __forceinline__ bool fen_key_equal(const DPos& a, const DPos& b) {
    return a.pawn == b.pawn && a.knight == b.knight && a.bishop == b.bishop && a.rook == b.rook && a.queen == b.queen &&
           a.king == b.king && a.white == b.white && a.meta == b.meta;
}

// stores the LEGAL ep square in the key bits of meta (chess.rs:52: HashMap<Chess,_> equality uses it)
__device__ __forceinline__ void set_key_bits(DPos& p, bool has_legal_ep) {
    int lep = has_legal_ep ? meta_ep(p.meta) : -1;
    p.meta = (p.meta & ~(0xFFULL << 48)) | ((u64)(lep + 1) << 48) | (1ULL << 55);
}
__device__ __forceinline__ bool same_position_key(const DPos& a, const DPos& b) {
    return a.pawn == b.pawn && a.knight == b.knight && a.bishop == b.bishop && a.rook == b.rook && a.queen == b.queen &&
           a.king == b.king && a.white == b.white && ((a.meta ^ b.meta) & META_KEY_MASK) == 0;
}

// move_to_index (chess.rs:73-116)
__device__ __forceinline__ int move_to_index(uint16_t mv, int turn) {
    int from = mv & 63, to = (mv >> 6) & 63;
    int file = from & 7, rank = turn ? 7 - (from >> 3) : from >> 3;
    int df = (to & 7) - file, dr = (turn ? 7 - (to >> 3) : to >> 3) - rank;
    int plane;
    if (df == 1 && dr == 2) plane = 0; else if (df == 2 && dr == 1) plane = 1;
    else if (df == 2 && dr == -1) plane = 2; else if (df == 1 && dr == -2) plane = 3;
    else if (df == -1 && dr == -2) plane = 4; else if (df == -2 && dr == -1) plane = 5;
    else if (df == -2 && dr == 1) plane = 6; else if (df == -1 && dr == 2) plane = 7;
    else if (df == 0) plane = dr > 0 ? 7 + dr : 35 - dr;
    else if (dr == 0) plane = df > 0 ? 21 + df : 49 - df;
    else if (df > 0) plane = dr > 0 ? 14 + dr : 28 + df;
    else plane = dr < 0 ? 42 - dr : 56 - df;
    return plane * 64 + rank * 8 + file;
}

// index_to_move (chess.rs:118-171) resolved against a generated legal-move list: the reference goes through a UCI
// string and shakmaty's legality check; queen is the only promotion reachable (chess.rs:165-167).
// returns the wire move or 0xFFFF (None)
__device__ __forceinline__ uint16_t index_to_move(int index, const DPos& p, const uint16_t* legal, int n_legal) {
    const int turn = meta_turn(p.meta);
    int plane = index >> 6, file = index & 7, crank = (index >> 3) & 7;
    int from_rank = turn ? 7 - crank : crank;
    int df, dr;
    if (plane < 8) {
        const int kdf[8] = {1, 2, 2, 1, -1, -2, -2, -1}, kdr[8] = {2, 1, -1, -2, -2, -1, 1, 2};
        df = kdf[plane]; dr = kdr[plane];
    } else if (plane < 15) { df = 0; dr = plane - 7; }
    else if (plane < 22) { df = plane - 14; dr = plane - 14; }
    else if (plane < 29) { df = plane - 21; dr = 0; }
    else if (plane < 36) { df = plane - 28; dr = 28 - plane; }
    else if (plane < 43) { df = 0; dr = 35 - plane; }
    else if (plane < 50) { df = 42 - plane; dr = 42 - plane; }
    else if (plane < 57) { df = 49 - plane; dr = 0; }
    else { df = 56 - plane; dr = plane - 56; }
    if (turn) dr = -dr;
    int dest_file = file + df, dest_rank = from_rank + dr;
    if (dest_file < 0 || dest_file > 7 || dest_rank < 0 || dest_rank > 7) return 0xFFFF;
    int from = from_rank * 8 + file, to = dest_rank * 8 + dest_file;
    const u64 fb = bit(from);
    const u64 occ = occupied(p);
    if (!(occ & fb)) return 0xFFFF;
    // UciMove::to_move: a king moving two files from e1/e8 along the back rank also denotes castling
    int alt_to = -1;
    if ((p.king & fb) && from == turn * 56 + 4 && (to >> 3) == (from >> 3) && (dest_file - file == 2 || file - dest_file == 2))
        alt_to = dest_file > file ? turn * 56 + 7 : turn * 56;
    const bool promo = (p.pawn & fb) && (dest_rank == 0 || dest_rank == 7);
    for (int i = 0; i < n_legal; i++) {
        uint16_t m = legal[i];
        int mf = m & 63, mt = (m >> 6) & 63, mp = (m >> 12) & 7, sp = m >> 15;
        if (mf != from) continue;
        if (mt == to && (promo ? mp == 4 : mp == 0)) return m;
        if (alt_to >= 0 && sp && mt == alt_to) return m;
    }
    return 0xFFFF;
}

// to_tensor (chess.rs:191-245): value of element [plane][rank][file] (rank already flipped for Black)
__device__ __forceinline__ float plane_value(const DPos& p, int plane, int sq_canon, int pseudo_ep, u64 ours, u64 theirs) {
    const int turn = meta_turn(p.meta);
    const int sq = turn ? (sq_canon ^ 56) : sq_canon;
    if (plane < 12) {
        const u64 side = plane < 6 ? ours : theirs;
        const int r = plane < 6 ? plane : plane - 6;
        const u64 bb = r == 0 ? p.pawn : r == 1 ? p.knight : r == 2 ? p.bishop : r == 3 ? p.rook : r == 4 ? p.queen : p.king;
        return (bb & side & bit(sq)) ? 1.0f : 0.0f;
    }
    if (plane < 16) {
        const int c = meta_castling(p.meta);
        const int us_r = (c >> (turn * 2)) & 3, th_r = (c >> ((turn ^ 1) * 2)) & 3;
        const int v = plane == 12 ? (us_r & 1) : plane == 13 ? (us_r >> 1) : plane == 14 ? (th_r & 1) : (th_r >> 1);
        return v ? 1.0f : 0.0f;
    }
    if (plane == 16) return sq == pseudo_ep ? 1.0f : 0.0f;
    if (plane == 17) return __fdiv_rn((float)meta_halfmoves(p.meta), 100.0f);
    return __fdiv_rn((float)meta_fullmoves(p.meta), 200.0f);
}

#endif  // __CUDACC__

}  // namespace azb
