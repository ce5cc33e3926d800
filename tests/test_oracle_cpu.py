"""The oracle against the golden vectors the path has (public perft tables, hand-derived codec vectors) and its own invariants."""
import numpy as np
import pytest

from helpers import SPECIAL_FENS, load_golden, orc, random_playouts, uci_to_wire


@pytest.mark.parametrize("entry", load_golden("perft.json")["positions"], ids=lambda e: e["name"])
def test_perft_known_answers(entry):
    pos = orc.from_fen(entry["fen"])
    budget = 6_000_000
    for depth, expected in enumerate(entry["nodes"], 1):
        if expected > budget:
            break
        assert orc.perft(pos, depth) == expected, (entry["name"], depth)


@pytest.mark.parametrize("vec", load_golden("codec.json")["vectors"], ids=lambda v: v["uci"])
def test_codec_vectors(vec):
    pos = orc.from_fen(vec["fen"])
    mv = uci_to_wire(pos, vec["uci"])
    assert orc.move_to_index(pos, mv) == vec["index"]
    back = orc.index_to_move(pos, vec["index"])
    assert back == mv


def test_codec_round_trip_and_underpromotion_collapse():
    positions, _ = random_playouts(300, seed=7)
    for pos in list(positions) + [orc.from_fen(f) for f in SPECIAL_FENS]:
        mv, ix = orc.legal_moves(pos)
        for m, i in zip(mv, ix):
            back = orc.index_to_move(pos, int(i))
            promo = (int(m) >> 12) & 7
            if promo in (0, 4):
                assert back == m
            else:  # under-promotions share the queen promotion's index (chess.rs:165-167)
                assert back == (int(m) & ~(7 << 12)) | (4 << 12)
        legal = set(int(i) for i in ix)
        for i in range(0, 4096, 37):
            if i not in legal:
                b = orc.index_to_move(pos, i)
                # the only non-listed indices that still decode are the king-two-squares spellings of castling
                assert b is None or (b >> 15) == 1


def test_planes():
    pos = orc.from_fen("rnbqkbnr/pppp1ppp/8/8/4pP2/8/PPPPP1PP/RNBQKBNR b KQkq f3 7 150")
    t = orc.to_tensor(pos)
    assert t.shape == (19, 8, 8)
    assert t[:12].sum() == 32
    assert t[0].sum() == 8 and t[6].sum() == 8
    # black to move: ranks are flipped, black pawn e4 -> canonical rank 4 (7-3), file 4
    assert t[0, 4, 4] == 1.0
    assert (t[12:16] == 1.0).all()
    assert t[16].sum() == 1.0 and t[16, 7 - 2, 5] == 1.0  # f3 is capturable by the e4 pawn
    assert np.all(t[17] == np.float32(7) / np.float32(100))
    assert np.all(t[18] == np.float32(150) / np.float32(200))
    # ep square that no pawn can capture is not shown (pseudo_legal_ep_square)
    pos2 = orc.from_fen("rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 1")
    assert orc.to_tensor(pos2)[16].sum() == 0.0


def test_outcomes_and_play_move_rules():
    assert orc.outcome(orc.from_fen("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1")) == 1
    assert orc.outcome(orc.from_fen("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1")) == 2
    assert orc.outcome(orc.from_fen("8/8/8/8/8/2k5/8/K1n5 w - - 0 1")) == 1
    assert orc.outcome(orc.from_fen("8/8/8/8/8/2k5/8/K1nn4 w - - 0 1")) == 0
    # threefold repetition by knight shuffles (chess.rs:52-60)
    pos = orc.startpos()
    hist = [pos.copy()]
    res = None
    for uci in ["g1f3", "g8f6", "f3g1", "f6g8", "g1f3", "g8f6", "f3g1", "f6g8"]:
        mv = uci_to_wire(pos, uci)
        idx = orc.move_to_index(pos, mv)
        pos, res = orc.play_move(pos, idx, np.array(hist, orc.POSITION_DTYPE))
        hist.append(pos.copy())
        if res != 0:
            break
    assert res == 1 and len(hist) == 9
    # illegal index
    _, r = orc.play_move(orc.startpos(), 0)
    assert r == -1
    # 50-move and 200-fullmove limits
    p = orc.from_fen("4k3/8/8/8/8/8/8/R3K3 w - - 99 60")
    _, r = orc.play_move(p, orc.move_to_index(p, uci_to_wire(p, "a1a2")))
    assert r == 1
    p = orc.from_fen("4k3/8/8/8/8/8/8/R3K3 b - - 0 199")
    _, r = orc.play_move(p, orc.move_to_index(p, uci_to_wire(p, "e8e7")))
    assert r == 1


def test_rng_and_dirichlet():
    L = orc.lib()
    assert L.orc_rng_u64(1, 2, 3, 4, 5) == L.orc_rng_u64(1, 2, 3, 4, 5)
    assert L.orc_rng_u64(1, 2, 3, 4, 5) != L.orc_rng_u64(1, 2, 3, 4, 6)
    xs = np.array([1e-12, 0.003, 0.5, 0.999, 1.0, 1.5, 7.0, 1e9])
    assert np.allclose([L.orc_det_log(float(x)) for x in xs], np.log(xs), rtol=1e-14, atol=1e-15)
    ys = np.array([-40.0, -3.3, -1e-9, 0.0, 0.7, 12.5])
    assert np.allclose([L.orc_det_exp(float(y)) for y in ys], np.exp(ys), rtol=1e-13)
    d = orc.dirichlet(42, 0, 0, 0.3, 20)
    assert abs(d.sum() - 1.0) < 1e-5 and (d >= 0).all()
    means = np.mean([orc.dirichlet(42, g, 0, 0.3, 20) for g in range(400)], axis=0)
    assert np.abs(means - 0.05).max() < 0.02
    # Gamma(0.3): E[x_i^2] of Dirichlet(0.3 x 20) = a(a+1)/(A(A+1)) = 0.39/42
    second = np.mean([orc.dirichlet(42, g, 1, 0.3, 20) ** 2 for g in range(2000)])
    assert abs(second - 0.39 / 42) < 0.0015


def test_mcts_invariants():
    prm = orc.make_params(num_simulations=64)
    ev = orc.make_evaluator("stub", stub_seed=5)
    root = orc.startpos()
    visits, scores, depth, evals = orc.search(root, prm, ev)
    assert visits.sum() == 64 and depth >= 1 and evals <= 65
    _, ix = orc.legal_moves(root)
    assert set(np.nonzero(visits)[0]) <= set(int(i) for i in ix)
    # first simulation picks argmax P over legal moves (q = 0, u = 3 P)
    prm1 = orc.make_params(num_simulations=1)
    v1, _, _, _ = orc.search(root, prm1, ev)
    pol, _ = orc.stub_eval(5, root)
    legal = np.array([int(i) for i in ix])
    assert np.argmax(v1) == legal[np.argmax(pol[legal])]
    # deterministic
    v2, s2, _, _ = orc.search(root, prm, ev)
    assert np.array_equal(visits, v2) and np.array_equal(scores, s2)
    # noise changes the search, keyed by (game, ply)
    vn, _, _, _ = orc.search(root, prm, ev, noise_game=3, noise_ply=0)
    vn2, _, _, _ = orc.search(root, prm, ev, noise_game=3, noise_ply=0)
    assert np.array_equal(vn, vn2) and vn.sum() == 64


def test_selfplay_episode_terminates_and_backfills():
    prm = orc.make_params(num_simulations=16)
    ev = orc.make_evaluator("stub", stub_seed=9)
    ep = orc.selfplay_episode(prm, ev, game_id=1, max_steps=512)
    st = ep["stats"]
    assert st.result in (1, 2, 3) and st.n_steps >= 2
    assert all(v.sum() == 16 for v in ep["visits"])
    fm = None
    # final values: +-(1 - fullmoves/400) with the sign of the side to move, or 0 for a draw
    fv = ep["final_value"]
    turn = np.where(ep["positions"]["turn"] == 0, 1.0, -1.0)
    if st.result == 1:
        assert np.all(fv == 0)
    else:
        winner = 1.0 if st.result == 2 else -1.0
        assert np.all(np.sign(fv) == turn * winner)
        assert np.all(np.abs(fv) == np.abs(fv[0])) and 0.5 <= abs(fv[0]) <= 1.0
    # with the shared evaluator cache the record is identical (the cache is semantically transparent)
    c = orc.cache_create()
    ep2 = orc.selfplay_episode(prm, ev, game_id=1, max_steps=512, cache=c)
    orc.cache_destroy(c)
    assert np.array_equal(ep["action"], ep2["action"]) and np.array_equal(ep["visits"], ep2["visits"])
    assert ep2["stats"].cache_hits > 0


def test_minimax_hand_derived_scores():
    """chess.rs:247-318 restated: values worked out by hand (material 100/320/330/500/900, mate 20000 + remaining depth)."""
    start = orc.from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1")
    for d in (1, 2, 3):
        assert not orc.minimax_scores(start, d).any() and len(orc.minimax_scores(start, d)) == 20
    mate = orc.from_fen("6k1/5ppp/8/8/8/8/8/R6K w - - 0 1")           # Ra8 mates; otherwise rook against three pawns
    mv, _ = orc.legal_moves(mate)
    s1, s2 = orc.minimax_scores(mate, 1), orc.minimax_scores(mate, 2)
    ra8 = [k for k, m in enumerate(mv) if (m & 63) == 0 and ((m >> 6) & 63) == 56][0]
    assert s1[ra8] == 20000 and s2[ra8] == 20001
    assert all(s1[k] == 200 for k in range(len(mv)) if k != ra8)
    kk = orc.from_fen("8/8/4k3/8/8/4K3/8/8 w - - 0 1")                 # dead position after any move: 0, never material
    assert not orc.minimax_scores(kk, 3).any()
    q = orc.from_fen("4k3/8/8/3q4/4P3/8/8/4K3 w - - 0 1")              # exd5 wins the queen; anything else loses the pawn
    mv, _ = orc.legal_moves(q)
    cap = [k for k, m in enumerate(mv) if (m & 63) == 28 and ((m >> 6) & 63) == 35][0]
    s1, s2 = orc.minimax_scores(q, 1), orc.minimax_scores(q, 2)
    assert s1[cap] == 100 and s2[cap] == 100
    assert all(s1[k] == -800 and s2[k] == -900 for k in range(len(mv)) if k != cap)
