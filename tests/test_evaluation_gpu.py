"""validation.rs evaluate on the engine: lockstep matches between players, colours alternating by game parity."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from alphazero_chess_b200 import evaluation as ev

pytestmark = pytest.mark.gpu


def test_matches_and_elo():
    w = az.random_weights(seed=42)
    with az.Engine(max_games=32, max_batch=64, num_simulations=24, seed=5) as e1, az.Engine(max_games=32, max_batch=64, num_simulations=24, seed=5) as e2:
        e1.load_weights(w)
        e2.load_weights(az.random_weights(seed=43))
        mcts, base, rnd = ev.MctsPlayer(e1), ev.BasePlayer(e2), ev.RandomPlayer(e1)
        r = ev.evaluate(mcts, rnd, e1, n_games=16, num_stochastic_moves=4, seed=1)
        assert r["unfinished"] == 0 and 0.0 <= r["winrate"] <= 1.0
        assert abs(r["p1_winrate"] + r["drawrate"] + r["p2_winrate"] - 1.0) < 1e-6
        assert abs(r["winrate"] - (r["p1_winrate"] + r["drawrate"] / 2)) < 1e-6
        r2 = ev.evaluate(mcts, rnd, e1, n_games=16, num_stochastic_moves=4, seed=1)
        assert np.array_equal(r["results"], r2["results"])           # keyed randomness: a match is reproducible
        rb = ev.evaluate(base, rnd, e1, n_games=8, num_stochastic_moves=2, seed=3)
        assert rb["unfinished"] == 0
        elos, matrix = ev.compute_elo_rankings([rnd, base], 1000.0, e1, n_games=8, seed=7)
        assert elos[0] == 1000.0 and matrix[0, 1] + matrix[1, 0] == pytest.approx(1.0)


def test_minimax_scores_match_oracle():
    """az_minimax against the oracle's restatement of chess.rs:247-318 on playout positions, mates and dead positions."""
    from helpers import orc, random_playouts
    roots, _ = random_playouts(40, seed=77, max_plies=90)
    fens = ["6k1/5ppp/8/8/8/8/8/R6K w - - 0 1",          # mate in one
            "8/8/4k3/8/8/4K3/8/8 w - - 0 1",             # bare kings: every child is a dead position
            "4k3/8/8/3q4/4P3/8/8/4K3 w - - 0 1",         # a hanging queen
            "7k/5Q2/6K1/8/8/8/8/8 w - - 0 1",            # mates and stalemates one ply away
            "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"]
    pos = np.concatenate([roots, np.array([orc.from_fen(f) for f in fens], az.POSITION_DTYPE).reshape(-1)])
    with az.Engine(max_games=64) as e:
        for depth in (1, 2, 3):
            scores, count = e.minimax(pos, depth)
            for i in range(len(pos)):
                want = orc.minimax_scores(pos[i], depth)
                assert count[i] == len(want), (i, depth)
                assert np.array_equal(scores[i, : count[i]], want), (i, depth)
                assert not scores[i, count[i]:].any()
        s4, c4 = e.minimax(pos[-5:-1], 4)
        for k, i in enumerate(range(len(pos) - 5, len(pos) - 1)):
            assert np.array_equal(s4[k, : c4[k]], orc.minimax_scores(pos[i], 4))
        mate = e.minimax(orc.from_fen(fens[0]), 2)[0][0]
        assert mate.max() == 20001                       # -( -20000 - remaining depth 1 )


def test_minimax_player_beats_random():
    with az.Engine(max_games=32) as e:
        mm, rnd = ev.MiniMaxPlayer(e, depth=2), ev.RandomPlayer(e)
        r = ev.evaluate(mm, rnd, e, n_games=12, seed=4)
        assert r["unfinished"] == 0
        assert r["winrate"] >= 0.6, r
        r2 = ev.evaluate(mm, rnd, e, n_games=12, seed=4)
        assert np.array_equal(r["results"], r2["results"])
