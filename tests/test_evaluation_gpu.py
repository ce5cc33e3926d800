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
