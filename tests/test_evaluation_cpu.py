"""ratings.rs Elo fit and the move-selection helpers of validation.rs (host math, no GPU)."""
import numpy as np

from alphazero_chess_b200 import evaluation as ev


def test_compute_elos_two_players():
    # a 75 % score corresponds to 400*log10(3) = 190.8 Elo; player 0 is the anchor (ratings.rs:122)
    elos = ev.compute_elos([[0.5, 0.25], [0.75, 0.5]], 1000.0)
    assert elos[0] == 1000.0 and abs(float(elos[1]) - (1000.0 + 400.0 * np.log10(3.0))) < 0.5


def test_compute_elos_three_players_ordering():
    m = np.array([[0.5, 0.4, 0.2], [0.6, 0.5, 0.35], [0.8, 0.65, 0.5]], np.float32)
    elos = ev.compute_elos(m, 0.0)
    assert elos[0] == 0.0 and elos[0] < elos[1] < elos[2]
    # fixed point: expected score equals actual score for the non-anchor players
    for i in (1, 2):
        exp = sum(1.0 / (1.0 + 10.0 ** ((elos[j] - elos[i]) / 400.0)) for j in range(3) if j != i)
        assert abs(exp - (m[i].sum() - 0.5)) < 1e-2


def test_move_selection_helpers():
    w = np.zeros(4096, np.float32)
    w[[5, 100, 4000]] = [0.25, 0.5, 0.25]
    assert ev._last_argmax(w) == 100
    w2 = w.copy(); w2[4000] = 0.5
    assert ev._last_argmax(w2) == 4000                 # ties: the LAST maximum (Iterator::max_by)
    assert ev._weighted_index(w, 0.0) == 5 and ev._weighted_index(w, 0.24) == 5
    assert ev._weighted_index(w, 0.26) == 100 and ev._weighted_index(w, 0.76) == 4000 and ev._weighted_index(w, 0.999999) == 4000
