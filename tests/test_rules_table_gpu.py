"""The CUDA engine (az_play_move, az_encode, and the search kernel's own rule path) against the hand-derived table
tests/golden/rules.json -- the same vectors the oracle is pinned with in tests/test_rules_table_cpu.py."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import load_golden, orc, uci_to_wire

pytestmark = pytest.mark.gpu
RULES = load_golden("rules.json")


@pytest.fixture(scope="module")
def eng():
    e = az.Engine(max_games=64, num_simulations=64)
    e.set_evaluator_stub(1, 5)
    yield e
    e.close()


def test_play_move_table(eng):
    cases = RULES["play_move"]
    pos = np.array([az.position_from_fen(c["fen"]) for c in cases], az.POSITION_DTYPE)
    idx = [orc.move_to_index(p, uci_to_wire(p, c["uci"])) for p, c in zip(pos, cases)]
    offs = np.arange(len(cases) + 1, dtype=np.uint32)
    _, res = eng.play_move(pos, idx, pos, offs)
    for c, r in zip(cases, res):
        assert r == c["result"], (c["fen"], c["uci"], int(r), c["why"])


def test_repetition_sequences(eng):
    for seq in RULES["sequences"]:
        pos = az.position_from_fen(seq["start"])
        hist = [pos.copy()]
        for ply, (uci, want) in enumerate(zip(seq["moves"], seq["results"])):
            idx = orc.move_to_index(pos, uci_to_wire(pos, uci))
            h = np.array(hist, az.POSITION_DTYPE)
            new_pos, res = eng.play_move(pos, [idx], h, np.array([0, len(h)], np.uint32))
            assert res[0] == want, (seq["name"], ply, uci, int(res[0]))
            pos = new_pos[0]
            hist.append(pos.copy())


def test_search_sees_the_same_draws(eng):
    """The search kernel has its own copy of the rules (draw_by_rules over history + path): one move before each sequence's
    drawn position, the drawing move's child is terminal, so it is never expanded -- its subtree depth stays 0 and its
    accumulated score is exactly 0 (value 0.0 per visit) -- and the whole search equals the oracle's."""
    for seq in RULES["sequences"]:
        pos = az.position_from_fen(seq["start"])
        hist = [pos.copy()]
        for uci in seq["moves"][:-1]:
            pos = orc.play_encoded(pos, uci_to_wire(pos, uci))
            hist.append(pos.copy())
        h = np.array(hist, az.POSITION_DTYPE)
        visits, scores, depth = eng.search(pos, num_simulations=64, history=h, hist_offsets=np.array([0, len(h)], np.uint32), want_scores=True)
        v, s, d, _ = orc.search(pos, orc.make_params(num_simulations=64), orc.make_evaluator("stub", stub_seed=5), history=h)
        assert np.array_equal(visits[0], v) and np.array_equal(scores[0], s) and depth[0] == d, seq["name"]
        idx = orc.move_to_index(pos, uci_to_wire(pos, seq["moves"][-1]))
        assert scores[0][idx] == 0.0, seq["name"]


def test_ep_plane(eng):
    for case in RULES["ep_plane"]:
        planes = eng.encode(az.position_from_fen(case["fen"]))[0]
        want = np.zeros((8, 8), np.float32)
        if case["plane16"]:
            want[case["plane16"][0], case["plane16"][1]] = 1.0
        assert np.array_equal(planes[16], want), (case["fen"], case["why"])


def test_move_order(eng):
    def uci(mv):
        f, t, promo = mv & 63, (mv >> 6) & 63, (mv >> 12) & 7
        return "abcdefgh"[f & 7] + str((f >> 3) + 1) + "abcdefgh"[t & 7] + str((t >> 3) + 1) + ["", "n", "b", "r", "q"][promo]

    cases = RULES["move_order"]
    pos = np.array([az.position_from_fen(c["fen"]) for c in cases], az.POSITION_DTYPE)
    moves, _, count = eng.movegen(pos)
    for i, c in enumerate(cases):
        assert [uci(int(m)) for m in moves[i, : count[i]]] == c["moves"], (c["fen"], c["why"])
