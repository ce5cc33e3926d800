"""The oracle against tests/golden/rules.json: hand-derived expectations (from the published rules of shakmaty 0.29.0 and
chess.rs:9-11,36-63,232) for insufficient material, mate / stalemate, the counter limits and the repetition key."""
import numpy as np

from helpers import load_golden, orc, uci_to_wire

RULES = load_golden("rules.json")


def test_play_move_table():
    for case in RULES["play_move"]:
        pos = orc.from_fen(case["fen"])
        idx = orc.move_to_index(pos, uci_to_wire(pos, case["uci"]))
        _, res = orc.play_move(pos, idx, np.array([pos], orc.POSITION_DTYPE))
        assert res == case["result"], (case["fen"], case["uci"], res, case["why"])


def test_repetition_sequences():
    for seq in RULES["sequences"]:
        pos = orc.from_fen(seq["start"])
        hist = [pos.copy()]
        for ply, (uci, want) in enumerate(zip(seq["moves"], seq["results"])):
            idx = orc.move_to_index(pos, uci_to_wire(pos, uci))
            pos, res = orc.play_move(pos, idx, np.array(hist, orc.POSITION_DTYPE))
            assert res == want, (seq["name"], ply, uci, res)
            hist.append(pos.copy())


def test_ep_plane():
    for case in RULES["ep_plane"]:
        planes = orc.to_tensor(orc.from_fen(case["fen"]))
        want = np.zeros((8, 8), np.float32)
        if case["plane16"]:
            want[case["plane16"][0], case["plane16"][1]] = 1.0
        assert np.array_equal(planes[16], want), (case["fen"], case["why"])


def _uci(mv):
    f, t, promo = mv & 63, (mv >> 6) & 63, (mv >> 12) & 7
    return "abcdefgh"[f & 7] + str((f >> 3) + 1) + "abcdefgh"[t & 7] + str((t >> 3) + 1) + ["", "n", "b", "r", "q"][promo]


def test_move_order():
    for case in RULES["move_order"]:
        mv, _ = orc.legal_moves(orc.from_fen(case["fen"]))
        assert [_uci(int(m)) for m in mv] == case["moves"], (case["fen"], case["why"])
