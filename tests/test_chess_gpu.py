"""chess.rs entry points of the CUDA library against the oracle (bit-exact: move lists in order, indices, positions, planes)."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import SPECIAL_FENS, load_golden, orc, random_playouts, uci_to_wire

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = az.Engine(max_games=64, max_batch=8192, num_simulations=32)
    yield e
    e.close()


@pytest.fixture(scope="module")
def corpus():
    positions, histories = random_playouts(3000, seed=42, max_plies=120)
    special = np.array([orc.from_fen(f) for f in SPECIAL_FENS], orc.POSITION_DTYPE)
    return np.concatenate([special, positions]), [np.array([s], orc.POSITION_DTYPE) for s in special] + histories


def test_position_helpers_match_oracle():
    for fen in SPECIAL_FENS:
        a, b = az.position_from_fen(fen), orc.from_fen(fen)
        assert a.tobytes() == b.tobytes()
    assert az.start_position().tobytes() == orc.startpos().tobytes()


def test_movegen_lists_identical(eng, corpus):
    positions, _ = corpus
    moves, index, count = eng.movegen(positions)
    for i, pos in enumerate(positions):
        mv, ix = orc.legal_moves(pos)
        assert count[i] == len(mv), i
        assert np.array_equal(moves[i, : count[i]], mv), i
        assert np.array_equal(index[i, : count[i]], ix), i
        assert np.all(moves[i, count[i]:] == az.MOVE_NONE)


def test_warp_cooperative_generator_matches(eng, corpus):
    """The search kernel's warp-cooperative generator must emit exactly the serial generator's list (and the oracle's)."""
    import ctypes

    positions, _ = corpus
    extra = ["8/8/8/8/1pPp4/8/8/K6k b - c3 0 1",            # two pawns can capture en passant
             "4k3/8/8/8/1pPp4/8/8/K3R3 b - c3 0 1",          # ... while in check from the rook: neither resolves it
             "k7/8/8/8/8/8/1p1p1p2/R1B1N2K b - - 0 1",       # promotions with and without capture, several pawns
             "r3k2r/8/8/8/8/8/8/R3K2R b KQkq - 0 1", "4k3/8/8/8/8/8/8/R3K2R w KQ - 0 1",
             "3rkr2/8/8/8/8/8/8/R3K2R w KQ - 0 1",            # both castling paths attacked
             "4k3/8/8/8/8/5n2/8/R3K2R w KQ - 0 1",            # knight check: no castling
             "Q2k4/8/8/8/8/8/8/4K3 b - - 0 1", "3k4/3Q4/8/8/8/8/8/3RK3 b - - 0 1"]
    pos = np.concatenate([positions, np.array([orc.from_fen(f) for f in extra], orc.POSITION_DTYPE)])
    n = len(pos)
    moves = np.empty((n, az.MAX_MOVES), np.uint16)
    count = np.empty(n, np.int32)
    p = np.ascontiguousarray(pos, az.POSITION_DTYPE)
    rc = az.lib().az_dbg_movegen_warp(eng._h, n, ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(moves.ctypes.data),
                                      ctypes.c_void_p(count.ctypes.data))
    assert rc == 0
    ref_moves, _, ref_count = eng.movegen(pos)
    assert np.array_equal(count, ref_count)
    assert np.array_equal(moves, ref_moves)
    for i in range(len(positions), n):
        mv, _ = orc.legal_moves(pos[i])
        assert np.array_equal(moves[i, : count[i]], mv), i


def test_movegen_empty_and_ragged(eng):
    m, ix, c = eng.movegen(np.zeros(0, az.POSITION_DTYPE))
    assert m.shape == (0, 256) and c.shape == (0,)
    one = eng.movegen(orc.from_fen("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1"))
    assert one[2][0] == 0
    with pytest.raises(az.EngineError):
        eng.movegen(np.zeros(8193, az.POSITION_DTYPE))


@pytest.mark.parametrize("entry", load_golden("perft.json")["positions"], ids=lambda e: e["name"])
def test_perft_known_answers(eng, entry):
    pos = az.position_from_fen(entry["fen"])
    for depth, expected in enumerate(entry["nodes"], 1):
        if expected > 250_000_000:
            break
        assert int(eng.perft(pos, depth)[0]) == expected, (entry["name"], depth)


def test_perft_batch_of_roots(eng, corpus):
    positions, _ = corpus
    roots = positions[:512]
    got = eng.perft(roots, 2)
    want = orc.perft_batch(roots, 2, 8)
    assert np.array_equal(got, want)
    got3 = eng.perft(roots[:64], 3)
    assert np.array_equal(got3, orc.perft_batch(roots[:64], 3, 8))


def test_play_move_matches_oracle(eng, corpus):
    positions, histories = corpus
    rng = np.random.default_rng(1)
    sel, actions = [], []
    for i, pos in enumerate(positions[:1500]):
        _, ix = orc.legal_moves(pos)
        if len(ix) == 0:
            continue
        a = int(ix[rng.integers(len(ix))]) if rng.random() < 0.9 else int(rng.integers(4096))
        sel.append(i)
        actions.append(a)
    pos_in = positions[sel]
    hist = np.concatenate([histories[i] for i in sel])
    offs = np.zeros(len(sel) + 1, np.uint32)
    offs[1:] = np.cumsum([len(histories[i]) for i in sel])
    new_pos, res = eng.play_move(pos_in, actions, hist, offs)
    n_illegal = 0
    for k, i in enumerate(sel):
        want_pos, want_res = orc.play_move(positions[i], actions[k], histories[i])
        assert res[k] == want_res, (k, actions[k])
        if want_res == -1:
            n_illegal += 1
            assert new_pos[k].tobytes() == positions[i].tobytes()
        else:
            assert new_pos[k].tobytes() == want_pos.tobytes(), k
    assert n_illegal > 0


def test_play_move_repetition_and_limits(eng):
    pos = orc.startpos()
    hist = [pos.copy()]
    res = None
    for uci in ["g1f3", "g8f6", "f3g1", "f6g8", "g1f3", "g8f6", "f3g1", "f6g8"]:
        idx = orc.move_to_index(pos, uci_to_wire(pos, uci))
        h = np.array(hist, az.POSITION_DTYPE)
        new_pos, r = eng.play_move(pos, [idx], h, np.array([0, len(hist)], np.uint32))
        pos, res = new_pos[0], int(r[0])
        hist.append(pos.copy())
        if res != 0:
            break
    assert res == az.DRAW and len(hist) == 9
    for fen, uci in [("4k3/8/8/8/8/8/8/R3K3 w - - 99 60", "a1a2"), ("4k3/8/8/8/8/8/8/R3K3 b - - 0 199", "e8e7")]:
        p = orc.from_fen(fen)
        _, r = eng.play_move(p, [orc.move_to_index(p, uci_to_wire(p, uci))])
        assert r[0] == az.DRAW


def test_codec(eng, corpus):
    for vec in load_golden("codec.json")["vectors"]:
        pos = orc.from_fen(vec["fen"])
        mv = uci_to_wire(pos, vec["uci"])
        assert eng.move_to_index(pos, [mv])[0] == vec["index"]
        assert eng.index_to_move(pos, [vec["index"]])[0] == mv
    positions, _ = corpus
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 4096, len(positions)).astype(np.uint16)
    got = eng.index_to_move(positions, idx)
    for i, pos in enumerate(positions):
        want = orc.index_to_move(pos, int(idx[i]))
        assert got[i] == (az.MOVE_NONE if want is None else want), i
    # the king-two-squares spelling of castling (UciMove::to_move)
    p = orc.from_fen("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1")
    assert eng.index_to_move(p, [23 * 64 + 4])[0] == uci_to_wire(p, "e1h1")


def test_planes_bit_exact(eng, corpus):
    positions, _ = corpus
    planes = eng.encode(positions)
    for i, pos in enumerate(positions):
        assert np.array_equal(planes[i], orc.to_tensor(pos)), i
